// Library-level entry points: ABI version, per-thread last error, device check.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace corrif {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace corrif

extern "C" {

int corrif_abi_version(void) { return CORRIF_ABI_VERSION; }

int corrif_sizeof_gemm_desc(void) { return (int)sizeof(corrif_gemm_desc); }

const char* corrif_last_error(void) { return corrif::g_last_error; }

int corrif_check_device(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    corrif::set_last_error("check_device: %s", cudaGetErrorString(e));
    return (int)e;
  }
  if (major != 10) {
    corrif::set_last_error("check_device: compute capability %d.x, kernels are sm_100a only", major);
    return CORRIF_EARCH;
  }
  return 0;
}

}  // extern "C"
