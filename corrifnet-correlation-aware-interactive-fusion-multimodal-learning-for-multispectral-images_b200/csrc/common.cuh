// Shared device/host helpers for libcorrif_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/corrif.h"

namespace corrif {

// ---- error plumbing -------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define CORRIF_REQUIRE(cond, ...)                         \
  do {                                                    \
    if (!(cond)) {                                        \
      ::corrif::set_last_error(__VA_ARGS__);              \
      return CORRIF_EINVAL;                               \
    }                                                     \
  } while (0)

// Returns the launch status of the kernel just enqueued (no sync).
static inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int num_sms();

// ---- warp reductions ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- 128-bit streaming accesses ---------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// read-once data: bypass L1 allocation
__device__ __forceinline__ float4 ld4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// ---- TF32 rounding --------------------------------------------------------------------------
// tcgen05.mma.kind::tf32 TRUNCATES fp32 operands to 10 mantissa bits (measured: signed relative bias
// -7e-4 on N(0,1) data).  Producers whose output is only read by tensor-core GEMMs therefore round to
// nearest at store time (cvt.rna.tf32), which removes the bias and halves the operand error.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 round_tf32_4(float4 v) {
  return make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
}

// ---- exact GELU (F.gelu default, erf form) and its derivative ---------------------------------
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// ---- Philox4x32-10 counter RNG: keep(seed, site, element) is a pure function -------------------
struct Philox4 { uint32_t x, y, z, w; };
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}
// Keep flags of the 4 consecutive elements [4*quad, 4*quad+4).  Element e is kept iff its 32-bit
// draw >= p * 2^32.
__device__ __forceinline__ void dropout_keep4(uint64_t seed, uint32_t site, uint64_t quad,
                                              uint32_t thresh, float scale, float (&m)[4]) {
  Philox4 r = philox4x32_10((uint32_t)quad, (uint32_t)(quad >> 32), site, 0x5EEDu,
                            (uint32_t)seed, (uint32_t)(seed >> 32));
  m[0] = r.x >= thresh ? scale : 0.f;
  m[1] = r.y >= thresh ? scale : 0.f;
  m[2] = r.z >= thresh ? scale : 0.f;
  m[3] = r.w >= thresh ? scale : 0.f;
}
static inline uint32_t dropout_threshold(float p) {
  double t = (double)p * 4294967296.0;
  if (t < 0) t = 0;
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}

}  // namespace corrif
