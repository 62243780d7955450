// Micro-benchmark: cycles per tcgen05.mma.kind::tf32 (M = 128, K = 8) instruction as a function of N and
// of where the A operand lives (shared memory "SS" vs tensor memory "TS").  One CTA per SM issues
// `iters` x 8 back-to-back MMAs into one accumulator, commits, waits, and reports clock64() deltas.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I<pkg>/csrc -o mma_probe tools/mma_probe.cu
// The operand contents are irrelevant (shared memory is left uninitialised); only the timing matters.
#include <cstdio>
#include <cstdlib>
#include "tc05.cuh"

namespace corrif { void set_last_error(const char*, ...) {} int num_sms() { return 148; } }
using namespace corrif::tc05;

template <int N, bool TS, bool B_MN>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t holder;
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&holder, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = holder;
  const uint32_t sA = sb, sB = sb + 64 * 1024;             // A: 128 x 64 fp32 (2 k-blocks), B: N x 64
  constexpr uint32_t idesc = idesc_tf32(N, false, B_MN);
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint64_t bd = B_MN ? smem_desc_mnmajor(sB + t * 1024, 64 * 128)
                                   : smem_desc_kmajor(sB + (t >> 2) * (N * 128) + (t & 3) * 32);
          if (TS) tcgen05_mma_tf32_ts(tmem, tmem + 256 + 8 * t, bd, idesc, 1u);
          else tcgen05_mma_tf32(tmem, smem_desc_kmajor(sA + (t >> 2) * (128 * 128) + (t & 3) * 32), bd, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) tcgen05_commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
    t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, bool TS, bool B_MN>
static void run(const char* name, long long* d_out) {
  const int iters = 2000, smem = 200 * 1024;
  cudaFuncSetAttribute(probe<N, TS, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148}) {
    probe<N, TS, B_MN><<<grid, 128, smem>>>(iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%-28s grid %3d: %7.1f cycles per MMA (128 x %3d x 8: %4.0f MAC/clk/SM)  %s\n", name, grid,
           (double)c / (iters * 8), N, 128.0 * N * 8 / ((double)c / (iters * 8)), e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long));
  run<64, false, false>("SS  N=64  B K-major", d_out);
  run<64, true, false>("TS  N=64  B K-major", d_out);
  run<64, true, true>("TS  N=64  B MN-major", d_out);
  run<128, false, false>("SS  N=128 B K-major", d_out);
  run<128, true, false>("TS  N=128 B K-major", d_out);
  run<256, false, false>("SS  N=256 B K-major", d_out);
  run<256, true, false>("TS  N=256 B K-major", d_out);
  return 0;
}
