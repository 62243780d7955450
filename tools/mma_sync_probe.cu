// Throughput probe for the warp-level tensor-core path on sm_100a: mma.sync.m16n8k8 tf32, (a) operands in
// registers (pipe rate), (b) A fragment re-loaded from shared memory for every MMA (the 3-D convolution's inner
// loop when C_out = 8: one A fragment per MMA), (c) A fragment shared by 2 / 4 MMAs (C_out = 16 / 32).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o tools/mma_sync_probe tools/mma_sync_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NB, bool FROM_SMEM>
__global__ void __launch_bounds__(256) probe(float* out, int iters) {
  __shared__ float tile[8][32 * 4 * 2 + 16];     // per warp: two 512-byte A fragments' worth, padded
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 32 * 4 * 2; i += 32) tile[warp][i] = (float)(i & 7) * 0.125f;
  __syncwarp();
  float acc[4][NB][4] = {};
  unsigned a[4] = {0x3f800000u, 0x3f000000u, 0x3e800000u, 0x3e000000u};
  unsigned b[NB][2];
  for (int n = 0; n < NB; ++n) { b[n][0] = 0x3f800000u + n; b[n][1] = 0x3f000000u + n; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {                 // 4 independent accumulator sets: hide the MMA latency
      if (FROM_SMEM) {
        const float* p = &tile[warp][((it + u) & 1) * 128];
        a[0] = __float_as_uint(p[lane]);      a[1] = __float_as_uint(p[lane + 32]);
        a[2] = __float_as_uint(p[lane + 64]); a[3] = __float_as_uint(p[lane + 96]);
      }
#pragma unroll
      for (int n = 0; n < NB; ++n) mma_tf32(acc[u][n], a, b[n]);
    }
  }
  float s = 0.f;
  for (int u = 0; u < 4; ++u) for (int n = 0; n < NB; ++n) for (int j = 0; j < 4; ++j) s += acc[u][n][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NB, bool FROM_SMEM>
static void run(const char* name, float* out, int sms) {
  const int iters = 4096, blocks = sms * 2, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<NB, FROM_SMEM><<<blocks, threads>>>(out, 64);
  cudaEventRecord(e0);
  probe<NB, FROM_SMEM><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = (double)blocks * (threads / 32) * iters * 4 * NB;
  const double tflops = mmas * 2.0 * 16 * 8 * 8 / (ms * 1e-3) / 1e12;
  const double clk_per_mma_sm = (ms * 1e-3) * 1.965e9 / (mmas / sms);
  printf("%-44s %8.3f ms  %7.1f TFLOP/s  %.2f clk per MMA per SM  (%s)\n", name, ms, tflops, clk_per_mma_sm,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 2 * 256);
  run<1, false>("regs only, 1 MMA per A fragment", out, sms);
  run<2, false>("regs only, 2 MMAs per A fragment", out, sms);
  run<1, true>("A from smem (4 LDS.32), 1 MMA per fragment", out, sms);
  run<2, true>("A from smem, 2 MMAs per fragment", out, sms);
  run<4, true>("A from smem, 4 MMAs per fragment", out, sms);
  cudaFree(out);
  return 0;
}
