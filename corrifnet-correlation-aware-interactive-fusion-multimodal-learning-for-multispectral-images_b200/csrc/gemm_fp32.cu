// Exact-fp32 GEMM on CUDA cores (FFMA).  This is the CHECKING precision of corrif_gemm: same
// descriptor, layouts, batching, split-K and epilogues as the tcgen05 kernel, so a parity failure
// can be attributed to TF32 rounding or to a kernel bug on the device itself.  It is not the hot
// path (about 1/20 of the tensor-core rate).
#include "gemm.cuh"

namespace corrif {

constexpr int FT = 64;   // tile M = tile N
constexpr int FK = 16;

__global__ void __launch_bounds__(256)
gemm_fp32_kernel(const float* __restrict__ A, const float* __restrict__ B, int64_t a_rs, int64_t a_ks,
                 int64_t b_rs, int64_t b_ks, int K, int batch_inner, int split_k, int64_t a_bo,
                 int64_t a_bi, int64_t b_bo, int64_t b_bi, int64_t d_bo, int64_t d_bi, int64_t bias_bo,
                 EpiArgs e, const DropArgs drop) {
  __shared__ float As[FK][FT + 4];
  __shared__ float Bs[FK][FT + 4];
  const int z = blockIdx.z;
  const int split = z % split_k, batch = z / split_k;
  const int bi = batch % batch_inner, bo = batch / batch_inner;
  epi_setup_dropout(e, drop, bo);
  if (e.bias) e.bias += bo * bias_bo;
  A += bo * a_bo + bi * a_bi;
  B += bo * b_bo + bi * b_bi;
  const int64_t doff = bo * d_bo + bi * d_bi;
  e.D += doff;
  if (e.residual) e.residual += doff;
  if (e.aux) e.aux += doff;

  const int kchunk = (((K + split_k - 1) / split_k) + FK - 1) / FK * FK;
  const int k_begin = split * kchunk, k_end = min(K, k_begin + kchunk);
  const int m0 = blockIdx.x * FT, n0 = blockIdx.y * FT;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += FK) {
    for (int i = threadIdx.x; i < FT * FK; i += 256) {
      // pick the fast index so that global reads are contiguous for either storage order
      int r, k;
      if (a_ks == 1) { k = i % FK; r = i / FK; } else { r = i % FT; k = i / FT; }
      const int gm = m0 + r, gk = k0 + k;
      As[k][r] = (gm < e.M && gk < k_end) ? A[(int64_t)gm * a_rs + (int64_t)gk * a_ks] : 0.f;
      if (b_ks == 1) { k = i % FK; r = i / FK; } else { r = i % FT; k = i / FT; }
      const int gn = n0 + r, gk2 = k0 + k;
      Bs[k][r] = (gn < e.N && gk2 < k_end) ? B[(int64_t)gn * b_rs + (int64_t)gk2 * b_ks] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n < e.N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m < e.M) epilogue_store4(e, m, n, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
  }
}

int gemm_fp32_launch(const corrif_gemm_desc& g, cudaStream_t stream) {
  EpiArgs e = make_epi_args(g);
  const DropArgs drop = make_drop_args(g);
  const int64_t a_rs = g.a_mn_major ? 1 : g.lda, a_ks = g.a_mn_major ? g.lda : 1;
  const int64_t b_rs = g.b_mn_major ? 1 : g.ldb, b_ks = g.b_mn_major ? g.ldb : 1;
  dim3 grid((g.M + FT - 1) / FT, (g.N + FT - 1) / FT, g.batch_outer * g.batch_inner * g.split_k);
  gemm_fp32_kernel<<<grid, 256, 0, stream>>>(g.A, g.B, a_rs, a_ks, b_rs, b_ks, g.K, g.batch_inner,
                                             g.split_k, g.a_bo, g.a_bi, g.b_bo, g.b_bi, g.d_bo,
                                             g.d_bi, g.bias_bo, e, drop);
  return launch_status("gemm_fp32");
}

}  // namespace corrif
