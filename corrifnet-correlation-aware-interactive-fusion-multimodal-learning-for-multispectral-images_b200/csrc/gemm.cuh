// Internal GEMM plumbing shared by the tcgen05 (TF32) kernel and the CUDA-core fp32 checking kernel.
#pragma once
#include "common.cuh"

namespace corrif {

// Everything the epilogue needs, already offset for the batch the CTA works on.
struct EpiArgs {
  float* D;
  const float* bias;
  const float* residual;
  float* aux;
  int64_t ldd, ldr, ldaux;
  int M, N;
  int mode;
  float alpha;
  bool round_tf32;   // round D to TF32 (nearest) because its only consumers are tensor-core GEMMs
};

// Epilogue on 4 consecutive columns [n, n+4) of row m (n % 4 == 0, all leading dims % 4 == 0).
__device__ __forceinline__ void epilogue_store4(const EpiArgs& e, int m, int n, float4 v) {
  v.x *= e.alpha; v.y *= e.alpha; v.z *= e.alpha; v.w *= e.alpha;
  float* d = e.D + (int64_t)m * e.ldd + n;
  switch (e.mode) {
    case CORRIF_EPI_STORE:
      break;
    case CORRIF_EPI_BIAS: {
      const float4 b = ld4(e.bias + n);
      v = make_float4(v.x + b.x, v.y + b.y, v.z + b.z, v.w + b.w);
    } break;
    case CORRIF_EPI_BIAS_GELU: {
      const float4 b = ld4(e.bias + n);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      st4(e.aux + (int64_t)m * e.ldaux + n, v);
      v = make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
    } break;
    case CORRIF_EPI_BIAS_RESIDUAL: {
      const float4 b = ld4(e.bias + n);
      const float4 r = ld4(e.residual + (int64_t)m * e.ldr + n);
      v = make_float4(v.x + b.x + r.x, v.y + b.y + r.y, v.z + b.z + r.z, v.w + b.w + r.w);
    } break;
    case CORRIF_EPI_MUL_DGELU: {
      const float4 u = ld4(e.aux + (int64_t)m * e.ldaux + n);
      v = make_float4(v.x * dgelu_erf(u.x), v.y * dgelu_erf(u.y), v.z * dgelu_erf(u.z),
                      v.w * dgelu_erf(u.w));
    } break;
    case CORRIF_EPI_ATOMIC_ADD:
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                   :: "l"(d), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      return;
    default:
      break;
  }
  if (e.mode == CORRIF_EPI_ATOMIC_ADD) return;
  if (e.round_tf32) v = round_tf32_4(v);
  st4(d, v);
}

int gemm_tf32_launch(const corrif_gemm_desc& g, cudaStream_t stream);
int gemm_fp32_launch(const corrif_gemm_desc& g, cudaStream_t stream);

}  // namespace corrif
