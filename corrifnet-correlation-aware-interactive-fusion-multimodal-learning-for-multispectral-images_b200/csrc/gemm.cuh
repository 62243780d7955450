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
  // fused dropout (thresh == 0: off)
  uint32_t drop_thresh;
  float drop_scale;            // 1/(1-p)^k
  uint64_t key_a, key_b;       // dropout_key(seed, site); key_b used iff two_sites
  bool two_sites;
};

// host-provided dropout request; the kernel resolves seed_dev and derives the keys once per thread
struct DropArgs {
  float p;
  uint32_t site_a, site_b;
  uint64_t seed;
  const uint64_t* seed_dev;
  uint32_t site_bo;            // site increment per outer batch index
};
inline DropArgs make_drop_args(const corrif_gemm_desc& g) {
  return DropArgs{g.drop_p, g.drop_site_a, g.drop_site_b, g.drop_seed, g.drop_seed_dev, g.drop_site_bo};
}
__device__ __forceinline__ void epi_setup_dropout(EpiArgs& e, const DropArgs& d, int bo = 0) {
  if (d.p <= 0.f) return;
  const uint64_t seed = d.seed + (d.seed_dev ? *d.seed_dev : 0ull);
  const float ks = 1.0f / (1.0f - d.p);
  e.two_sites = d.site_b != CORRIF_NO_SITE;
  e.key_a = dropout_key(seed, d.site_a + (uint32_t)bo * d.site_bo);
  e.key_b = e.two_sites ? dropout_key(seed, d.site_b + (uint32_t)bo * d.site_bo) : 0ull;
  e.drop_scale = e.two_sites ? ks * ks : ks;
  double t = (double)d.p * 65536.0 + 0.5;
  e.drop_thresh = t > 65535.0 ? 65535u : (uint32_t)t;
}

inline EpiArgs make_epi_args(const corrif_gemm_desc& g) {
  EpiArgs e{g.D, g.bias, g.residual, g.aux, g.ldd, g.ldr, g.ldaux, g.M, g.N, g.epilogue, g.alpha,
            (g.flags & CORRIF_GEMM_ROUND_TF32) != 0, 0u, 1.0f, 0ull, 0ull, false};
  return e;
}

// keep(site_a) * keep(site_b) * scale for the 4 elements starting at linear index m*N + n
__device__ __forceinline__ float4 epi_dropout4(const EpiArgs& e, int m, int n, float4 v) {
  const uint64_t quad = ((uint64_t)m * (uint64_t)e.N + (uint64_t)n) >> 2;
  uint32_t km = dropout_keepmask4(e.key_a, quad, e.drop_thresh);
  if (e.two_sites) km &= dropout_keepmask4(e.key_b, quad, e.drop_thresh);
  v.x = (km & 1u) ? v.x * e.drop_scale : 0.f; v.y = (km & 2u) ? v.y * e.drop_scale : 0.f;
  v.z = (km & 4u) ? v.z * e.drop_scale : 0.f; v.w = (km & 8u) ? v.w * e.drop_scale : 0.f;
  return v;
}

// Epilogue on 4 consecutive columns [n, n+4) of row m (n % 4 == 0, all leading dims % 4 == 0), in two
// phases so that callers can issue the global READS of several rows (residual / saved pre-activation)
// back to back before any store: with a fused load->compute->store per row the compiler cannot hoist
// the loads over the stores and every row pays a full DRAM round trip (measured 2.4x on the
// bias+residual GEMMs).
__device__ __forceinline__ float4 epilogue_prefetch4(const EpiArgs& e, int m, int n) {
  if (e.mode == CORRIF_EPI_BIAS_RESIDUAL) return ld4_stream(e.residual + (int64_t)m * e.ldr + n);
  if (e.mode == CORRIF_EPI_MUL_DGELU) return ld4_stream(e.aux + (int64_t)m * e.ldaux + n);
  return make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ void epilogue_apply4(const EpiArgs& e, int m, int n, float4 v, float4 pre) {
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
      if (e.drop_thresh) v = epi_dropout4(e, m, n, v);
    } break;
    case CORRIF_EPI_BIAS_RESIDUAL: {
      const float4 b = ld4(e.bias + n);
      v = make_float4(v.x + b.x, v.y + b.y, v.z + b.z, v.w + b.w);
      if (e.drop_thresh) v = epi_dropout4(e, m, n, v);
      v = make_float4(v.x + pre.x, v.y + pre.y, v.z + pre.z, v.w + pre.w);
    } break;
    case CORRIF_EPI_MUL_DGELU: {
      v = make_float4(v.x * dgelu_erf(pre.x), v.y * dgelu_erf(pre.y), v.z * dgelu_erf(pre.z),
                      v.w * dgelu_erf(pre.w));
      if (e.drop_thresh) v = epi_dropout4(e, m, n, v);
    } break;
    case CORRIF_EPI_ATOMIC_ADD:
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                   :: "l"(d), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      return;
    default:
      break;
  }
  if (e.round_tf32) v = round_tf32_4(v);
  st4(d, v);
}

__device__ __forceinline__ void epilogue_store4(const EpiArgs& e, int m, int n, float4 v) {
  epilogue_apply4(e, m, n, v, epilogue_prefetch4(e, m, n));
}

int gemm_tf32_launch(const corrif_gemm_desc& g, cudaStream_t stream);
int gemm_fp32_launch(const corrif_gemm_desc& g, cudaStream_t stream);

}  // namespace corrif
