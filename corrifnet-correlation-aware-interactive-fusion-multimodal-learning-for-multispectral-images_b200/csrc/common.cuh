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
// For values that go STRAIGHT into a tensor-core operand (registers -> TMEM): adding half a TF32 ulp
// and letting the MMA truncate the low 13 bits IS round-to-nearest (ties away, like cvt.rna) - one
// integer add instead of the four instructions cvt.rna.tf32 expands to on sm_100 (add, Inf test,
// select, mask).  Only for finite values (P in [0,1], dS); the low bits are left dirty.
__device__ __forceinline__ float round_tf32_operand(float x) {
  return __uint_as_float(__float_as_uint(x) + 0x1000u);
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

// ---- counter-based dropout RNG -------------------------------------------------------------------
// keep(seed, site, element) is a pure function, so the backward regenerates (or the attention forward
// saves as bits) exactly the forward's decisions.  The per-(seed, site) key is a SplitMix64 finaliser
// (computed once per thread); the per-element stream is a 5-round Philox2x32-style counter hash
// (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3": L' = mulhi(R, M) ^ L ^ k, R' = mullo(R, M),
// k += golden) of the quad index: 64 bits = four 16-bit draws for the 4 consecutive elements
// [4*quad, 4*quad+4).  One IMAD.WIDE + one 3-input LOP3 + one add per round (~4 instructions per
// element) against ~7 for a SplitMix64 per quad and ~18 for Philox4x32-10 - it matters inside the
// fused attention kernel and the GEMM epilogues, whose element-wise warps are the critical resource.
// tests/test_rng_stats.py checks keep rate, lag / cross-site correlation and row-count variance of
// this exact function.  Element e is kept iff its 16-bit draw >= round(p * 65536).
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t dropout_key(uint64_t seed, uint32_t site) {
  return splitmix64(seed ^ ((uint64_t)(site + 1u) * 0xD6E8FEB86659FD93ull));
}
__host__ __device__ __forceinline__ uint64_t dropout_bits(uint64_t key, uint64_t quad) {
  uint32_t c0 = (uint32_t)quad, c1 = (uint32_t)(quad >> 32) ^ (uint32_t)(key >> 32), k = (uint32_t)key;
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const uint64_t p = (uint64_t)c0 * 0xD256D193u;
    c0 = (uint32_t)(p >> 32) ^ c1 ^ k;
    c1 = (uint32_t)p;
    k += 0x9E3779B9u;
  }
  return ((uint64_t)c1 << 32) | c0;
}
// Same stream with the five round keys (k + r * golden) and the key's upper word precomputed: inside
// the attention forward the compiler re-derived them for every quad (5 extra adds per hash).
struct DropRoundKeys { uint32_t k[5]; uint32_t hi; };
__device__ __forceinline__ DropRoundKeys dropout_round_keys(uint64_t key) {
  DropRoundKeys rk;
#pragma unroll
  for (int r = 0; r < 5; ++r) rk.k[r] = (uint32_t)key + (uint32_t)r * 0x9E3779B9u;
  rk.hi = (uint32_t)(key >> 32);
  return rk;
}
__device__ __forceinline__ uint64_t dropout_bits(const DropRoundKeys& rk, uint64_t quad) {
  uint32_t c0 = (uint32_t)quad, c1 = (uint32_t)(quad >> 32) ^ rk.hi;
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const uint64_t p = (uint64_t)c0 * 0xD256D193u;
    c0 = (uint32_t)(p >> 32) ^ c1 ^ rk.k[r];
    c1 = (uint32_t)p;
  }
  return ((uint64_t)c1 << 32) | c0;
}
__device__ __forceinline__ uint32_t dropout_keepmask4(const DropRoundKeys& rk, uint64_t quad, uint32_t thresh) {
  const uint64_t r = dropout_bits(rk, quad);
  const uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
  return ((lo & 0xFFFFu) >= thresh ? 1u : 0u) | ((lo >> 16) >= thresh ? 2u : 0u) |
         ((hi & 0xFFFFu) >= thresh ? 4u : 0u) | ((hi >> 16) >= thresh ? 8u : 0u);
}
__device__ __forceinline__ void dropout_keep4(uint64_t key, uint64_t quad, uint32_t thresh, float scale,
                                              float (&m)[4]) {
  const uint64_t r = dropout_bits(key, quad);
  const uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
  m[0] = (lo & 0xFFFFu) >= thresh ? scale : 0.f;
  m[1] = (lo >> 16) >= thresh ? scale : 0.f;
  m[2] = (hi & 0xFFFFu) >= thresh ? scale : 0.f;
  m[3] = (hi >> 16) >= thresh ? scale : 0.f;
}
// Same decisions as dropout_keep4 as a 4-bit mask (bit e = element 4*quad+e is kept).
__device__ __forceinline__ uint32_t dropout_keepmask4(uint64_t key, uint64_t quad, uint32_t thresh) {
  const uint64_t r = dropout_bits(key, quad);
  const uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
  return ((lo & 0xFFFFu) >= thresh ? 1u : 0u) | ((lo >> 16) >= thresh ? 2u : 0u) |
         ((hi & 0xFFFFu) >= thresh ? 4u : 0u) | ((hi >> 16) >= thresh ? 8u : 0u);
}
static inline uint32_t dropout_threshold(float p) {
  double t = (double)p * 65536.0 + 0.5;
  if (t < 0) t = 0;
  if (t > 65535.0) t = 65535.0;
  return (uint32_t)t;
}

}  // namespace corrif
