// Inter-modal correlation (InterFormer) forward/backward, mmvit4.py:481-507.  The reference hard-wires M = 3
// modalities (mmvit4.py:15, 394-396); the kernels are templates over M (2..6) so that the same code serves
// BASELINE.json configs[4] (every 3-band group of the 20-band cube as its own modality: M = 6, 3584 tokens).
//
// Reference semantics: scores of query modality X are flattened to
// [3, B*C*S], soft-maxed over the 3 key modalities and re-viewed as [B, 3C, S].  The view keeps the
// element position (s,c) but re-interprets the (key modality m, batch b) pair: with
// j = 3*b' + i = m*B + b, output sample b' multiplies v_i[b'] by A_X[m,b].
//
// HBM-bound, no contraction: 128-bit coalesced accesses along the channel dim, everything else in
// registers.  The forward gathers (thread = output element), the backward is organised per SCORE
// sample (thread = (b,s,c)) because j is a bijection: every dv_i[b'] has exactly one producer, so
// no atomics are needed.
#include "common.cuh"

namespace corrif {


struct f4 { float v[4]; };
__device__ __forceinline__ f4 ldf4(const float* p) {
  const float4 t = ld4(p);
  return f4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ void stf4(float* p, const f4& a) { st4(p, make_float4(a.v[0], a.v[1], a.v[2], a.v[3])); }

// qkv [M][B][S][3C]; skip [M][B][S][C]; tokens [B][(M+1)S][C]
template <int IC_M>
__global__ void __launch_bounds__(128)
inter_corr_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ skip,
                      float* __restrict__ tokens, int B, int S, int C) {
  const float IC_RSQRT_M = rsqrtf((float)IC_M);            // 1/sqrt(M), mmvit4.py:484 (correctly rounded for M = 3, 4)
  const int cq = C / 4;
  const int64_t total = (int64_t)B * S * cq;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // sample index faster than the token index: the q, k of sample b are needed by the three output
  // samples (m B + b) / 3, a third of the batch apart; with all samples of a token row in flight
  // together those re-reads hit L2 (DRAM traffic was 1.7x the algorithmic bytes with b' slowest)
  const int c = (int)(idx % cq) * 4;
  const int bp = (int)((idx / cq) % B);
  const int s = (int)(idx / ((int64_t)cq * B));
  const int64_t C3 = 3 * (int64_t)C;
  const int64_t mod_stride = (int64_t)B * S * C3;

  f4 acc[IC_M];
#pragma unroll
  for (int X = 0; X < IC_M; ++X) acc[X] = ldf4(skip + (((int64_t)X * B + bp) * S + s) * C + c);

#pragma unroll
  for (int i = 0; i < IC_M; ++i) {
    const int j = IC_M * bp + i;
    const int m = j / B, b = j % B;
    const int64_t row_b = ((int64_t)b * S + s) * C3 + c;     // offset inside one modality
    const int64_t row_bp = ((int64_t)bp * S + s) * C3 + c;
    f4 k[IC_M];
#pragma unroll
    for (int mm = 0; mm < IC_M; ++mm) k[mm] = ldf4(qkv + mm * mod_stride + row_b + C);
    const f4 v = ldf4(qkv + i * mod_stride + row_bp + 2 * C);
#pragma unroll
    for (int X = 0; X < IC_M; ++X) {
      const f4 q = ldf4(qkv + X * mod_stride + row_b);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float sc[IC_M], mx = -3.0e38f;
#pragma unroll
        for (int mm = 0; mm < IC_M; ++mm) { sc[mm] = q.v[e] * k[mm].v[e] * IC_RSQRT_M; mx = fmaxf(mx, sc[mm]); }
        float sum = 0.f, em = 0.f;
#pragma unroll
        for (int mm = 0; mm < IC_M; ++mm) {
          const float ex = __expf(sc[mm] - mx);
          sum += ex;
          em = mm == m ? ex : em;
        }
        acc[X].v[e] += em / sum * v.v[e];
      }
    }
  }
#pragma unroll
  for (int X = 0; X < IC_M; ++X)
    stf4(tokens + ((int64_t)bp * (IC_M + 1) * S + (int64_t)X * S + s) * C + c, acc[X]);
}

// g = dL/dtokens [B][(M+1)S][C]; dqkv [M][B][S][3C]
template <int IC_M>
__global__ void __launch_bounds__(128)
inter_corr_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ g,
                      float* __restrict__ dqkv, int B, int S, int C, int g_group_major) {
  const float IC_RSQRT_M = rsqrtf((float)IC_M);
  const int cq = C / 4;
  const int64_t total = (int64_t)B * S * cq;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % cq) * 4;
  const int s = (int)((idx / cq) % S);
  const int b = (int)(idx / ((int64_t)cq * S));
  const int64_t C3 = 3 * (int64_t)C;
  const int64_t mod_stride = (int64_t)B * S * C3;
  const int64_t row_b = ((int64_t)b * S + s) * C3 + c;

  f4 q[IC_M], k[IC_M];
#pragma unroll
  for (int X = 0; X < IC_M; ++X) {
    q[X] = ldf4(qkv + X * mod_stride + row_b);
    k[X] = ldf4(qkv + X * mod_stride + row_b + C);
  }
  // A[X][m] = softmax over m of q_X * k_m / sqrt(3)
  f4 A[IC_M][IC_M];
#pragma unroll
  for (int X = 0; X < IC_M; ++X)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float sc[IC_M], mx = -3.0e38f;
#pragma unroll
      for (int mm = 0; mm < IC_M; ++mm) { sc[mm] = q[X].v[e] * k[mm].v[e] * IC_RSQRT_M; mx = fmaxf(mx, sc[mm]); }
      float sum = 0.f;
#pragma unroll
      for (int mm = 0; mm < IC_M; ++mm) { sc[mm] = __expf(sc[mm] - mx); sum += sc[mm]; }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int mm = 0; mm < IC_M; ++mm) A[X][mm].v[e] = sc[mm] * inv;
    }
  // dA[X][m] = g_X[b'] * v_i[b'] with (b', i) = divmod(m*B + b, 3); dv_i[b'] = sum_X A[X][m]*g_X[b']
  f4 dA[IC_M][IC_M];
#pragma unroll
  for (int m = 0; m < IC_M; ++m) {
    const int j = m * B + b;
    const int bp = j / IC_M, i = j % IC_M;
    const int64_t row_bp = ((int64_t)bp * S + s) * C3 + c;
    const f4 v = ldf4(qkv + i * mod_stride + row_bp + 2 * C);
    f4 dv = {{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int X = 0; X < IC_M; ++X) {
      // g is [B][(M+1)S][C] (the token layout) or, group-major, [M+1][B][S][C] (as LayerNorm-backward can
      // write it, see corrif_layernorm_bwd_regroup)
      const f4 gx = ldf4(g + (g_group_major ? (((int64_t)X * B + bp) * S + s)
                                            : ((int64_t)bp * (IC_M + 1) * S + (int64_t)X * S + s)) * C + c);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dA[X][m].v[e] = gx.v[e] * v.v[e];
        dv.v[e] += A[X][m].v[e] * gx.v[e];
      }
    }
    stf4(dqkv + i * mod_stride + row_bp + 2 * C, dv);
  }
  // ds[X][m] = A*(dA - sum_m' A*dA)/sqrt(3);  dq_X = sum_m ds*k_m;  dk_m = sum_X ds*q_X
  f4 dk[IC_M];
#pragma unroll
  for (int m = 0; m < IC_M; ++m) dk[m] = f4{{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
  for (int X = 0; X < IC_M; ++X) {
    f4 dq = {{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float dot = 0.f;
#pragma unroll
      for (int mm = 0; mm < IC_M; ++mm) dot += A[X][mm].v[e] * dA[X][mm].v[e];
#pragma unroll
      for (int m = 0; m < IC_M; ++m) {
        const float ds = A[X][m].v[e] * (dA[X][m].v[e] - dot) * IC_RSQRT_M;
        dq.v[e] += ds * k[m].v[e];
        dk[m].v[e] += ds * q[X].v[e];
      }
    }
    stf4(dqkv + X * mod_stride + row_b, dq);
  }
#pragma unroll
  for (int m = 0; m < IC_M; ++m) stf4(dqkv + m * mod_stride + row_b + C, dk[m]);
}

}  // namespace corrif

using namespace corrif;

template <int M>
static void launch_fwd(const float* qkv, const float* skip, float* tokens, int B, int S, int C, cudaStream_t st) {
  const int64_t total = (int64_t)B * S * (C / 4);
  inter_corr_fwd_kernel<M><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(qkv, skip, tokens, B, S, C);
}
template <int M>
static void launch_bwd(const float* qkv, const float* g, float* dqkv, int B, int S, int C, int gm, cudaStream_t st) {
  const int64_t total = (int64_t)B * S * (C / 4);
  inter_corr_bwd_kernel<M><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(qkv, g, dqkv, B, S, C, gm);
}

extern "C" {

int corrif_inter_corr_fwd(const float* qkv, const float* skip, float* tokens, int32_t M, int32_t B,
                          int32_t S, int32_t C, void* stream) {
  CORRIF_REQUIRE(M >= 2 && M <= 6, "inter_corr: M must be 2..6 (got %d)", M);
  CORRIF_REQUIRE(qkv && skip && tokens && B > 0 && S > 0 && C > 0 && C % 4 == 0,
                 "inter_corr_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  switch (M) {
    case 2: launch_fwd<2>(qkv, skip, tokens, B, S, C, st); break;
    case 3: launch_fwd<3>(qkv, skip, tokens, B, S, C, st); break;
    case 4: launch_fwd<4>(qkv, skip, tokens, B, S, C, st); break;
    case 5: launch_fwd<5>(qkv, skip, tokens, B, S, C, st); break;
    default: launch_fwd<6>(qkv, skip, tokens, B, S, C, st); break;
  }
  return launch_status("inter_corr_fwd");
}

int corrif_inter_corr_bwd_layout(const float* qkv, const float* g_tokens, float* dqkv, int32_t M,
                                 int32_t B, int32_t S, int32_t C, int32_t g_group_major, void* stream) {
  CORRIF_REQUIRE(M >= 2 && M <= 6, "inter_corr: M must be 2..6 (got %d)", M);
  CORRIF_REQUIRE(qkv && g_tokens && dqkv && B > 0 && S > 0 && C > 0 && C % 4 == 0,
                 "inter_corr_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  switch (M) {
    case 2: launch_bwd<2>(qkv, g_tokens, dqkv, B, S, C, g_group_major, st); break;
    case 3: launch_bwd<3>(qkv, g_tokens, dqkv, B, S, C, g_group_major, st); break;
    case 4: launch_bwd<4>(qkv, g_tokens, dqkv, B, S, C, g_group_major, st); break;
    case 5: launch_bwd<5>(qkv, g_tokens, dqkv, B, S, C, g_group_major, st); break;
    default: launch_bwd<6>(qkv, g_tokens, dqkv, B, S, C, g_group_major, st); break;
  }
  return launch_status("inter_corr_bwd");
}

int corrif_inter_corr_bwd(const float* qkv, const float* g_tokens, float* dqkv, int32_t M,
                          int32_t B, int32_t S, int32_t C, void* stream) {
  return corrif_inter_corr_bwd_layout(qkv, g_tokens, dqkv, M, B, S, C, 0, stream);
}

}  // extern "C"
