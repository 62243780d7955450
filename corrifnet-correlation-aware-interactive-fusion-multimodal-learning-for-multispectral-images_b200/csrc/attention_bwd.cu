// Fused multi-head self-attention backward for sm_100a (autograd of mmvit4.py:307-312).
//
// Given dO, the saved qkv, O and lse (log2-domain log-sum-exp from the forward), produce dqkv with the
// same strided [B*N, 3C] layout.  P is recomputed from Q, K and lse; nothing of size N x N touches HBM.
//   delta[q] = sum_d dO[q,d] * O[q,d]                                   (pre-pass, HBM bound)
//   P   = 2^(S*scale*log2e - lse[q]),  S = Q K^T
//   dP  = (dO V^T) * keep/(1-p)                                         (keep bits saved by the forward)
//   dS  = P * (dP - delta[q]) * scale
//   dQ  = dS K        dK = dS^T Q        dV = (P*keep/(1-p))^T dO
// Two kernels so that every tensor-core operand the threads have to write is K-major with
// "thread = accumulator row":
//   attn_bwd_dq : CTA owns a 128-row query tile (TMEM lane = q), loops over 64-key tiles:
//                 S, dP -> dS (stored back to TMEM) -> dQ += dS . K_j   (K_j also staged MN-major)
//   attn_bwd_dkv: CTA owns a 128-row key tile (TMEM lane = kv), loops over 64-query tiles and computes
//                 the TRANSPOSED products S^T = K_j Q_i^T, dP^T = V_j dO_i^T directly, so P^T and dS^T
//                 come out with kv as the row: dV += P^T . dO_i, dK += dS^T . Q_i  (Q_i, dO_i also MN-major)
// fp32 tensor-core operands cannot share one smem image between K-major and MN-major use (MN-major
// 32-bit operands only exist in the 128B_BASE32B swizzle), hence the duplicate TMA loads.
#include "tc05.cuh"

namespace corrif {
namespace attn {
using namespace tc05;

constexpr int HD = 64;
constexpr int TB = 128;   // owner tile rows (TMEM lanes)
constexpr int TL = 64;    // loop tile rows

// 8 MMAs over the 64-wide head dim: both operands K-major tiles of [rows x 64] stored as two
// [rows x 128 B] k-blocks.
__device__ __forceinline__ void mma_headdim(uint32_t tm, uint32_t sA, int rowsA, uint32_t sB, int rowsB,
                                            uint32_t idesc) {
#pragma unroll
  for (int t = 0; t < HD / 8; ++t) {
    const uint64_t ad = smem_desc_kmajor(sA + (t >> 2) * (rowsA * 128) + (t & 3) * 32);
    const uint64_t bd = smem_desc_kmajor(sB + (t >> 2) * (rowsB * 128) + (t & 3) * 32);
    tcgen05_mma_tf32(tm, ad, bd, idesc, t > 0 ? 1u : 0u);
  }
}
// 8 MMAs over a 64-long contraction: A = thread-written K-major [128 x 64] (two 16 KB k-blocks),
// B = MN-major [n = 64, k = 64] staged as two [64 x 128 B] chunks.
__device__ __forceinline__ void mma_tile64(uint32_t tm, uint32_t sA, uint32_t sBmn, uint32_t idesc,
                                           bool accumulate) {
#pragma unroll
  for (int t = 0; t < TL / 8; ++t) {
    const uint64_t ad = smem_desc_kmajor(sA + (t >> 2) * (TB * 128) + (t & 3) * 32);
    const uint64_t bd = smem_desc_mnmajor(sBmn + t * 1024, TL * 128);
    tcgen05_mma_tf32(tm, ad, bd, idesc, (accumulate || t > 0) ? 1u : 0u);
  }
}
__device__ __forceinline__ void tma_tile(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int col0,
                                         int row0, int rows) {
  tma_load_2d(dst, map, bar, col0, row0);
  tma_load_2d(dst + rows * 128, map, bar, col0 + 32, row0);
}
__device__ __forceinline__ void st_swz(uint32_t tile, int row, int col4, float4 v) {
  // K-major SWIZZLE_128B image of a [128 x 64] fp32 tile: k-block = col/32, 16-B chunk ^= (row & 7)
  const uint32_t addr = tile + (col4 >> 3) * (TB * 128) + row * 128 + ((uint32_t)((col4 & 7) ^ (row & 7)) << 4);
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct BwdArgs {
  const float* lse;          // [B*H, N]  log2 domain
  const float* delta;        // [B*H, N]
  const uint32_t* maskbits;  // [B*H, N, N/32] or nullptr (no dropout)
  float* dqkv;               // [B*N, 3C]
  int N, H;
  float scale, scale_log2e, keep_scale;
};

// ------------------------------------------------------------------------------------------------
// delta pre-pass: one warp per token row of [rows, H*64]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_delta_kernel(const float* __restrict__ O, const float* __restrict__ dO, float* __restrict__ delta,
                  int64_t rows, int N, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int C = H * HD;
  const int64_t b = row / N, q = row % N;
  for (int c0 = 0; c0 < C; c0 += 128) {
    const float4 o = ld4_stream(O + row * C + c0 + lane * 4), d = ld4_stream(dO + row * C + c0 + lane * 4);
    float s = (o.x * d.x + o.y * d.y) + (o.z * d.z + o.w * d.w);
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);   // 16 lanes = 1 head
    if ((lane & 15) == 0) {
      const int h = c0 / HD + (lane >> 4);
      delta[(b * H + h) * N + q] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Both backward kernels: warp 0 = TMA producer, warp 1 = tcgen05 issuer, warps 2..9 = element-wise
// (thread = (accumulator row = TMEM lane, column half g)); no cross-thread reduction is needed in the
// backward because lse and delta are given.
// ------------------------------------------------------------------------------------------------
constexpr int BWD_THREADS = 320;

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// dQ kernel.  S/dP double-buffered in TMEM; dS goes back to the tensor core THROUGH TMEM (tcgen05.st,
// A-operand-in-TMEM MMA), which frees 32 KB of shared memory for a third K/V stage: with two stages a
// stage was re-loaded (48 KB from L2) right before its next use and ~2500 cycles of TMA latency were
// exposed on every tile (ncu: >50 % of stall samples in the two mbarrier waits).
// ------------------------------------------------------------------------------------------------
namespace dq {
constexpr int KV_STAGES = 3;
constexpr int OFF_Q = 0, OFF_DO = OFF_Q + TB * 256;
constexpr int OFF_STAGE = OFF_DO + TB * 256;
constexpr int STAGE_BYTES = 3 * TL * 256;   // K (K-major), K (MN-major), V (K-major)
constexpr int SMEM_BYTES = OFF_STAGE + KV_STAGES * STAGE_BYTES + 1024;
// TMEM columns: buffer u in {0,1}: S [128u, +64) dP [128u+64, +64); dQ [256,320); dS operand u: [320+64u, +64)
constexpr uint32_t TMEM_COLS = 512;
}  // namespace dq

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmKk, const __grid_constant__ CUtensorMap tmKmn,
                   const BwdArgs a) {
  using namespace dq;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t own_full, kv_full[KV_STAGES], kv_free[KV_STAGES], sdp_full[2], sdp_free[2],
      ds_full[2], ds_free[2], fin;
  __shared__ uint32_t tmem_holder;
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = sb + OFF_Q, sDO = sb + OFF_DO;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, bh = blockIdx.y;
  const int b = bh / a.H, h = bh % a.H, C = a.H * HD;
  const int q_row0 = b * a.N + qt * TB, kv_row0 = b * a.N;
  const int ntiles = a.N / TL;

  if (warp == 0 && lane == 0) {
    mbar_init(&own_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_free[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sdp_full[s], 1); mbar_init(&sdp_free[s], 256);
      mbar_init(&ds_full[s], 256); mbar_init(&ds_free[s], 1);
    }
    mbar_init(&fin, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_holder, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint32_t tDQ = tmem + 256, tDS = tmem + 320;

  if (warp == 0 && lane == 0) {
    mbar_expect_tx(&own_full, 2 * TB * 256);
    tma_tile(sQ, &tmQ, &own_full, h * HD, q_row0, TB);
    tma_tile(sDO, &tmDO, &own_full, h * HD, q_row0, TB);
    for (int j = 0; j < ntiles; ++j) {
      const int s = j % KV_STAGES;
      const uint32_t ph = (uint32_t)(j / KV_STAGES) & 1u;
      mbar_wait(&kv_free[s], ph ^ 1u);
      mbar_expect_tx(&kv_full[s], STAGE_BYTES);
      const uint32_t st = sb + OFF_STAGE + s * STAGE_BYTES;
      tma_tile(st, &tmKk, &kv_full[s], C + h * HD, kv_row0 + j * TL, TL);              // K  K-major
      tma_tile(st + TL * 256, &tmKmn, &kv_full[s], C + h * HD, kv_row0 + j * TL, TL);  // K  MN-major
      tma_tile(st + 2 * TL * 256, &tmKk, &kv_full[s], 2 * C + h * HD, kv_row0 + j * TL, TL);  // V K-major
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t id_s = idesc_tf32(TL, false, false);   // [128 x 64] = own . loop^T over d
    constexpr uint32_t id_q = idesc_tf32(HD, false, true);    // [128 x 64] = dS . K (K MN-major)
    mbar_wait(&own_full, 0);
    auto issue_sdp = [&](int j) {
      const int u = j & 1, s = j % KV_STAGES;
      const uint32_t st = sb + OFF_STAGE + s * STAGE_BYTES;
      mbar_wait(&kv_full[s], (uint32_t)(j / KV_STAGES) & 1u);
      mbar_wait(&sdp_free[u], ((uint32_t)(j >> 1) & 1u) ^ 1u);   // element-wise done with buffer u (tile j-2)
      tcgen05_fence_after();
      mma_headdim(tmem + 128 * u, sQ, TB, st, TL, id_s);                       // S  = Q  K_j^T
      mma_headdim(tmem + 128 * u + 64, sDO, TB, st + 2 * TL * 256, TL, id_s);  // dP = dO V_j^T
      tcgen05_commit(&sdp_full[u]);
    };
    issue_sdp(0);
    for (int j = 0; j < ntiles; ++j) {
      if (j + 1 < ntiles) issue_sdp(j + 1);                    // overlaps the element-wise work of tile j
      const int u = j & 1, s = j % KV_STAGES;
      mbar_wait(&ds_full[u], (uint32_t)(j >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t kmn = sb + OFF_STAGE + s * STAGE_BYTES + TL * 256;
#pragma unroll
      for (int t = 0; t < TL / 8; ++t)                          // dQ += dS(TMEM) . K_j
        tcgen05_mma_tf32_ts(tDQ, tDS + 64 * u + 8 * t, smem_desc_mnmajor(kmn + t * 1024, TL * 128), id_q,
                            (j > 0 || t > 0) ? 1u : 0u);
      tcgen05_commit(&ds_free[u]);
      tcgen05_commit(&kv_free[s]);
    }
    tcgen05_commit(&fin);
  } else if (warp >= 2) {
    const int quad = warp & 3, g = (warp - 2) >> 2, row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int q = qt * TB + row;
    const float lse = a.lse[(int64_t)bh * a.N + q], dl = a.delta[(int64_t)bh * a.N + q];
    const uint32_t* mrow = a.maskbits ? a.maskbits + ((int64_t)bh * a.N + q) * (a.N / 32) : nullptr;
    uint32_t rs[32], rp[32];
    for (int j = 0; j < ntiles; ++j) {
      const int u = j & 1;
      const uint32_t ph2 = (uint32_t)(j >> 1) & 1u;
      const uint32_t bits = mrow ? mrow[j * 2 + g] : 0xffffffffu;
      mbar_wait(&sdp_full[u], ph2);
      tcgen05_fence_after();
      tmem_ld32_nowait(tmem + 128 * u + lane_addr + g * 32, rs);
      tmem_ld32_nowait(tmem + 128 * u + 64 + lane_addr + g * 32, rp);
      tmem_wait_ld();
      tcgen05_fence_before();
      mbar_arrive(&sdp_free[u]);                               // S/dP buffer u may be refilled
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float p = ex2_approx(__uint_as_float(rs[c]) * a.scale_log2e - lse);
        const float dp = ((bits >> c) & 1u) ? __uint_as_float(rp[c]) * a.keep_scale : 0.f;
        rs[c] = __float_as_uint(round_tf32(p * (dp - dl) * a.scale));
      }
      mbar_wait(&ds_free[u], ph2 ^ 1u);                        // dQ MMA of tile j-2 has read dS buffer u
      tcgen05_fence_after();
      tmem_st32(tDS + 64 * u + lane_addr + g * 32, rs);
      tcgen05_fence_before();
      mbar_arrive(&ds_full[u]);
    }
    mbar_wait(&fin, 0);
    tcgen05_fence_after();
    float* orow = a.dqkv + (int64_t)(q_row0 + row) * (3 * C) + h * HD + g * 32;
    tmem_ld32(tDQ + lane_addr + g * 32, rs);
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4)
      st4(orow + 4 * q4, make_float4(__uint_as_float(rs[4 * q4]), __uint_as_float(rs[4 * q4 + 1]),
                                     __uint_as_float(rs[4 * q4 + 2]), __uint_as_float(rs[4 * q4 + 3])));
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// dK / dV kernel.  S^T/dP^T double-buffered in TMEM, P^T and dS^T handed to the tensor core through
// TMEM, both query-tile images (K-major for S^T/dP^T, MN-major for dV/dK) double-buffered in smem.
// ------------------------------------------------------------------------------------------------
namespace dkv {
constexpr int OFF_K = 0, OFF_V = OFF_K + TB * 256;
constexpr int OFF_KM = OFF_V + TB * 256;            // 2 stages x {Q_i K-major, dO_i K-major}
constexpr int KM_BYTES = 2 * TL * 256;
constexpr int OFF_MN = OFF_KM + 2 * KM_BYTES;       // 2 stages x {Q_i MN-major, dO_i MN-major}
constexpr int MN_BYTES = 2 * TL * 256;
constexpr int SMEM_BYTES = OFF_MN + 2 * MN_BYTES + 1024;
// TMEM columns: buffer u: S^T [128u,+64) dP^T [128u+64,+64); dV [256,320) dK [320,384);
//               P^T operand [384,448); dS^T operand [448,512)
constexpr uint32_t TMEM_COLS = 512;
}  // namespace dkv

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmKVk,   // qkv,  box {32,128} K-major (K_j, V_j)
                    const __grid_constant__ CUtensorMap tmQk,    // qkv,  box {32, 64} K-major (Q_i)
                    const __grid_constant__ CUtensorMap tmQmn,   // qkv,  box {32, 64} MN-major (Q_i)
                    const __grid_constant__ CUtensorMap tmDOk,   // dO,   box {32, 64} K-major
                    const __grid_constant__ CUtensorMap tmDOmn,  // dO,   box {32, 64} MN-major
                    const BwdArgs a) {
  using namespace dkv;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t own_full, km_full[2], km_free[2], mn_full[2], mn_free[2], st_full[2], st_free[2],
      pds_full, pds_free, fin;
  __shared__ uint32_t tmem_holder;
  __shared__ float s_lse[2][TL], s_delta[2][TL];
  __shared__ uint32_t s_bits[2][TL][4];              // keep bits of (query c, key word w) for this tile
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = sb + OFF_K, sV = sb + OFF_V;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, bh = blockIdx.y;
  const int b = bh / a.H, h = bh % a.H, C = a.H * HD;
  const int kv_row0 = b * a.N + kt * TB, q_base = b * a.N;
  const int ntiles = a.N / TL;

  if (warp == 0 && lane == 0) {
    mbar_init(&own_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&km_full[s], 1); mbar_init(&km_free[s], 1); mbar_init(&mn_full[s], 1); mbar_init(&mn_free[s], 1);
      mbar_init(&st_full[s], 1); mbar_init(&st_free[s], 256);
    }
    mbar_init(&pds_full, 256); mbar_init(&pds_free, 1); mbar_init(&fin, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_holder, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint32_t tDV = tmem + 256, tDK = tmem + 320, tPT = tmem + 384, tDST = tmem + 448;

  if (warp == 0 && lane == 0) {
    mbar_expect_tx(&own_full, 2 * TB * 256);
    tma_tile(sK, &tmKVk, &own_full, C + h * HD, kv_row0, TB);
    tma_tile(sV, &tmKVk, &own_full, 2 * C + h * HD, kv_row0, TB);
    for (int i = 0; i < ntiles; ++i) {
      const int s = i & 1;
      const uint32_t ph = (uint32_t)(i >> 1) & 1u;
      mbar_wait(&km_free[s], ph ^ 1u);
      mbar_expect_tx(&km_full[s], KM_BYTES);
      tma_tile(sb + OFF_KM + s * KM_BYTES, &tmQk, &km_full[s], h * HD, q_base + i * TL, TL);
      tma_tile(sb + OFF_KM + s * KM_BYTES + TL * 256, &tmDOk, &km_full[s], h * HD, q_base + i * TL, TL);
      mbar_wait(&mn_free[s], ph ^ 1u);
      mbar_expect_tx(&mn_full[s], MN_BYTES);
      tma_tile(sb + OFF_MN + s * MN_BYTES, &tmQmn, &mn_full[s], h * HD, q_base + i * TL, TL);
      tma_tile(sb + OFF_MN + s * MN_BYTES + TL * 256, &tmDOmn, &mn_full[s], h * HD, q_base + i * TL, TL);
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t id_s = idesc_tf32(TL, false, false);
    constexpr uint32_t id_g = idesc_tf32(HD, false, true);
    mbar_wait(&own_full, 0);
    auto issue_st = [&](int i) {
      const int u = i & 1;
      const uint32_t ph = (uint32_t)(i >> 1) & 1u;
      const uint32_t km = sb + OFF_KM + u * KM_BYTES;
      mbar_wait(&km_full[u], ph);
      mbar_wait(&st_free[u], ph ^ 1u);
      tcgen05_fence_after();
      mma_headdim(tmem + 128 * u, sK, TB, km, TL, id_s);                   // S^T  = K_j Q_i^T
      mma_headdim(tmem + 128 * u + 64, sV, TB, km + TL * 256, TL, id_s);   // dP^T = V_j dO_i^T
      tcgen05_commit(&km_free[u]);
      tcgen05_commit(&st_full[u]);
    };
    issue_st(0);
    for (int i = 0; i < ntiles; ++i) {
      if (i + 1 < ntiles) issue_st(i + 1);
      const int s = i & 1;
      mbar_wait(&pds_full, (uint32_t)i & 1u);
      mbar_wait(&mn_full[s], (uint32_t)(i >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t mn = sb + OFF_MN + s * MN_BYTES;
#pragma unroll
      for (int t = 0; t < TL / 8; ++t)                                      // dV += P^T(TMEM)  dO_i
        tcgen05_mma_tf32_ts(tDV, tPT + 8 * t, smem_desc_mnmajor(mn + TL * 256 + t * 1024, TL * 128), id_g,
                            (i > 0 || t > 0) ? 1u : 0u);
#pragma unroll
      for (int t = 0; t < TL / 8; ++t)                                      // dK += dS^T(TMEM) Q_i
        tcgen05_mma_tf32_ts(tDK, tDST + 8 * t, smem_desc_mnmajor(mn + t * 1024, TL * 128), id_g,
                            (i > 0 || t > 0) ? 1u : 0u);
      tcgen05_commit(&mn_free[s]);
      tcgen05_commit(&pds_free);
    }
    tcgen05_commit(&fin);
  } else if (warp >= 2) {
    const int quad = warp & 3, g = (warp - 2) >> 2, row = quad * 32 + lane;   // kv row == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int t256 = threadIdx.x - 64;                           // 0..255 among the 8 warps
    const int words = a.N / 32;
    uint32_t rs[32], rp[32];
    for (int i = 0; i < ntiles; ++i) {
      const int u = i & 1;
      const uint32_t ph2 = (uint32_t)(i >> 1) & 1u;
      // per-column statistics and keep bits of this query tile -> smem (double-buffered by tile parity)
      if (t256 < TL) s_lse[u][t256] = a.lse[(int64_t)bh * a.N + i * TL + t256];
      else if (t256 < 2 * TL) s_delta[u][t256 - TL] = a.delta[(int64_t)bh * a.N + i * TL + (t256 - TL)];
      if (a.maskbits)
        s_bits[u][t256 >> 2][t256 & 3] =
            a.maskbits[((int64_t)bh * a.N + i * TL + (t256 >> 2)) * words + kt * 4 + (t256 & 3)];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&st_full[u], ph2);
      tcgen05_fence_after();
      tmem_ld32_nowait(tmem + 128 * u + lane_addr + g * 32, rs);
      tmem_ld32_nowait(tmem + 128 * u + 64 + lane_addr + g * 32, rp);
      tmem_wait_ld();
      tcgen05_fence_before();
      mbar_arrive(&st_free[u]);
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int qc = g * 32 + c;                               // query column inside the tile
        const float p = ex2_approx(__uint_as_float(rs[c]) * a.scale_log2e - s_lse[u][qc]);
        float keep = 1.0f;
        if (a.maskbits) keep = ((s_bits[u][qc][quad] >> lane) & 1u) ? a.keep_scale : 0.f;
        const float pd = p * keep;
        rs[c] = __float_as_uint(round_tf32(pd));
        rp[c] = __float_as_uint(round_tf32((pd * __uint_as_float(rp[c]) - p * s_delta[u][qc]) * a.scale));
      }
      mbar_wait(&pds_free, ((uint32_t)i & 1u) ^ 1u);            // dV/dK MMAs of tile i-1 have read the operands
      tcgen05_fence_after();
      tmem_st32(tPT + lane_addr + g * 32, rs);
      tmem_st32(tDST + lane_addr + g * 32, rp);
      tcgen05_fence_before();
      mbar_arrive(&pds_full);
    }
    mbar_wait(&fin, 0);
    tcgen05_fence_after();
    float* krow = a.dqkv + (int64_t)(kv_row0 + row) * (3 * C) + C + h * HD + g * 32;
    float* vrow = krow + C;
    tmem_ld32_nowait(tDV + lane_addr + g * 32, rs);
    tmem_ld32_nowait(tDK + lane_addr + g * 32, rp);
    tmem_wait_ld();
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      st4(vrow + 4 * q4, make_float4(__uint_as_float(rs[4 * q4]), __uint_as_float(rs[4 * q4 + 1]),
                                     __uint_as_float(rs[4 * q4 + 2]), __uint_as_float(rs[4 * q4 + 3])));
      st4(krow + 4 * q4, make_float4(__uint_as_float(rp[4 * q4]), __uint_as_float(rp[4 * q4 + 1]),
                                     __uint_as_float(rp[4 * q4 + 2]), __uint_as_float(rp[4 * q4 + 3])));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace attn
}  // namespace corrif

using namespace corrif;

extern "C" int corrif_attention_bwd(const float* qkv, const float* O, const float* dO, const float* lse,
                                    const uint32_t* maskbits, float* delta, float* dqkv, int32_t B,
                                    int32_t N, int32_t H, int32_t D, float scale, float p_drop,
                                    void* stream) {
  using namespace corrif::attn;
  CORRIF_REQUIRE(qkv && O && dO && lse && delta && dqkv && B > 0, "attention_bwd: null/empty");
  CORRIF_REQUIRE(D == HD, "attention_bwd: head_dim must be 64 (got %d)", D);
  CORRIF_REQUIRE(H > 0 && N > 0 && N % TB == 0, "attention_bwd: N must be a multiple of 128 (got %d)", N);
  CORRIF_REQUIRE((int64_t)B * H <= 65535, "attention_bwd: B*H too large");
  CORRIF_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "attention_bwd: p_drop");
  CORRIF_REQUIRE(p_drop == 0.f || maskbits != nullptr, "attention_bwd: dropout needs the forward's maskbits");
  const int C = H * D;
  const uint64_t rows = (uint64_t)B * N;
  cudaStream_t st = (cudaStream_t)stream;
  attn_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(O, dO, delta, (int64_t)rows, N, H);
  int rc = launch_status("attention_delta");
  if (rc) return rc;

  CUtensorMap q128, do128, k64, k64mn, kv128, q64, q64mn, do64, do64mn;
  if ((rc = tc05::encode_map(&q128, qkv, 3 * C, rows, 3 * C, 32, TB, false))) return rc;
  if ((rc = tc05::encode_map(&do128, dO, C, rows, C, 32, TB, false))) return rc;
  if ((rc = tc05::encode_map(&k64, qkv, 3 * C, rows, 3 * C, 32, TL, false))) return rc;
  if ((rc = tc05::encode_map(&k64mn, qkv, 3 * C, rows, 3 * C, 32, TL, true))) return rc;
  kv128 = q128; q64 = k64; q64mn = k64mn;
  if ((rc = tc05::encode_map(&do64, dO, C, rows, C, 32, TL, false))) return rc;
  if ((rc = tc05::encode_map(&do64mn, dO, C, rows, C, 32, TL, true))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dq::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv::SMEM_BYTES);
    if (e != cudaSuccess) { set_last_error("attention_bwd: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  BwdArgs a;
  a.lse = lse; a.delta = delta; a.maskbits = p_drop > 0.f ? maskbits : nullptr; a.dqkv = dqkv;
  a.N = N; a.H = H; a.scale = scale; a.scale_log2e = scale * 1.4426950408889634f;
  a.keep_scale = 1.0f / (1.0f - p_drop);
  dim3 grid(N / TB, B * H);
  attn_bwd_dq_kernel<<<grid, BWD_THREADS, dq::SMEM_BYTES, st>>>(q128, do128, k64, k64mn, a);
  if ((rc = launch_status("attention_bwd_dq"))) return rc;
  attn_bwd_dkv_kernel<<<grid, BWD_THREADS, dkv::SMEM_BYTES, st>>>(kv128, q64, q64mn, do64, do64mn, a);
  return launch_status("attention_bwd_dkv");
}
