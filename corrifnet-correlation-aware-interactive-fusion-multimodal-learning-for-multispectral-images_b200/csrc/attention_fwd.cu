// Fused multi-head self-attention forward for sm_100a (SelfAttention.forward, mmvit4.py:307-312):
//   O[b, q, h*64 + :] = dropout(softmax(Q K^T * scale)) V      per (batch b, head h), head_dim 64
// Q, K, V are strided views of the qkv GEMM output [B*N, 3*C] (no reshape/permute copies); the
// N x N score matrix never touches HBM (the reference materialises [B,8,N,N]: 134 MB per sample at
// N = 2048).
//
// One CTA = one 128-row query tile of one (b, h); it walks the keys in tiles of 64:
//   warp 8      TMA producer: Q once, then K_j / V_j tiles (SWIZZLE_128B K-major for Q and K;
//               V is the B operand of P.V with the contraction (key) index slow in memory, i.e.
//               MN-major -> 128B_ATOM_32B boxes)
//   warp 9      tcgen05 issuer: S = Q K_j^T (TF32, fp32 accumulate, 64 TMEM columns), and, once the
//               softmax warps have stored P_j back into TMEM, O_j = P_j V_j with the A operand read
//               from TMEM (no shared-memory round trip for P; the freed 32 KB double-buffer K/V)
//   warps 0..7  online softmax, two warpgroups: thread = (query row = TMEM lane, column half g).
//               Group g owns score columns [32g, 32g+32) of the tile and output columns [32g, 32g+32):
//               one tcgen05.ld of S, row max exchanged between the halves through shared memory, reference
//               exponent / row sum in the log2 domain (FFMA2 + ex2.approx + FADD2), counter-RNG dropout
//               (keep bits saved for the backward), P_j stored to TMEM (tcgen05.st).  O accumulates in
//               TMEM over all key tiles; it is rescaled there only when a row maximum outgrows the
//               reference exponent by 2^8 (lazy rescale), and read once at the end.
// Shared memory is ~97 KB and TMEM 128 columns per CTA, so two CTAs share an SM and one's softmax
// overlaps the other's MMAs.  lse (log2-domain log-sum-exp) is saved for the backward.
#include <stdlib.h>
#include "tc05.cuh"

namespace corrif {
namespace attn {
using namespace tc05;

constexpr int TQ = 128, TK = 64, HD = 64;
constexpr int Q_BYTES = TQ * HD * 4, K_BYTES = TK * HD * 4, V_BYTES = TK * HD * 4;
constexpr int KV_BYTES = K_BYTES + V_BYTES;
constexpr int KV_STAGES = 3;                          // {K_j, V_j} ring
constexpr int OFF_KV = 0;
constexpr int SMEM_BYTES = OFF_KV + KV_STAGES * KV_BYTES + 1024;
// TMEM columns.  S: [0,64)  O_j: [64,128)  P_j (A operand of P.V): [128,192)  Q (A operand of Q.K^T): [192,256)
// Q lives in TMEM for the CTA's lifetime: with both operands in shared memory a 128 x 64 x 8 TF32 MMA
// has to fetch 6 KB (48 clk at 128 B/clk) for 32 clk of math; with A in TMEM only K_j's 2 KB remain.
constexpr uint32_t TMEM_COLS = 256;
constexpr float RESCALE_LOG2 = 8.0f;          // lazy rescale threshold (log2 domain)
// TS-mode kind::tf32 reads the low 13 mantissa bits of P as zero.  For p = 2^x the mantissa f is
// log-uniform on [1,2), so the mean relative truncation loss is 2^-11 * E[1/f] = 2^-11 * 0.7213 = 3.522e-4.
constexpr float TRUNC_COMP_SCALE = 1.0f + 3.522e-4f;
constexpr float TRUNC_COMP = 5.0806e-4f;      // log2(TRUNC_COMP_SCALE)

struct FwdArgs {
  const float* qkv;
  float* O;
  float* lse;
  uint32_t* maskbits;   // [B*H, N, N/32] keep bits (written when dropout is on), or nullptr
  int N, H;
  int64_t ldo;
  float scale_log2e;
  uint32_t thresh;
  float keep_scale;
  uint64_t seed;
  const uint64_t* seed_dev;
  uint32_t site;
  int group_batches;          // > 0: batch b is module b / group_batches (own dropout site)
  uint32_t group_site_stride;
  int round_out;
};

constexpr int NUM_THREADS = 320;     // 8 softmax warps, producer warp, MMA warp
// The scheduler arbitrates highest-warp-id-first, so the single-thread issuers sit above the softmax warps.
constexpr int PROD_WARP = 8, MMA_WARP = 9;

// DROP is a template parameter so that the dropout code is straight-line (no branch per float4 group).
// PRE (with DROP): the keep bits were produced ahead of time by attn_keepbits_kernel (same decisions) and are
// only READ here, one word per thread and tile, fetched a tile ahead: the counter hash, the 16-bit compares and
// the keep-bit word are 4.3 of the 12.5 instructions per score element of this issue-bound loop, and they do
// not depend on the data - the engine runs them on a second stream under the GEMMs that precede the attention.
template <bool DROP, bool PRE>
__global__ void __launch_bounds__(NUM_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, kv_full[KV_STAGES], kv_free[KV_STAGES], s_full, s_free, p_full, o_full;
  __shared__ uint32_t tmem_holder;
  __shared__ float s_max[2][2][TQ];    // [tile parity][column half][row]
  __shared__ float s_sum[2][TQ];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, bh = blockIdx.y;
  const int b = bh / a.H, h = bh % a.H;
  const int C = a.H * HD;
  const int q_row0 = b * a.N + qt * TQ;     // row in the [B*N, 3C] qkv matrix
  const int kv_row0 = b * a.N;
  const int ntiles = a.N / TK;

  // The softmax threads fetch their half of their query row BEFORE the setup barrier: the global-load latency
  // (ncu: ~9 % of the kernel's stall samples sat on it at N = 2048, more at N = 512) then overlaps barrier
  // initialisation and the TMEM allocation instead of following them.
  uint32_t r[32];
  uint64_t seed_off = 0;
  if (warp < 8) {
    const float* qrow = a.qkv + (int64_t)(q_row0 + (warp & 3) * 32 + lane) * (3 * C) + h * HD + (warp >> 2) * 32;
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      const float4 v = ld4(qrow + 4 * q4);
      r[4 * q4] = __float_as_uint(v.x); r[4 * q4 + 1] = __float_as_uint(v.y);
      r[4 * q4 + 2] = __float_as_uint(v.z); r[4 * q4 + 3] = __float_as_uint(v.w);
    }
    if (DROP && a.seed_dev != nullptr) seed_off = *a.seed_dev;
  }
  if (warp == PROD_WARP && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&tmV) : "memory");
    mbar_init(&q_full, 256); mbar_init(&s_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_free[s], 1); }
    mbar_init(&s_free, 256); mbar_init(&p_full, 256); mbar_init(&o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) tmem_alloc(&tmem_holder, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint32_t tS = tmem, tO = tmem + 64, tP = tmem + 128, tQ = tmem + 192;

  if (warp == PROD_WARP && lane == 0) {
    // ===================== TMA producer =====================
    for (int j = 0; j < ntiles; ++j) {
      const int s = j % KV_STAGES;
      mbar_wait(&kv_free[s], ((uint32_t)(j / KV_STAGES) & 1u) ^ 1u);
      mbar_expect_tx(&kv_full[s], KV_BYTES);
      const uint32_t sK = sbase + OFF_KV + s * KV_BYTES, sV = sK + K_BYTES;
      tma_load_2d(sK, &tmK, &kv_full[s], C + h * HD, kv_row0 + j * TK);
      tma_load_2d(sK + TK * 128, &tmK, &kv_full[s], C + h * HD + 32, kv_row0 + j * TK);
      tma_load_2d(sV, &tmV, &kv_full[s], 2 * C + h * HD, kv_row0 + j * TK);
      tma_load_2d(sV + TK * 128, &tmV, &kv_full[s], 2 * C + h * HD + 32, kv_row0 + j * TK);
    }
  } else if (warp == MMA_WARP) {
    // ===================== tcgen05 issuer (uniform warp, one elected lane: see elect_one()) =========
    constexpr uint32_t idesc_s = idesc_tf32(TK, false, false);   // S[128 x 64] = Q . K^T
    constexpr uint32_t idesc_o = idesc_tf32(HD, false, true);    // O[128 x 64] = P . V (V MN-major)
    mbar_wait(&q_full, 0);                           // the softmax warps have put Q into TMEM
    // S_{j+1} is issued BEFORE waiting for P_j: the softmax warps hand S back (s_free) as soon as they
    // have it in registers, so the next scores are computed under their exp / dropout work instead of
    // behind P_j . V_j (measured: ~500 of 3500 cycles per tile were spent waiting for S).
    auto issue_s = [&](int j) {
      const int s = j % KV_STAGES;
      const uint32_t sK = sbase + OFF_KV + s * KV_BYTES;
      mbar_wait(&kv_full[s], (uint32_t)(j / KV_STAGES) & 1u);
      mbar_wait(&s_free, ((uint32_t)j & 1u) ^ 1u);   // softmax finished reading S_{j-1}
      tcgen05_fence_after();
      if (elect_one()) {                           // S = Q(TMEM) . K_j^T
        mma8_ts_kmajor<TK>(tS, tQ, desc_lo_kmajor(sK), idesc_s, false);
        tcgen05_commit(&s_full);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < ntiles; ++j) {
      const uint32_t ph = (uint32_t)j & 1u;
      const int s = j % KV_STAGES;
      const uint32_t sV = sbase + OFF_KV + s * KV_BYTES + K_BYTES;
      if (j + 1 < ntiles) issue_s(j + 1);
      mbar_wait(&p_full, ph);                      // P_j in TMEM, O_{j-1} already consumed
      tcgen05_fence_after();
      if (elect_one()) {                           // O_j = P_j(TMEM) . V_j
        mma8_ts_mnmajor(tO, tP, desc_lo_mnmajor(sV, TK * 128), idesc_o, j > 0);   // O accumulates in TMEM
        tcgen05_commit(&kv_free[s]);
        tcgen05_commit(&o_full);
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ===================== online softmax (8 warps) =====================
    const int quad = warp & 3;                               // TMEM lane quadrant this warp may access
    const int g = warp >> 2;                                 // column half
    const int row = quad * 32 + lane;                        // query row in the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int q_in_head = qt * TQ + row;
    uint64_t key = 0;
    const int grp = a.group_batches > 0 ? b / a.group_batches : 0;
    const int bh_rng = a.group_batches > 0 ? (b - grp * a.group_batches) * a.H + h : bh;
    if (DROP)
      key = dropout_key(a.seed + seed_off, a.site + (uint32_t)grp * a.group_site_stride);
    const DropRoundKeys rk = dropout_round_keys(key);
    const uint64_t drop_row = ((uint64_t)bh_rng * a.N + q_in_head) * (uint64_t)a.N;
    float m = -INFINITY, l = 0.f;      // m: the reference exponent in use (log2 domain), l: row sum of 2^(s - m)
    {   // my half of my query row (loaded above) -> TMEM (qkv is already TF32-rounded by its producer)
      tmem_st32(tQ + lane_addr + g * 32, r);
      tcgen05_fence_before();
      mbar_arrive(&q_full);
    }
    const uint64_t sc2 = pack2(a.scale_log2e, a.scale_log2e);
    const uint32_t th_lo = a.thresh, th_hi = a.thresh << 16;
    const uint32_t* mrow = a.maskbits + ((int64_t)bh * a.N + q_in_head) * (a.N / 32) + g;
    uint32_t bits_next = (DROP && PRE) ? mrow[0] : 0u;
    for (int j = 0; j < ntiles; ++j) {
      const uint32_t ph = (uint32_t)j & 1u;
      const uint32_t bits_cur = bits_next;
      if (DROP && PRE && j + 1 < ntiles) bits_next = mrow[(j + 1) * (TK / 32)];   // no arithmetic on it before the next tile
      mbar_wait(&s_full, ph);
      tcgen05_fence_after();
      tmem_ld32(tS + lane_addr + g * 32, r);                 // my 32 score columns (kept in registers)
      tcgen05_fence_before();
      mbar_arrive(&s_free);                                  // S is in registers: Q K_{j+1}^T may overwrite it
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(r[c]));
      s_max[ph][g][row] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, s_max[ph][g ^ 1][row]) * a.scale_log2e;
      // Lazy rescale: O_j accumulates in TMEM across tiles; the reference exponent m only moves when the
      // row maximum outgrows it by 2^RESCALE_LOG2 (P then stays <= 256, l <= N * 256), which after the
      // first tile is rare - the old per-tile "(acc + O_j) * alpha" fold cost 2 of the ~22 instructions per
      // score element plus a TMEM load of O every tile.  Both column halves of a row see the same m and mx,
      // so they take the same per-row decision; a warp enters when any of its rows does (alpha = 1 elsewhere).
      const bool grow = mx > m + RESCALE_LOG2;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mx : m;
        const float alpha = ex2_approx(m - m_new);           // first tile: 2^(-inf) = 0
        l *= alpha;
        m = m_new;
        if (j > 0) {
          mbar_wait(&o_full, ph ^ 1u);                       // P_{j-1} . V_{j-1} has landed in O
          tcgen05_fence_after();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t o16[16];
            tmem_ld16(tO + lane_addr + g * 32 + hh * 16, o16);
#pragma unroll
            for (int c = 0; c < 16; ++c) o16[c] = __float_as_uint(__uint_as_float(o16[c]) * alpha);
            tmem_st16(tO + lane_addr + g * 32 + hh * 16, o16);
          }
          tcgen05_fence_before();
        }
      }
      // p' = 2^(s*scale*log2e - m + TRUNC_COMP), partial row sum, dropout, publish my k-block of P_j.
      // P is only ever a TS-mode tensor-core operand, which TRUNCATES to TF32: instead of rounding every
      // element (one integer add each) all of P is scaled up by 2^TRUNC_COMP = 1 + E[truncation loss]
      // (free: folded into the exponent offset) and taken out again in the final 1 / l.
      const float nm = TRUNC_COMP - m;
      const uint64_t nm2 = pack2(nm, nm);
      uint64_t rs_a = pack2(0.f, 0.f), rs_b = rs_a;
      uint32_t keepbits = 0u;
      const uint64_t q0 = (drop_row + (uint64_t)(j * TK + g * 32)) >> 2;
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        float p0, p1, p2, p3;
        unpack2(fma2(pack2u(r[4 * q4 + 0], r[4 * q4 + 1]), sc2, nm2), p0, p1);
        unpack2(fma2(pack2u(r[4 * q4 + 2], r[4 * q4 + 3]), sc2, nm2), p2, p3);
        p0 = ex2_approx(p0); p1 = ex2_approx(p1); p2 = ex2_approx(p2); p3 = ex2_approx(p3);
        rs_a = add2(rs_a, pack2(p0, p1));
        rs_b = add2(rs_b, pack2(p2, p3));
        if (DROP && PRE) {
          p0 = (bits_cur & (1u << (4 * q4))) ? p0 : 0.f; p1 = (bits_cur & (2u << (4 * q4))) ? p1 : 0.f;
          p2 = (bits_cur & (4u << (4 * q4))) ? p2 : 0.f; p3 = (bits_cur & (8u << (4 * q4))) ? p3 : 0.f;
        }
        if (DROP && !PRE) {       // 1/(1-p) is applied once to the output, not per element
          const uint64_t hbits = dropout_bits(rk, q0 + q4);          // same decisions as dropout_keepmask4
          const uint32_t lo = (uint32_t)hbits, hi = (uint32_t)(hbits >> 32);
          const bool k0 = (lo & 0xFFFFu) >= th_lo, k1 = lo >= th_hi, k2 = (hi & 0xFFFFu) >= th_lo, k3 = hi >= th_hi;
          p0 = k0 ? p0 : 0.f; p1 = k1 ? p1 : 0.f; p2 = k2 ? p2 : 0.f; p3 = k3 ? p3 : 0.f;
          if (k0) keepbits |= 1u << (4 * q4);
          if (k1) keepbits |= 2u << (4 * q4);
          if (k2) keepbits |= 4u << (4 * q4);
          if (k3) keepbits |= 8u << (4 * q4);
        }
        r[4 * q4 + 0] = __float_as_uint(p0); r[4 * q4 + 1] = __float_as_uint(p1);
        r[4 * q4 + 2] = __float_as_uint(p2); r[4 * q4 + 3] = __float_as_uint(p3);
      }
      if (DROP && !PRE && a.maskbits != nullptr)
        a.maskbits[((int64_t)bh * a.N + q_in_head) * (a.N / 32) + j * (TK / 32) + g] = keepbits;
      {
        float s0, s1;
        unpack2(add2(rs_a, rs_b), s0, s1);
        l += s0 + s1;
      }
      if (j > 0) {                                  // P_{j-1} . V_{j-1} must have finished reading P_{j-1}; it was
        mbar_wait(&o_full, ph ^ 1u);                // issued a whole exp / dropout pass ago, so this does not block
        tcgen05_fence_after();
      }
      tmem_st32(tP + lane_addr + g * 32, r);        // P_j -> TMEM
      tcgen05_fence_before();
      mbar_arrive(&p_full);
    }
    // O = sum_j P_j V_j sits in TMEM; combine the two halves' partial row sums
    mbar_wait(&o_full, (uint32_t)(ntiles - 1) & 1u);
    tcgen05_fence_after();
    tmem_ld32(tO + lane_addr + g * 32, r);
    s_sum[g][row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += s_sum[g ^ 1][row];
    const float inv = (DROP ? a.keep_scale : 1.0f) * TRUNC_COMP_SCALE / l;
    float* orow = a.O + (int64_t)(q_row0 + row) * a.ldo + h * HD + g * 32;
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      float4 v = make_float4(__uint_as_float(r[4 * q4]) * inv, __uint_as_float(r[4 * q4 + 1]) * inv,
                             __uint_as_float(r[4 * q4 + 2]) * inv, __uint_as_float(r[4 * q4 + 3]) * inv);
      if (a.round_out) v = round_tf32_4(v);
      st4(orow + 4 * q4, v);
    }
    if (g == 0) a.lse[(int64_t)bh * a.N + q_in_head] = m + log2f(l) - TRUNC_COMP;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, TMEM_COLS);
}

// The forward's dropout decisions as a stand-alone pass: word w = ((b*H + h)*N + q)*(N/32) + kw holds the keep
// bits of keys [32 kw, 32 kw + 32) of query q - exactly what attn_fwd_kernel<true, false> would store.
__global__ void __launch_bounds__(256)
attn_keepbits_kernel(uint32_t* __restrict__ maskbits, int N, int H, int64_t total_words, uint32_t thresh, uint64_t seed,
                     const uint64_t* __restrict__ seed_dev, uint32_t site, int group_batches, uint32_t group_site_stride) {
  if (seed_dev != nullptr) seed += *seed_dev;
  const int words = N / 32;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // two words (16 independent hash chains) per thread and sweep: when the pass runs beside a GEMM it gets one
  // block per SM, and 8 warps only fill the ALU pipe if each of them carries enough independent work
  for (int64_t w0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w0 < total_words; w0 += 2 * stride) {
    uint32_t bits[2] = {0u, 0u};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t w = w0 + u * stride;
      if (w < total_words) {
        const int kw = (int)(w % words);
        const int64_t rowidx = w / words;
        const int q = (int)(rowidx % N), bh = (int)(rowidx / N);
        const int b = bh / H, h = bh % H;
        const int grp = group_batches > 0 ? b / group_batches : 0;
        const int bh_rng = group_batches > 0 ? (b - grp * group_batches) * H + h : bh;
        const DropRoundKeys rk = dropout_round_keys(dropout_key(seed, site + (uint32_t)grp * group_site_stride));
        const uint64_t q0 = ((((uint64_t)bh_rng * N + q) * (uint64_t)N) + (uint64_t)kw * 32u) >> 2;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) bits[u] |= dropout_keepmask4(rk, q0 + q4, thresh) << (4 * q4);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (w0 + u * stride < total_words) maskbits[w0 + u * stride] = bits[u];
  }
}

}  // namespace attn
}  // namespace corrif

using namespace corrif;

extern "C" int corrif_attention_keepbits(uint32_t* maskbits, int32_t B, int32_t N, int32_t H, float p_drop, uint64_t seed,
                                         const uint64_t* seed_dev, uint32_t site, int32_t group_batches,
                                         uint32_t group_site_stride, int32_t max_blocks, void* stream) {
  using namespace corrif::attn;
  CORRIF_REQUIRE(maskbits && B > 0 && H > 0 && N > 0 && N % 32 == 0, "attention_keepbits: null/empty or N % 32 != 0");
  CORRIF_REQUIRE(p_drop > 0.f && p_drop < 1.f, "attention_keepbits: 0 < p_drop < 1");
  CORRIF_REQUIRE(group_batches >= 0 && (group_batches == 0 || B % group_batches == 0),
                 "attention_keepbits: group_batches must divide B");
  const int64_t total = (int64_t)B * H * N * (N / 32);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = max_blocks > 0 ? max_blocks : (int64_t)num_sms() * 8;   // e.g. one block per SM beside a GEMM
  if (blocks > cap) blocks = cap;
  attn_keepbits_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(maskbits, N, H, total, dropout_threshold(p_drop), seed,
                                                                         seed_dev, site, group_batches, group_site_stride);
  return launch_status("attention_keepbits");
}

static int attention_fwd_launch(const float* qkv, float* O, float* lse, uint32_t* maskbits,
                                int32_t B, int32_t N, int32_t H, int32_t D, float scale,
                                float p_drop, uint64_t seed,
                                const uint64_t* seed_dev, uint32_t site, int32_t group_batches,
                                uint32_t group_site_stride, int32_t round_tf32, bool premasked, void* stream) {
  using namespace corrif::attn;
  CORRIF_REQUIRE(qkv && O && lse && B > 0, "attention_fwd: null/empty");
  CORRIF_REQUIRE(D == HD, "attention_fwd: head_dim must be 64 (got %d)", D);
  CORRIF_REQUIRE(H > 0 && N > 0 && N % TQ == 0, "attention_fwd: N must be a multiple of 128 (got %d)", N);
  CORRIF_REQUIRE((int64_t)B * H <= 65535, "attention_fwd: B*H too large");
  CORRIF_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "attention_fwd: p_drop");
  CORRIF_REQUIRE(p_drop == 0.f || maskbits != nullptr, "attention_fwd: dropout needs a maskbits buffer");
  CORRIF_REQUIRE(group_batches >= 0 && (group_batches == 0 || B % group_batches == 0),
                 "attention_fwd: group_batches must divide B");
  CORRIF_REQUIRE(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)O % 16 == 0), "attention_fwd: alignment");
  const int C = H * D;
  CUtensorMap tq, tk, tv;
  int st = tc05::encode_map(&tq, qkv, 3 * C, (uint64_t)B * N, 3 * C, 32, TQ, false);
  if (st) return st;
  st = tc05::encode_map(&tk, qkv, 3 * C, (uint64_t)B * N, 3 * C, 32, TK, false);
  if (st) return st;
  st = tc05::encode_map(&tv, qkv, 3 * C, (uint64_t)B * N, 3 * C, 32, TK, true);
  if (st) return st;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { set_last_error("attention_fwd: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  FwdArgs a;
  a.qkv = qkv; a.O = O; a.lse = lse; a.maskbits = maskbits; a.N = N; a.H = H; a.ldo = C;
  a.scale_log2e = scale * 1.4426950408889634f;
  a.thresh = p_drop > 0.f ? dropout_threshold(p_drop) : 0u;
  a.keep_scale = 1.0f / (1.0f - p_drop);
  a.seed = seed; a.seed_dev = seed_dev; a.site = site; a.round_out = round_tf32;
  a.group_batches = group_batches; a.group_site_stride = group_site_stride;
  dim3 grid(N / TQ, B * H);
  if (a.thresh != 0u && premasked) attn_fwd_kernel<true, true><<<grid, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, a);
  else if (a.thresh != 0u) attn_fwd_kernel<true, false><<<grid, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, a);
  else attn_fwd_kernel<false, false><<<grid, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, a);
  return launch_status("attention_fwd");
}

extern "C" int corrif_attention_fwd(const float* qkv, float* O, float* lse, uint32_t* maskbits,
                                    int32_t B, int32_t N, int32_t H, int32_t D, float scale,
                                    float p_drop, uint64_t seed,
                                    const uint64_t* seed_dev, uint32_t site, int32_t group_batches,
                                    uint32_t group_site_stride, int32_t round_tf32, void* stream) {
  return attention_fwd_launch(qkv, O, lse, maskbits, B, N, H, D, scale, p_drop, seed, seed_dev, site, group_batches,
                              group_site_stride, round_tf32, false, stream);
}

extern "C" int corrif_attention_fwd_premasked(const float* qkv, float* O, float* lse, const uint32_t* maskbits,
                                              int32_t B, int32_t N, int32_t H, int32_t D, float scale,
                                              float p_drop, int32_t round_tf32, void* stream) {
  CORRIF_REQUIRE(p_drop > 0.f && maskbits != nullptr, "attention_fwd_premasked: needs p_drop > 0 and the keep bits");
  return attention_fwd_launch(qkv, O, lse, const_cast<uint32_t*>(maskbits), B, N, H, D, scale, p_drop, 0, nullptr, 0, 0, 0,
                              round_tf32, true, stream);
}
