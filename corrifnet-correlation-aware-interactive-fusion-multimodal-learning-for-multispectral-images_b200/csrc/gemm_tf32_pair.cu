// tcgen05 TF32 GEMM, CTA-pair variant (cta_group::2) with a TMA epilogue.  sm_100a only.
//
// Why pairs.  A single-CTA 128 x 256 TF32 tile needs 48 KB of fp32 operands per 512 tensor-core
// cycles = 96 B/clk/SM, but the L2 -> SM fabric delivers ~43 B/clk/SM chip-wide (ncu: tensor pipe
// capped near 40 %).  Here the two CTAs of a cluster (one TPC) share a 256 x 256 tile: each stages its
// own 128 rows of A and HALF of B (128 rows); the leader's tcgen05.mma.cta_group::2 (M = 256) reads
// both halves of B out of the two shared memories, and every CTA ends up with its own 128 x 256
// accumulator in its own TMEM.  Operand traffic per MAC drops by a third and a stage is 32 KB.
//
// Why a TMA epilogue.  With K = 512 a tile's main loop is ~8 k cycles; the row-per-thread / smem-
// transposed epilogue of gemm_tf32.cu costs ~12 k cycles per tile (~70 instructions per float4:
// 64-bit addressing, bounds checks, mode switch) and was the bottleneck (measured with the
// CORRIF_GEMM_TIMING counters below).  Here each epilogue warp turns a 32 x 16 accumulator chunk into
// a swizzled 2 KB shared-memory box and one elected lane hands it to the TMA unit:
//   plain / bias / GELU outputs      cp.async.bulk.tensor store
//   split-K weight gradients         cp.reduce.async.bulk.tensor .add (no per-element atomics)
//   residual / saved pre-activation  cp.async.bulk.tensor LOAD into the same box several chunks ahead,
//                                    combined in place, stored from there
// Edge tiles need no code: TMA clips stores and zero-fills loads.
// Shared memory (227 KB) is split at launch time between the operand ring and the epilogue boxes:
// store-only epilogues run 6 stages + 2 boxes per warp (the K = 512 main loops stalled on the
// release -> TMA -> full round trip with 4), loading epilogues 4 stages + 6 boxes (4 chunks in flight).
//
// Barriers (S = smem stage, T = TMEM accumulator buffer, double-buffered):
//   full[S]    leader only: its producer arms 2 x STAGE bytes, both CTAs' loads credit it
//   empty[S]   one per CTA, released for both by the leader's multicast tcgen05.commit
//   tfull[T]   one per CTA (multicast commit): this CTA's accumulator is complete
//   tempty[T]  leader only: 8 epilogue warps x 2 CTAs arrive (the peer's through DSMEM)
//   ldbar      per epilogue warp x 3 boxes: residual / aux chunk has landed
#include "gemm_tc.cuh"

namespace corrif {
namespace tc {

constexpr int PAIR_THREADS = 320;      // producer warp, MMA warp, 8 epilogue warps
constexpr int PAIR_BN = 256;
constexpr int MAX_STAGES = 6, MAX_BOXES = 6;
constexpr int CW = 16;                 // epilogue chunk: 32 rows x 16 columns
constexpr int BOX_BYTES = 32 * CW * 4; // 2 KB, rows of 64 B, SWIZZLE_64B
constexpr int PAIR_B_BYTES = (PAIR_BN / 2) * ROW_BYTES;
constexpr int PAIR_STAGE_BYTES = A_BYTES + PAIR_B_BYTES;
constexpr int PAIR_SMEM = 6 * PAIR_STAGE_BYTES + 8 * 2 * BOX_BYTES + 1024;   // == 5 + 4 boxes == 4 + 6 boxes
static_assert(PAIR_STAGE_BYTES == 8 * 2 * BOX_BYTES, "one stage must equal two boxes per epilogue warp");

struct PairArgs {
  KernelArgs k;
  int64_t ldd;
  int has_in;                 // epilogue reads a second operand (residual or saved pre-activation)
  int bpc;                    // boxes per chunk: 2 for BIAS_GELU (D and the pre-activation), else 1
  int stages, nbox;           // smem split: operand ring depth / epilogue boxes per warp
  unsigned long long* dbg;    // CORRIF_GEMM_TIMING=1: per-role wait cycles of cluster 0, else null
};

template <bool A_MN, bool B_MN>
__device__ __forceinline__ constexpr uint32_t make_idesc_pair() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
         ((uint32_t)(PAIR_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               :: "l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               :: "l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}


// GELU / GELU' for the TF32 path: erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the
// TF32 operand rounding), sharing exp(-x^2/2) between the erf tail and the Gaussian density.  About a
// third of the instructions of erff() + expf(), which is what bounded the GELU epilogues.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& ez) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  ez = ex2_approx(-1.4426950408889634f * z * z);                  // exp(-x^2 / 2)
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  q = 0.5f * q * t * ez;                                          // 0.5 * (1 - erf(z))
  cdf = x >= 0.f ? 1.0f - q : q;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float cdf, ez;
  gelu_parts(x, cdf, ez);
  return x * cdf;
}
__device__ __forceinline__ float dgelu_fast(float x) {
  float cdf, ez;
  gelu_parts(x, cdf, ez);
  return fmaf(x * 0.39894228040143267794f, ez, cdf);
}

// One 32 x 16 chunk, thread = row: straight-line code per (epilogue mode, number of dropout sites) so
// that the four float4 groups (and their RNG chains) interleave - with two epilogue warps per
// scheduler the epilogue is latency-bound, not throughput-bound.
template <int MODE, int SITES>
__device__ __forceinline__ void epi_chunk(const EpiArgs& e, const uint32_t (&r)[CW], const float4 (&bias)[CW / 4],
                                          uint32_t box, uint32_t box2, uint32_t my_row, uint32_t my_xor,
                                          int m, int n) {
  float4 v[CW / 4], in[CW / 4];
  uint32_t km[CW / 4];
#pragma unroll
  for (int j = 0; j < CW / 4; ++j) {
    if (MODE == CORRIF_EPI_BIAS_RESIDUAL || MODE == CORRIF_EPI_MUL_DGELU)
      in[j] = lds4(box + my_row + (((uint32_t)j ^ my_xor) << 4));
    if (SITES > 0) {
      const uint64_t quad = ((uint64_t)m * (uint64_t)e.N + (uint64_t)(n + 4 * j)) >> 2;
      km[j] = dropout_keepmask4(e.key_a, quad, e.drop_thresh);
      if (SITES > 1) km[j] &= dropout_keepmask4(e.key_b, quad, e.drop_thresh);
    }
  }
#pragma unroll
  for (int j = 0; j < CW / 4; ++j) {
    v[j] = make_float4(__uint_as_float(r[4 * j]) * e.alpha, __uint_as_float(r[4 * j + 1]) * e.alpha,
                       __uint_as_float(r[4 * j + 2]) * e.alpha, __uint_as_float(r[4 * j + 3]) * e.alpha);
    if (MODE == CORRIF_EPI_BIAS || MODE == CORRIF_EPI_BIAS_GELU || MODE == CORRIF_EPI_BIAS_RESIDUAL) {
      v[j].x += bias[j].x; v[j].y += bias[j].y; v[j].z += bias[j].z; v[j].w += bias[j].w;
    }
    if (MODE == CORRIF_EPI_BIAS_GELU) {
      sts4(box2 + my_row + (((uint32_t)j ^ my_xor) << 4), v[j]);
      v[j] = make_float4(gelu_fast(v[j].x), gelu_fast(v[j].y), gelu_fast(v[j].z), gelu_fast(v[j].w));
    }
    if (MODE == CORRIF_EPI_MUL_DGELU) {
      v[j].x *= dgelu_fast(in[j].x); v[j].y *= dgelu_fast(in[j].y);
      v[j].z *= dgelu_fast(in[j].z); v[j].w *= dgelu_fast(in[j].w);
    }
    if (SITES > 0) {
      v[j].x = (km[j] & 1u) ? v[j].x * e.drop_scale : 0.f; v[j].y = (km[j] & 2u) ? v[j].y * e.drop_scale : 0.f;
      v[j].z = (km[j] & 4u) ? v[j].z * e.drop_scale : 0.f; v[j].w = (km[j] & 8u) ? v[j].w * e.drop_scale : 0.f;
    }
    if (MODE == CORRIF_EPI_BIAS_RESIDUAL) {
      v[j].x += in[j].x; v[j].y += in[j].y; v[j].z += in[j].z; v[j].w += in[j].w;
    }
    if (e.round_tf32) v[j] = round_tf32_4(v[j]);
    sts4(box + my_row + (((uint32_t)j ^ my_xor) << 4), v[j]);
  }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
gemm_tf32_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmIn,
                      const __grid_constant__ CUtensorMap tmAux, const PairArgs pa, const int mt, const int nt,
                      const int total_tiles) {
  constexpr int BN = PAIR_BN, BNH = BN / 2, STAGE_BYTES = PAIR_STAGE_BYTES;
  const int STAGES = pa.stages;
  constexpr uint32_t TMEM_COLS = 512;                // two 256-column accumulator buffers
  const KernelArgs& args = pa.k;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t ld_bar[8][MAX_BOXES];
  __shared__ uint32_t tmem_base_holder;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();           // 0 = leader (issues the MMAs)
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int total_kb = (args.K + BK - 1) / BK;
  const int kb_per_split = (total_kb + args.split_k - 1) / args.split_k;   // host guarantees: no empty split
  const int tiles_per_z = mt * nt;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&tmD) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 16); }
    for (int w = 0; w < 8; ++w)
      for (int b = 0; b < MAX_BOXES; ++b) mbar_init(&ld_bar[w][b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_pair(&tmem_base_holder, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                // peer barriers initialised, both TMEMs allocated
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_holder;

  // tile -> (z, 256-row m block, n block, first k-block, k-blocks); n fastest: CTAs running together
  // share the A rows (read from DRAM once); B (weights, <= 3 MB) stays L2-resident anyway
  auto decode = [&](int tile, int& z, int& m0, int& n0, int& kb_begin, int& num_kb) {
    z = tile / tiles_per_z;
    const int rem = tile - z * tiles_per_z;
    n0 = (rem % nt) * BN;
    m0 = (rem / nt) * (2 * BM);
    const int split = z % args.split_k;
    kb_begin = split * kb_per_split;
    num_kb = min(total_kb, kb_begin + kb_per_split) - kb_begin;
  };

  if (warp == 0 && lane == 0) {
    // ================= TMA producer (both CTAs: own A rows, own half of B) =================
    const uint32_t full0 = mapa_shared(smem_u32(&full_bar[0]), 0);    // the leader's full barriers
    int s = 0;
    uint32_t ph = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      int z, m0, n0, kb_begin, num_kb;
      decode(tile, z, m0, n0, kb_begin, num_kb);
      const int batch = z / args.split_k;
      const int bi = batch % args.batch_inner, bo = batch / args.batch_inner;
      const int64_t a_off = bo * args.a_bo + bi * args.a_bi;
      const int64_t b_off = bo * args.b_bo + bi * args.b_bi;
      const int a_c0 = (int)(a_off % args.lda), a_c1 = (int)(a_off / args.lda);
      const int b_c0 = (int)(b_off % args.ldb), b_c1 = (int)(b_off / args.ldb);
      const int mr = m0 + (int)rank * BM, nr = n0 + (int)rank * BNH;
      for (int i = 0; i < num_kb; ++i) {
        const long long t0 = pa.dbg ? clock64() : 0;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        if (pa.dbg && blockIdx.x < 2) atomicAdd(&pa.dbg[rank], (unsigned long long)(clock64() - t0));
        if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
        const uint32_t fb = full0 + (uint32_t)s * 8u;
        const int k0 = (kb_begin + i) * BK;
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
        if (!A_MN) {
          tma_load_2d_pair(sa, &tmA, fb, a_c0 + k0, a_c1 + mr);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j)
            tma_load_2d_pair(sa + j * (BK * ROW_BYTES), &tmA, fb, a_c0 + mr + 32 * j, a_c1 + k0);
        }
        if (!B_MN) {
          tma_load_2d_pair(sb, &tmB, fb, b_c0 + k0, b_c1 + nr);
        } else {
#pragma unroll
          for (int j = 0; j < BNH / 32; ++j)
            tma_load_2d_pair(sb + j * (BK * ROW_BYTES), &tmB, fb, b_c0 + nr + 32 * j, b_c1 + k0);
        }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ================= MMA issuer (leader CTA only) =================
    constexpr uint32_t idesc = make_idesc_pair<A_MN, B_MN>();
    uint32_t tc = 0, ph = 0;
    int s = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++tc) {
      int z, m0, n0, kb_begin, num_kb;
      decode(tile, z, m0, n0, kb_begin, num_kb);
      const uint32_t buf = tc & 1u;
      const long long t0 = pa.dbg ? clock64() : 0;
      mbar_wait(&tempty_bar[buf], ((tc >> 1) & 1u) ^ 1u);      // both CTAs' epilogues drained this buffer
      if (pa.dbg && blockIdx.x == 0) atomicAdd(&pa.dbg[2], (unsigned long long)(clock64() - t0));
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + buf * BN;
      for (int i = 0; i < num_kb; ++i) {
        const long long t1 = pa.dbg ? clock64() : 0;
        mbar_wait(&full_bar[s], ph);
        if (pa.dbg && blockIdx.x == 0) atomicAdd(&pa.dbg[3], (unsigned long long)(clock64() - t1));
        tcgen05_fence_after();
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t ad = make_smem_desc<A_MN>(sa + k * (A_MN ? 1024 : UMMA_K * 4));
          const uint64_t bd = make_smem_desc<B_MN>(sb + k * (B_MN ? 1024 : UMMA_K * 4));
          tcgen05_mma_tf32_pair(tacc, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        tcgen05_commit_pair(&empty_bar[s], 3);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      tcgen05_commit_pair(&tfull_bar[buf], 3);
    }
  } else if (warp >= 2) {
    // ================= epilogue (both CTAs, own 128 rows x BN columns) =================
    const int ew = warp - 2, quad = warp & 3, half = ew >> 2;
    const int nbox = pa.nbox, bpc = pa.bpc;
    const uint32_t boxes = epi_base + (uint32_t)(ew * nbox) * BOX_BYTES;
    const uint32_t my_row = (uint32_t)lane * 64u, my_xor = (uint32_t)((lane >> 1) & 3);   // SWIZZLE_64B
    const uint32_t tempty0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    EpiArgs e = args.epi;
    epi_setup_dropout(e, args.drop);
    const float* const bias0 = e.bias;
    int cur_bo = 0;
    const int mode = e.mode;
    const int sites = e.drop_thresh == 0u ? 0 : (e.two_sites ? 2 : 1);
    const bool has_in = pa.has_in != 0;
    const bool has_bias = mode == CORRIF_EPI_BIAS || mode == CORRIF_EPI_BIAS_GELU || mode == CORRIF_EPI_BIAS_RESIDUAL;
    float4 bias_tile = make_float4(0.f, 0.f, 0.f, 0.f);   // lane l: bias of chunk l / 4, float4 group l % 4
    const int pfd = nbox - 2;                         // prefetch distance of the residual / aux loads

    // This warp's chunk sequence: per tile, chunk ci covers columns nb + 32 ci .. +16 (the two warps
    // of a lane quadrant interleave), rows m .. m+32; chunks beyond N are not visited.
    struct It { int tile, ci, nvalid, m, nb, dc0, dc1, bo; };
    auto load_tile = [&](It& t) {
      while (t.tile < total_tiles) {
        int z, m0, n0, kb_begin, num_kb;
        decode(t.tile, z, m0, n0, kb_begin, num_kb);
        t.nb = n0 + half * CW;
        const int nv = (e.N - t.nb + 2 * CW - 1) / (2 * CW);
        t.nvalid = nv < 0 ? 0 : (nv > BN / (2 * CW) ? BN / (2 * CW) : nv);
        if (t.nvalid > 0) {
          const int batch = z / args.split_k;
          const int bi = batch % args.batch_inner, bo = batch / args.batch_inner;
          const int64_t doff = bo * args.d_bo + bi * args.d_bi;
          t.m = m0 + (int)rank * BM + quad * 32;
          t.dc0 = (int)(doff % pa.ldd) + t.nb;
          t.dc1 = (int)(doff / pa.ldd) + t.m;
          t.bo = bo;
          t.ci = 0;
          return;
        }
        t.tile += num_clusters;
      }
    };
    auto advance = [&](It& t) {
      if (++t.ci >= t.nvalid) { t.tile += num_clusters; load_tile(t); }
    };
    It cur{cluster_id, 0, 0, 0, 0, 0, 0, 0};
    load_tile(cur);
    It pf = cur;
    uint32_t cc = 0, tc = 0;
    int cur_tile = -1;
    int box_i = 0, pf_box = 0;                        // box of chunk cc / of the next chunk to prefetch
    uint32_t ld_ph = 0;                               // phase of ld_bar[box_i]
    if (has_in) {
      for (int d = 0; d < pfd && pf.tile < total_tiles; ++d) {
        if (lane == 0) {
          mbar_expect_tx(&ld_bar[ew][pf_box], BOX_BYTES);
          tma_load_2d(boxes + pf_box * BOX_BYTES, &tmIn, &ld_bar[ew][pf_box], pf.dc0 + pf.ci * 2 * CW, pf.dc1);
        }
        advance(pf);
        if (++pf_box == nbox) pf_box = 0;
      }
    }
    long long t_wait = 0, t_work = 0, t_ld = 0, t_tm = 0;
    const long long t_loop0 = pa.dbg ? clock64() : 0;
    while (cur.tile < total_tiles) {
      const uint32_t box = boxes + (uint32_t)box_i * BOX_BYTES;
      const int box2_i = box_i + 1 == nbox ? 0 : box_i + 1;
      const uint32_t box2 = boxes + (uint32_t)box2_i * BOX_BYTES;      // BIAS_GELU: pre-activation out
      if (lane == 0) {
        // the box about to be refilled must have been read by its TMA store.  Loading / two-output
        // epilogues refill the box of chunk cc-2 (one store group may stay pending); store-only ones
        // rotate through all nbox boxes, so nbox-1 groups may stay pending - the stores queue behind
        // the operand loads in the TMA unit and need more than two chunks of time to drain.
        if (has_in || bpc == 2 || nbox == 2) bulk_wait_read<1>();
        else if (nbox == 4) bulk_wait_read<3>();
        else bulk_wait_read<5>();
        if (has_in && pf.tile < total_tiles) {
          mbar_expect_tx(&ld_bar[ew][pf_box], BOX_BYTES);
          tma_load_2d(boxes + pf_box * BOX_BYTES, &tmIn, &ld_bar[ew][pf_box], pf.dc0 + pf.ci * 2 * CW, pf.dc1);
        }
      }
      if (has_in && pf.tile < total_tiles) {
        advance(pf);
        if (++pf_box == nbox) pf_box = 0;
      }
      __syncwarp();
      const long long t0 = pa.dbg ? clock64() : 0;
      if (cur.tile != cur_tile) {                     // first chunk of a new tile: accumulator ready?
        if (cur_tile >= 0) ++tc;
        cur_tile = cur.tile;
        if (cur.bo != cur_bo) {                        // batched problems: own bias row / dropout sites
          cur_bo = cur.bo;
          if (bias0) e.bias = bias0 + (int64_t)cur_bo * args.bias_bo;
          if (args.drop.site_bo) epi_setup_dropout(e, args.drop, cur_bo);
        }
        if (has_bias) {                                // this warp's 8 x 16 bias values of the tile, one load
          const int nbias = cur.nb + (lane >> 2) * 2 * CW + (lane & 3) * 4;
          bias_tile = nbias < e.N ? __ldg(reinterpret_cast<const float4*>(e.bias + nbias)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(&tfull_bar[tc & 1u], (tc >> 1) & 1u);
        tcgen05_fence_after();
      }
      const long long t1 = pa.dbg ? clock64() : 0;
      uint32_t r[16];
      tmem_ld16(tmem_base + (tc & 1u) * BN + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * CW + cur.ci * 2 * CW), r);
      if (cur.ci == cur.nvalid - 1) {                 // last chunk of this tile: hand the buffer back
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty0 + (tc & 1u) * 8u);
      }
      const long long t2 = pa.dbg ? clock64() : 0;
      if (has_in) mbar_wait(&ld_bar[ew][box_i], ld_ph);
      const long long t3 = pa.dbg ? clock64() : 0;
      const int m = cur.m + lane;
      const int nc = cur.nb + cur.ci * 2 * CW;
      float4 bias[CW / 4];
      if (has_bias) {
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {                      // lane 4 ci + j holds this chunk's j-th group
          const int src = cur.ci * (CW / 4) + j;
          bias[j] = make_float4(__shfl_sync(0xffffffffu, bias_tile.x, src), __shfl_sync(0xffffffffu, bias_tile.y, src),
                                __shfl_sync(0xffffffffu, bias_tile.z, src), __shfl_sync(0xffffffffu, bias_tile.w, src));
        }
      }
      switch (mode * 4 + sites) {
        case CORRIF_EPI_BIAS * 4: epi_chunk<CORRIF_EPI_BIAS, 0>(e, r, bias, box, box2, my_row, my_xor, m, nc); break;
        case CORRIF_EPI_BIAS_GELU * 4: epi_chunk<CORRIF_EPI_BIAS_GELU, 0>(e, r, bias, box, box2, my_row, my_xor, m, nc); break;
        case CORRIF_EPI_BIAS_GELU * 4 + 1: epi_chunk<CORRIF_EPI_BIAS_GELU, 1>(e, r, bias, box, box2, my_row, my_xor, m, nc); break;
        case CORRIF_EPI_BIAS_RESIDUAL * 4: epi_chunk<CORRIF_EPI_BIAS_RESIDUAL, 0>(e, r, bias, box, box2, my_row, my_xor, m, nc); break;
        case CORRIF_EPI_BIAS_RESIDUAL * 4 + 1: epi_chunk<CORRIF_EPI_BIAS_RESIDUAL, 1>(e, r, bias, box, box2, my_row, my_xor, m, nc); break;
        case CORRIF_EPI_BIAS_RESIDUAL * 4 + 2: epi_chunk<CORRIF_EPI_BIAS_RESIDUAL, 2>(e, r, bias, box, box2, my_row, my_xor, m, nc); break;
        case CORRIF_EPI_MUL_DGELU * 4: epi_chunk<CORRIF_EPI_MUL_DGELU, 0>(e, r, bias, box, box2, my_row, my_xor, m, nc); break;
        case CORRIF_EPI_MUL_DGELU * 4 + 1: epi_chunk<CORRIF_EPI_MUL_DGELU, 1>(e, r, bias, box, box2, my_row, my_xor, m, nc); break;
        default: epi_chunk<CORRIF_EPI_STORE, 0>(e, r, bias, box, box2, my_row, my_xor, m, nc); break;   // STORE, ATOMIC_ADD
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        const int c0 = cur.dc0 + cur.ci * 2 * CW;
        if (mode == CORRIF_EPI_ATOMIC_ADD) tma_reduce_add_2d(&tmD, box, c0, cur.dc1);
        else tma_store_2d(&tmD, box, c0, cur.dc1);
        if (mode == CORRIF_EPI_BIAS_GELU) tma_store_2d(&tmAux, box2, c0, cur.dc1);
        bulk_commit();
      }
      if (pa.dbg) { t_wait += t1 - t0; t_work += clock64() - t1; t_ld += t3 - t2; t_tm += t2 - t1; }
      ++cc;
      box_i += bpc;
      if (box_i >= nbox) { box_i -= nbox; ld_ph ^= 1u; }
      advance(cur);
    }
    if (lane == 0) bulk_wait_read<0>();
    if (pa.dbg && blockIdx.x == 0 && warp == 2 && lane == 0) {
      atomicAdd(&pa.dbg[4], (unsigned long long)t_wait);
      atomicAdd(&pa.dbg[5], (unsigned long long)t_work);
      atomicAdd(&pa.dbg[6], (unsigned long long)cc);
      atomicAdd(&pa.dbg[7], (unsigned long long)t_ld);
      atomicAdd(&pa.dbg[8], (unsigned long long)t_tm);
      atomicAdd(&pa.dbg[9], (unsigned long long)(clock64() - t_loop0));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                // nobody leaves while the peer may still touch it
  if (warp == 1) tmem_dealloc_pair(tmem_base, TMEM_COLS);
}

template <bool A_MN, bool B_MN>
static int launch_pair_variant(const corrif_gemm_desc& g, const CUtensorMap (&tm)[5], cudaStream_t stream) {
  auto kern = gemm_tf32_pair_kernel<A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM);
    if (e != cudaSuccess) { set_last_error("gemm_tf32(pair): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  PairArgs pa;
  KernelArgs& a = pa.k;
  a.epi = make_epi_args(g);
  a.drop = make_drop_args(g);
  a.K = g.K; a.batch_inner = g.batch_inner;
  // no empty trailing split: the epilogue's chunk sequence assumes every tile is computed
  const int total_kb = (g.K + BK - 1) / BK;
  const int kb_per = (total_kb + g.split_k - 1) / g.split_k;
  a.split_k = (total_kb + kb_per - 1) / kb_per;
  a.a_bo = g.a_bo; a.a_bi = g.a_bi; a.b_bo = g.b_bo; a.b_bi = g.b_bi; a.d_bo = g.d_bo; a.d_bi = g.d_bi;
  a.lda = g.lda; a.ldb = g.ldb; a.bias_bo = g.bias_bo;
  pa.ldd = g.ldd;
  pa.has_in = (g.epilogue == CORRIF_EPI_BIAS_RESIDUAL || g.epilogue == CORRIF_EPI_MUL_DGELU) ? 1 : 0;
  pa.bpc = g.epilogue == CORRIF_EPI_BIAS_GELU ? 2 : 1;
  // smem split (see the header): every combination fills the same PAIR_SMEM bytes
  pa.stages = pa.has_in ? 4 : (pa.bpc == 2 ? 5 : 6);
  pa.nbox = pa.has_in ? 6 : (pa.bpc == 2 ? 4 : 2);
  static const char* cfg_env = getenv("CORRIF_PAIR_STAGES");        // tuning override: 4, 5 or 6
  if (cfg_env) {
    const int st_ = atoi(cfg_env);
    if (st_ >= 4 && st_ <= 6 && (!pa.has_in || st_ <= 5) && (pa.bpc == 1 || st_ <= 5)) {
      pa.stages = st_; pa.nbox = 2 + 2 * (6 - st_);
    }
  }
  static const bool timing = getenv("CORRIF_GEMM_TIMING") != nullptr;
  static unsigned long long* dbg = nullptr;
  pa.dbg = nullptr;
  if (timing) {
    if (!dbg) cudaMalloc(&dbg, 16 * sizeof(unsigned long long));
    cudaMemsetAsync(dbg, 0, 16 * sizeof(unsigned long long), stream);
    pa.dbg = dbg;
  }
  const int mt = (g.M + 2 * BM - 1) / (2 * BM), nt = (g.N + PAIR_BN - 1) / PAIR_BN;
  const int total = mt * nt * g.batch_outer * g.batch_inner * a.split_k;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(PAIR_THREADS, 1, 1);
  cfg.dynamicSmemBytes = PAIR_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // The static tile striding is only balanced if every cluster of the grid is resident at once; how
  // many CTA pairs the GPCs can host is a property of the part (it is NOT always SMs / 2).
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cfg.gridDim = dim3(num_sms(), 1, 1);
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = num_sms() / 2; }
    max_clusters = n < num_sms() / 2 ? n : num_sms() / 2;
    if (const char* ov = getenv("CORRIF_PAIR_CLUSTERS")) { const int v = atoi(ov); if (v > 0 && v <= max_clusters) max_clusters = v; }
    if (timing) fprintf(stderr, "[pair] max co-resident clusters: %d\n", max_clusters);
  }
  const int clusters = total < max_clusters ? total : max_clusters;
  cfg.gridDim = dim3(2 * clusters, 1, 1);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm[0], tm[1], tm[2], tm[3], tm[4], pa, mt, nt, total);
  if (e != cudaSuccess) { set_last_error("gemm_tf32(pair): launch: %s", cudaGetErrorString(e)); return (int)e; }
  if (timing) {
    unsigned long long h[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[pair M%d N%d K%d split%d epi%d] cluster0 cycles: prod wait empty %llu/%llu  mma wait tempty %llu"
            "  mma wait full %llu  epi wait acc %llu  epi work %llu (tmem ld %llu, wait in %llu)  chunks %llu  epi loop total %llu\n", g.M, g.N, g.K, a.split_k,
            g.epilogue, h[0], h[1], h[2], h[3], h[4], h[5], h[8], h[7], h[6], h[9]);
  }
  return launch_status("gemm_tf32_pair");
}

}  // namespace tc

bool gemm_tf32_pair_supported(const corrif_gemm_desc& g) {
  if (g.N <= 128 || g.M <= 128) return false;
  if (g.batch_outer * g.batch_inner > 1) {
    // batched problems address D through one map over the whole buffer: TMA cannot clip per batch
    if (g.M % 256 != 0 || g.N % 32 != 0) return false;
    // ... and every problem's D must start at column 0 of a row of that map
    if (g.d_bo % g.ldd != 0 || g.d_bi % g.ldd != 0 || g.N > g.ldd) return false;
    if (g.residual && g.ldr != g.ldd) return false;
    if (g.aux && g.ldaux != g.ldd) return false;
  }
  return true;
}

int gemm_tf32_pair_launch(const corrif_gemm_desc& g, cudaStream_t stream) {
  using namespace tc;
  const bool batched = g.batch_outer * g.batch_inner > 1;
  auto span = [&](int64_t bo_stride, int64_t bi_stride) {
    return (int64_t)(g.batch_outer - 1) * bo_stride + (int64_t)(g.batch_inner - 1) * bi_stride;
  };
  CUtensorMap tm[5];
  int st;
  {
    const int64_t extra = span(g.a_bo, g.a_bi);
    const int64_t rows = g.a_mn_major ? g.K : g.M, cols = g.a_mn_major ? g.M : g.K;
    const uint64_t dim0 = batched ? (uint64_t)g.lda : (uint64_t)cols;
    const uint64_t dim1 = (uint64_t)(rows + (batched ? (extra + g.lda - 1) / g.lda : 0));
    if ((st = encode_map(&tm[0], g.A, dim0, dim1, g.lda, 32, g.a_mn_major ? BK : BM, g.a_mn_major != 0))) return st;
  }
  {
    const int64_t extra = span(g.b_bo, g.b_bi);
    const int64_t rows = g.b_mn_major ? g.K : g.N, cols = g.b_mn_major ? g.N : g.K;
    const uint64_t dim0 = batched ? (uint64_t)g.ldb : (uint64_t)cols;
    const uint64_t dim1 = (uint64_t)(rows + (batched ? (extra + g.ldb - 1) / g.ldb : 0));
    if ((st = encode_map(&tm[1], g.B, dim0, dim1, g.ldb, 32, g.b_mn_major ? BK : PAIR_BN / 2, g.b_mn_major != 0)))
      return st;
  }
  {
    // output-side maps: boxes of 32 rows x 16 columns, SWIZZLE_64B
    const int64_t extra = span(g.d_bo, g.d_bi);
    auto out_map = [&](CUtensorMap* m, const float* base, int64_t ld) {
      const uint64_t dim0 = batched ? (uint64_t)ld : (uint64_t)g.N;
      const uint64_t dim1 = (uint64_t)(g.M + (batched ? (extra + ld - 1) / ld : 0));
      return encode_map_swz(m, base, dim0, dim1, ld, CW, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    };
    if ((st = out_map(&tm[2], g.D, g.ldd))) return st;
    tm[3] = tm[2];
    tm[4] = tm[2];
    if (g.epilogue == CORRIF_EPI_BIAS_RESIDUAL) { if ((st = out_map(&tm[3], g.residual, g.ldr))) return st; }
    if (g.epilogue == CORRIF_EPI_MUL_DGELU) { if ((st = out_map(&tm[3], g.aux, g.ldaux))) return st; }
    if (g.epilogue == CORRIF_EPI_BIAS_GELU) { if ((st = out_map(&tm[4], g.aux, g.ldaux))) return st; }
  }
  if (!g.a_mn_major && !g.b_mn_major) return launch_pair_variant<false, false>(g, tm, stream);
  if (!g.a_mn_major && g.b_mn_major) return launch_pair_variant<false, true>(g, tm, stream);
  if (g.a_mn_major && !g.b_mn_major) return launch_pair_variant<true, false>(g, tm, stream);
  return launch_pair_variant<true, true>(g, tm, stream);
}

}  // namespace corrif
