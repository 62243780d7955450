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
// CORRIF_GEMM_TIMING counters below).  Here each epilogue warp turns a 32 x 32 accumulator chunk into
// a swizzled 4 KB shared-memory box and one elected lane hands it to the TMA unit:
//   plain / bias / GELU outputs      cp.async.bulk.tensor store
//   split-K weight gradients         cp.reduce.async.bulk.tensor .add (no per-element atomics)
//   residual / saved pre-activation  cp.async.bulk.tensor LOAD into the same box one chunk ahead,
//                                    combined in place, stored from there
// Edge tiles need no code: TMA clips stores and zero-fills loads.
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
constexpr int PAIR_STAGES = 4;
constexpr int BOX_BYTES = 32 * 32 * 4; // one epilogue chunk: 32 rows x 128 B, SWIZZLE_128B
constexpr int EPI_BOXES = 3;           // per warp: rotating boxes (prefetch distance 1, store in flight)
constexpr int PAIR_B_BYTES = (PAIR_BN / 2) * ROW_BYTES;
constexpr int PAIR_STAGE_BYTES = A_BYTES + PAIR_B_BYTES;
constexpr int PAIR_SMEM = PAIR_STAGES * PAIR_STAGE_BYTES + 8 * EPI_BOXES * BOX_BYTES + 1024;

struct PairArgs {
  KernelArgs k;
  int64_t ldd;
  int has_in;                 // epilogue reads a second operand (residual or saved pre-activation)
  int has_aux_out;            // BIAS_GELU: also stores the pre-activation
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

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
gemm_tf32_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmIn,
                      const __grid_constant__ CUtensorMap tmAux, const PairArgs pa, const int mt, const int nt,
                      const int total_tiles) {
  constexpr int BN = PAIR_BN, BNH = BN / 2, STAGES = PAIR_STAGES, STAGE_BYTES = PAIR_STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 512;                // two 256-column accumulator buffers
  const KernelArgs& args = pa.k;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t ld_bar[8][EPI_BOXES];
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
      for (int b = 0; b < EPI_BOXES; ++b) mbar_init(&ld_bar[w][b], 1);
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
    uint32_t it = 0;
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
      for (int i = 0; i < num_kb; ++i, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
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
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ================= MMA issuer (leader CTA only) =================
    constexpr uint32_t idesc = make_idesc_pair<A_MN, B_MN>();
    uint32_t it = 0, tc = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++tc) {
      int z, m0, n0, kb_begin, num_kb;
      decode(tile, z, m0, n0, kb_begin, num_kb);
      const uint32_t buf = tc & 1u;
      const long long t0 = pa.dbg ? clock64() : 0;
      mbar_wait(&tempty_bar[buf], ((tc >> 1) & 1u) ^ 1u);      // both CTAs' epilogues drained this buffer
      if (pa.dbg && blockIdx.x == 0) atomicAdd(&pa.dbg[2], (unsigned long long)(clock64() - t0));
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + buf * BN;
      for (int i = 0; i < num_kb; ++i, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
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
      }
      tcgen05_commit_pair(&tfull_bar[buf], 3);
    }
  } else if (warp >= 2) {
    // ================= epilogue (both CTAs, own 128 rows x BN columns) =================
    const int ew = warp - 2, quad = warp & 3, half = ew >> 2;
    const uint32_t boxes = epi_base + (uint32_t)ew * (EPI_BOXES * BOX_BYTES);
    const uint32_t my_row = (uint32_t)lane * 128u, my_xor = (uint32_t)(lane & 7);
    const uint32_t tempty0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    EpiArgs e = args.epi;
    epi_setup_dropout(e, args.drop);
    const int mode = e.mode;
    const bool has_in = pa.has_in != 0;
    constexpr int NCH = BN / 64;                      // chunks of 32 columns per warp and tile

    // chunk iterator over this warp's (tile, ci) sequence, skipping chunks that lie beyond N
    struct Chunk { int tile, ci, m, n, dc0, dc1; };
    auto locate = [&](Chunk& ch) {                    // fills coordinates; advances past invalid chunks
      while (ch.tile < total_tiles) {
        int z, m0, n0, kb_begin, num_kb;
        decode(ch.tile, z, m0, n0, kb_begin, num_kb);
        const int n = n0 + (half + 2 * ch.ci) * 32;
        if (n < e.N) {
          const int batch = z / args.split_k;
          const int bi = batch % args.batch_inner, bo = batch / args.batch_inner;
          const int64_t doff = bo * args.d_bo + bi * args.d_bi;
          ch.m = m0 + (int)rank * BM + quad * 32;
          ch.n = n;
          ch.dc0 = (int)(doff % pa.ldd) + n;
          ch.dc1 = (int)(doff / pa.ldd) + ch.m;
          return;
        }
        if (++ch.ci == NCH) { ch.ci = 0; ch.tile += num_clusters; }
      }
    };
    auto advance = [&](Chunk ch) {
      if (++ch.ci == NCH) { ch.ci = 0; ch.tile += num_clusters; }
      locate(ch);
      return ch;
    };
    Chunk cur{cluster_id, 0, 0, 0, 0, 0};
    locate(cur);
    uint32_t cc = 0, tc = 0;
    int cur_tile = -1;
    if (has_in && cur.tile < total_tiles && lane == 0) {
      mbar_expect_tx(&ld_bar[ew][0], BOX_BYTES);
      tma_load_2d(boxes, &tmIn, &ld_bar[ew][0], cur.dc0, cur.dc1);
    }
    long long t_wait = 0, t_work = 0;
    while (cur.tile < total_tiles) {
      const Chunk nxt = advance(cur);
      const uint32_t b = cc % EPI_BOXES;
      const uint32_t box = boxes + b * BOX_BYTES;
      if (lane == 0) {
        // box (cc+1)%3 was last read by the store of chunk cc-2; box cc%3 by the store of chunk cc-3
        if (pa.has_aux_out) bulk_wait_read<0>(); else bulk_wait_read<1>();
        if (has_in && nxt.tile < total_tiles) {
          const uint32_t nb = (cc + 1) % EPI_BOXES;
          mbar_expect_tx(&ld_bar[ew][nb], BOX_BYTES);
          tma_load_2d(boxes + nb * BOX_BYTES, &tmIn, &ld_bar[ew][nb], nxt.dc0, nxt.dc1);
        }
      }
      __syncwarp();
      const long long t0 = pa.dbg ? clock64() : 0;
      if (cur.tile != cur_tile) {                     // first chunk of a new tile: accumulator ready?
        if (cur_tile >= 0) ++tc;
        cur_tile = cur.tile;
        mbar_wait(&tfull_bar[tc & 1u], (tc >> 1) & 1u);
        tcgen05_fence_after();
      }
      const long long t1 = pa.dbg ? clock64() : 0;
      uint32_t r[32];
      tmem_ld32(tmem_base + (tc & 1u) * BN + ((uint32_t)(quad * 32) << 16) + (uint32_t)((half + 2 * cur.ci) * 32), r);
      if (nxt.tile != cur.tile) {                     // last chunk of this tile: hand the buffer back
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty0 + (tc & 1u) * 8u);
      }
      if (has_in) mbar_wait(&ld_bar[ew][b], (cc / EPI_BOXES) & 1u);
      const int m = cur.m + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = cur.n + 4 * j;
        float4 v = make_float4(__uint_as_float(r[4 * j]) * e.alpha, __uint_as_float(r[4 * j + 1]) * e.alpha,
                               __uint_as_float(r[4 * j + 2]) * e.alpha, __uint_as_float(r[4 * j + 3]) * e.alpha);
        const uint32_t sa = box + my_row + (((uint32_t)j ^ my_xor) << 4);
        if (mode == CORRIF_EPI_BIAS || mode == CORRIF_EPI_BIAS_GELU || mode == CORRIF_EPI_BIAS_RESIDUAL) {
          const float4 bv = n < e.N ? __ldg(reinterpret_cast<const float4*>(e.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
        }
        if (mode == CORRIF_EPI_BIAS_GELU) {
          sts4(boxes + ((cc + 1) % EPI_BOXES) * BOX_BYTES + my_row + (((uint32_t)j ^ my_xor) << 4), v);
          v = make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
          if (e.drop_thresh) v = epi_dropout4(e, m, n, v);
        } else if (mode == CORRIF_EPI_BIAS_RESIDUAL) {
          if (e.drop_thresh) v = epi_dropout4(e, m, n, v);
          const float4 rv = lds4(sa);
          v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
        } else if (mode == CORRIF_EPI_MUL_DGELU) {
          const float4 u = lds4(sa);
          v.x *= dgelu_erf(u.x); v.y *= dgelu_erf(u.y); v.z *= dgelu_erf(u.z); v.w *= dgelu_erf(u.w);
          if (e.drop_thresh) v = epi_dropout4(e, m, n, v);
        }
        if (e.round_tf32) v = round_tf32_4(v);
        sts4(sa, v);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (mode == CORRIF_EPI_ATOMIC_ADD) tma_reduce_add_2d(&tmD, box, cur.dc0, cur.dc1);
        else tma_store_2d(&tmD, box, cur.dc0, cur.dc1);
        if (mode == CORRIF_EPI_BIAS_GELU)
          tma_store_2d(&tmAux, boxes + ((cc + 1) % EPI_BOXES) * BOX_BYTES, cur.dc0, cur.dc1);
        bulk_commit();
      }
      if (pa.dbg) { t_wait += t1 - t0; t_work += clock64() - t1; }
      ++cc;
      cur = nxt;
    }
    if (lane == 0) bulk_wait_read<0>();
    if (pa.dbg && blockIdx.x == 0 && warp == 2 && lane == 0) {
      atomicAdd(&pa.dbg[4], (unsigned long long)t_wait);
      atomicAdd(&pa.dbg[5], (unsigned long long)t_work);
      atomicAdd(&pa.dbg[6], (unsigned long long)cc);
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
  a.drop = DropArgs{g.drop_p, g.drop_site_a, g.drop_site_b, g.drop_seed, g.drop_seed_dev};
  a.K = g.K; a.batch_inner = g.batch_inner;
  // no empty trailing split: the epilogue's chunk sequence assumes every tile is computed
  const int total_kb = (g.K + BK - 1) / BK;
  const int kb_per = (total_kb + g.split_k - 1) / g.split_k;
  a.split_k = (total_kb + kb_per - 1) / kb_per;
  a.a_bo = g.a_bo; a.a_bi = g.a_bi; a.b_bo = g.b_bo; a.b_bi = g.b_bi; a.d_bo = g.d_bo; a.d_bi = g.d_bi;
  a.lda = g.lda; a.ldb = g.ldb;
  pa.ldd = g.ldd;
  pa.has_in = (g.epilogue == CORRIF_EPI_BIAS_RESIDUAL || g.epilogue == CORRIF_EPI_MUL_DGELU) ? 1 : 0;
  pa.has_aux_out = g.epilogue == CORRIF_EPI_BIAS_GELU ? 1 : 0;
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
  const int max_clusters = num_sms() / 2;
  const int clusters = total < max_clusters ? total : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters, 1, 1);
  cfg.blockDim = dim3(PAIR_THREADS, 1, 1);
  cfg.dynamicSmemBytes = PAIR_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm[0], tm[1], tm[2], tm[3], tm[4], pa, mt, nt, total);
  if (e != cudaSuccess) { set_last_error("gemm_tf32(pair): launch: %s", cudaGetErrorString(e)); return (int)e; }
  if (timing) {
    unsigned long long h[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[pair M%d N%d K%d split%d epi%d] cluster0 cycles: prod wait empty %llu/%llu  mma wait tempty %llu"
            "  mma wait full %llu  epi wait acc %llu  epi work %llu  chunks %llu\n", g.M, g.N, g.K, a.split_k,
            g.epilogue, h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  }
  return launch_status("gemm_tf32_pair");
}

}  // namespace tc

bool gemm_tf32_pair_supported(const corrif_gemm_desc& g) {
  if (g.N <= 128 || g.M <= 128) return false;
  if (g.batch_outer * g.batch_inner > 1) {
    // batched problems address D through one map over the whole buffer: TMA cannot clip per batch
    if (g.M % 256 != 0 || g.N % 32 != 0) return false;
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
    // output-side maps: 32 x 32 boxes, SWIZZLE_128B
    const int64_t extra = span(g.d_bo, g.d_bi);
    auto out_map = [&](CUtensorMap* m, const float* base, int64_t ld) {
      const uint64_t dim0 = batched ? (uint64_t)ld : (uint64_t)g.N;
      const uint64_t dim1 = (uint64_t)(g.M + (batched ? (extra + ld - 1) / ld : 0));
      return encode_map(m, base, dim0, dim1, ld, 32, 32, false);
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
