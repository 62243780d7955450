// tcgen05 TF32 GEMM for sm_100a:  D = epilogue(alpha * A . B^T), fp32 in HBM, fp32 accumulate in TMEM.
//
//   * operands stay fp32 in HBM; TMA (cp.async.bulk.tensor, SWIZZLE_128B) stages [rows x 32 floats]
//     boxes into shared memory, tcgen05.mma.kind::tf32 reads them as TF32 (top 19 bits);
//   * one 128 x BN output tile per CTA, accumulator = BN TMEM columns x 128 lanes;
//   * warp roles: warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer
//     (one elected lane), warps 2..5 = epilogue (tcgen05.ld 32x32b, one TMEM lane quadrant each);
//   * smem ring of STAGES {A,B} slots guarded by full/empty mbarriers; tcgen05.commit releases a
//     slot when the MMAs that read it retire and signals the epilogue after the last k-block;
//   * both operands may be K-major (row . row) or MN-major (transposed in memory): MN-major tiles
//     are staged as [k][32 floats] boxes and described to the tensor core with a_major/b_major = 1,
//     so weight gradients (dY^T . X) and P^T/V^T products need no transposition pass;
//   * smem per CTA is kept under half an SM so two CTAs are co-resident and one tile's epilogue
//     overlaps the other's main loop.
//
// Descriptor encodings follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor"
// tables (cross-checked against cute/arch/mma_sm100_desc.hpp).
#include "gemm_tc.cuh"

namespace corrif {
namespace tc {

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 2)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const KernelArgs args) {
  constexpr int B_BYTES = BN * ROW_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;   // BN in {64,128,256}: already a power of two

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_holder;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B: 1024-B aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- which problem / tile / k-range -------------------------------------------------------
  const int z = blockIdx.z;
  const int split = z % args.split_k, batch = z / args.split_k;
  const int bi = batch % args.batch_inner, bo = batch / args.batch_inner;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int total_kb = (args.K + BK - 1) / BK;
  const int kb_per_split = (total_kb + args.split_k - 1) / args.split_k;
  const int kb_begin = split * kb_per_split;
  const int kb_end = min(total_kb, kb_begin + kb_per_split);
  const int num_kb = kb_end - kb_begin;           // may be <= 0 for a trailing split: nothing to add

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_base_holder)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_holder;

  if (num_kb > 0) {
    if (warp == 0 && lane == 0) {
      // ================= TMA producer =================
      const int64_t a_off = bo * args.a_bo + bi * args.a_bi;
      const int64_t b_off = bo * args.b_bo + bi * args.b_bi;
      const int a_c0 = (int)(a_off % args.lda), a_c1 = (int)(a_off / args.lda);
      const int b_c0 = (int)(b_off % args.ldb), b_c1 = (int)(b_off / args.ldb);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        const int k0 = (kb_begin + i) * BK;
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
        if (!A_MN) {
          tma_load_2d(sa, &tmA, &full_bar[s], a_c0 + k0, a_c1 + m0);          // box {32 k, 128 rows}
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j)                                    // box {32 rows, 32 k}
            tma_load_2d(sa + j * (BK * ROW_BYTES), &tmA, &full_bar[s], a_c0 + m0 + 32 * j, a_c1 + k0);
        }
        if (!B_MN) {
          tma_load_2d(sb, &tmB, &full_bar[s], b_c0 + k0, b_c1 + n0);          // box {32 k, BN rows}
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j)
            tma_load_2d(sb + j * (BK * ROW_BYTES), &tmB, &full_bar[s], b_c0 + n0 + 32 * j, b_c1 + k0);
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = make_idesc<BN, A_MN, B_MN>();
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: step 32 B inside the 128-B swizzle row; MN-major: step one 8-k group (1024 B)
          const uint64_t ad = make_smem_desc<A_MN>(sa + k * (A_MN ? 1024 : UMMA_K * 4));
          const uint64_t bd = make_smem_desc<B_MN>(sb + k * (B_MN ? 1024 : UMMA_K * 4));
          tcgen05_mma_tf32(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        tcgen05_commit(&empty_bar[s]);          // frees the slot when these MMAs have read it
      }
      tcgen05_commit(&tmem_full_bar);           // accumulator complete
    } else if (warp >= 2) {
      // ================= epilogue: TMEM -> registers -> global =================
      mbar_wait(&tmem_full_bar, 0);
      tcgen05_fence_after();
      EpiArgs e = args.epi;
      epi_setup_dropout(e, args.drop, bo);
      if (e.bias) e.bias += bo * args.bias_bo;
      const int64_t doff = bo * args.d_bo + bi * args.d_bi;
      e.D += doff;
      if (e.residual) e.residual += doff;
      if (e.aux) e.aux += doff;
      const int quad = warp & 3;                // TMEM lanes [32*quad, 32*quad+32)
      const int m = m0 + quad * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), r);
        if (m < e.M) {
          float4 pre[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int n = n0 + c * 32 + j * 4;
            pre[j] = n < e.N ? epilogue_prefetch4(e, m, n) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int n = n0 + c * 32 + j * 4;
            if (n < e.N)
              epilogue_apply4(e, m, n, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), pre[j]);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ================================================================================================
// v2: persistent kernel.  One CTA per SM walks the tile list (m fastest, so CTAs running together
// share the B tile in L2).  The accumulator is double-buffered in TMEM (2 x BN columns): while the
// epilogue warps drain tile t, the producer/MMA warps are already running the main loop of tile t+1.
// The smem ring runs continuously across tiles.  The epilogue transposes each 32x32 accumulator
// chunk through shared memory so that global accesses are 128-byte row segments (8 lanes x float4
// per row, 4 rows per instruction) instead of one row per thread.
// ================================================================================================
constexpr int EPI_PAD = 36;   // floats per staged row (32 + 4: float4-aligned, bank-conflict free)
constexpr int P_THREADS = 320;  // producer warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant)

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_tf32_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const KernelArgs args, const int mt, const int nt, const int total_tiles) {
  constexpr int B_BYTES = BN * ROW_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_holder;
  __shared__ __align__(16) float stage_buf[8][32][EPI_PAD];

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = (args.K + BK - 1) / BK;
  const int kb_per_split = (total_kb + args.split_k - 1) / args.split_k;
  const int tiles_per_z = mt * nt;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_holder, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_holder;

  // tile -> (z, m, n) and its k-block range
  auto decode = [&](int tile, int& z, int& m0, int& n0, int& kb_begin, int& num_kb) {
    z = tile / tiles_per_z;
    const int rem = tile - z * tiles_per_z;
    n0 = (rem % nt) * BN;            // n fastest: CTAs running together share the A rows (read from
    m0 = (rem / nt) * BM;            // DRAM once); B (weights, <= 3 MB) stays L2-resident anyway
    const int split = z % args.split_k;
    kb_begin = split * kb_per_split;
    num_kb = min(total_kb, kb_begin + kb_per_split) - kb_begin;
  };

  if (warp == 0 && lane == 0) {
    // ================= TMA producer =================
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int z, m0, n0, kb_begin, num_kb;
      decode(tile, z, m0, n0, kb_begin, num_kb);
      const int batch = z / args.split_k;
      const int bi = batch % args.batch_inner, bo = batch / args.batch_inner;
      const int64_t a_off = bo * args.a_bo + bi * args.a_bi;
      const int64_t b_off = bo * args.b_bo + bi * args.b_bi;
      const int a_c0 = (int)(a_off % args.lda), a_c1 = (int)(a_off / args.lda);
      const int b_c0 = (int)(b_off % args.ldb), b_c1 = (int)(b_off / args.ldb);
      for (int i = 0; i < num_kb; ++i, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        const int k0 = (kb_begin + i) * BK;
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
        if (!A_MN) {
          tma_load_2d(sa, &tmA, &full_bar[s], a_c0 + k0, a_c1 + m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j)
            tma_load_2d(sa + j * (BK * ROW_BYTES), &tmA, &full_bar[s], a_c0 + m0 + 32 * j, a_c1 + k0);
        }
        if (!B_MN) {
          tma_load_2d(sb, &tmB, &full_bar[s], b_c0 + k0, b_c1 + n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j)
            tma_load_2d(sb + j * (BK * ROW_BYTES), &tmB, &full_bar[s], b_c0 + n0 + 32 * j, b_c1 + k0);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc<BN, A_MN, B_MN>();
    uint32_t it = 0, tc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int z, m0, n0, kb_begin, num_kb;
      decode(tile, z, m0, n0, kb_begin, num_kb);
      if (num_kb <= 0) continue;
      const uint32_t buf = tc & 1u;
      mbar_wait(&tempty_bar[buf], ((tc >> 1) & 1u) ^ 1u);      // epilogue has drained this buffer
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + buf * BN;
      for (int i = 0; i < num_kb; ++i, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t ad = make_smem_desc<A_MN>(sa + k * (A_MN ? 1024 : UMMA_K * 4));
          const uint64_t bd = make_smem_desc<B_MN>(sb + k * (B_MN ? 1024 : UMMA_K * 4));
          tcgen05_mma_tf32(tacc, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        tcgen05_commit(&empty_bar[s]);
      }
      tcgen05_commit(&tfull_bar[buf]);
      ++tc;
    }
  } else if (warp >= 2) {
    // ================= epilogue =================
    const int quad = warp & 3;                                 // TMEM lane quadrant of this warp
    const int half = (warp - 2) >> 2;                          // which half of the column chunks
    float (*stg)[EPI_PAD] = stage_buf[warp - 2];
    const int r_sub = lane >> 3, c4 = (lane & 7) * 4;          // coalesced mapping: 8 lanes per row
    uint32_t tc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int z, m0, n0, kb_begin, num_kb;
      decode(tile, z, m0, n0, kb_begin, num_kb);
      if (num_kb <= 0) continue;
      const int batch = z / args.split_k;
      const int bi = batch % args.batch_inner, bo = batch / args.batch_inner;
      EpiArgs e = args.epi;
      epi_setup_dropout(e, args.drop, bo);
      if (e.bias) e.bias += bo * args.bias_bo;
      const int64_t doff = bo * args.d_bo + bi * args.d_bi;
      e.D += doff;
      if (e.residual) e.residual += doff;
      if (e.aux) e.aux += doff;
      const uint32_t buf = tc & 1u;
      // Software-pipelined epilogue: the residual / saved pre-activation of a chunk is fetched one
      // chunk ahead (the first one even before the accumulator is ready), so DRAM latency hides
      // behind the main loop and the previous chunk's math instead of stalling every chunk.
      constexpr int NCH = BN / 64;                             // chunks handled by this warp
      const int mrow = m0 + quad * 32 + r_sub;
      auto prefetch = [&](int c, float4 (&pre)[8]) {
        const int n = n0 + c * 32 + c4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = mrow + 4 * i;
          pre[i] = (m < e.M && n < e.N) ? epilogue_prefetch4(e, m, n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      float4 pre[8], pre_next[8];
      prefetch(half, pre);
      mbar_wait(&tfull_bar[buf], (tc >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = half + 2 * ci;
        uint32_t r[32];
        tmem_ld32(tacc + (uint32_t)(c * 32), r);
        if (ci + 1 < NCH) prefetch(c + 2, pre_next);
        if (n0 + c * 32 < e.N) {                      // warp-uniform: skip fully out-of-range chunks
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(&stg[lane][4 * j]) =
                make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          __syncwarp();
          const int n = n0 + c * 32 + c4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = r_sub + 4 * i;
            const int m = m0 + quad * 32 + rr;
            if (m < e.M && n < e.N)
              epilogue_apply4(e, m, n, *reinterpret_cast<const float4*>(&stg[rr][c4]), pre[i]);
          }
          __syncwarp();
        }
        if (ci + 1 < NCH) {
#pragma unroll
          for (int i = 0; i < 8; ++i) pre[i] = pre_next[i];
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tempty_bar[buf]);
      ++tc;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
static int launch_persistent_variant(const corrif_gemm_desc& g, const CUtensorMap& ta, const CUtensorMap& tb,
                                     cudaStream_t stream) {
  constexpr int smem = STAGES * (A_BYTES + BN * ROW_BYTES) + 1024;
  auto kern = gemm_tf32_persistent_kernel<BN, STAGES, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_last_error("gemm_tf32(v2): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  KernelArgs a;
  a.epi = make_epi_args(g);
  a.drop = make_drop_args(g);
  a.K = g.K; a.batch_inner = g.batch_inner; a.split_k = g.split_k;
  a.a_bo = g.a_bo; a.a_bi = g.a_bi; a.b_bo = g.b_bo; a.b_bi = g.b_bi; a.d_bo = g.d_bo; a.d_bi = g.d_bi;
  a.lda = g.lda; a.ldb = g.ldb; a.bias_bo = g.bias_bo;
  const int mt = (g.M + BM - 1) / BM, nt = (g.N + BN - 1) / BN;
  const int total = mt * nt * g.batch_outer * g.batch_inner * g.split_k;
  const int grid = total < num_sms() ? total : num_sms();
  kern<<<grid, P_THREADS, smem, stream>>>(ta, tb, a, mt, nt, total);
  return launch_status("gemm_tf32_persistent");
}

template <int BN, int STAGES>
static int launch_persistent_bn(const corrif_gemm_desc& g, const CUtensorMap& ta, const CUtensorMap& tb,
                                cudaStream_t stream) {
  if (!g.a_mn_major && !g.b_mn_major) return launch_persistent_variant<BN, STAGES, false, false>(g, ta, tb, stream);
  if (!g.a_mn_major && g.b_mn_major) return launch_persistent_variant<BN, STAGES, false, true>(g, ta, tb, stream);
  if (g.a_mn_major && !g.b_mn_major) return launch_persistent_variant<BN, STAGES, true, false>(g, ta, tb, stream);
  return launch_persistent_variant<BN, STAGES, true, true>(g, ta, tb, stream);
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
static int launch_variant(const corrif_gemm_desc& g, const CUtensorMap& ta, const CUtensorMap& tb,
                          cudaStream_t stream) {
  constexpr int smem = STAGES * (A_BYTES + BN * ROW_BYTES) + 1024;
  auto kern = gemm_tf32_kernel<BN, STAGES, A_MN, B_MN>;
  static bool configured = false;   // per template instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_last_error("gemm_tf32: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  KernelArgs a;
  a.epi = make_epi_args(g);
  a.drop = make_drop_args(g);
  a.K = g.K; a.batch_inner = g.batch_inner; a.split_k = g.split_k;
  a.a_bo = g.a_bo; a.a_bi = g.a_bi; a.b_bo = g.b_bo; a.b_bi = g.b_bi; a.d_bo = g.d_bo; a.d_bi = g.d_bi;
  a.lda = g.lda; a.ldb = g.ldb; a.bias_bo = g.bias_bo;
  dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN, g.batch_outer * g.batch_inner * g.split_k);
  kern<<<grid, NUM_THREADS, smem, stream>>>(ta, tb, a);
  return launch_status("gemm_tf32");
}

template <int BN, int STAGES>
static int launch_bn(const corrif_gemm_desc& g, const CUtensorMap& ta, const CUtensorMap& tb,
                     cudaStream_t stream) {
  if (!g.a_mn_major && !g.b_mn_major) return launch_variant<BN, STAGES, false, false>(g, ta, tb, stream);
  if (!g.a_mn_major && g.b_mn_major) return launch_variant<BN, STAGES, false, true>(g, ta, tb, stream);
  if (g.a_mn_major && !g.b_mn_major) return launch_variant<BN, STAGES, true, false>(g, ta, tb, stream);
  return launch_variant<BN, STAGES, true, true>(g, ta, tb, stream);
}

}  // namespace tc

// Tile width: the widest BN (more MMA work per byte staged: a 128 x BN tile moves (128+BN)*128 B per
// 128*BN*32 MACs) that still leaves about one tile per SM.
// Tile width (measured with tools/gemm_bench.py, profiles/r01_gemm_shapes.txt): 128 x 256 tiles move
// 25 % fewer operand bytes per MAC than 128 x 128, but only pay off when there are at least two full
// waves of them; below that 128 x 128 keeps more SMs busy.
static int choose_bn(const corrif_gemm_desc& g) {
  static const char* force = getenv("CORRIF_GEMM_BN");          // tuning override
  if (force) { const int v = atoi(force); if (v == 64 || v == 128 || v == 256) return g.N <= 64 ? 64 : v; }
  if (g.N <= 64) return 64;
  const int64_t z = (int64_t)g.batch_outer * g.batch_inner * g.split_k;
  const int64_t mt = (g.M + 127) / 128;
  if (g.N >= 256 && mt * ((g.N + 255) / 256) * z >= 2 * num_sms()) return 256;
  return 128;
}

int gemm_tf32_launch(const corrif_gemm_desc& g, cudaStream_t stream) {
  using namespace tc;
  // Split-K weight gradients are many short, latency-bound CTAs: two co-resident non-persistent CTAs
  // per SM (v1) beat one persistent CTA there; everything else runs the persistent kernel (v2).
  static const bool force_v1 = getenv("CORRIF_GEMM_V1") != nullptr;    // bring-up A/B switches
  static const bool no_pair = getenv("CORRIF_GEMM_NOPAIR") != nullptr;
  // CTA-pair 256 x 256 tiles (v3, gemm_tf32_pair.cu) whenever the problem is wide enough to fill them:
  // they need a third less operand traffic per MAC than any single-CTA tile, which is what bounds
  // TF32 GEMMs here.
  if (!force_v1 && !no_pair && gemm_tf32_pair_supported(g)) return gemm_tf32_pair_launch(g, stream);
  const bool use_v1 = force_v1 || g.split_k > 1;
  const int BN = use_v1 ? (g.N <= 64 ? 64 : 128) : choose_bn(g);
  const bool batched = g.batch_outer * g.batch_inner > 1;
  // Extent of the memory an operand's tensor map must cover.  Un-batched: exact logical extent, so
  // TMA zero-fills ragged M/N/K edges.  Batched: the whole buffer reachable through the offsets;
  // the contraction dim then has no zero fill, hence K % 32 == 0 is required (checked by caller).
  auto span = [&](int64_t bo_stride, int64_t bi_stride) {
    return (int64_t)(g.batch_outer - 1) * bo_stride + (int64_t)(g.batch_inner - 1) * bi_stride;
  };
  CUtensorMap ta, tb;
  int st;
  {
    const int64_t extra = span(g.a_bo, g.a_bi);
    const int64_t rows = g.a_mn_major ? g.K : g.M, cols = g.a_mn_major ? g.M : g.K;
    const uint64_t dim0 = batched ? (uint64_t)g.lda : (uint64_t)cols;
    const uint64_t dim1 = (uint64_t)(rows + (batched ? (extra + g.lda - 1) / g.lda : 0));
    st = encode_map(&ta, g.A, dim0, dim1, g.lda, 32, g.a_mn_major ? BK : BM, g.a_mn_major != 0);
    if (st) return st;
  }
  {
    const int64_t extra = span(g.b_bo, g.b_bi);
    const int64_t rows = g.b_mn_major ? g.K : g.N, cols = g.b_mn_major ? g.N : g.K;
    const uint64_t dim0 = batched ? (uint64_t)g.ldb : (uint64_t)cols;
    const uint64_t dim1 = (uint64_t)(rows + (batched ? (extra + g.ldb - 1) / g.ldb : 0));
    st = encode_map(&tb, g.B, dim0, dim1, g.ldb, 32, g.b_mn_major ? BK : BN, g.b_mn_major != 0);
    if (st) return st;
  }
  if (use_v1) {
    if (BN == 64) return launch_bn<64, 4>(g, ta, tb, stream);
    return launch_bn<128, 3>(g, ta, tb, stream);
  }
  if (BN == 64) return launch_persistent_bn<64, 6>(g, ta, tb, stream);
  if (BN == 128) return launch_persistent_bn<128, 5>(g, ta, tb, stream);
  return launch_persistent_bn<256, 3>(g, ta, tb, stream);
}

}  // namespace corrif

using namespace corrif;

extern "C" int corrif_gemm(const corrif_gemm_desc* d, void* stream) {
  CORRIF_REQUIRE(d != nullptr, "gemm: null descriptor");
  const corrif_gemm_desc& g = *d;
  CORRIF_REQUIRE(g.A && g.B && g.D, "gemm: null operand");
  CORRIF_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: M,N,K must be positive (%d,%d,%d)", g.M, g.N, g.K);
  CORRIF_REQUIRE(g.batch_outer >= 1 && g.batch_inner >= 1 && g.split_k >= 1, "gemm: batch/split >= 1");
  CORRIF_REQUIRE((int64_t)g.batch_outer * g.batch_inner * g.split_k <= 65535, "gemm: grid.z too large");
  CORRIF_REQUIRE(g.N % 4 == 0 && g.ldd % 4 == 0 && g.lda % 4 == 0 && g.ldb % 4 == 0,
                 "gemm: N and leading dims must be multiples of 4");
  CORRIF_REQUIRE(((uintptr_t)g.A % 16 == 0) && ((uintptr_t)g.B % 16 == 0) && ((uintptr_t)g.D % 16 == 0),
                 "gemm: operands must be 16-byte aligned");
  CORRIF_REQUIRE(g.a_bo % 4 == 0 && g.a_bi % 4 == 0 && g.b_bo % 4 == 0 && g.b_bi % 4 == 0 &&
                 g.d_bo % 4 == 0 && g.d_bi % 4 == 0, "gemm: batch offsets must be multiples of 4");
  CORRIF_REQUIRE(g.epilogue >= CORRIF_EPI_STORE && g.epilogue <= CORRIF_EPI_ATOMIC_ADD, "gemm: epilogue");
  CORRIF_REQUIRE(g.split_k == 1 || g.epilogue == CORRIF_EPI_ATOMIC_ADD,
                 "gemm: split_k > 1 requires CORRIF_EPI_ATOMIC_ADD");
  if (g.epilogue == CORRIF_EPI_BIAS || g.epilogue == CORRIF_EPI_BIAS_GELU ||
      g.epilogue == CORRIF_EPI_BIAS_RESIDUAL)
    CORRIF_REQUIRE(g.bias != nullptr && (uintptr_t)g.bias % 16 == 0, "gemm: bias missing/unaligned");
  if (g.epilogue == CORRIF_EPI_BIAS_RESIDUAL)
    CORRIF_REQUIRE(g.residual != nullptr && g.ldr % 4 == 0 && (uintptr_t)g.residual % 16 == 0,
                   "gemm: residual missing/unaligned");
  if (g.drop_p != 0.f) {
    CORRIF_REQUIRE(g.drop_p > 0.f && g.drop_p < 1.f, "gemm: drop_p");
    CORRIF_REQUIRE(g.epilogue == CORRIF_EPI_BIAS_RESIDUAL || g.epilogue == CORRIF_EPI_BIAS_GELU ||
                   g.epilogue == CORRIF_EPI_MUL_DGELU, "gemm: fused dropout needs BIAS_RESIDUAL, BIAS_GELU or MUL_DGELU");
    CORRIF_REQUIRE(g.ldd == g.N, "gemm: fused dropout needs ldd == N");
  }
  if (g.epilogue == CORRIF_EPI_BIAS_GELU || g.epilogue == CORRIF_EPI_MUL_DGELU)
    CORRIF_REQUIRE(g.aux != nullptr && g.ldaux % 4 == 0 && (uintptr_t)g.aux % 16 == 0,
                   "gemm: aux missing/unaligned");
  if (g.precision == CORRIF_GEMM_FP32) return gemm_fp32_launch(g, (cudaStream_t)stream);
  CORRIF_REQUIRE(g.precision == CORRIF_GEMM_TF32, "gemm: unknown precision %d", g.precision);
  if (g.batch_outer * g.batch_inner > 1)
    CORRIF_REQUIRE(g.K % 32 == 0, "gemm(tf32): batched problems need K %% 32 == 0 (K=%d)", g.K);
  if (g.split_k > 1)
    CORRIF_REQUIRE(g.K % 32 == 0, "gemm(tf32): split_k needs K %% 32 == 0 (K=%d)", g.K);
  return gemm_tf32_launch(g, (cudaStream_t)stream);
}
