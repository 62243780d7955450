// Channels-last 3x3x3 convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA): forward and data
// gradient of the decoder's / early-fusion blocks (mmvit4.py:29-45, 47-56, 222-292).
//
// "Line convolution".  A row of 128 voxels along x of one (sample, z, y) - a LINE - is the M dimension of the MMA.
// The three x-taps are moved into N:
//
//     P[x, (dx, co)] = sum over (dz, dy) sum over ci  in[z + dz, y + dy, x, ci] * W[co][ci][dz, dy, dx]
//     out[x, co]     = P[x - 1, (-1, co)] + P[x, (0, co)] + P[x + 1, (+1, co)]
//
// so one input line, staged ONCE in shared memory by TMA exactly as it lies in HBM ([x][channels], K-major with the
// hardware swizzle), is the A operand of every tap: no im2col operand and no shifted shared-memory view exists - the
// y / z shifts are a choice of TMEM accumulator, the x shift is two warp shuffles in the epilogue.  An input line of
// plane zv feeds, per y-tap, the accumulators of the output planes zv-1, zv, zv+1; they sit side by side in TMEM and
// the weight tile stacks the three z-taps in the same order, so ONE MMA with N = 3 * NPAD (NPAD = 3 * Cout rounded to
// 16) serves all three.  Every MMA accumulates; the epilogue zeroes an accumulator as soon as it has read it.
// Replicate padding (mmvit4.py:225-236) is a clamped line coordinate in y / z and "use your own value" at the two x
// borders; its adjoint (the data gradient of a replicate-padded convolution, for which the warp-level path needs a
// separate border kernel) is the same with the mirrored weight tap on the clamped axis.
//
// Lines narrower than 128 voxels (W = 64, 32, 16) stack the same (z, y) line of R = 128 / W SAMPLES in one tile:
// taps never cross samples, so a tap shift moves the whole tile.  A concatenation of sources runs as uniform 8 / 16 /
// 32-channel K chunks (one TMA box each); wide outputs run as chunks of 32 / 16 / 8 output channels whose weights are
// resident in shared memory, all at once or one chunk at a time.
//
// CTA = 11 warps, one CTA per SM: a TMA producer (ring of up to 12 line slots), two MMA-issuing warps (the ring slots
// are dealt between them; a single issuing thread was the limit), two groups of four epilogue warps (TMEM quadrant =
// warp % 4).  A work item is (output-channel chunk, z range, sample group, strip of T = 4 / 2 / 1 lines in y); the
// issuers walk the input planes of the range in order and keep four output planes x T lines of accumulators in TMEM
// (512 columns): three being accumulated, one being drained.  The epilogue adds the two x neighbours (shuffles; the
// warp boundaries through shared memory), bias, ReLU, stores channels-last and keeps the InstanceNorm statistics
// (sum, sum of squares per (sample, channel)) in registers until the item ends.
//
// Measured (batch 8, 128^3, profiles/r03_*): 32 -> 8 channels forward 0.67 ms (warp-level kernel: 2.21; HBM floor at
// the measured 6.55 TB/s: 0.42), 8 -> 32 data gradient 1.15 ms incl. the padding adjoint (1.97), 64^3 64 -> 16
// forward 0.23 ms (0.76).
//
// Operand precision: tcgen05.mma.kind::tf32 truncates fp32 operands to 10 mantissa bits.  Weights are rounded to
// nearest when packed; activations arrive by TMA untouched, so their truncation (relative bias -2^-11 E[1/mantissa]
// = -3.52e-4, the same constant the attention kernels use for P) is compensated by scaling the packed weights by
// 1 + 3.52e-4: zero-mean error with the RMS of round-to-nearest (measured 3.0e-4 relative L2 against fp64).
#include <stdlib.h>
#include "tc05.cuh"

namespace corrif {
namespace convtc {
using namespace tc05;

constexpr int MAXKC = 12;         // A sub-tiles (<= 32 channels of one source) per line
#ifndef CORRIF_TC_NISSUE
#define CORRIF_TC_NISSUE 2
#endif
constexpr int NISSUE = CORRIF_TC_NISSUE;         // MMA-issuing warps (lines are dealt round-robin)
constexpr int NTHREADS = 320 + 32 * (NISSUE - 1);   // warp 0: TMA producer, warp 1 and warps 10..: MMA issuers, warps 2..9: two epilogue groups of four
constexpr int ACC_SLOTS = 4;      // output planes in flight in TMEM
constexpr int MAX_RING = 12;
constexpr float TRUNC_COMP = 1.0f + 3.52e-4f;

struct KChunk {
  int src;      // source tensor (tensor map index)
  int c0;       // first channel inside that source
  int cs;       // first channel in concatenation order
  int swb;      // bytes per row = 4 * channels of the chunk = swizzle span (32, 64 or 128)
  int a_off;    // byte offset of the [128 x swb] tile inside a ring slot
  int w_off;    // byte offset of the [3 * NPAD x swb] tile (z-taps +1, 0, -1 stacked) inside one y-tap's weight block
};

struct Plan {
  int ok;
  int R, T, ZL, CC, NPAD, NKC, NS, SWB, w_resident;
  int n_nchunks, n_zchunks, n_bgroups, n_strips;
  long long total_items;
  int slot_bytes, tap_bytes, w_bytes;
  int smem_bytes;
  KChunk kc[MAXKC];
};

struct Args {
  int B, D, H, W, R, T, ZL, NKC, NS;
  int n_nchunks, n_zchunks, n_bgroups, n_strips, total_items;
  int pad_mode, relu, Cout, w_resident;
  int debug;    // CORRIF_CONV_TC_DEBUG bit 0: epilogue only hands the accumulators back, 1: no MMAs, 2: no TMA loads (timing experiments)
  int slot_bytes, tap_bytes, w_bytes;
  KChunk kc[MAXKC];
  const float* wpk;
  const float* bias;
  float* out;
  long long ldo;
  double* stats;
};

// plan for one output-channel chunk width CC (32, 16 or 8)
static Plan make_plan_cc(const corrif_conv3d_desc& d, int nsm, int CC) {
  Plan p{};
  p.ok = 0;
  p.R = 128 / d.W;
  p.CC = CC;
  p.n_nchunks = d.Cout / p.CC;
  p.NPAD = (3 * p.CC + 15) / 16 * 16;                 // 32, 48, 96
  p.T = 512 / (ACC_SLOTS * p.NPAD);                   // 4, 2, 1
  if (p.T > 4) p.T = 4;
  if (p.T > d.H) p.T = d.H;
  int nk = 0, cs = 0, a_off = 0, w_off = 0;
  // K chunks: every chunk has the same width (the kernel is templated on it) - the widest of 32 / 16 / 8 channels
  // that divides every source, so cat(24, 8) runs as four 8-channel chunks and cat(48, 16) as four 16-channel ones
  int cw = 32;
  for (int s = 0; s < d.nsrc; ++s)
    while (cw >= 8 && d.src[s].C % cw) cw >>= 1;
  if (cw < 8) return p;
  p.SWB = cw * 4;
  for (int s = 0; s < d.nsrc; ++s) {
    for (int c0 = 0; c0 < d.src[s].C; c0 += cw) {
      if (nk == MAXKC) return p;
      p.kc[nk] = KChunk{s, c0, cs, cw * 4, a_off, w_off};
      a_off += 128 * cw * 4;
      w_off += (3 * p.NPAD * cw * 4 + 1023) / 1024 * 1024;
      cs += cw;
      ++nk;
    }
  }
  p.NKC = nk;
  p.slot_bytes = a_off;
  p.tap_bytes = w_off;
  p.w_bytes = 3 * w_off;
  const int budget = 232448 - 4096 - 1024;            // 227 KB per CTA minus static shared memory and the alignment slack
  // the weights of every output-channel chunk stay in shared memory if they fit beside a ring of >= 4 line slots;
  // otherwise one chunk at a time (reloaded when a CTA's items move on to the next chunk)
  long long wtot = (long long)p.n_nchunks * p.w_bytes;
  p.w_resident = 1;
  if (wtot + 4ll * p.slot_bytes > budget) { wtot = p.w_bytes; p.w_resident = 0; }
  if (wtot + 3ll * p.slot_bytes > budget) return p;
  long long ns = (budget - wtot) / p.slot_bytes;
  p.NS = (int)(ns > MAX_RING ? MAX_RING : ns);
  p.smem_bytes = 1024 + (int)wtot + p.NS * p.slot_bytes;
  p.n_bgroups = d.B / p.R;
  p.n_strips = (d.H + p.T - 1) / p.T;
  // z range per item: every range re-reads two halo planes, and the items should fill whole rounds of the SMs
  long long best_cost = -1;
  for (int zl = d.D < 4 ? d.D : 4; zl <= d.D; ++zl) {
    const long long nz = (d.D + zl - 1) / zl;
    const long long items = (long long)p.n_nchunks * nz * p.n_bgroups * p.n_strips;
    const long long rounds = (items + nsm - 1) / nsm;
    const long long cost = rounds * (zl + 2);
    if (best_cost < 0 || cost <= best_cost) { best_cost = cost; p.ZL = zl; p.n_zchunks = (int)nz; p.total_items = items; }
  }
  if (p.total_items >= (1ll << 31)) return p;
  p.ok = 1;
  return p;
}

// The widest output-channel chunk (32, 16, 8) that divides Cout and whose weights fit in shared memory: wide chunks
// read the input fewer times, narrow ones need less shared memory per chunk (64 -> 320 channels at 16^3 runs as 20
// chunks of 16).
static Plan make_plan(const corrif_conv3d_desc& d, int nsm) {
  Plan none{};
  none.ok = 0;
  if (d.ksize != 3) return none;
  if (!(d.W == 16 || d.W == 32 || d.W == 64 || d.W == 128)) return none;
  if (d.B % (128 / d.W)) return none;
  if (d.Cout > 512 || d.Cout % 8) return none;
  for (int cc = 32; cc >= 8; cc >>= 1) {
    if (d.Cout % cc) continue;
    const Plan p = make_plan_cc(d, nsm, cc);
    if (p.ok) return p;
  }
  return none;
}

// ---- device helpers --------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :: "r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tmem_ld_nw(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_nw(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_nw(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// zero N consecutive TMEM columns of this warp's 32 lanes (no wait)
template <int N>
__device__ __forceinline__ void tmem_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  if constexpr (N == 8) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" :: "r"(taddr), "r"(z) : "memory");
  } else if constexpr (N == 16) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                 :: "r"(taddr), "r"(z) : "memory");
  } else {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" :: "r"(taddr), "r"(z) : "memory");
  }
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory descriptor, rows of `swb` bytes (= the swizzle span), 8-row groups dense
__device__ __forceinline__ uint32_t desc_hi_swz(int swb) {
  const uint32_t type = swb == 128 ? 2u : (swb == 64 ? 4u : 6u);
  return (uint32_t)((8 * swb) >> 4) | (1u << 14) | (type << 29);
}

struct Item { int nc, zb, ze, bg, y0, ylast; };
__device__ __forceinline__ Item decode_item(const Args& a, int item) {
  Item it;
  const int strip = item % a.n_strips; item /= a.n_strips;
  it.bg = item % a.n_bgroups; item /= a.n_bgroups;
  const int zc = item % a.n_zchunks;
  it.nc = item / a.n_zchunks;
  it.y0 = strip * a.T;
  it.ylast = min(it.y0 + a.T, a.H) - 1;
  it.zb = zc * a.ZL;
  it.ze = min(it.zb + a.ZL, a.D);
  return it;
}

template <int CC, int SWB, bool STATS>
__global__ void __launch_bounds__(NTHREADS, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                 const __grid_constant__ CUtensorMap tm2, const Args a) {
  constexpr int NPAD = (3 * CC + 15) / 16 * 16;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_RING];
  __shared__ __align__(8) uint64_t empty_bar[MAX_RING];
  __shared__ __align__(8) uint64_t acc_full[ACC_SLOTS];
  __shared__ __align__(8) uint64_t acc_empty[ACC_SLOTS];
  __shared__ uint32_t tmem_base_holder;
  // [group][parity][quadrant][0: last row's dx=-1 part, 1: first row's dx=+1 part][channel]
  __shared__ __align__(16) float xchg[2][2][4][2][CC == 32 ? 16 : CC];
  __shared__ float s_bias[512];

  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_w = sbase;
  const uint32_t s_ring = sbase + (uint32_t)((a.w_resident ? a.n_nchunks : 1) * a.w_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = a.D, H = a.H, W = a.W, T = a.T, NS = a.NS;
  const int pad = a.pad_mode;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&tm0) : "memory");
    for (int s = 0; s < MAX_RING; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC_SLOTS; ++s) { mbar_init(&acc_full[s], NISSUE); mbar_init(&acc_empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_holder, 512);
  for (int i = threadIdx.x; i < a.Cout && i < 512; i += NTHREADS) s_bias[i] = a.bias ? __ldg(a.bias + i) : 0.f;
  // packed weights -> shared memory (the image is already swizzled): every chunk once, or chunk `nc` on demand
  auto stage_weights = [&](int first_chunk, int nchunks) {
    const float4* src = reinterpret_cast<const float4*>(a.wpk) + (size_t)first_chunk * (a.w_bytes / 16);
    const int n16 = nchunks * a.w_bytes / 16;
    constexpr int U = 8;
    for (int base = threadIdx.x; base < n16; base += NTHREADS * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * NTHREADS;
        v[u] = i < n16 ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * NTHREADS;
        if (i < n16)
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(s_w + i * 16), "f"(v[u].x), "f"(v[u].y),
                       "f"(v[u].z), "f"(v[u].w) : "memory");
      }
    }
    fence_proxy_async();
  };
  if (a.w_resident) stage_weights(0, a.n_nchunks);
  // Non-resident weights: when a CTA's next item belongs to another output-channel chunk, every warp meets at a
  // CTA barrier (the epilogue warps get there only after the last plane of the previous item, i.e. after every MMA
  // that reads the old chunk has completed), the chunk is replaced, and a second barrier releases the issuers.
  int cur_nc = -1;
  auto switch_chunk = [&](int nc) {
    if (a.w_resident || nc == cur_nc) return;
    __syncthreads();
    stage_weights(nc, 1);
    __syncthreads();
    cur_nc = nc;
  };
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  // Every MMA accumulates (one MMA spans the accumulators of up to three output planes, which are at different
  // stages of their sums, so there is no per-MMA "first contribution" flag): the accumulators start at zero and the
  // epilogue zeroes a tile again as soon as it has read it.
  if (warp >= 2 && warp < 6) {
    const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
    for (int c = 0; c < 512; c += 32) tmem_zero<32>(t0 + c);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (warp == 0) {
    // ================= TMA producer: one ring slot per (virtual plane, virtual line) =================
    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
      const Item it = decode_item(a, item);
      switch_chunk(it.nc);
      for (int zv = it.zb - 1; zv <= it.ze; ++zv) {
        const bool zoob = zv < 0 || zv >= D;
        if (zoob && pad == CORRIF_PAD_ZEROS) continue;
        const int za = min(max(zv, 0), D - 1);
        for (int yv = it.y0 - 1; yv <= it.ylast + 1; ++yv) {
          const bool yoob = yv < 0 || yv >= H;
          if (yoob && pad == CORRIF_PAD_ZEROS) continue;
          const int ya = min(max(yv, 0), H - 1);
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[s], (a.debug & 4) ? 0u : (uint32_t)a.slot_bytes);
            const uint32_t dst0 = s_ring + (uint32_t)(s * a.slot_bytes);
            for (int k = 0; k < ((a.debug & 4) ? 0 : a.NKC); ++k) {
              const KChunk kc = a.kc[k];
              const CUtensorMap* tm = kc.src == 0 ? &tm0 : (kc.src == 1 ? &tm1 : &tm2);
              for (int r = 0; r < a.R; ++r)
                tma_load_4d(dst0 + (uint32_t)(kc.a_off + r * W * SWB), tm, &full_bar[s], kc.c0, 0, ya,
                            (it.bg * a.R + r) * D + za);
            }
          }
          __syncwarp();
          if (++s == NS) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1 || warp >= 10) {
    // ================= MMA issuers =================
    // Every MMA accumulates, so MMAs of different lines need no order among themselves: the ring slots are dealt
    // round-robin to NISSUE warps (the single issuing thread, not the tensor pipe, was the limit: ~160 mostly
    // uniform-datapath instructions per line at ~8 cycles each).  Each issuer commits its own MMAs to the plane's
    // acc_full barrier (count NISSUE).
    const int iss = warp == 1 ? 0 : warp - 9;
    // One thread issues every MMA.  An input line feeds, per y-tap, the accumulators of the (up to) three output
    // planes zv-1, zv, zv+1: they sit side by side in TMEM (column = (line * 4 + plane slot) * NPAD) and the weight
    // tile stacks the z-taps +1, 0, -1 in the same order, so ONE MMA with N = 3 * NPAD serves all three (two MMAs
    // when the slot ring wraps or 3 * NPAD > 256).  Measured with N = NPAD MMAs, one per (dz, dy): the tensor pipe
    // needed ~44 cycles per MMA (the 4 KB A operand is re-read from shared memory for every tap) and the issuing
    // thread ~700 cycles of bookkeeping per line; both fall by ~3x with the merged form.
    constexpr int NKS = SWB / 32;                                            // k-steps (8 channels) per chunk
    constexpr uint32_t HI = (uint32_t)((8 * SWB) >> 4) | (1u << 14) | ((SWB == 128 ? 2u : (SWB == 64 ? 4u : 6u)) << 29);
    constexpr uint32_t ATILE16 = 128 * SWB / 16;                             // strides in 16-byte units
    constexpr uint32_t BLK16 = NPAD * SWB / 16;
    constexpr uint32_t WTILE16 = ((3 * NPAD * SWB + 1023) / 1024 * 1024) / 16;
    constexpr int MAXRUN = 3 * NPAD <= 256 ? 3 : 2;
    const bool adj = pad == CORRIF_PAD_REPLICATE_ADJOINT, zeros = pad == CORRIF_PAD_ZEROS;
    const uint32_t wy16 = (uint32_t)a.tap_bytes >> 4;
    const uint32_t slot16 = (uint32_t)a.slot_bytes >> 4;
    const uint32_t ring16 = desc_lo_kmajor(s_ring);
    const int NKC = a.NKC;
    int s = 0;
    uint32_t ph = 0, pc_base = 0;
    for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
      const Item it = decode_item(a, item);
      switch_chunk(it.nc);
      const uint32_t w_item16 = desc_lo_kmajor(s_w + (uint32_t)((a.w_resident ? it.nc : 0) * a.w_bytes));
      for (int zv = it.zb - 1; zv <= it.ze; ++zv) {
        const int zo = zv + 1;                             // the output plane this input plane touches first
        if (zo < it.ze) {
          const uint32_t pc = pc_base + (uint32_t)(zo - it.zb);
          mbar_wait(&acc_empty[pc & 3u], ((pc >> 2) & 1u) ^ 1u);   // drained and zeroed by the epilogue
          tcgen05_fence_after();
        }
        const bool zoob = zv < 0 || zv >= D;
        if (!(zoob && zeros)) {
          // output planes fed by this input plane -> one or two runs of adjacent accumulator slots
          const int zlo = max(zv - 1, it.zb), zhi = min(zv + 1, it.ze - 1);
          const int len = zhi - zlo + 1;
          const int slot0 = (int)((pc_base + (uint32_t)(zlo - it.zb)) & 3u);
          // weight block (z-tap) of plane zlo: block b holds dz = 1 - b; a clamped plane of the adjoint mirrors it
          const int blk0 = (zoob && adj) ? 1 + (zv - zlo) : 1 - (zv - zlo);
          const int len1 = min(min(len, 4 - slot0), MAXRUN), len2 = len - len1;
          const uint32_t col1 = (uint32_t)slot0 * NPAD, col2 = (uint32_t)((slot0 + len1) & 3) * NPAD;
          const uint32_t brow1 = (uint32_t)blk0 * BLK16, brow2 = (uint32_t)(blk0 + len1) * BLK16;
          const uint32_t idesc1 = idesc_tf32(NPAD * len1, false, false), idesc2 = idesc_tf32(NPAD * (len2 > 0 ? len2 : 1), false, false);
          for (int yv = it.y0 - 1; yv <= it.ylast + 1; ++yv) {
            const bool yoob = yv < 0 || yv >= H;
            if (yoob && zeros) continue;
            // a ring slot always belongs to the same issuer: a waiter that skipped one use of a barrier would see
            // the phase parity of two uses ago and fall through before the line has landed
            const bool mine = (s % NISSUE) == iss;
            if (mine) {
            mbar_wait(&full_bar[s], ph);
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t a16 = ring16 + (uint32_t)s * slot16;
              const int dylo = max(-1, yv - it.ylast), dyhi = min(1, yv - it.y0);
              for (int dy = dylo; dy <= dyhi; ++dy) {
                const int wy = (yoob && adj) ? -dy : dy;
                const uint32_t d0 = tmem_base + (uint32_t)(yv - dy - it.y0) * (4u * NPAD);
                const uint32_t w16 = w_item16 + (uint32_t)(wy + 1) * wy16;
                for (int k = 0; k < NKC; ++k) {
#pragma unroll
                  for (int ks = 0; ks < NKS; ++ks)
                    tcgen05_mma_tf32(d0 + col1, desc_from(a16 + (uint32_t)k * ATILE16 + 2u * ks, HI),
                                     desc_from(w16 + brow1 + (uint32_t)k * WTILE16 + 2u * ks, HI), idesc1, 1u);
                }
                if (len2 > 0) {
                  for (int k = 0; k < NKC; ++k) {
#pragma unroll
                    for (int ks = 0; ks < NKS; ++ks)
                      tcgen05_mma_tf32(d0 + col2, desc_from(a16 + (uint32_t)k * ATILE16 + 2u * ks, HI),
                                       desc_from(w16 + brow2 + (uint32_t)k * WTILE16 + 2u * ks, HI), idesc2, 1u);
                  }
                }
              }
              tcgen05_commit(&empty_bar[s]);               // the slot is free once these MMAs have read it
            }
            __syncwarp();
            }
            if (++s == NS) { s = 0; ph ^= 1u; }
          }
        }
        const int zd = zv - 1;                             // complete: all three input planes have been applied
        if (zd >= it.zb) {
          if (elect_one()) tcgen05_commit(&acc_full[(pc_base + (uint32_t)(zd - it.zb)) & 3u]);
          __syncwarp();
        }
      }
      pc_base += (uint32_t)(it.ze - it.zb);
    }
  } else {
    // ================= epilogue: x-tap sum, bias, ReLU, store, InstanceNorm statistics =================
    // Two groups of four warps (TMEM quadrant = warp % 4).  Up to 16 output channels per tile the groups take
    // alternate tiles; a 32-channel tile is split between them by channel halves (one thread = one voxel keeps 3 x
    // its channels in registers, and a 96-column tile per thread was both spilling and 2x slower than the MMAs).
    constexpr int CCG = CC == 32 ? 16 : CC;
    constexpr bool SPLIT = CC == 32;
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int x = row & (W - 1), rsub = row / W;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int choff = SPLIT ? grp * CCG : 0;
    const bool xfirst = x == 0, xlast = x == W - 1;
    uint32_t pc_base = 0, tcnt = 0;
    for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
      const Item it = decode_item(a, item);
      switch_chunk(it.nc);
      const int b = it.bg * a.R + rsub;
      const int n0 = it.nc * CC + choff;
      const int nm = it.ylast - it.y0 + 1;
      float bias[CCG];
#pragma unroll
      for (int j = 0; j < CCG; ++j) bias[j] = s_bias[n0 + j];
      // the statistics live in registers for a whole item
      float ssum[STATS ? CCG : 1], ssq[STATS ? CCG : 1];
#pragma unroll
      for (int j = 0; j < (STATS ? CCG : 1); ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
      for (int z = it.zb; z < it.ze; ++z) {
        const uint32_t pc = pc_base + (uint32_t)(z - it.zb), slot = pc & 3u;
        mbar_wait(&acc_full[slot], (pc >> 2) & 1u);
        tcgen05_fence_after();
        const int m0 = (a.debug & 1) ? nm : (SPLIT ? 0 : (((z - it.zb) * nm + grp) & 1));
        for (int m = m0; m < nm; m += SPLIT ? 1 : 2, ++tcnt) {
          const uint32_t taddr = tmem_base + lane_addr + ((uint32_t)m * 4u + slot) * NPAD + (uint32_t)choff;
          uint32_t p0[CCG], p1[CCG], p2[CCG];
          tmem_ld_nw(taddr, p0);
          tmem_ld_nw(taddr + CC, p1);
          tmem_ld_nw(taddr + 2 * CC, p2);
          tmem_ld_wait();
          tmem_zero<CCG>(taddr);                            // hand the accumulator back empty
          tmem_zero<CCG>(taddr + CC);
          tmem_zero<CCG>(taddr + 2 * CC);
          float (*xb)[2][CCG] = xchg[grp][tcnt & 1u];
          if (lane == 31) {
#pragma unroll
            for (int j = 0; j < CCG; j += 4)
              *reinterpret_cast<uint4*>(&xb[quad][0][j]) = make_uint4(p0[j], p0[j + 1], p0[j + 2], p0[j + 3]);
          }
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < CCG; j += 4)
              *reinterpret_cast<uint4*>(&xb[quad][1][j]) = make_uint4(p2[j], p2[j + 1], p2[j + 2], p2[j + 3]);
          }
          if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
          // left / right x-neighbours: a shuffle for 30 of 32 lanes; the warp-boundary lanes read the other warp's
          // row from shared memory and the two lanes at the ends of a line apply the padding rule - both as
          // branches only the affected lanes take (per-element selects made this loop 31 instructions per value)
          float l[CCG], r[CCG];
#pragma unroll
          for (int j = 0; j < CCG; ++j) {
            l[j] = __shfl_up_sync(0xffffffffu, __uint_as_float(p0[j]), 1);
            r[j] = __shfl_down_sync(0xffffffffu, __uint_as_float(p2[j]), 1);
          }
          if (xfirst) {
#pragma unroll
            for (int j = 0; j < CCG; ++j)
              l[j] = pad == CORRIF_PAD_ZEROS ? 0.f : __uint_as_float(pad == CORRIF_PAD_REPLICATE ? p0[j] : p2[j]);
          } else if (lane == 0) {
#pragma unroll
            for (int j = 0; j < CCG; j += 4) {
              const float4 t = *reinterpret_cast<const float4*>(&xb[quad - 1][0][j]);
              l[j] = t.x; l[j + 1] = t.y; l[j + 2] = t.z; l[j + 3] = t.w;
            }
          }
          if (xlast) {
#pragma unroll
            for (int j = 0; j < CCG; ++j)
              r[j] = pad == CORRIF_PAD_ZEROS ? 0.f : __uint_as_float(pad == CORRIF_PAD_REPLICATE ? p2[j] : p0[j]);
          } else if (lane == 31) {
#pragma unroll
            for (int j = 0; j < CCG; j += 4) {
              const float4 t = *reinterpret_cast<const float4*>(&xb[quad + 1][1][j]);
              r[j] = t.x; r[j + 1] = t.y; r[j + 2] = t.z; r[j + 3] = t.w;
            }
          }
          float o[CCG];
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < CCG; ++j) o[j] = fmaxf((__uint_as_float(p1[j]) + bias[j]) + (l[j] + r[j]), 0.f);
          } else {
#pragma unroll
            for (int j = 0; j < CCG; ++j) o[j] = (__uint_as_float(p1[j]) + bias[j]) + (l[j] + r[j]);
          }
          if constexpr (STATS) {
#pragma unroll
            for (int j = 0; j < CCG; ++j) { ssum[j] += o[j]; ssq[j] = fmaf(o[j], o[j], ssq[j]); }
          }
          float* op = a.out + ((((long long)b * D + z) * H + (it.y0 + m)) * W + x) * a.ldo + n0;
#pragma unroll
          for (int j = 0; j < CCG; j += 4) if (!(a.debug & 8)) st4(op + j, make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]));
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);
      }
      pc_base += (uint32_t)(it.ze - it.zb);
      if constexpr (STATS) if (a.stats) {
        // lanes of one sample: the whole warp (W >= 32) or one half of it (W = 16)
        const int span = W < 32 ? W : 32;
#pragma unroll
        for (int j = 0; j < CCG; ++j) {
          float s1 = ssum[j], s2 = ssq[j];
          for (int o2 = span >> 1; o2 > 0; o2 >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o2);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o2);
          }
          if ((lane & (span - 1)) == 0) {
            double* sp = a.stats + ((long long)b * a.Cout + n0 + j) * 2;
            atomicAdd(sp, (double)s1);
            atomicAdd(sp + 1, (double)s2);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Weight operand image: [output-channel chunk][y-tap][K chunk][z-tap block: dz = +1, 0, -1][NPAD rows (dx, co)][swb bytes],
// each [3 * NPAD x swb] tile laid out K-major with the swizzle of its chunk and padded to 1 KB, i.e. exactly what the
// kernel copies to shared memory.  The z-taps are stacked in the order of the output planes they feed (zv-1, zv, zv+1).
__global__ void pack_tc_kernel(const float* __restrict__ w, float* __restrict__ wpk, int wCin, int flip, int CC,
                               int NPAD, int NKC, int tap_bytes, long long total, const Args a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int tapf = tap_bytes / 4;
  const int nc = (int)(i / (3ll * tapf));
  const int rem = (int)(i - (long long)nc * 3 * tapf);
  const int wyi = rem / tapf;
  const int o = (rem - wyi * tapf) * 4;
  float v = 0.f;
  for (int k = 0; k < NKC; ++k) {
    const KChunk kc = a.kc[k];
    const int ot = o - kc.w_off;
    if (ot < 0 || ot >= 3 * NPAD * kc.swb) continue;
    const int ol = ot ^ (((ot >> 7) & (kc.swb / 16 - 1)) << 4);      // physical -> logical offset (the XOR is an involution)
    const int n = ol / kc.swb, kk = (ol % kc.swb) / 4;
    const int blk = n / NPAD, nn = n - blk * NPAD;
    const int dxi = nn / CC, col = nn - dxi * CC;
    if (dxi < 3) {
      const int wz = 1 - blk;
      const int no = nc * CC + col, kch = kc.cs + kk, tap27 = ((wz + 1) * 3 + wyi) * 3 + dxi;
      if (!flip) v = w[((long long)no * wCin + kch) * 27 + tap27];
      else v = w[((long long)kch * wCin + no) * 27 + (26 - tap27)];      // dX = conv(dY, W^T mirrored)
    }
    break;
  }
  wpk[i] = round_tf32(v * TRUNC_COMP);
}

static int encode_line_map(CUtensorMap* map, const corrif_vol_src& s, int B, int D, int H, int W, int ch) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point not found"); return CORRIF_EDRIVER; }
  cuuint64_t dims[4] = {(cuuint64_t)s.C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * D};
  cuuint64_t strides[3] = {(cuuint64_t)s.ld * 4, (cuuint64_t)s.ld * 4 * W, (cuuint64_t)s.ld * 4 * W * H};
  cuuint32_t box[4] = {(cuuint32_t)ch, (cuuint32_t)W, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle swz = ch == 32 ? CU_TENSOR_MAP_SWIZZLE_128B
                                          : (ch == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)s.p, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("conv3d_tc: cuTensorMapEncodeTiled failed (%d): base %p C %d ld %lld volume %dx%dx%dx%d", (int)r,
                   (const void*)s.p, s.C, (long long)s.ld, B, D, H, W);
    return CORRIF_EDRIVER;
  }
  return 0;
}

static void fill_args(Args& a, const corrif_conv3d_desc& d, const Plan& p) {
  a.B = d.B; a.D = d.D; a.H = d.H; a.W = d.W; a.R = p.R; a.T = p.T; a.ZL = p.ZL; a.NKC = p.NKC; a.NS = p.NS;
  a.n_nchunks = p.n_nchunks; a.n_zchunks = p.n_zchunks; a.n_bgroups = p.n_bgroups; a.n_strips = p.n_strips;
  a.total_items = (int)p.total_items;
  a.pad_mode = d.pad_mode; a.relu = d.relu; a.Cout = d.Cout; a.w_resident = p.w_resident;
  static const int debug = getenv("CORRIF_CONV_TC_DEBUG") ? atoi(getenv("CORRIF_CONV_TC_DEBUG")) : 0;
  a.debug = debug;
  a.slot_bytes = p.slot_bytes; a.tap_bytes = p.tap_bytes; a.w_bytes = p.w_bytes;
  for (int k = 0; k < MAXKC; ++k) a.kc[k] = p.kc[k];
  a.wpk = d.wpk; a.bias = d.bias; a.out = d.out; a.ldo = d.ldo; a.stats = d.stats;
}

template <int CC, int SWB, bool STATS>
static int launch(const CUtensorMap* tm, const Args& a, const Plan& p, cudaStream_t stream) {
  auto kern = conv3d_tc_kernel<CC, SWB, STATS>;
  static int configured = 0;
  if (configured < p.smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes);
    if (e != cudaSuccess) { set_last_error("conv3d_tc: smem attribute (%d B): %s", p.smem_bytes, cudaGetErrorString(e)); return (int)e; }
    configured = p.smem_bytes;
  }
  const int nsm = num_sms();
  const unsigned grid = (unsigned)(p.total_items < nsm ? p.total_items : nsm);
  kern<<<grid, NTHREADS, p.smem_bytes, stream>>>(tm[0], tm[1], tm[2], a);
  return launch_status("conv3d_tc");
}

static int check_desc(const corrif_conv3d_desc& d, const char* what) {
  CORRIF_REQUIRE(d.nsrc >= 1 && d.nsrc <= 3, "%s: nsrc must be 1..3", what);
  CORRIF_REQUIRE(d.B > 0 && d.D > 0 && d.H > 0 && d.W > 0, "%s: empty volume", what);
  CORRIF_REQUIRE(d.pad_mode >= CORRIF_PAD_ZEROS && d.pad_mode <= CORRIF_PAD_REPLICATE_ADJOINT, "%s: pad_mode", what);
  int csum = 0;
  for (int i = 0; i < d.nsrc; ++i) {
    CORRIF_REQUIRE(d.src[i].C > 0 && d.src[i].C % 4 == 0 && d.src[i].ld >= d.src[i].C && d.src[i].ld % 4 == 0,
                   "%s: source %d: C (%d) and ld (%lld) must be multiples of 4, ld >= C", what, i, d.src[i].C,
                   (long long)d.src[i].ld);
    csum += d.src[i].C;
  }
  CORRIF_REQUIRE(csum == d.Cin, "%s: Cin (%d) != sum of source channels (%d)", what, d.Cin, csum);
  CORRIF_REQUIRE(d.Cout > 0 && d.Cout % 8 == 0, "%s: Cout (%d) must be a multiple of 8", what, d.Cout);
  CORRIF_REQUIRE((long long)d.B * d.D * d.H * d.W < (1ll << 31), "%s: volume too large", what);
  return 0;
}

}  // namespace convtc
}  // namespace corrif

using namespace corrif;
using namespace corrif::convtc;

extern "C" int corrif_conv3d_tc_supported(const corrif_conv3d_desc* desc) {
  if (!desc) return 0;
  const corrif_conv3d_desc& d = *desc;
  if (d.nsrc < 1 || d.nsrc > 3 || d.B <= 0 || d.D <= 0 || d.H <= 0 || d.W <= 0 || d.Cout <= 0) return 0;
  int csum = 0;
  for (int i = 0; i < d.nsrc; ++i) csum += d.src[i].C;
  if (csum != d.Cin || (long long)d.B * d.D * d.H * d.W >= (1ll << 31)) return 0;
  return make_plan(d, 148).ok;
}

extern "C" int64_t corrif_conv3d_tc_pack_floats(const corrif_conv3d_desc* desc) {
  if (!corrif_conv3d_tc_supported(desc)) return 0;
  const Plan p = make_plan(*desc, 148);
  return (int64_t)p.n_nchunks * p.w_bytes / 4;
}

extern "C" int corrif_conv3d_tc_pack_weights(const corrif_conv3d_desc* desc, const float* w, float* wpk,
                                             int32_t transpose_flip, void* stream) {
  CORRIF_REQUIRE(desc && w && wpk, "conv3d_tc_pack_weights: null pointer");
  const corrif_conv3d_desc& d = *desc;
  int rc = check_desc(d, "conv3d_tc_pack_weights");
  if (rc) return rc;
  const Plan p = make_plan(d, 148);
  CORRIF_REQUIRE(p.ok, "conv3d_tc_pack_weights: shape not supported by the tcgen05 line convolution");
  Args a{};
  fill_args(a, d, p);
  const long long total = (long long)p.n_nchunks * p.w_bytes / 4;
  // w is the forward weight [Cout_fwd][Cin_fwd][27]; for the data gradient the descriptor's Cout is Cin_fwd
  const int wCin = transpose_flip ? d.Cout : d.Cin;
  pack_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, wpk, wCin, transpose_flip, p.CC,
                                                                                     p.NPAD, p.NKC, p.tap_bytes, total, a);
  return launch_status("conv3d_tc_pack_weights");
}

extern "C" int corrif_conv3d_tc_fwd(const corrif_conv3d_desc* desc, void* stream) {
  CORRIF_REQUIRE(desc != nullptr, "conv3d_tc_fwd: null descriptor");
  const corrif_conv3d_desc& d = *desc;
  int rc = check_desc(d, "conv3d_tc_fwd");
  if (rc) return rc;
  CORRIF_REQUIRE(d.out && ((uintptr_t)d.out % 16) == 0 && d.ldo >= d.Cout && d.ldo % 4 == 0, "conv3d_tc_fwd: output volume");
  CORRIF_REQUIRE(d.wpk != nullptr && ((uintptr_t)d.wpk % 16) == 0, "conv3d_tc_fwd: packed weights missing / unaligned");
  for (int i = 0; i < d.nsrc; ++i)
    CORRIF_REQUIRE(d.src[i].p && ((uintptr_t)d.src[i].p % 16) == 0, "conv3d_tc_fwd: source %d null / unaligned", i);
  const Plan p = make_plan(d, num_sms());
  CORRIF_REQUIRE(p.ok, "conv3d_tc_fwd: shape not supported (ksize 3, W in {16,32,64,128}, B %% (128/W) == 0, source "
                       "channels 8, 16 or a multiple of 32, Cout 8, 16 or a multiple of 32, weights resident in shared memory)");
  // the packed image depends on the plan only through quantities make_plan derives from the channel split
  CUtensorMap tm[3];
  for (int i = 0; i < 3; ++i) {
    rc = encode_line_map(&tm[i], d.src[i < d.nsrc ? i : 0], d.B, d.D, d.H, d.W, p.SWB / 4);
    if (rc) return rc;
  }
  Args a{};
  fill_args(a, d, p);
  cudaStream_t st = (cudaStream_t)stream;
  switch (p.CC * 1000 + p.SWB) {
    case 8032: return launch<8, 32, true>(tm, a, p, st);
    case 8064: return launch<8, 64, true>(tm, a, p, st);
    case 8128: return launch<8, 128, true>(tm, a, p, st);
    case 16032: return launch<16, 32, true>(tm, a, p, st);
    case 16064: return launch<16, 64, true>(tm, a, p, st);
    case 16128: return launch<16, 128, true>(tm, a, p, st);
    case 32032: return a.stats ? launch<32, 32, true>(tm, a, p, st) : launch<32, 32, false>(tm, a, p, st);
    case 32064: return a.stats ? launch<32, 64, true>(tm, a, p, st) : launch<32, 64, false>(tm, a, p, st);
    default: return a.stats ? launch<32, 128, true>(tm, a, p, st) : launch<32, 128, false>(tm, a, p, st);
  }
}
