// Weight gradient of the channels-last 3x3x3 convolution on the 5th-generation tensor cores (tcgen05 + TMEM):
//
//     dW[co][ci][dz, dy, dx] += sum over voxels  x[b, clamp_or_zero(z + dz, y + dy, x + dx), ci] * g[b, z, y, x, co]
//
// for the decoder's big layers (mmvit4.py:222-292: 8 / 16 output channels, up to 64 input channels, 128^3 / 64^3
// voxels), where the warp-level kernel of conv3d_wgrad.cu spends 5.4 of the 58 ms of a micro-batch step.
//
// The reduction runs over voxels (K = x), and a channels-last line [x][channels] is exactly an MN-major operand tile
// ([k][32 floats] rows, the tensor core's SWIZZLE_128B_BASE32B layout: the four 32-byte chunks of a row are XOR-ed with
// the row index mod 4).  TMA cannot write it for the narrow sources of a concatenation (8 / 16 / 24 channels), so
// eight producer warps place the lines with 16-byte cp.async copies at the swizzled addresses - asynchronous, no
// registers, a ring of five or six planes (>= 96 KB in flight per SM) - and the arrival of a plane is an mbarrier the
// copies themselves trigger (cp.async.mbarrier.arrive):
//
//   A (one per input plane): 4 boxes = the 4 consecutive input lines y0-1 .. y0+2, each [64 voxels][32 channels]
//   B (one per gradient plane, written by four more warps): per gradient line y0, y0+1 one row per voxel holding the
//     three x-shifted copies of g:  [g[x+1] | g[x] | g[x-1] | 0]  (+ the clamped tap of replicate padding)
//
// One warp issues D_dz[(j, ci), (gl, dx, co)] += A(z + dz)^T . B(z): M = 128, N = 64 or 128, K = 64 voxels per step,
// both operands MN-major, three accumulators (dz) that live in TMEM for the WHOLE kernel.  Row block j against
// gradient line gl is the y-tap dy = j - 1 - gl (two of the eight (j, gl) blocks are discarded).  Only at the very end
// does a CTA read its accumulators and add them to dW with atomics.  Each input plane is staged once and used by three
// steps.  Inputs reach the tensor core untouched (truncated to TF32): the truncation's mean (-3.52e-4, see conv3d_tc.cu)
// is taken out of dW in the epilogue; the gradient tiles are rounded to nearest when they are written.
// (First version: register transposition into K-major tiles - correct, but with one or two planes of loads in flight
// per SM it was latency-bound at 1.7 ms for the 32 -> 8 layer.)
#include <stdlib.h>
#include "tc05.cuh"

namespace corrif {
namespace wgtc {
using namespace tc05;

constexpr int NT_WARPS = 8;                        // warps 0..7 stage the input planes (cp.async), warp 8 issues the MMAs,
constexpr int NG_WARPS = 4;                        // warps 9..12 build the gradient tiles
constexpr int NISSUE = 2;                          // MMA-issuing warps: warp 8 and warp 13 (each takes half of the k-steps)
constexpr int NTHREADS = 32 * (NT_WARPS + NISSUE + NG_WARPS);
constexpr int XH = 64;                             // voxels along x per step
constexpr int A_SLOT = 4 * XH * 128;               // 4 lines x [64 voxels][32 channels]
constexpr int BOX = XH * 128;                      // one [64 voxels][32 floats] box
constexpr int NBUF = 2;
constexpr float TRUNC_COMP = 1.0f + 3.52e-4f;
constexpr int MAX_SRC = 3;

struct Src { const float* p; int C; long long ld; };

struct Args {
  Src src[MAX_SRC];
  int nsrc;
  int B, D, H, W;
  int CinTot, ci0, ncin, CO;
  int replicate, debug;
  int n_xh, n_yp, ZL, n_zc, total_items;
  const float* g;
  long long ldg;
  float* dW;
};

struct Item { int b, y0, x0, zb, ze; };
__device__ __forceinline__ Item decode_item(const Args& a, int item) {
  Item it;
  const int xh = item % a.n_xh; item /= a.n_xh;
  const int yp = item % a.n_yp; item /= a.n_yp;
  it.b = item % a.B;
  const int zc = item / a.B;
  it.x0 = xh * XH; it.y0 = yp * 2;
  it.zb = zc * a.ZL;
  it.ze = min(it.zb + a.ZL, a.D);
  return it;
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" :: "r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" :: "r"(taddr), "r"(z) : "memory");
}
// TF32 round-to-nearest of a value on its way into an MMA operand (the tensor core drops the low 13 bits)
__device__ __forceinline__ uint32_t rnd(float v) { return __float_as_uint(v) + 0x1000u; }

template <int CO>
__global__ void __launch_bounds__(NTHREADS, 1) conv3d_wgrad_tc_kernel(const Args a) {
  constexpr int NB = 4 * CO;                       // floats per gradient-line row: 3 x-shifts x CO (+ CO of padding)
  constexpr int GBOX = NB / 32;                    // 32-float boxes per gradient line (1 or 2)
  constexpr int BROWS = 2 * NB;                    // N of the MMA: two gradient lines
  constexpr int B_SLOT = 2 * GBOX * BOX;
  constexpr int NA = CO == 8 ? 6 : 5;              // ring of input planes: three live, the rest in flight
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t fullA[NA], emptyA[NA], fullB[NBUF], emptyB[NBUF], done_bar;
  __shared__ uint32_t tmem_base_holder;

  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = sbase, sB = sbase + NA * A_SLOT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = a.D, H = a.H, W = a.W;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NA; ++s) { mbar_init(&fullA[s], 32 * NT_WARPS); mbar_init(&emptyA[s], NISSUE); }
    for (int s = 0; s < NBUF; ++s) { mbar_init(&fullB[s], NG_WARPS); mbar_init(&emptyB[s], NISSUE); }
    mbar_init(&done_bar, NISSUE);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NT_WARPS) tmem_alloc(&tmem_base_holder, 512);
  // padding (channels beyond this pass, the unused quarter of every gradient row) stays zero for the whole kernel:
  // clear the operand rings once
  for (uint32_t i = threadIdx.x; i < (uint32_t)(NA * A_SLOT + NBUF * B_SLOT) / 16; i += NTHREADS) sts128(sbase + i * 16, 0, 0, 0, 0);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  if (warp < 4) {                                  // accumulators start at zero: every MMA accumulates
    const uint32_t t0 = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 512; c += 32) tmem_zero32(t0 + c);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (warp < NT_WARPS) {
    // ================= input planes: cp.async at the swizzled addresses of the MN-major tile =================
    // a plane = 4 lines x 64 voxels x 8 channel quads = 2048 16-byte pieces, 8 per thread: thread = (voxel, quad),
    // so the 8 lanes of a quarter warp copy the 128 contiguous bytes of one voxel
    const int cq = threadIdx.x & 7, xb = threadIdx.x >> 3;         // quad 0..7, voxel 0..31 (and + 32)
    const float* sp = nullptr;                       // the thread's channel quad -> source tensor
    long long sld = 0;
    if (4 * cq < a.ncin) {
      int c = a.ci0 + 4 * cq;
#pragma unroll
      for (int s = 0; s < MAX_SRC; ++s)
        if (s < a.nsrc && sp == nullptr) {
          if (c < a.src[s].C) { sp = a.src[s].p + c; sld = a.src[s].ld; }
          else c -= a.src[s].C;
        }
    }
    // 32-byte chunk (cq >> 1) of row x lands at chunk ((cq >> 1) ^ (x & 3)); x and x + 32 share x & 3
    const uint32_t doff = (uint32_t)xb * 128u + ((((uint32_t)cq >> 1) ^ ((uint32_t)xb & 3u)) << 5) + ((uint32_t)cq & 1u) * 16u;
    uint32_t ia = 0;
    for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
      const Item it = decode_item(a, item);
      const int nplanes = it.ze - it.zb + 2;
      for (int k = 0; k < nplanes; ++k, ++ia) {
        const uint32_t slot = ia % NA;
        mbar_wait(&emptyA[slot], ((ia / NA) & 1u) ^ 1u);
        if (sp != nullptr && !(a.debug & 2)) {
          int pz = it.zb - 1 + k;
          bool zin = pz >= 0 && pz < D;
          if (a.replicate) { pz = min(max(pz, 0), D - 1); zin = true; }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int py = it.y0 - 1 + j;
            bool in = zin && py >= 0 && py < H;
            if (a.replicate) { py = min(max(py, 0), H - 1); in = zin; }
            const float* lp = in ? sp + ((((long long)it.b * D + pz) * H + py) * W + it.x0 + xb) * sld : sp;
            const uint32_t dst = sA + slot * A_SLOT + (uint32_t)j * BOX + doff;
            const int nbytes = in ? 16 : 0;          // out of the volume with zero padding: zero-fill
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(lp), "r"(nbytes) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst + 32u * 128u), "l"(in ? lp + 32 * sld : sp), "r"(nbytes) : "memory");
          }
        }
        // the barrier counts this thread once its copies have landed
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(&fullA[slot])) : "memory");
      }
    }
  } else if (warp > NT_WARPS && warp <= NT_WARPS + NG_WARPS) {
    // ================= gradient rows: one thread per voxel of the two gradient lines =================
    // The loads of the next plane are issued before the current one is stored (two register sets that swap roles):
    // otherwise every step pays a full memory latency here, which the two-slot ring cannot hide.
    const int tid = threadIdx.x - 32 * (NT_WARPS + 1);
    const int gl = tid >> 6, xl = tid & 63;
    struct GV { float4 m[CO / 4], c[CO / 4], p[CO / 4]; };          // g[x - 1], g[x], g[x + 1]
    int item = blockIdx.x, z = 0, x = 0;
    Item it = decode_item(a, item);
    bool more = item < a.total_items;
    if (more) z = it.zb;
    auto load_next = [&](GV& v) -> bool {                           // load the next plane of the sequence, if any
      if (!more) return false;
      x = it.x0 + xl;
      const float* gp = a.g + ((((long long)it.b * D + z) * H + it.y0 + gl) * W + x) * a.ldg;
#pragma unroll
      for (int c = 0; c < CO / 4; ++c) {
        v.c[c] = ld4(gp + 4 * c);
        v.m[c] = x > 0 ? ld4(gp - a.ldg + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        v.p[c] = x < W - 1 ? ld4(gp + a.ldg + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (++z == it.ze) {
        item += gridDim.x;
        more = item < a.total_items;
        if (more) { it = decode_item(a, item); z = it.zb; }
      }
      return true;
    };
    // x of a thread never changes (x0 + xl with the same xl; x0 only matters through the line ends)
    auto store = [&](uint32_t ib, const GV& v, int xs) {
      const uint32_t slot = ib & 1u;
      mbar_wait(&emptyB[slot], ((ib >> 1) & 1u) ^ 1u);
      // replicate padding reads x[clamp(x + dx)]: at the two ends of a line the clamped tap lands on the voxel itself
      const float em = (a.replicate && xs == 0) ? 1.f : 0.f, ep = (a.replicate && xs == W - 1) ? 1.f : 0.f;
      // row xl of the line's box(es): float n = dxi * CO + co sits in 32-byte chunk n / 8 (swizzled with xl & 3)
      const uint32_t rowb = sB + slot * B_SLOT + (uint32_t)(gl * GBOX) * BOX + (uint32_t)xl * 128u;
#pragma unroll
      for (int dxi = 0; dxi < 3; ++dxi) {
#pragma unroll
        for (int c = 0; c < CO / 4; ++c) {
          const float4 g0 = v.c[c], gs = dxi == 0 ? v.p[c] : v.m[c];
          const float e = dxi == 0 ? em : ep;
          uint32_t o[4];
          if (dxi == 1) { o[0] = rnd(g0.x); o[1] = rnd(g0.y); o[2] = rnd(g0.z); o[3] = rnd(g0.w); }
          else { o[0] = rnd(gs.x + e * g0.x); o[1] = rnd(gs.y + e * g0.y); o[2] = rnd(gs.z + e * g0.z); o[3] = rnd(gs.w + e * g0.w); }
          const uint32_t n = (uint32_t)(dxi * CO + 4 * c), box = n >> 5, ch = (n >> 3) & 3u;
          sts128(rowb + box * BOX + ((ch ^ ((uint32_t)xl & 3u)) << 5) + (n & 7u) * 4u, o[0], o[1], o[2], o[3]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&fullB[slot]);
    };
    GV va, vb;
    uint32_t ib = 0;
    int xa = 0, xbb = 0;
    bool ha = load_next(va);
    xa = x;
    while (ha) {
      const bool hb = load_next(vb);
      xbb = x;
      store(ib++, va, xa);
      if (!hb) break;
      ha = load_next(va);
      xa = x;
      store(ib++, vb, xbb);
    }
  } else {
    // ================= MMA issuers =================
    // Every MMA accumulates, so the two issuing warps need no order between them: each takes half of the k-steps of
    // every (plane, z-tap) product and commits its own MMAs to the slot barriers (count NISSUE).  One thread issuing
    // all 24 MMAs of a step needed ~2.1 k cycles per step against ~1.3 k of tensor-pipe time (ncu: pipe 40 % active).
    constexpr uint32_t idesc = idesc_tf32(BROWS, true, true);
    constexpr int KS_PER = (XH / 8) / NISSUE;
    const int iss = warp == NT_WARPS ? 0 : 1;
    const uint32_t a16 = desc_lo_mnmajor(sA, BOX), b16 = desc_lo_mnmajor(sB, BOX);
    uint32_t wslot = 0, phmask = 0, ib = 0;              // next ring slot to wait for; expected parity per slot
    auto wait_plane = [&]() -> uint32_t {
      const uint32_t sl = wslot;
      mbar_wait(&fullA[sl], (phmask >> sl) & 1u);
      phmask ^= 1u << sl;
      wslot = wslot + 1 == NA ? 0 : wslot + 1;
      return sl;
    };
    for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
      const Item it = decode_item(a, item);
      const int nsteps = it.ze - it.zb;
      uint32_t s0 = wait_plane(), s1 = wait_plane();
      for (int s = 0; s < nsteps; ++s, ++ib) {
        const uint32_t s2 = wait_plane();
        mbar_wait(&fullB[ib & 1u], (ib >> 1) & 1u);
        fence_proxy_async();                               // the planes were written through the generic proxy (cp.async)
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t bs = b16 + (ib & 1u) * (uint32_t)(B_SLOT >> 4) + (uint32_t)(iss * KS_PER * (1024 >> 4));
          const uint32_t sl[3] = {s0, s1, s2};
#pragma unroll
          for (int dzi = 0; dzi < 3; ++dzi) {
            if (a.debug & 1) break;
            const uint32_t as = a16 + sl[dzi] * (uint32_t)(A_SLOT >> 4) + (uint32_t)(iss * KS_PER * (1024 >> 4));
            const uint32_t d_tmem = tmem_base + (uint32_t)(dzi * BROWS);
#pragma unroll
            for (int ks = 0; ks < KS_PER; ++ks)            // 8 voxels (k rows) per MMA = 1 KB further in both tiles
              tcgen05_mma_tf32(d_tmem, desc_from(as + (uint32_t)(ks * (1024 >> 4)), DESC_HI_MNMAJOR),
                               desc_from(bs + (uint32_t)(ks * (1024 >> 4)), DESC_HI_MNMAJOR), idesc, 1u);
          }
          tcgen05_commit(&emptyB[ib & 1u]);
          tcgen05_commit(&emptyA[s0]);                     // plane z - 1 is not read again
          if (s == nsteps - 1) {
            tcgen05_commit(&emptyA[s1]);
            tcgen05_commit(&emptyA[s2]);
          }
        }
        __syncwarp();
        s0 = s1; s1 = s2;
      }
    }
    if (elect_one()) tcgen05_commit(&done_bar);
    __syncwarp();
  }

  // ================= epilogue: the CTA's partial weight gradient -> dW =================
  if (warp < 4) {
    mbar_wait(&done_bar, 0);
    tcgen05_fence_after();
    const int row = warp * 32 + lane, j = row >> 5, ci = row & 31;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int dzi = 0; dzi < 3; ++dzi)
      for (int g2 = 0; g2 < 2; ++g2) {
        const int dy = j - 1 - g2;                                     // warp-uniform (j = warp)
        for (int n0 = 0; n0 < 3 * CO; n0 += 8) {
          uint32_t r[8];
          tmem_ld8(taddr + (uint32_t)(dzi * BROWS + g2 * NB + n0), r);
          if (dy < -1 || dy > 1 || ci >= a.ncin) continue;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int n = n0 + e, dxi = n / CO, co = n - dxi * CO;
            atomicAdd(a.dW + ((long long)co * a.CinTot + a.ci0 + ci) * 27 + dzi * 9 + (dy + 1) * 3 + dxi,
                      __uint_as_float(r[e]) * TRUNC_COMP);
          }
        }
      }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == NT_WARPS) tmem_dealloc(tmem_base, 512);
}

static bool supported(const corrif_conv3d_desc& d) {
  if (d.ksize != 3 || d.nsrc < 1 || d.nsrc > 3) return false;
  if (d.W % XH || d.H % 2 || d.B <= 0 || d.D <= 0) return false;
  if (!(d.Cout == 8 || d.Cout == 16) || d.Cin <= 0 || d.Cin > 64 || d.Cin % 4) return false;
  int csum = 0;
  for (int i = 0; i < d.nsrc; ++i) {
    if (d.src[i].C <= 0 || d.src[i].C % 4 || d.src[i].ld % 4 || d.src[i].ld < d.src[i].C) return false;
    csum += d.src[i].C;
  }
  return csum == d.Cin && (long long)d.B * d.D * d.H * d.W < (1ll << 31);
}

}  // namespace wgtc
}  // namespace corrif

using namespace corrif;
using namespace corrif::wgtc;

extern "C" int corrif_conv3d_wgrad_tc_supported(const corrif_conv3d_desc* desc) {
  return desc != nullptr && supported(*desc) ? 1 : 0;
}

extern "C" int corrif_conv3d_wgrad_tc(const corrif_conv3d_desc* desc, const float* g, int64_t ldg, float* dW,
                                      void* stream) {
  CORRIF_REQUIRE(desc != nullptr && g != nullptr && dW != nullptr, "conv3d_wgrad_tc: null pointer");
  const corrif_conv3d_desc& d = *desc;
  CORRIF_REQUIRE(supported(d), "conv3d_wgrad_tc: shape not supported (ksize 3, W %% 64 == 0, H even, Cout 8 or 16, Cin <= 64)");
  CORRIF_REQUIRE(d.pad_mode == CORRIF_PAD_ZEROS || d.pad_mode == CORRIF_PAD_REPLICATE, "conv3d_wgrad_tc: pad_mode");
  CORRIF_REQUIRE(((uintptr_t)g % 16) == 0 && ldg % 4 == 0 && ldg >= d.Cout, "conv3d_wgrad_tc: gradient volume unaligned / ld < Cout");
  for (int i = 0; i < d.nsrc; ++i)
    CORRIF_REQUIRE(d.src[i].p && ((uintptr_t)d.src[i].p % 16) == 0, "conv3d_wgrad_tc: source %d null / unaligned", i);
  Args a{};
  for (int i = 0; i < MAX_SRC; ++i) {
    a.src[i].p = i < d.nsrc ? d.src[i].p : nullptr;
    a.src[i].C = i < d.nsrc ? d.src[i].C : 0;
    a.src[i].ld = i < d.nsrc ? d.src[i].ld : 0;
  }
  a.nsrc = d.nsrc; a.B = d.B; a.D = d.D; a.H = d.H; a.W = d.W; a.CinTot = d.Cin; a.CO = d.Cout;
  a.replicate = d.pad_mode == CORRIF_PAD_REPLICATE;
  a.g = g; a.ldg = ldg; a.dW = dW;
  static const int debug = getenv("CORRIF_WGRAD_TC_DEBUG") ? atoi(getenv("CORRIF_WGRAD_TC_DEBUG")) : 0;
  a.debug = debug;
  a.n_xh = d.W / XH; a.n_yp = d.H / 2;
  const int nsm = num_sms();
  // z range per item: every range re-transposes two halo planes; the items should fill whole rounds of the SMs
  long long best = -1;
  for (int zl = d.D < 4 ? d.D : 4; zl <= d.D; ++zl) {
    const long long nz = (d.D + zl - 1) / zl;
    const long long items = nz * d.B * a.n_yp * a.n_xh;
    const long long cost = ((items + nsm - 1) / nsm) * (zl + 2);
    if (best < 0 || cost <= best) { best = cost; a.ZL = zl; a.n_zc = (int)nz; a.total_items = (int)items; }
  }
  const int b_slot = 2 * (d.Cout / 8) * BOX;
  const int smem = 1024 + (d.Cout == 8 ? 6 : 5) * A_SLOT + NBUF * b_slot;
  auto kern8 = conv3d_wgrad_tc_kernel<8>;
  auto kern16 = conv3d_wgrad_tc_kernel<16>;
  static int configured8 = 0, configured16 = 0;
  int& configured = d.Cout == 8 ? configured8 : configured16;
  if (configured < smem) {
    cudaError_t e = d.Cout == 8 ? cudaFuncSetAttribute(kern8, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                                : cudaFuncSetAttribute(kern16, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_last_error("conv3d_wgrad_tc: smem attribute (%d B): %s", smem, cudaGetErrorString(e)); return (int)e; }
    configured = smem;
  }
  const unsigned grid = (unsigned)(a.total_items < nsm ? a.total_items : nsm);
  for (int ci0 = 0; ci0 < d.Cin; ci0 += 32) {          // 32 input channels per pass (rows of the MMA's M dimension)
    a.ci0 = ci0;
    a.ncin = d.Cin - ci0 < 32 ? d.Cin - ci0 : 32;
    if (d.Cout == 8) kern8<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
    else kern16<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
    int rc = launch_status("conv3d_wgrad_tc");
    if (rc) return rc;
  }
  return 0;
}
