// Shared pieces of the channels-last 3-D convolution kernels (conv3d_fwd.cu, conv3d_wgrad.cu): tile geometry,
// shared-memory layout of the staged input window, the multi-source channels-last loader and the warp-level
// TF32 MMA.
//
// Which convolution runs where.  The 3x3x3 layers at 128^3 / 64^3 voxels - forward and data gradient, where most of the
// time is - run on the tcgen05 "line convolution" of conv3d_tc.cu (a 128-voxel row is the M dimension, the x-taps are
// folded into N, the y / z taps are a choice of TMEM accumulator).  The kernels in this header keep
//   * the weight gradient (reduction over voxels: both operands would have to be K(= voxel)-contiguous, which
//     channels-last lines are not),
//   * the 1x1x1 convolutions and the 3x3x3 layers whose packed weights do not fit in shared memory (>= 64 channels on
//     both sides, 32^3 voxels and below) or whose geometry the line kernel does not take (ragged W, channel counts
//     that are not 8 / 16 / a multiple of 32),
// as warp-level mma.sync over an input window staged once in shared memory with the 27 taps read as shifted views of
// it (A fragments reused across taps in registers: 10 staged rows serve 3 dy-taps x 4 row pairs).
// Measured mma.sync m16n8k8 TF32 rate on B200: 270 TFLOP/s (profiles/r02a_mma_sync_probe.txt).
#pragma once
#include "common.cuh"

namespace corrif {
namespace conv {

constexpr int TX = 8, TY = 8, TZ = 4;                 // output tile: one warp per z-slice, 8 x 8 voxels each
constexpr int HX = TX + 2, HY = TY + 2, HZ = TZ + 2;  // staged input window (one-voxel halo)
constexpr int TILE_VOX = TX * TY * TZ;                // 256
constexpr int WIN_VOX = HX * HY * HZ;                 // 600
constexpr int NTHREADS = 128;
// shared-memory window: [channel group of 4][voxel][4 floats]; the group stride is padded by one voxel so that
// (stride / 4) % 32 == 4: the 8 lanes of a quarter-warp that store the 8 groups of one voxel, and the 32 lanes of
// a weight-gradient A-fragment load (two groups x four voxels), all hit distinct banks
constexpr int CGS3 = (WIN_VOX + 1) * 16;              // bytes per channel group, 3x3x3 kernels
constexpr int CGS1 = (TILE_VOX + 1) * 16;             // 1x1x1 kernels: the tile is 256 consecutive voxels
constexpr int MAX_SRC = 3;

struct Src {
  const float* p;       // channels-last volume [B, D, H, W, C] with voxel stride ld (>= C)
  int C;
  long long ld;
};

struct Geom {
  int B, D, H, W;
  int tiles_x, tiles_y, tiles_z;
};

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm(   // not volatile: a pure function of its operands, the scheduler may interleave it with the fragment loads
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Shared-memory fragment loads as ordinary C++ loads through the shared window (ordered against __syncthreads by
// the compiler, free to be scheduled ahead of the MMAs that do not depend on them - volatile asm pinned every load
// and every MMA in program order).
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  return *reinterpret_cast<const uint32_t*>(__cvta_shared_to_generic(addr));
}
__device__ __forceinline__ void lds64(uint32_t addr, uint32_t& a, uint32_t& b) {
  const uint2 v = *reinterpret_cast<const uint2*>(__cvta_shared_to_generic(addr));
  a = v.x; b = v.y;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// 4 consecutive channels [c, c+4) of voxel `vox` of the channel-concatenation of the sources (c % 4 == 0;
// every source has C % 4 == 0); zeros beyond the last source.  The caller rounds to TF32 (the MMA truncates).
__device__ __forceinline__ float4 load_cat4_raw(const Src (&src)[MAX_SRC], int nsrc, long long vox, int c) {
#pragma unroll
  for (int s = 0; s < MAX_SRC; ++s) {
    if (s < nsrc) {
      if (c < src[s].C) return ld4(src[s].p + vox * src[s].ld + c);
      c -= src[s].C;
    }
  }
  return make_float4(0.f, 0.f, 0.f, 0.f);
}

// Round-to-nearest TF32 for values on their way into an MMA operand: add half a TF32 ulp and let the tensor core
// truncate the low 13 bits (ties away from zero, like cvt.rna; finite inputs only; the low bits stay dirty).
// One integer add per element instead of the four instructions cvt.rna.tf32 expands to.
__device__ __forceinline__ float4 rnd4(float4 v) {
  v.x = __uint_as_float(__float_as_uint(v.x) + 0x1000u); v.y = __uint_as_float(__float_as_uint(v.y) + 0x1000u);
  v.z = __uint_as_float(__float_as_uint(v.z) + 0x1000u); v.w = __uint_as_float(__float_as_uint(v.w) + 0x1000u);
  return v;
}

// Stage the input window of one tile for channels [c0, c0 + kc) into shared memory.
//   KS == 3: window voxel (hz,hy,hx) = input voxel (z0+hz-1, y0+hy-1, x0+hx-1), out-of-volume coordinates are
//            clamped (replicate padding, mmvit4.py:225-236 pad_type='replicate') or read as zeros (the RFM blocks'
//            default zero padding, mmvit4.py:47-56)
//   KS == 1: the tile is 256 consecutive voxels of sample b starting at v0
// A thread keeps ONE (x, channel group) for the whole window and walks its 60 (z, y) lines, so the source tensor of
// the concatenation, its stride and the clamped x are resolved once per pass instead of once per 16 bytes (the
// first version spent 70 % of the kernel's instructions on per-item index arithmetic).  Loads are issued in batches
// of 8-10 BEFORE any of them is stored (a load -> store pair per iteration pays a DRAM round trip per 16 bytes).
template <int KS>
__device__ __forceinline__ void stage_window(uint32_t smem_in, const Src (&src)[MAX_SRC], int nsrc, const Geom& g,
                                             int b, int z0, int y0, int x0, long long v0, long long nvox,
                                             int c0, int kc, bool replicate) {
  const int gshift = kc == 32 ? 3 : (kc == 16 ? 2 : 1);      // channel groups per voxel = kc / 4 (8, 4 or 2)
  const int groups = 1 << gshift;
  constexpr int CGS = KS == 3 ? CGS3 : CGS1;
  // the thread's channel group -> source tensor (fixed for the pass)
  const int cg = threadIdx.x & (groups - 1);
  const float* sp = nullptr;
  long long sld = 0;
  {
    int c = c0 + cg * 4;
#pragma unroll
    for (int s = 0; s < MAX_SRC; ++s)
      if (s < nsrc && sp == nullptr) {
        if (c < src[s].C) { sp = src[s].p + c; sld = src[s].ld; }
        else c -= src[s].C;
      }
  }
  if constexpr (KS == 3) {
    const int per = HX << gshift;                        // threads per window line
    const int lines_per_iter = NTHREADS / per;           // 6 / 3 / 1
    const int rl = threadIdx.x / per, hx = (threadIdx.x - rl * per) >> gshift;
    const bool active = rl < lines_per_iter;
    int x = x0 + hx - 1;
    const bool xin = x >= 0 && x < g.W;
    x = min(max(x, 0), g.W - 1);
    const uint32_t sdst = smem_in + cg * CGS + hx * 16;
    constexpr int U = 10, LINES = HZ * HY;
    for (int r0 = rl; r0 < LINES; r0 += lines_per_iter * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = r0 + u * lines_per_iter;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active && r < LINES && sp != nullptr) {
          const int hz = r / HY, hy = r - hz * HY;
          int z = z0 + hz - 1, y = y0 + hy - 1;
          bool inside = xin && z >= 0 && z < g.D && y >= 0 && y < g.H;
          if (replicate) { z = min(max(z, 0), g.D - 1); y = min(max(y, 0), g.H - 1); inside = true; }
          if (inside) v[u] = ld4(sp + (long long)(((b * g.D + z) * g.H + y) * g.W + x) * sld);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = r0 + u * lines_per_iter;
        if (active && r < LINES) sts128(sdst + r * (HX * 16), rnd4(v[u]));
      }
    }
  } else {
    constexpr int U = 8;
    const int vl = threadIdx.x >> gshift, vstep = NTHREADS >> gshift;     // voxels advance by vstep per item
    const uint32_t sdst = smem_in + cg * CGS;
    for (int h0 = vl; h0 < TILE_VOX; h0 += vstep * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int hv = h0 + u * vstep;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hv < TILE_VOX && sp != nullptr && v0 + hv < nvox) v[u] = ld4(sp + ((long long)b * nvox + v0 + hv) * sld);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int hv = h0 + u * vstep;
        if (hv < TILE_VOX) sts128(sdst + hv * 16, rnd4(v[u]));
      }
    }
  }
}

// ---- asynchronous staging (cp.async): global -> shared without a register round trip, zero-filled when invalid --
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid, bool via_l1 = false) {
  const int n = valid ? 16 : 0;
  if (via_l1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(n) : "memory");
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// The 3x3x3 window of stage_window<3>, issued as cp.async copies (same thread -> (x, channel group) mapping).  The
// data lands un-rounded; consumers add half a TF32 ulp when they load their fragments (see rnd_u32).
__device__ __forceinline__ void stage_window3_async(uint32_t smem_in, const Src (&src)[MAX_SRC], int nsrc, const Geom& g,
                                                    int b, int z0, int y0, int x0, int c0, int kc, bool replicate,
                                                    bool via_l1 = false) {
  const int gshift = kc == 32 ? 3 : (kc == 16 ? 2 : 1);
  const int groups = 1 << gshift;
  const int cg = threadIdx.x & (groups - 1);
  const float* sp = nullptr;
  long long sld = 0;
  {
    int c = c0 + cg * 4;
#pragma unroll
    for (int s = 0; s < MAX_SRC; ++s)
      if (s < nsrc && sp == nullptr) {
        if (c < src[s].C) { sp = src[s].p + c; sld = src[s].ld; }
        else c -= src[s].C;
      }
  }
  const int per = HX << gshift;
  const int lines_per_iter = NTHREADS / per;
  const int rl = threadIdx.x / per, hx = (threadIdx.x - rl * per) >> gshift;
  if (rl >= lines_per_iter) return;
  int x = x0 + hx - 1;
  const bool xin = x >= 0 && x < g.W;
  x = min(max(x, 0), g.W - 1);
  const uint32_t sdst = smem_in + cg * CGS3 + hx * 16;
  const float* any = src[0].p;                       // a valid address for the zero-fill form
#pragma unroll 4
  for (int r = rl; r < HZ * HY; r += lines_per_iter) {
    const int hz = r / HY, hy = r - hz * HY;
    int z = z0 + hz - 1, y = y0 + hy - 1;
    bool inside = xin && z >= 0 && z < g.D && y >= 0 && y < g.H;
    if (replicate) { z = min(max(z, 0), g.D - 1); y = min(max(y, 0), g.H - 1); inside = true; }
    inside = inside && sp != nullptr;
    const float* p = inside ? sp + (long long)(((b * g.D + z) * g.H + y) * g.W + x) * sld : any;
    cp_async16(sdst + r * (HX * 16), p, inside, via_l1);
  }
}
__device__ __forceinline__ void stage_linear_async(uint32_t smem_dst, const float4* __restrict__ gsrc, int n16) {
  for (int i = threadIdx.x; i < n16; i += NTHREADS) cp_async16(smem_dst + i * 16, gsrc + i, true);
}
// round-to-nearest TF32 of a raw fp32 bit pattern headed for an MMA operand (the MMA drops the low 13 bits)
__device__ __forceinline__ uint32_t rnd_u32(uint32_t v) { return v + 0x1000u; }

// batched copy of n16 16-byte words global -> shared (packed weights), loads first, then stores
__device__ __forceinline__ void stage_linear(uint32_t smem_dst, const float4* __restrict__ gsrc, int n16) {
  constexpr int U = 8;
  for (int base = threadIdx.x; base < n16; base += NTHREADS * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * NTHREADS;
      v[u] = i < n16 ? __ldg(gsrc + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * NTHREADS;
      if (i < n16) sts128(smem_dst + i * 16, v[u]);
    }
  }
}

// Tiling plan shared by the packer and the kernels: NB = 8-channel output blocks per CTA, KC = input channels
// staged per pass.  Sized so that TWO window buffers plus the packed weights of a pass leave room for 2-3 CTAs per
// SM (3x3x3: 16 channels per pass for one output block, 8 for wider ones).  KC never exceeds the channel count
// rounded up to 8 / 16 / 32 (no passes over zero padding).
struct Plan { int NB, KC; };
inline Plan conv_plan(int Cin, int Cout, int ks) {
  static const int cand[6] = {8, 6, 4, 3, 2, 1};
  Plan p{1, 32};
  for (int i = 0; i < 6; ++i)
    if ((Cout / 8) % cand[i] == 0) { p.NB = cand[i]; break; }
  if (ks == 3) p.KC = p.NB == 1 ? 16 : 8;
  else p.KC = 32;
  const int need = Cin <= 8 ? 8 : (Cin <= 16 ? 16 : 32);
  if (p.KC > need) p.KC = need;
  return p;
}
inline int wgrad_nb(int Cout) {
  static const int cand[4] = {4, 3, 2, 1};
  for (int i = 0; i < 4; ++i)
    if ((Cout / 8) % cand[i] == 0) return cand[i];
  return 1;
}

}  // namespace conv
}  // namespace corrif
