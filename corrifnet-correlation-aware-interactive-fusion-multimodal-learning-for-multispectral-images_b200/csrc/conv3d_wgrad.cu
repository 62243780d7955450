// Weight gradient of the channels-last 3-D convolution:
//
//   dW[co][ci][tap] += sum_{b, v} x[b, clamp_or_zero(v + tap), ci] * g[b, v, co]        (g = d pre-activation)
//
// As a GEMM the contraction runs over up to 16.8 M voxels and the result is tiny (27*Cin x Cout), so the kernel is
// persistent: a CTA owns one pass of 16 input channels x one block of NB*8 output channels, walks the voxel tiles
// with a grid stride, stages each tile's input window and gradient tile in shared memory and keeps its partial
// dW in REGISTERS across all tiles; it touches global memory with one atomic add per element at the very end
// (a few hundred CTAs x 27*16*NB*8 elements instead of one atomic per tile).  Warps split the 27 taps (3x3x3) or
// the voxel steps (1x1x1).  Fragment mapping (m16n8k8, A = x^T [16 ci x 8 voxels], B = g [8 voxels x 8 co]): the 8
// voxels of a step are the 8 x-neighbours of one line; k = 0..3 are the even, k = 4..7 the odd ones, which makes
// the 32 lanes of every A load hit 32 distinct banks of the [channel group][voxel][4] window.
#include <stdlib.h>
#include "conv3d.cuh"

namespace corrif {
namespace conv {

constexpr int WKC = 16;                       // input channels per pass (one m16 block)

struct WgradArgs {
  Src src[MAX_SRC];
  int nsrc;
  Geom g;
  int Cin, Cout;
  int replicate;
  const float* grad;      // [B, D, H, W, Cout], voxel stride ldg
  long long ldg;
  float* dW;              // torch layout [Cout][Cin][taps]
  int total_tiles;
  int via_l1;             // cp.async through L1 (.ca) instead of L2 only (.cg): A/B switch CORRIF_WGRAD_CA
};

template <int KS, int NB>
__global__ void __launch_bounds__(NTHREADS, NB <= 2 ? 3 : 2) conv3d_wgrad_kernel(const WgradArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int TAPS = KS == 3 ? 27 : 1;
  constexpr int CGS = KS == 3 ? CGS3 : CGS1;
  constexpr int MYTAPS = KS == 3 ? 7 : 1;            // taps per warp: warp w owns taps w, w+4, ...
  constexpr int GLD = NB * 8 + 4;                    // padded row of the gradient tile (bank-conflict free B loads)
  const uint32_t s_in = smem_addr(smem);
  const uint32_t s_g = s_in + (WKC / 4) * CGS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int pass = blockIdx.y, nt = blockIdx.z;
  const long long nvox = (long long)a.g.D * a.g.H * a.g.W;
  const int per = KS == 3 ? a.g.tiles_x * a.g.tiles_y * a.g.tiles_z : (int)((nvox + TILE_VOX - 1) / TILE_VOX);

  float acc[MYTAPS][NB][4];
#pragma unroll
  for (int i = 0; i < MYTAPS; ++i)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][nb][j] = 0.f;

  for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
    const int b = tile / per;
    int rem = tile - b * per;
    int z0 = 0, y0 = 0, x0 = 0;
    long long v0 = 0;
    if constexpr (KS == 3) {
      x0 = (rem % a.g.tiles_x) * TX; rem /= a.g.tiles_x;
      y0 = (rem % a.g.tiles_y) * TY;
      z0 = (rem / a.g.tiles_y) * TZ;
    } else {
      v0 = (long long)rem * TILE_VOX;
    }
    __syncthreads();
    stage_window<KS>(s_in, a.src, a.nsrc, a.g, b, z0, y0, x0, v0, nvox, pass * WKC, WKC, a.replicate != 0);
    // gradient tile: 256 voxels x NB*8 channels, zeros outside the volume (those voxels must not contribute);
    // loads batched ahead of the stores like the window's
    {
      constexpr int U = 8, TOTAL = TILE_VOX * NB * 2;
      for (int base = threadIdx.x; base < TOTAL; base += NTHREADS * U) {
        float4 val[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * NTHREADS;
          val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < TOTAL) {
            const int q = i % (NB * 2), v = i / (NB * 2);
            long long vox;
            bool ok;
            if constexpr (KS == 3) {
              const int x = x0 + (v & 7), y = y0 + ((v >> 3) & 7), z = z0 + (v >> 6);
              ok = z < a.g.D && y < a.g.H && x < a.g.W;
              vox = (((long long)b * a.g.D + z) * a.g.H + y) * a.g.W + x;
            } else {
              ok = v0 + v < nvox;
              vox = (long long)b * nvox + v0 + v;
            }
            if (ok) val[u] = ld4(a.grad + vox * a.ldg + nt * NB * 8 + q * 4);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * NTHREADS;
          if (i < TOTAL) sts128(s_g + ((i / (NB * 2)) * GLD + (i % (NB * 2)) * 4) * 4, rnd4(val[u]));
        }
      }
    }
    __syncthreads();
    // ---- accumulate: k-steps of 8 voxels (one line of the tile) ------------------------------------
    const uint32_t a_lane = s_in + (g >> 2) * CGS + (g & 3) * 4;      // rows g: channel g; rows g+8: + 2 groups
    if constexpr (KS == 3) {
#pragma unroll 1
      for (int line = 0; line < TZ * TY; ++line) {
        const int zz = line >> 3, yy = line & 7;
        uint32_t Bf[NB][2];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const uint32_t gb = s_g + ((line * 8 + 2 * t) * GLD + nb * 8 + g) * 4;
          Bf[nb][0] = lds32(gb);
          Bf[nb][1] = lds32(gb + GLD * 4);
        }
#pragma unroll
        for (int i = 0; i < MYTAPS; ++i) {
          const int tap = warp + 4 * i;
          if (tap < TAPS) {
            const int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;
            const uint32_t ad = a_lane + ((((zz + dz) * HY + yy + dy) * HX) + 2 * t + dx) * 16;
            const uint32_t a0 = lds32(ad), a1 = lds32(ad + 2 * CGS), a2 = lds32(ad + 16), a3 = lds32(ad + 2 * CGS + 16);
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) mma_tf32(acc[i][nb], a0, a1, a2, a3, Bf[nb][0], Bf[nb][1]);
          }
        }
      }
    } else {
#pragma unroll 1
      for (int step = warp * 8; step < warp * 8 + 8; ++step) {     // warps split the 32 voxel steps of the tile
        uint32_t Bf[NB][2];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const uint32_t gb = s_g + ((step * 8 + 2 * t) * GLD + nb * 8 + g) * 4;
          Bf[nb][0] = lds32(gb);
          Bf[nb][1] = lds32(gb + GLD * 4);
        }
        const uint32_t ad = a_lane + (step * 8 + 2 * t) * 16;
        const uint32_t a0 = lds32(ad), a1 = lds32(ad + 2 * CGS), a2 = lds32(ad + 16), a3 = lds32(ad + 2 * CGS + 16);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) mma_tf32(acc[0][nb], a0, a1, a2, a3, Bf[nb][0], Bf[nb][1]);
      }
    }
  }
  // ---- one atomic add per element: rows = ci (g, g+8), cols = co (2t, 2t+1) ---------------------------
#pragma unroll
  for (int i = 0; i < MYTAPS; ++i) {
    const int tap = KS == 3 ? warp + 4 * i : 0;
    if (tap < TAPS) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ci = pass * WKC + g + (j >> 1) * 8;
          const int co = (nt * NB + nb) * 8 + 2 * t + (j & 1);
          if (ci < a.Cin) atomicAdd(a.dW + ((long long)co * a.Cin + ci) * TAPS + tap, acc[i][nb][j]);
        }
    }
  }
}

// ---- 3x3x3, second version: input lines reused across the dy taps in registers -----------------------------
// The kernel above loads one A fragment (x^T of one shifted line: 4 shared-memory loads) per MMA - 4.3 loads per
// MMA, shared-memory bound at 43 TFLOP/s on the 128^3 layers.  Here warp w owns the 8 output lines of z-slice w
// and keeps their gradient fragments in 16 registers; it walks the 3 x 10 x 3 (dz, window line, dx) A fragments
// once, and each of them is multiplied with the gradient lines y = line - dy of all three dy taps: 360 + 16 loads
// feed 216 MMAs (1.7 per MMA).  All 27 taps accumulate in registers (108); 8 output channels per CTA (grid.z).
// Window and gradient tile are double-buffered with cp.async (the next tile travels while this one is multiplied);
// fragments are rounded to TF32 as they are loaded.
__global__ void __launch_bounds__(NTHREADS, 2) conv3d_wgrad3_kernel(const WgradArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int GLD = 12;                            // 8 channels + 4 padding floats per gradient voxel
  constexpr uint32_t WIN = (WKC / 4) * CGS3, GT = TILE_VOX * GLD * 4;
  const uint32_t s_in0 = smem_addr(smem);            // two window buffers, then two gradient-tile buffers
  const uint32_t s_g0 = s_in0 + 2 * WIN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int pass = blockIdx.y, nt = blockIdx.z;
  const int per = a.g.tiles_x * a.g.tiles_y * a.g.tiles_z;

  float acc[27][4];
#pragma unroll
  for (int i = 0; i < 27; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // cp.async double buffering over this CTA's tiles: tile i+1 is in flight while tile i is multiplied
  auto issue = [&](int tile, int buf) {
    const int b = tile / per;
    int rem = tile - b * per;
    const int x0 = (rem % a.g.tiles_x) * TX; rem /= a.g.tiles_x;
    const int y0 = (rem % a.g.tiles_y) * TY;
    const int z0 = (rem / a.g.tiles_y) * TZ;
    stage_window3_async(s_in0 + buf * WIN, a.src, a.nsrc, a.g, b, z0, y0, x0, pass * WKC, WKC, a.replicate != 0, a.via_l1 != 0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {                     // gradient tile: 256 voxels x 8 channels, zeros outside the volume
      const int i = threadIdx.x + u * NTHREADS;
      const int q = i & 1, v = i >> 1;
      const int x = x0 + (v & 7), y = y0 + ((v >> 3) & 7), z = z0 + (v >> 6);
      const bool ok = z < a.g.D && y < a.g.H && x < a.g.W;
      const float* p = ok ? a.grad + (long long)(((b * a.g.D + z) * a.g.H + y) * a.g.W + x) * a.ldg + nt * 8 + q * 4 : a.grad;
      cp_async16(s_g0 + buf * GT + (v * GLD + q * 4) * 4, p, ok, a.via_l1 != 0);
    }
  };
  int tile = blockIdx.x, buf = 0;
  if (tile < a.total_tiles) issue(tile, 0);
  cp_async_commit();
  for (; tile < a.total_tiles; tile += gridDim.x, buf ^= 1) {
    if (tile + (int)gridDim.x < a.total_tiles) issue(tile + gridDim.x, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t s_in = s_in0 + buf * WIN, s_g = s_g0 + buf * GT;
    uint32_t Bf[8][2];
#pragma unroll
    for (int yy = 0; yy < 8; ++yy) {
      const uint32_t gb = s_g + (((warp * 8 + yy) * 8 + 2 * t) * GLD + g) * 4;
      Bf[yy][0] = rnd_u32(lds32(gb));
      Bf[yy][1] = rnd_u32(lds32(gb + GLD * 4));
    }
    const uint32_t a_lane = s_in + (g >> 2) * CGS3 + (g & 3) * 4 + (2 * t) * 16;
#pragma unroll
    for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
      for (int hy = 0; hy < HY; ++hy) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const uint32_t ad = a_lane + ((((warp + dz) * HY + hy) * HX) + dx) * 16;
          const uint32_t a0 = rnd_u32(lds32(ad)), a1 = rnd_u32(lds32(ad + 2 * CGS3)), a2 = rnd_u32(lds32(ad + 16)),
                         a3 = rnd_u32(lds32(ad + 2 * CGS3 + 16));
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int yy = hy - dy;
            if (yy >= 0 && yy < 8) mma_tf32(acc[(dz * 3 + dy) * 3 + dx], a0, a1, a2, a3, Bf[yy][0], Bf[yy][1]);
          }
        }
      }
    }
    __syncthreads();                                  // this buffer is refilled by the next iteration's issue
  }
  cp_async_wait<0>();
#pragma unroll
  for (int tap = 0; tap < 27; ++tap)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = pass * WKC + g + (j >> 1) * 8;
      const int co = nt * 8 + 2 * t + (j & 1);
      if (ci < a.Cin) atomicAdd(a.dW + ((long long)co * a.Cin + ci) * 27 + tap, acc[tap][j]);
    }
}

static int launch_wgrad3(const WgradArgs& a0, cudaStream_t stream) {
  WgradArgs a = a0;
  const int smem = 2 * ((WKC / 4) * CGS3 + TILE_VOX * 12 * 4);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3d_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_last_error("conv3d_wgrad: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  a.total_tiles = a.g.tiles_x * a.g.tiles_y * a.g.tiles_z * a.g.B;
  const int passes = (a.Cin + WKC - 1) / WKC, ntiles = a.Cout / 8;
  int px = (4 * num_sms() + passes * ntiles - 1) / (passes * ntiles);     // two rounds of the 2 resident CTAs per SM
  px = px < 1 ? 1 : (px > a.total_tiles ? a.total_tiles : px);
  dim3 grid((unsigned)px, (unsigned)passes, (unsigned)ntiles);
  conv3d_wgrad3_kernel<<<grid, NTHREADS, smem, stream>>>(a);
  return launch_status("conv3d_wgrad");
}

template <int KS, int NB>
static int launch(const WgradArgs& a0, cudaStream_t stream) {
  WgradArgs a = a0;
  constexpr int CGS = KS == 3 ? CGS3 : CGS1;
  const int smem = (WKC / 4) * CGS + TILE_VOX * (NB * 8 + 4) * 4;
  auto kern = conv3d_wgrad_kernel<KS, NB>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_last_error("conv3d_wgrad: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    configured = true;
  }
  const long long nvox = (long long)a.g.D * a.g.H * a.g.W;
  const int per = KS == 3 ? a.g.tiles_x * a.g.tiles_y * a.g.tiles_z : (int)((nvox + TILE_VOX - 1) / TILE_VOX);
  a.total_tiles = per * a.g.B;
  const int passes = (a.Cin + WKC - 1) / WKC, ntiles = a.Cout / (NB * 8);
  int px = (6 * num_sms() + passes * ntiles - 1) / (passes * ntiles);     // two rounds of the ~3 resident CTAs per SM
  px = px < 1 ? 1 : (px > a.total_tiles ? a.total_tiles : px);
  dim3 grid((unsigned)px, (unsigned)passes, (unsigned)ntiles);
  kern<<<grid, NTHREADS, smem, stream>>>(a);
  return launch_status("conv3d_wgrad");
}

template <int KS>
static int launch_nb(int NB, const WgradArgs& a, cudaStream_t s) {
  switch (NB) {
    case 1: return launch<KS, 1>(a, s);
    case 2: return launch<KS, 2>(a, s);
    case 3: return launch<KS, 3>(a, s);
    case 4: return launch<KS, 4>(a, s);
  }
  set_last_error("conv3d_wgrad: unsupported channel block %d", NB);
  return CORRIF_EINVAL;
}

// ---- replicate-padding part of the data gradient -----------------------------------------------------------
// Forward reads x[clamp(v + tap)].  The zero-padded data gradient (conv3d_fwd with mirrored weights) covers every
// (v, tap) whose v + tap lies inside the volume; what is left are the pairs with v + tap OUTSIDE in at least one
// axis, which the clamp folds onto border voxels u = clamp(v + tap):  dx[u] += sum_co g[v][co] * W[co][ci][tap].
// Per axis the v with clamp(v + t) == u are  u - t  (inside the volume: already counted unless another axis is
// outside) and  u itself when (t == -1 and u == 0) or (t == +1 and u == n-1)  (outside: the clamped read).
// One thread per (border voxel, input channel), channels fastest.
__device__ __forceinline__ void border_voxel(uint32_t r, int D, int H, int W, int& z, int& y, int& x) {
  // z faces, then y faces of the remaining slab, then x faces of the remaining core (32-bit arithmetic throughout:
  // the first version's 64-bit divisions cost more than the gradient itself)
  const uint32_t zf = (uint32_t)H * W, nz = D >= 2 ? 2 : 1, nyf = H >= 2 ? 2 : 1, nxf = W >= 2 ? 2 : 1;
  const uint32_t Dm = D - nz, Hm = H - nyf;
  if (r < nz * zf) {
    z = r < zf ? 0 : D - 1; r -= r < zf ? 0 : zf; y = (int)(r / W); x = (int)(r - (uint32_t)y * W);
    return;
  }
  r -= nz * zf;
  const uint32_t yfaces = Dm * nyf * W;
  if (r < yfaces) {
    const uint32_t zz = r / (nyf * W);
    r -= zz * (nyf * W);
    z = 1 + (int)zz; y = r < (uint32_t)W ? 0 : H - 1; x = (int)(r < (uint32_t)W ? r : r - W);
    return;
  }
  r -= yfaces;
  const uint32_t zz = r / (Hm * nxf);
  r -= zz * (Hm * nxf);
  z = 1 + (int)zz; y = 1 + (int)(r / nxf); x = (r % nxf) == 0 ? 0 : W - 1;
}

__global__ void __launch_bounds__(256) dgrad_border_kernel(const float* __restrict__ grad, long long ldg,
                                                           const float* __restrict__ w, float* __restrict__ dx,
                                                           long long ldx, int B, int D, int H, int W, int Cin, int Cout,
                                                           uint32_t nborder, uint32_t total) {
  // one thread per (border voxel, 4 input channels)
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= total) return;
  const uint32_t Q = (uint32_t)Cin >> 2;
  const uint32_t bu = tid / Q;
  const int ci = (int)(tid - bu * Q) * 4;
  const int b = (int)(bu / nborder);
  int uz, uy, ux;
  border_voxel(bu - (uint32_t)b * nborder, D, H, W, uz, uy, ux);
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int tz = -1; tz <= 1; ++tz) {
    const bool oz = (tz == -1 && uz == 0) || (tz == 1 && uz == D - 1);
    const int vz_in = uz - tz;
    const bool iz = vz_in >= 0 && vz_in < D;
#pragma unroll 1
    for (int ty = -1; ty <= 1; ++ty) {
      const bool oy = (ty == -1 && uy == 0) || (ty == 1 && uy == H - 1);
      const int vy_in = uy - ty;
      const bool iy = vy_in >= 0 && vy_in < H;
#pragma unroll 1
      for (int tx = -1; tx <= 1; ++tx) {
        const bool ox = (tx == -1 && ux == 0) || (tx == 1 && ux == W - 1);
        if (!(oz || oy || ox)) continue;                    // no clamped read reaches u through this tap
        const int vx_in = ux - tx;
        const bool ix = vx_in >= 0 && vx_in < W;
        const int tap = ((tz + 1) * 3 + (ty + 1)) * 3 + (tx + 1);
        const float* wt = w + (uint32_t)(tap * Cout) * (uint32_t)Cin + ci;          // w is taps-major: [27][Cout][Cin]
        // 2 x 2 x 2 choices (inside / clamped) per axis; at least one clamped
#pragma unroll
        for (int cz = 0; cz < 2; ++cz) {
          if (cz ? !oz : !iz) continue;
#pragma unroll
          for (int cy = 0; cy < 2; ++cy) {
            if (cy ? !oy : !iy) continue;
#pragma unroll
            for (int cx = 0; cx < 2; ++cx) {
              if (cx ? !ox : !ix) continue;
              if (!(cz | cy | cx)) continue;                // fully inside: the zero-padded pass has it
              const int vz = cz ? uz : vz_in, vy = cy ? uy : vy_in, vx = cx ? ux : vx_in;
              const float* gp = grad + (long long)(((b * D + vz) * H + vy) * W + vx) * ldg;
              for (int co = 0; co < Cout; co += 4) {
                const float4 gv = ld4(gp + co);
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + co * Cin));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + (co + 1) * Cin));
                const float4 w2 = __ldg(reinterpret_cast<const float4*>(wt + (co + 2) * Cin));
                const float4 w3 = __ldg(reinterpret_cast<const float4*>(wt + (co + 3) * Cin));
                sum.x += gv.x * w0.x + gv.y * w1.x + gv.z * w2.x + gv.w * w3.x;
                sum.y += gv.x * w0.y + gv.y * w1.y + gv.z * w2.y + gv.w * w3.y;
                sum.z += gv.x * w0.z + gv.y * w1.z + gv.z * w2.z + gv.w * w3.z;
                sum.w += gv.x * w0.w + gv.y * w1.w + gv.z * w2.w + gv.w * w3.w;
              }
            }
          }
        }
      }
    }
  }
  float* o = dx + (long long)(((b * D + uz) * H + uy) * W + ux) * ldx + ci;
  float4 cur = ld4(o);
  cur.x += sum.x; cur.y += sum.y; cur.z += sum.z; cur.w += sum.w;
  st4(o, cur);
}

}  // namespace conv
}  // namespace corrif

using namespace corrif;
using namespace corrif::conv;

int corrif_conv_check_desc(const corrif_conv3d_desc& d, const char* what);

extern "C" int corrif_conv3d_wgrad(const corrif_conv3d_desc* desc, const float* g, int64_t ldg, float* dW, void* stream) {
  CORRIF_REQUIRE(desc != nullptr && g != nullptr && dW != nullptr, "conv3d_wgrad: null pointer");
  const corrif_conv3d_desc& d = *desc;
  int rc = corrif_conv_check_desc(d, "conv3d_wgrad");
  if (rc) return rc;
  CORRIF_REQUIRE(((uintptr_t)g % 16) == 0 && ldg % 4 == 0 && ldg >= d.Cout, "conv3d_wgrad: gradient volume unaligned / ld < Cout");
  WgradArgs a;
  for (int i = 0; i < MAX_SRC; ++i) {
    a.src[i].p = i < d.nsrc ? d.src[i].p : nullptr;
    a.src[i].C = i < d.nsrc ? d.src[i].C : 0;
    a.src[i].ld = i < d.nsrc ? d.src[i].ld : 0;
  }
  a.nsrc = d.nsrc;
  a.g = Geom{d.B, d.D, d.H, d.W, (d.W + TX - 1) / TX, (d.H + TY - 1) / TY, (d.D + TZ - 1) / TZ};
  a.Cin = d.Cin; a.Cout = d.Cout; a.replicate = d.pad_mode == CORRIF_PAD_REPLICATE;
  a.grad = g; a.ldg = ldg; a.dW = dW; a.total_tiles = 0;
  // cp.async through L1 (.ca) measured 3-7 % faster than L2-only (.cg) on the 128^3 layers; CORRIF_WGRAD_CG reverts
  static const int via_l1 = getenv("CORRIF_WGRAD_CG") == nullptr;
  a.via_l1 = via_l1;
  const int NB = wgrad_nb(d.Cout);
  static const bool old3 = getenv("CORRIF_WGRAD_V1") != nullptr;      // A/B switch: first 3x3x3 kernel
  if (d.ksize == 3 && !old3) return launch_wgrad3(a, (cudaStream_t)stream);
  if (d.ksize == 3) return launch_nb<3>(NB, a, (cudaStream_t)stream);
  return launch_nb<1>(NB, a, (cudaStream_t)stream);
}

extern "C" int corrif_conv3d_dgrad_border(const float* g, int64_t ldg, const float* w, float* dx, int64_t ldx,
                                          int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                                          void* stream) {
  CORRIF_REQUIRE(g && w && dx, "conv3d_dgrad_border: null pointer");
  CORRIF_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3d_dgrad_border: empty problem");
  const long long nz = D >= 2 ? 2 : 1, nyf = H >= 2 ? 2 : 1, nxf = W >= 2 ? 2 : 1;
  const long long Dm = D - nz, Hm = H - nyf;
  const long long nborder = nz * H * W + Dm * nyf * W + Dm * Hm * nxf;
  const long long threads_total = nborder * B * (Cin / 4);
  CORRIF_REQUIRE(Cout % 4 == 0 && ldg % 4 == 0 && ((uintptr_t)g % 16) == 0, "conv3d_dgrad_border: Cout / ldg must be multiples of 4");
  CORRIF_REQUIRE(Cin % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)dx % 16) == 0 && ((uintptr_t)w % 16) == 0, "conv3d_dgrad_border: Cin / ldx must be multiples of 4, pointers 16-byte aligned");
  CORRIF_REQUIRE(threads_total < (1ll << 31) && (long long)B * D * H * W < (1ll << 31), "conv3d_dgrad_border: problem too large");
  const long long blocks = (threads_total + 255) / 256;
  dgrad_border_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, ldg, w, dx, ldx, B, D, H, W, Cin, Cout,
                                                                       (uint32_t)nborder, (uint32_t)threads_total);
  return launch_status("conv3d_dgrad_border");
}
