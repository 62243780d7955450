// Channels-last 3-D convolution, forward and data gradient (see conv3d.cuh for the design rationale).
//
//   out[b, v, co] = act( bias[co] + sum_{tap, ci} W[co][ci][tap] * x[b, clamp_or_zero(v + tap), ci] )
//
// One CTA = one 8 x 8 x 4 voxel tile (one warp per z-slice) x one block of NB*8 output channels.  Per pass over
// KC input channels the CTA stages the 10 x 10 x 6 input window (replicate / zero padding resolved by the loader:
// no padded copy of the tensor exists) and the packed weights of that pass in shared memory; two or three CTAs
// are resident per SM, so one CTA's staging overlaps its neighbours' math.  Inside a warp the 64 voxels of the
// slice are four m16 blocks (rows g -> x = g of line y = 2*mb, rows g+8 -> line y = 2*mb+1); for a fixed
// (dz, dx, 8-channel step) the ten input lines y-1 .. y+8 are loaded ONCE into registers and serve all three dy
// taps of all four blocks: 20 shared-memory loads feed 12*NB MMAs.
// The epilogue adds the bias, applies ReLU, stores channels-last and reduces the per-(sample, channel) sum and sum
// of squares of what it stored (InstanceNorm statistics) through warp shuffles -> shared -> one double atomic per
// channel and CTA.
#include <stdlib.h>
#include "conv3d.cuh"

namespace corrif {
namespace conv {

struct FwdArgs {
  Src src[MAX_SRC];
  int nsrc;
  Geom g;
  int Cin, Cout, KC;
  int replicate, relu;
  const float* wpk;
  const float* bias;
  float* out;
  long long ldo;
  double* stats;
};

// packed weights: [cout tile][pass][tap][8-channel step][nb][lane][2]
//   lane = g*4 + t holds B[k = t][n = g] and B[k = t + 4][n = g] of the (8 x 8) block
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ wpk, int Cin, int Cout, int taps,
                                    int NB, int KC, int transpose_flip, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // logical problem of the kernel that will consume this: K = input channels, N = output channels
  const int Kc = transpose_flip ? Cout : Cin, Nc = transpose_flip ? Cin : Cout;
  const int steps = KC / 8, passes = (Kc + KC - 1) / KC;
  long long r = i;
  const int j = (int)(r % 2); r /= 2;
  const int lane = (int)(r % 32); r /= 32;
  const int nb = (int)(r % NB); r /= NB;
  const int step = (int)(r % steps); r /= steps;
  const int tap = (int)(r % taps); r /= taps;
  const int pass = (int)(r % passes); r /= passes;
  const int nt = (int)r;
  const int k = pass * KC + step * 8 + (lane & 3) + 4 * j;
  const int n = (nt * NB + nb) * 8 + (lane >> 2);
  float v = 0.f;
  if (k < Kc && n < Nc) {
    if (!transpose_flip) v = w[((long long)n * Cin + k) * taps + tap];
    else v = w[((long long)k * Cin + n) * taps + (taps - 1 - tap)];   // dX = conv(dY, W^T mirrored)
  }
  wpk[i] = round_tf32(v);
}

template <int KS, int NB>
__global__ void __launch_bounds__(NTHREADS, NB == 1 ? 4 : (NB == 2 ? 3 : 2)) conv3d_fwd_kernel(const FwdArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float s_stat[NB * 8 * 2];
  constexpr int TAPS = KS == 3 ? 27 : 1;
  constexpr int CGS = KS == 3 ? CGS3 : CGS1;
  const int KC = a.KC, steps = KC / 8, groups = KC / 4;
  const uint32_t s_in = smem_addr(smem);
  const uint32_t s_w = s_in + groups * CGS;
  const int w_floats = TAPS * steps * NB * 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nt = blockIdx.y;
  const long long nvox = (long long)a.g.D * a.g.H * a.g.W;
  const int per = KS == 3 ? a.g.tiles_x * a.g.tiles_y * a.g.tiles_z : (int)((nvox + TILE_VOX - 1) / TILE_VOX);
  const int total_tiles = per * a.g.B;
  const int passes = (a.Cin + KC - 1) / KC;
  // Persistent over tiles (grid stride): with a single pass over the input channels the packed weights are staged
  // once per CTA and stay in shared memory for all of its tiles.
  for (int tile0 = blockIdx.x; tile0 < total_tiles; tile0 += gridDim.x) {
  // tile coordinates
  int b, z0 = 0, y0 = 0, x0 = 0;
  long long v0 = 0;
  {
    int tile = tile0;
    b = tile / per; tile -= b * per;
    if (KS == 3) {
      x0 = (tile % a.g.tiles_x) * TX; tile /= a.g.tiles_x;
      y0 = (tile % a.g.tiles_y) * TY;
      z0 = (tile / a.g.tiles_y) * TZ;
    } else {
      v0 = (long long)tile * TILE_VOX;
    }
  }

  float acc[4][NB][4];
#pragma unroll
  for (int mb = 0; mb < 4; ++mb)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[mb][nb][j] = 0.f;

  for (int pass = 0; pass < passes; ++pass) {
    __syncthreads();                                   // previous pass / tile fully consumed
    stage_window<KS>(s_in, a.src, a.nsrc, a.g, b, z0, y0, x0, v0, nvox, pass * KC, KC, a.replicate != 0);
    if (passes > 1 || tile0 == (int)blockIdx.x)
      stage_linear(s_w, reinterpret_cast<const float4*>(a.wpk + ((long long)nt * passes + pass) * w_floats), w_floats / 4);
    __syncthreads();
    for (int step = 0; step < steps; ++step) {
      const uint32_t in0 = s_in + (2 * step) * CGS + t * 4;
      if constexpr (KS == 3) {
#pragma unroll 1
        for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            uint32_t A[10][2];
#pragma unroll
            for (int r = 0; r < 10; ++r) {
              const uint32_t ad = in0 + ((((warp + dz) * HY + r) * HX) + g + dx) * 16;
              A[r][0] = lds32(ad);
              A[r][1] = lds32(ad + CGS);
            }
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int tap = (dz * 3 + dy) * 3 + dx;
              uint32_t Bf[NB][2];
#pragma unroll
              for (int nb = 0; nb < NB; ++nb)
                lds64(s_w + (((tap * steps + step) * NB + nb) * 64 + lane * 2) * 4, Bf[nb][0], Bf[nb][1]);
#pragma unroll
              for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
                  mma_tf32(acc[mb][nb], A[2 * mb + dy][0], A[2 * mb + 1 + dy][0], A[2 * mb + dy][1],
                           A[2 * mb + 1 + dy][1], Bf[nb][0], Bf[nb][1]);
            }
          }
        }
      } else {
        uint32_t A[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint32_t ad = in0 + (warp * 64 + r * 8 + g) * 16;
          A[r][0] = lds32(ad);
          A[r][1] = lds32(ad + CGS);
        }
        uint32_t Bf[NB][2];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
          lds64(s_w + ((step * NB + nb) * 64 + lane * 2) * 4, Bf[nb][0], Bf[nb][1]);
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
          for (int nb = 0; nb < NB; ++nb)
            mma_tf32(acc[mb][nb], A[2 * mb][0], A[2 * mb + 1][0], A[2 * mb][1], A[2 * mb + 1][1], Bf[nb][0], Bf[nb][1]);
      }
    }
  }

  // ---- epilogue: bias, ReLU, channels-last store, InstanceNorm statistics ---------------------------
  if (a.stats != nullptr) {
    for (int i = threadIdx.x; i < NB * 16; i += NTHREADS) s_stat[i] = 0.f;
  }
  __syncthreads();
  float cs[NB][2], cq[NB][2];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) { cs[nb][0] = cs[nb][1] = cq[nb][0] = cq[nb][1] = 0.f; }
#pragma unroll
  for (int mb = 0; mb < 4; ++mb) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      long long vox;
      bool ok;
      if (KS == 3) {
        const int z = z0 + warp, y = y0 + 2 * mb + h, x = x0 + g;
        ok = z < a.g.D && y < a.g.H && x < a.g.W;
        vox = (((long long)b * a.g.D + z) * a.g.H + y) * a.g.W + x;
      } else {
        const long long vv = v0 + warp * 64 + mb * 16 + h * 8 + g;
        ok = vv < nvox;
        vox = (long long)b * nvox + vv;
      }
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int co = (nt * NB + nb) * 8 + 2 * t;
        float v0f = acc[mb][nb][2 * h], v1f = acc[mb][nb][2 * h + 1];
        if (a.bias != nullptr) { v0f += __ldg(a.bias + co); v1f += __ldg(a.bias + co + 1); }
        if (a.relu) { v0f = fmaxf(v0f, 0.f); v1f = fmaxf(v1f, 0.f); }
        if (ok) {
          *reinterpret_cast<float2*>(a.out + vox * a.ldo + co) = make_float2(v0f, v1f);
          cs[nb][0] += v0f; cs[nb][1] += v1f;
          cq[nb][0] += v0f * v0f; cq[nb][1] += v1f * v1f;
        }
      }
    }
  }
  if (a.stats != nullptr) {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float s = cs[nb][j], q = cq[nb][j];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {             // over the 8 rows g (lanes with equal t)
          s += __shfl_xor_sync(0xffffffffu, s, o);
          q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g == 0) {
          atomicAdd(&s_stat[(nb * 8 + 2 * t + j) * 2], s);
          atomicAdd(&s_stat[(nb * 8 + 2 * t + j) * 2 + 1], q);
        }
      }
    __syncthreads();
    for (int i = threadIdx.x; i < NB * 16; i += NTHREADS) {
      const int co = nt * NB * 8 + (i >> 1);
      atomicAdd(a.stats + ((long long)b * a.Cout + co) * 2 + (i & 1), (double)s_stat[i]);
    }
  }
  }   // tile loop
}

// ---- 3x3x3, software-pipelined: cp.async double buffering of the window (and of the weights when they change per
// stage) -----------------------------------------------------------------------------------------------------
// The kernel above alternates "stage" and "compute" inside a CTA and relies on the other resident CTAs to cover the
// staging latency; here a CTA walks its sequence of (tile, channel pass) stages with TWO window buffers: the
// cp.async copies of stage s+1 are in flight while the warps run the MMAs of stage s.  The packed weights stay
// resident in shared memory when all passes fit (Cin <= 32 with one output block: the 128^3 / 64^3 layers), else
// they are double-buffered with the window.  Fragments are rounded to TF32 when loaded (one integer add each).
template <int NB>
__global__ void __launch_bounds__(NTHREADS, NB == 1 ? 2 : (NB == 2 ? 3 : 2)) conv3d_fwd3_async_kernel(const FwdArgs a, const int w_resident) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float s_stat[NB * 8 * 2];
  const int KC = a.KC, steps = KC / 8, groups = KC / 4;
  const int w_floats = 27 * steps * NB * 64;
  const int passes = (a.Cin + KC - 1) / KC;
  const uint32_t win_bytes = groups * CGS3;
  const uint32_t s_in0 = smem_addr(smem);
  const uint32_t s_w0 = s_in0 + 2 * win_bytes;       // resident: [pass][w]; else two buffers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nt = blockIdx.y;
  const int per = a.g.tiles_x * a.g.tiles_y * a.g.tiles_z;
  const int total_tiles = per * a.g.B;
  const int my_tiles = total_tiles > (int)blockIdx.x ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nstages = my_tiles * passes;
  const float4* wbase = reinterpret_cast<const float4*>(a.wpk + (long long)nt * passes * w_floats);

  auto tile_coords = [&](int stage, int& b, int& z0, int& y0, int& x0, int& pass) {
    const int ti = stage / passes;
    pass = stage - ti * passes;
    int tile = blockIdx.x + ti * gridDim.x;
    b = tile / per; tile -= b * per;
    x0 = (tile % a.g.tiles_x) * TX; tile /= a.g.tiles_x;
    y0 = (tile % a.g.tiles_y) * TY;
    z0 = (tile / a.g.tiles_y) * TZ;
  };
  auto issue = [&](int stage) {
    int b, z0, y0, x0, pass;
    tile_coords(stage, b, z0, y0, x0, pass);
    stage_window3_async(s_in0 + (stage & 1) * win_bytes, a.src, a.nsrc, a.g, b, z0, y0, x0, pass * KC, KC, a.replicate != 0);
    if (!w_resident) stage_linear_async(s_w0 + (stage & 1) * (w_floats * 4), wbase + (long long)pass * (w_floats / 4), w_floats / 4);
  };

  if (nstages > 0) {
    if (w_resident) stage_linear_async(s_w0, wbase, passes * (w_floats / 4));
    issue(0);
  }
  cp_async_commit();

  float acc[4][NB][4];
  for (int stage = 0; stage < nstages; ++stage) {
    int b, z0, y0, x0, pass;
    tile_coords(stage, b, z0, y0, x0, pass);
    if (stage + 1 < nstages) issue(stage + 1);
    cp_async_commit();
    cp_async_wait<1>();                               // everything but the newest group: stage `stage` has landed
    __syncthreads();
    if (pass == 0) {
#pragma unroll
      for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[mb][nb][j] = 0.f;
    }
    const uint32_t s_in = s_in0 + (stage & 1) * win_bytes;
    const uint32_t s_w = w_resident ? s_w0 + pass * (w_floats * 4) : s_w0 + (stage & 1) * (w_floats * 4);
    for (int step = 0; step < steps; ++step) {
      const uint32_t in0 = s_in + (2 * step) * CGS3 + t * 4;
#pragma unroll 1
      for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          uint32_t A[10][2];
#pragma unroll
          for (int r = 0; r < 10; ++r) {
            const uint32_t ad = in0 + ((((warp + dz) * HY + r) * HX) + g + dx) * 16;
            A[r][0] = rnd_u32(lds32(ad));
            A[r][1] = rnd_u32(lds32(ad + CGS3));
          }
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int tap = (dz * 3 + dy) * 3 + dx;
            uint32_t Bf[NB][2];
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
              lds64(s_w + (((tap * steps + step) * NB + nb) * 64 + lane * 2) * 4, Bf[nb][0], Bf[nb][1]);
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
              for (int nb = 0; nb < NB; ++nb)
                mma_tf32(acc[mb][nb], A[2 * mb + dy][0], A[2 * mb + 1 + dy][0], A[2 * mb + dy][1],
                         A[2 * mb + 1 + dy][1], Bf[nb][0], Bf[nb][1]);
          }
        }
      }
    }
    if (pass == passes - 1) {
      // ---- epilogue of the tile (same as the first kernel) ----
      if (a.stats != nullptr)
        for (int i = threadIdx.x; i < NB * 16; i += NTHREADS) s_stat[i] = 0.f;
      __syncthreads();
      float cs[NB][2], cq[NB][2];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) { cs[nb][0] = cs[nb][1] = cq[nb][0] = cq[nb][1] = 0.f; }
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int z = z0 + warp, y = y0 + 2 * mb + h, x = x0 + g;
          const bool ok = z < a.g.D && y < a.g.H && x < a.g.W;
          const long long vox = ((b * a.g.D + z) * a.g.H + y) * a.g.W + x;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int co = (nt * NB + nb) * 8 + 2 * t;
            float v0f = acc[mb][nb][2 * h], v1f = acc[mb][nb][2 * h + 1];
            if (a.bias != nullptr) { v0f += __ldg(a.bias + co); v1f += __ldg(a.bias + co + 1); }
            if (a.relu) { v0f = fmaxf(v0f, 0.f); v1f = fmaxf(v1f, 0.f); }
            if (ok) {
              *reinterpret_cast<float2*>(a.out + vox * a.ldo + co) = make_float2(v0f, v1f);
              cs[nb][0] += v0f; cs[nb][1] += v1f;
              cq[nb][0] += v0f * v0f; cq[nb][1] += v1f * v1f;
            }
          }
        }
      }
      if (a.stats != nullptr) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float s = cs[nb][j], q = cq[nb][j];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
              s += __shfl_xor_sync(0xffffffffu, s, o);
              q += __shfl_xor_sync(0xffffffffu, q, o);
            }
            if (g == 0) {
              atomicAdd(&s_stat[(nb * 8 + 2 * t + j) * 2], s);
              atomicAdd(&s_stat[(nb * 8 + 2 * t + j) * 2 + 1], q);
            }
          }
        __syncthreads();
        for (int i = threadIdx.x; i < NB * 16; i += NTHREADS) {
          const int co = nt * NB * 8 + (i >> 1);
          atomicAdd(a.stats + ((long long)b * a.Cout + co) * 2 + (i & 1), (double)s_stat[i]);
        }
      }
    }
    __syncthreads();                                  // buffer (stage & 1) may be refilled by the next iteration's issue
  }
  cp_async_wait<0>();
}

template <int NB>
static int launch_async(const FwdArgs& a, cudaStream_t stream) {
  const int steps = a.KC / 8, groups = a.KC / 4;
  const int w_bytes = 27 * steps * NB * 64 * 4;
  const int passes = (a.Cin + a.KC - 1) / a.KC;
  const int w_resident = passes * w_bytes <= 32 * 1024;
  const int smem = 2 * groups * CGS3 + (w_resident ? passes * w_bytes : 2 * w_bytes);
  auto kern = conv3d_fwd3_async_kernel<NB>;
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_last_error("conv3d_fwd(async): smem attribute (%d B): %s", smem, cudaGetErrorString(e)); return (int)e; }
    configured = smem;
  }
  const long long total = (long long)a.g.tiles_x * a.g.tiles_y * a.g.tiles_z * a.g.B;
  const int ntiles = a.Cout / (NB * 8);
  long long gx = (long long)num_sms() * 16 / ntiles;
  gx = gx < 1 ? 1 : (gx > total ? total : gx);
  kern<<<dim3((unsigned)gx, (unsigned)ntiles, 1), NTHREADS, smem, stream>>>(a, w_resident);
  return launch_status("conv3d_fwd(async)");
}

template <int KS, int NB>
static int launch(const FwdArgs& a, cudaStream_t stream) {
  constexpr int TAPS = KS == 3 ? 27 : 1;
  constexpr int CGS = KS == 3 ? CGS3 : CGS1;
  const int smem = (a.KC / 4) * CGS + TAPS * (a.KC / 8) * NB * 64 * 4;
  auto kern = conv3d_fwd_kernel<KS, NB>;
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_last_error("conv3d_fwd: smem attribute (%d B): %s", smem, cudaGetErrorString(e)); return (int)e; }
    configured = smem;
  }
  const long long nvox = (long long)a.g.D * a.g.H * a.g.W;
  const long long per = KS == 3 ? (long long)a.g.tiles_x * a.g.tiles_y * a.g.tiles_z : (nvox + TILE_VOX - 1) / TILE_VOX;
  const long long total = per * a.g.B;
  const int ntiles = a.Cout / (NB * 8);
  // enough CTAs for ~8 rounds of the resident set (2-3 per SM): balances the tail, amortises the weight staging
  long long gx = (long long)num_sms() * 32 / ntiles;
  gx = gx < 1 ? 1 : (gx > total ? total : gx);
  dim3 grid((unsigned)gx, (unsigned)ntiles, 1);
  kern<<<grid, NTHREADS, smem, stream>>>(a);
  return launch_status("conv3d_fwd");
}

template <int KS>
static int launch_nb(int NB, const FwdArgs& a, cudaStream_t s) {
  switch (NB) {
    case 1: return launch<KS, 1>(a, s);
    case 2: return launch<KS, 2>(a, s);
    case 3: return launch<KS, 3>(a, s);
    case 4: return launch<KS, 4>(a, s);
    case 6: return launch<KS, 6>(a, s);
    case 8: return launch<KS, 8>(a, s);
  }
  set_last_error("conv3d_fwd: unsupported channel block %d", NB);
  return CORRIF_EINVAL;
}

}  // namespace conv
}  // namespace corrif

using namespace corrif;
using namespace corrif::conv;

static int check_volume_ptr(const void* p, long long ld, int C, const char* what) {
  CORRIF_REQUIRE(p != nullptr, "%s: null pointer", what);
  CORRIF_REQUIRE(((uintptr_t)p % 16) == 0, "%s: pointer must be 16-byte aligned", what);
  CORRIF_REQUIRE(C > 0 && C % 4 == 0 && ld >= C && ld % 4 == 0, "%s: C (%d) and ld (%lld) must be multiples of 4, ld >= C", what, C, ld);
  return 0;
}

int corrif_conv_check_desc(const corrif_conv3d_desc& d, const char* what) {
  CORRIF_REQUIRE(d.nsrc >= 1 && d.nsrc <= 3, "%s: nsrc must be 1..3", what);
  CORRIF_REQUIRE(d.B > 0 && d.D > 0 && d.H > 0 && d.W > 0, "%s: empty volume", what);
  CORRIF_REQUIRE(d.ksize == 1 || d.ksize == 3, "%s: ksize must be 1 or 3 (got %d)", what, d.ksize);
  CORRIF_REQUIRE(d.pad_mode == CORRIF_PAD_ZEROS || d.pad_mode == CORRIF_PAD_REPLICATE, "%s: pad_mode", what);
  int csum = 0;
  for (int i = 0; i < d.nsrc; ++i) {
    int rc = check_volume_ptr(d.src[i].p, d.src[i].ld, d.src[i].C, what);
    if (rc) return rc;
    csum += d.src[i].C;
  }
  CORRIF_REQUIRE(csum == d.Cin, "%s: Cin (%d) != sum of source channels (%d)", what, d.Cin, csum);
  CORRIF_REQUIRE(d.Cin % 8 == 0 && d.Cout % 8 == 0 && d.Cout > 0, "%s: Cin (%d) and Cout (%d) must be multiples of 8", what, d.Cin, d.Cout);
  CORRIF_REQUIRE((long long)d.B * d.D * d.H * d.W < (1ll << 31), "%s: volume too large", what);
  return 0;
}

extern "C" int corrif_sizeof_conv3d_desc(void) { return (int)sizeof(corrif_conv3d_desc); }

extern "C" int64_t corrif_conv3d_pack_floats(int32_t Cin, int32_t Cout, int32_t ksize) {
  if (Cin <= 0 || Cout <= 0 || Cin % 8 || Cout % 8 || (ksize != 1 && ksize != 3)) return 0;
  // large enough for either orientation (forward: K = Cin, N = Cout; data gradient: swapped)
  const int taps = ksize == 3 ? 27 : 1;
  int64_t best = 0;
  for (int flip = 0; flip < 2; ++flip) {
    const int Kc = flip ? Cout : Cin, Nc = flip ? Cin : Cout;
    const Plan p = conv_plan(Kc, Nc, ksize);
    const int64_t passes = (Kc + p.KC - 1) / p.KC, ntiles = Nc / (p.NB * 8);
    const int64_t n = ntiles * passes * taps * (p.KC / 8) * p.NB * 64;
    best = n > best ? n : best;
  }
  return best;
}

extern "C" int corrif_conv3d_pack_weights(const float* w, float* wpk, int32_t Cin, int32_t Cout, int32_t ksize,
                                          int32_t transpose_flip, void* stream) {
  CORRIF_REQUIRE(w && wpk, "conv3d_pack_weights: null pointer");
  CORRIF_REQUIRE(Cin > 0 && Cout > 0 && Cin % 8 == 0 && Cout % 8 == 0, "conv3d_pack_weights: Cin/Cout must be multiples of 8");
  CORRIF_REQUIRE(ksize == 1 || ksize == 3, "conv3d_pack_weights: ksize must be 1 or 3");
  const int taps = ksize == 3 ? 27 : 1;
  const int Kc = transpose_flip ? Cout : Cin, Nc = transpose_flip ? Cin : Cout;
  const Plan p = conv_plan(Kc, Nc, ksize);
  const long long passes = (Kc + p.KC - 1) / p.KC, ntiles = Nc / (p.NB * 8);
  const long long total = ntiles * passes * taps * (p.KC / 8) * p.NB * 64;
  pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w, wpk, Cin, Cout, taps, p.NB, p.KC, transpose_flip, total);
  return launch_status("conv3d_pack_weights");
}

extern "C" int corrif_conv3d_fwd(const corrif_conv3d_desc* desc, void* stream) {
  CORRIF_REQUIRE(desc != nullptr, "conv3d_fwd: null descriptor");
  const corrif_conv3d_desc& d = *desc;
  int rc = corrif_conv_check_desc(d, "conv3d_fwd");
  if (rc) return rc;
  rc = check_volume_ptr(d.out, d.ldo, d.Cout, "conv3d_fwd(out)");
  if (rc) return rc;
  CORRIF_REQUIRE(d.wpk != nullptr && ((uintptr_t)d.wpk % 16) == 0, "conv3d_fwd: packed weights missing / unaligned");
  const Plan p = conv_plan(d.Cin, d.Cout, d.ksize);
  FwdArgs a;
  for (int i = 0; i < MAX_SRC; ++i) {
    a.src[i].p = i < d.nsrc ? d.src[i].p : nullptr;
    a.src[i].C = i < d.nsrc ? d.src[i].C : 0;
    a.src[i].ld = i < d.nsrc ? d.src[i].ld : 0;
  }
  a.nsrc = d.nsrc;
  a.g = Geom{d.B, d.D, d.H, d.W, (d.W + TX - 1) / TX, (d.H + TY - 1) / TY, (d.D + TZ - 1) / TZ};
  a.Cin = d.Cin; a.Cout = d.Cout; a.KC = p.KC;
  a.replicate = d.pad_mode == CORRIF_PAD_REPLICATE; a.relu = d.relu;
  a.wpk = d.wpk; a.bias = d.bias; a.out = d.out; a.ldo = d.ldo; a.stats = d.stats;
  // A/B switch.  Measured on B200 (tools/conv_bench.py, batch 8): the double-buffered kernel needs 2 x window of shared
  // memory, i.e. 2 CTAs per SM instead of 4, and loses: 32 -> 8 channels at 128^3 2.77 ms vs 2.21 ms, 64 -> 16 at 64^3
  // 0.91 vs 0.76 ms.  With the loader's instruction count fixed the kernel is co-limited by shared-memory fragment
  // loads (1.9 per MMA) and the MMA pipe, which more resident warps overlap better than a deeper copy pipeline.
  static const bool use_async = getenv("CORRIF_CONV_ASYNC") != nullptr;
  if (d.ksize == 3 && p.NB <= 4 && use_async) {
    switch (p.NB) {
      case 1: return launch_async<1>(a, (cudaStream_t)stream);
      case 2: return launch_async<2>(a, (cudaStream_t)stream);
      case 3: return launch_async<3>(a, (cudaStream_t)stream);
      default: return launch_async<4>(a, (cudaStream_t)stream);
    }
  }
  if (d.ksize == 3) return launch_nb<3>(p.NB, a, (cudaStream_t)stream);
  return launch_nb<1>(p.NB, a, (cudaStream_t)stream);
}
