// Element-wise / resampling kernels on channels-last volumes [B, D, H, W, C] (voxel stride ld): InstanceNorm3d after
// ReLU (apply, and the two-pass backward of ReLU -> InstanceNorm), bias-gradient column sums, trilinear
// (align_corners=True) and nearest resize, forward and gather-form backward.  All accesses are 128-bit; a thread
// always works on the same 4 channels, so per-channel statistics live in registers and leave through one block
// reduction + one double atomic per (sample, channel quad element) and block.
#include "common.cuh"

namespace corrif {
namespace vol {

constexpr int THREADS = 256;

__device__ __forceinline__ float4 f4(float v) { return make_float4(v, v, v, v); }

// grid: (blocks over voxels, B).  Every thread owns channel quad q = tid % Q and voxels (tid / Q) + k * (THREADS / Q)...
// requires THREADS % Q == 0 (Q = C / 4 in {2,4,6,8,12,16,24,32,48}: handled by using the largest multiple of Q <= THREADS).
struct Lane { int q; long long v, vstep; bool active; };
__device__ __forceinline__ Lane lane_of(int Q, long long blocks_x) {
  const int usable = (THREADS / Q) * Q;
  Lane l;
  l.active = (int)threadIdx.x < usable;
  l.q = threadIdx.x % Q;
  const int rows = usable / Q;
  l.v = (long long)blockIdx.x * rows + threadIdx.x / Q;
  l.vstep = blocks_x * rows;
  return l;
}

// block-reduce 4 + 4 doubles held per thread for channel quad q and add them to out[(b*C + 4q + e)*2 + {0,1}]
__device__ __forceinline__ void reduce_pairs_to_global(double (&s)[4], double (&t)[4], int Q, int q, bool active,
                                                       double* out, int b, int C) {
  // Few channels (Q a power of two <= 32, every thread active): lanes of a warp that own the same quad differ by
  // multiples of Q, so a butterfly over the offsets Q .. 16 folds a warp, and Q threads then add the 8 warps' partials.
  // The general path below lets Q threads walk all 256 partials serially - for the 8-channel volumes at 128^3 (Q = 2,
  // 16 k blocks) that tail was most of a block's life.
  if (Q <= 32 && (Q & (Q - 1)) == 0 && THREADS % Q == 0) {
    __shared__ double shw[THREADS / 32][32][8];
    double v[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[e] = active ? s[e] : 0.0; v[4 + e] = active ? t[e] : 0.0; }
    for (int o = Q; o < 32; o <<= 1)
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += __shfl_xor_sync(0xffffffffu, v[e], o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < Q)
#pragma unroll
      for (int e = 0; e < 8; ++e) shw[warp][lane][e] = v[e];
    __syncthreads();
    if ((int)threadIdx.x < Q) {
      double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int w = 0; w < THREADS / 32; ++w)
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += shw[w][threadIdx.x][e];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        atomicAdd(out + ((long long)b * C + 4 * threadIdx.x + e) * 2, a[e]);
        atomicAdd(out + ((long long)b * C + 4 * threadIdx.x + e) * 2 + 1, a[4 + e]);
      }
    }
    return;
  }
  __shared__ double sh[THREADS][8];
#pragma unroll
  for (int e = 0; e < 4; ++e) { sh[threadIdx.x][e] = active ? s[e] : 0.0; sh[threadIdx.x][4 + e] = active ? t[e] : 0.0; }
  __syncthreads();
  if ((int)threadIdx.x < Q) {
    const int usable = (THREADS / Q) * Q;
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < usable; i += Q)
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] += sh[i][e];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      atomicAdd(out + ((long long)b * C + 4 * threadIdx.x + e) * 2, a[e]);
      atomicAdd(out + ((long long)b * C + 4 * threadIdx.x + e) * 2 + 1, a[4 + e]);
    }
  }
}

// y = (r - mean) * rstd in place; statistics from the convolution's (sum, sum of squares)
__global__ void __launch_bounds__(THREADS) instnorm_apply_kernel(float* x, long long ld, const double* __restrict__ stats,
                                                                 float* mean_out, float* rstd_out, long long nvox, int C,
                                                                 float eps) {
  const int Q = C / 4, b = blockIdx.y;
  const Lane l = lane_of(Q, gridDim.x);
  if (!l.active) return;
  float mu[4], rs[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const double s = stats[((long long)b * C + 4 * l.q + e) * 2], ss = stats[((long long)b * C + 4 * l.q + e) * 2 + 1];
    const double m = s / (double)nvox;
    double var = ss / (double)nvox - m * m;
    var = var > 0.0 ? var : 0.0;
    mu[e] = (float)m;
    rs[e] = (float)(1.0 / sqrt(var + (double)eps));
  }
  if (blockIdx.x == 0 && threadIdx.x < Q) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mean_out[(long long)b * C + 4 * l.q + e] = mu[e];
      rstd_out[(long long)b * C + 4 * l.q + e] = rs[e];
    }
  }
  float* base = x + (long long)b * nvox * ld + 4 * l.q;
  for (long long v = l.v; v < nvox; v += 4 * l.vstep) {        // 4 independent loads in flight per thread
    float4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v + u * l.vstep < nvox) r[u] = ld4(base + (v + u * l.vstep) * ld);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v + u * l.vstep < nvox) {
        r[u].x = (r[u].x - mu[0]) * rs[0]; r[u].y = (r[u].y - mu[1]) * rs[1];
        r[u].z = (r[u].z - mu[2]) * rs[2]; r[u].w = (r[u].w - mu[3]) * rs[3];
        st4(base + (v + u * l.vstep) * ld, r[u]);
      }
  }
}

// sums[b][c] += (sum dy, sum dy * y)
__global__ void __launch_bounds__(THREADS) instnorm_bwd_stats_kernel(const float* __restrict__ dy, long long lddy,
                                                                     const float* __restrict__ y, long long ldy,
                                                                     double* sums, long long nvox, int C) {
  const int Q = C / 4, b = blockIdx.y;
  const Lane l = lane_of(Q, gridDim.x);
  double s[4] = {0, 0, 0, 0}, t[4] = {0, 0, 0, 0};
  if (l.active) {
    const float* pd = dy + (long long)b * nvox * lddy + 4 * l.q;
    const float* py = y + (long long)b * nvox * ldy + 4 * l.q;
    float fs[4] = {0, 0, 0, 0}, ft[4] = {0, 0, 0, 0};
    int n = 0;
#pragma unroll 4
    for (long long v = l.v; v < nvox; v += l.vstep) {
      const float4 d = ld4_stream(pd + v * lddy), yy = ld4_stream(py + v * ldy);
      fs[0] += d.x; fs[1] += d.y; fs[2] += d.z; fs[3] += d.w;
      ft[0] += d.x * yy.x; ft[1] += d.y * yy.y; ft[2] += d.z * yy.z; ft[3] += d.w * yy.w;
      if (++n == 64) {                      // flush the fp32 partials into doubles every 64 voxels
#pragma unroll
        for (int e = 0; e < 4; ++e) { s[e] += fs[e]; t[e] += ft[e]; fs[e] = ft[e] = 0.f; }
        n = 0;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { s[e] += fs[e]; t[e] += ft[e]; }
  }
  reduce_pairs_to_global(s, t, Q, l.q, l.active, sums, b, C);
}

// g = [r > 0] * rstd * (dy - sum_dy/n - y * sum_dyy/n);  dbias[c] += sum g.   relu == 0: no mask.
// mean == nullptr: no normalisation at all (plain ReLU backward is not needed by the model; rejected by the launcher)
__global__ void __launch_bounds__(THREADS) instnorm_relu_bwd_apply_kernel(
    const float* __restrict__ dy, long long lddy, const float* __restrict__ y, long long ldy,
    const float* __restrict__ mean, const float* __restrict__ rstd, const double* __restrict__ sums, float* g,
    long long ldg, float* dbias, long long nvox, int C, int relu) {
  const int Q = C / 4, b = blockIdx.y;
  const Lane l = lane_of(Q, gridDim.x);
  float acc[4] = {0, 0, 0, 0};
  if (l.active) {
    float rs[4], thr[4], m1[4], m2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long ch = (long long)b * C + 4 * l.q + e;
      rs[e] = rstd[ch];
      thr[e] = (0.f - mean[ch]) * rs[e];           // the forward's value of a clamped-to-zero activation
      m1[e] = (float)(sums[ch * 2] / (double)nvox);
      m2[e] = (float)(sums[ch * 2 + 1] / (double)nvox);
    }
    const float* pd = dy + (long long)b * nvox * lddy + 4 * l.q;
    const float* py = y + (long long)b * nvox * ldy + 4 * l.q;
    float* pg = g + (long long)b * nvox * ldg + 4 * l.q;
    for (long long v0 = l.v; v0 < nvox; v0 += 2 * l.vstep) {      // two voxels per iteration: 4 loads in flight
      float4 d[2], yy[2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (v0 + u * l.vstep < nvox) { d[u] = ld4(pd + (v0 + u * l.vstep) * lddy); yy[u] = ld4_stream(py + (v0 + u * l.vstep) * ldy); }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long v = v0 + u * l.vstep;
        if (v >= nvox) continue;
        float4 o;
        o.x = rs[0] * (d[u].x - m1[0] - yy[u].x * m2[0]); o.y = rs[1] * (d[u].y - m1[1] - yy[u].y * m2[1]);
        o.z = rs[2] * (d[u].z - m1[2] - yy[u].z * m2[2]); o.w = rs[3] * (d[u].w - m1[3] - yy[u].w * m2[3]);
        if (relu) {
          o.x = yy[u].x > thr[0] ? o.x : 0.f; o.y = yy[u].y > thr[1] ? o.y : 0.f;
          o.z = yy[u].z > thr[2] ? o.z : 0.f; o.w = yy[u].w > thr[3] ? o.w : 0.f;
        }
        st4(pg + v * ldg, o);
        acc[0] += o.x; acc[1] += o.y; acc[2] += o.z; acc[3] += o.w;
      }
    }
  }
  if (dbias != nullptr) {
    __shared__ float sh[THREADS][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sh[threadIdx.x][e] = l.active ? acc[e] : 0.f;
    __syncthreads();
    if ((int)threadIdx.x < Q) {
      const int usable = (THREADS / Q) * Q;
      float a[4] = {0, 0, 0, 0};
      for (int i = threadIdx.x; i < usable; i += Q)
#pragma unroll
        for (int e = 0; e < 4; ++e) a[e] += sh[i][e];
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(dbias + 4 * threadIdx.x + e, a[e]);
    }
  }
}

// ---- train-mode BatchNorm3d (+ residual add) (+ ReLU) on a channels-last volume --------------------------------
// The encoders' Bottleneck3D (mmvit4.py:196-212): conv -> BN -> ReLU, and conv -> BN -> (+identity) -> ReLU.  Statistics
// are per channel over ALL samples and voxels (the volume is treated as one sample of B*D*H*W rows).  Forward: the
// per-channel (sum, sum of squares) come from instnorm_bwd_stats_kernel(x, x); this kernel normalises, scales, adds the
// residual, clamps, and writes mean / biased variance / rstd.  Backward: g = dy * [y > 0]; sums = (sum g, sum g * xhat)
// (= d beta, d gamma); dx = gamma * rstd * (g - sum_g / n - xhat * sum_gx / n); d residual = g.
__global__ void __launch_bounds__(THREADS) bn_apply_kernel(const float* __restrict__ x, long long ldx,
                                                           const double* __restrict__ stats, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ res,
                                                           long long ldr, float* __restrict__ y, long long ldy, float* mean_out,
                                                           float* var_out, float* rstd_out, long long rows, int C, float eps,
                                                           int relu) {
  const int Q = C / 4;
  const Lane l = lane_of(Q, gridDim.x);
  if (!l.active) return;
  float sc[4], sh[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = 4 * l.q + e;
    const double m = stats[c * 2] / (double)rows;
    double var = stats[c * 2 + 1] / (double)rows - m * m;
    var = var > 0.0 ? var : 0.0;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    if (blockIdx.x == 0 && (int)threadIdx.x < Q) { mean_out[c] = (float)m; var_out[c] = (float)var; rstd_out[c] = rs; }
    sc[e] = rs * gamma[c];
    sh[e] = beta[c] - (float)m * sc[e];
  }
  for (long long v = l.v; v < rows; v += 2 * l.vstep) {
    float4 a[2], r[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long vv = v + u * l.vstep;
      if (vv < rows) {
        a[u] = ld4_stream(x + vv * ldx + 4 * l.q);
        r[u] = res != nullptr ? ld4_stream(res + vv * ldr + 4 * l.q) : f4(0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long vv = v + u * l.vstep;
      if (vv >= rows) continue;
      float4 o;
      o.x = a[u].x * sc[0] + sh[0] + r[u].x; o.y = a[u].y * sc[1] + sh[1] + r[u].y;
      o.z = a[u].z * sc[2] + sh[2] + r[u].z; o.w = a[u].w * sc[3] + sh[3] + r[u].w;
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      st4(y + vv * ldy + 4 * l.q, o);
    }
  }
}

__global__ void __launch_bounds__(THREADS) bn_bwd_stats_kernel(const float* __restrict__ dy, long long lddy,
                                                               const float* __restrict__ y, long long ldy,
                                                               const float* __restrict__ x, long long ldx,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               double* sums, long long rows, int C, int relu) {
  const int Q = C / 4;
  const Lane l = lane_of(Q, gridDim.x);
  double s[4] = {0, 0, 0, 0}, t[4] = {0, 0, 0, 0};
  if (l.active) {
    float mu[4], rs[4], fs[4] = {0, 0, 0, 0}, ft[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 4; ++e) { mu[e] = mean[4 * l.q + e]; rs[e] = rstd[4 * l.q + e]; }
    int n = 0;
#pragma unroll 2
    for (long long v = l.v; v < rows; v += l.vstep) {
      float4 g = ld4_stream(dy + v * lddy + 4 * l.q);
      const float4 xx = ld4_stream(x + v * ldx + 4 * l.q);
      if (relu) {
        const float4 yy = ld4_stream(y + v * ldy + 4 * l.q);
        g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f; g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
      }
      fs[0] += g.x; fs[1] += g.y; fs[2] += g.z; fs[3] += g.w;
      ft[0] += g.x * (xx.x - mu[0]) * rs[0]; ft[1] += g.y * (xx.y - mu[1]) * rs[1];
      ft[2] += g.z * (xx.z - mu[2]) * rs[2]; ft[3] += g.w * (xx.w - mu[3]) * rs[3];
      if (++n == 64) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { s[e] += fs[e]; t[e] += ft[e]; fs[e] = ft[e] = 0.f; }
        n = 0;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { s[e] += fs[e]; t[e] += ft[e]; }
  }
  reduce_pairs_to_global(s, t, Q, l.q, l.active, sums, 0, C);
}

__global__ void __launch_bounds__(THREADS) bn_bwd_apply_kernel(const float* __restrict__ dy, long long lddy,
                                                               const float* __restrict__ y, long long ldy,
                                                               const float* __restrict__ x, long long ldx,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma, const double* __restrict__ sums,
                                                               float* __restrict__ dx, long long lddx, float* __restrict__ dres,
                                                               long long lddr, float* dgamma, float* dbeta, long long rows, int C,
                                                               int relu) {
  const int Q = C / 4;
  const Lane l = lane_of(Q, gridDim.x);
  if (!l.active) return;
  float mu[4], rs[4], k[4], m1[4], m2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = 4 * l.q + e;
    mu[e] = mean[c]; rs[e] = rstd[c]; k[e] = gamma[c] * rs[e];
    m1[e] = (float)(sums[c * 2] / (double)rows);
    m2[e] = (float)(sums[c * 2 + 1] / (double)rows);
    if (blockIdx.x == 0 && (int)threadIdx.x < Q) { dbeta[c] = (float)sums[c * 2]; dgamma[c] = (float)sums[c * 2 + 1]; }
  }
  for (long long v = l.v; v < rows; v += l.vstep) {
    float4 g = ld4_stream(dy + v * lddy + 4 * l.q);
    const float4 xx = ld4_stream(x + v * ldx + 4 * l.q);
    if (relu) {
      const float4 yy = ld4_stream(y + v * ldy + 4 * l.q);
      g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f; g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
    }
    if (dres != nullptr) st4(dres + v * lddr + 4 * l.q, g);
    float4 o;
    o.x = k[0] * (g.x - m1[0] - (xx.x - mu[0]) * rs[0] * m2[0]); o.y = k[1] * (g.y - m1[1] - (xx.y - mu[1]) * rs[1] * m2[1]);
    o.z = k[2] * (g.z - m1[2] - (xx.z - mu[2]) * rs[2] * m2[2]); o.w = k[3] * (g.w - m1[3] - (xx.w - mu[3]) * rs[3] * m2[3]);
    st4(dx + v * lddx + 4 * l.q, o);
  }
}

__global__ void __launch_bounds__(THREADS) colsum_kernel(const float* __restrict__ g, long long ldg, float* dbias,
                                                         long long rows, int C) {
  const int Q = C / 4;
  const Lane l = lane_of(Q, gridDim.x);
  float acc[4] = {0, 0, 0, 0};
  if (l.active)
    for (long long v = l.v; v < rows; v += l.vstep) {
      const float4 d = ld4_stream(g + v * ldg + 4 * l.q);
      acc[0] += d.x; acc[1] += d.y; acc[2] += d.z; acc[3] += d.w;
    }
  __shared__ float sh[THREADS][4];
#pragma unroll
  for (int e = 0; e < 4; ++e) sh[threadIdx.x][e] = l.active ? acc[e] : 0.f;
  __syncthreads();
  if ((int)threadIdx.x < Q) {
    const int usable = (THREADS / Q) * Q;
    float a[4] = {0, 0, 0, 0};
    for (int i = threadIdx.x; i < usable; i += Q)
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] += sh[i][e];
#pragma unroll
    for (int e = 0; e < 4; ++e) atomicAdd(dbias + 4 * threadIdx.x + e, a[e]);
  }
}

// ---- resampling --------------------------------------------------------------------------------------------
// torch's area_pixel_compute_source_index for align_corners=True: src = dst * (in - 1) / (out - 1) (0 if out == 1)
__device__ __forceinline__ float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }
__device__ __forceinline__ void ac_src(int dst, float scale, int in, int& i0, int& i1, float& w1) {
  const float s = scale * (float)dst;
  i0 = (int)s;
  i0 = i0 > in - 1 ? in - 1 : i0;
  i1 = i0 < in - 1 ? i0 + 1 : i0;
  w1 = s - (float)i0;
}

__global__ void __launch_bounds__(THREADS) trilinear_fwd_kernel(const float* __restrict__ x, long long ldx, float* y,
                                                                long long ldy, int C, int Di, int Hi, int Wi, int Do,
                                                                int Ho, int Wo, long long total) {
  const int Q = C / 4;
  const float sz = ac_scale(Di, Do), sy = ac_scale(Hi, Ho), sx = ac_scale(Wi, Wo);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % Q);
    long long v = i / Q;
    const int xo = (int)(v % Wo); v /= Wo;
    const int yo = (int)(v % Ho); v /= Ho;
    const int zo = (int)(v % Do);
    const long long b = v / Do;
    int z0, z1, y0, y1, x0, x1;
    float wz, wy, wx;
    ac_src(zo, sz, Di, z0, z1, wz); ac_src(yo, sy, Hi, y0, y1, wy); ac_src(xo, sx, Wi, x0, x1, wx);
    const float* base = x + 4 * q;
    auto at = [&](int z, int yy, int xx) { return ld4(base + ((((long long)b * Di + z) * Hi + yy) * Wi + xx) * ldx); };
    float4 o = f4(0.f);
    auto fma4 = [&](float w, float4 a) { o.x += w * a.x; o.y += w * a.y; o.z += w * a.z; o.w += w * a.w; };
    fma4((1 - wz) * (1 - wy) * (1 - wx), at(z0, y0, x0)); fma4((1 - wz) * (1 - wy) * wx, at(z0, y0, x1));
    fma4((1 - wz) * wy * (1 - wx), at(z0, y1, x0));       fma4((1 - wz) * wy * wx, at(z0, y1, x1));
    fma4(wz * (1 - wy) * (1 - wx), at(z1, y0, x0));       fma4(wz * (1 - wy) * wx, at(z1, y0, x1));
    fma4(wz * wy * (1 - wx), at(z1, y1, x0));             fma4(wz * wy * wx, at(z1, y1, x1));
    st4(y + ((((long long)b * Do + zo) * Ho + yo) * Wo + xo) * ldy + 4 * q, o);
  }
}

// gather-form backward: input index i receives from outputs o with floor(scale*o) in {i-1, i}; the candidate range
// is bracketed from the inverse map and every candidate re-evaluates the forward's own index / weight arithmetic
__device__ __forceinline__ void ac_range(int i, float scale, int in, int out, int& lo, int& hi) {
  if (scale <= 0.f) { lo = 0; hi = out - 1; return; }         // out == 1 or in == 1: everything reads index 0
  const float inv = 1.f / scale;
  lo = (int)floorf((float)(i - 1) * inv) - 1;
  hi = (int)ceilf((float)(i + 1) * inv) + 1;
  lo = lo < 0 ? 0 : lo;
  hi = hi > out - 1 ? out - 1 : hi;
}
__device__ __forceinline__ float ac_weight(int o, int i, float scale, int in) {
  int i0, i1;
  float w1;
  ac_src(o, scale, in, i0, i1, w1);
  float w = 0.f;
  if (i0 == i) w += 1.f - w1;
  if (i1 == i) w += w1;                                         // i0 == i1 at the last index: weights add to 1
  return w;
}

// per-axis candidate list of one input index: first output index and up to MAXC weights (zero = no contribution)
constexpr int MAXC = 8;
__device__ __forceinline__ int ac_candidates(int i, float scale, int in, int out, float (&w)[MAXC]) {
  int lo, hi;
  ac_range(i, scale, in, out, lo, hi);
  // trim leading non-contributors so that the window of MAXC candidates starts at the first real one
  while (lo < hi && ac_weight(lo, i, scale, in) == 0.f) ++lo;
#pragma unroll
  for (int k = 0; k < MAXC; ++k) w[k] = (lo + k <= hi) ? ac_weight(lo + k, i, scale, in) : 0.f;
  return lo;
}

// Gather-form backward.  With more than MAXC contributing outputs per axis (strong down-sampling, e.g. the
// encoder's 64 -> 8 pooling has at most 2 / scale = 18) the generic loop below is used instead.
__global__ void __launch_bounds__(THREADS) trilinear_bwd_kernel(const float* __restrict__ dy, long long lddy, float* dx,
                                                                long long lddx, int C, int Di, int Hi, int Wi, int Do,
                                                                int Ho, int Wo, long long total, int small) {
  const int Q = C / 4;
  const float sz = ac_scale(Di, Do), sy = ac_scale(Hi, Ho), sx = ac_scale(Wi, Wo);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % Q);
    long long v = i / Q;
    const int xi = (int)(v % Wi); v /= Wi;
    const int yi = (int)(v % Hi); v /= Hi;
    const int zi = (int)(v % Di);
    const long long b = v / Di;
    float4 o = f4(0.f);
    if (small) {
      float wz[MAXC], wy[MAXC], wx[MAXC];
      const int zl = ac_candidates(zi, sz, Di, Do, wz), yl = ac_candidates(yi, sy, Hi, Ho, wy),
                xl = ac_candidates(xi, sx, Wi, Wo, wx);
#pragma unroll
      for (int a = 0; a < MAXC; ++a) {
        if (wz[a] == 0.f) continue;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          if (wy[c] == 0.f) continue;
          const float wzy = wz[a] * wy[c];
          const float* row = dy + ((((long long)b * Do + zl + a) * Ho + yl + c) * Wo + xl) * lddy + 4 * q;
#pragma unroll
          for (int e = 0; e < MAXC; ++e) {
            if (wx[e] == 0.f) continue;
            const float4 d = ld4(row + (long long)e * lddy);
            const float w = wzy * wx[e];
            o.x += w * d.x; o.y += w * d.y; o.z += w * d.z; o.w += w * d.w;
          }
        }
      }
    } else {
      int zl, zh, yl, yh, xl, xh;
      ac_range(zi, sz, Di, Do, zl, zh); ac_range(yi, sy, Hi, Ho, yl, yh); ac_range(xi, sx, Wi, Wo, xl, xh);
      for (int zo = zl; zo <= zh; ++zo) {
        const float wz = ac_weight(zo, zi, sz, Di);
        if (wz == 0.f) continue;
        for (int yo = yl; yo <= yh; ++yo) {
          const float wy = ac_weight(yo, yi, sy, Hi);
          if (wy == 0.f) continue;
          for (int xo = xl; xo <= xh; ++xo) {
            const float wx = ac_weight(xo, xi, sx, Wi);
            if (wx == 0.f) continue;
            const float4 d = ld4(dy + ((((long long)b * Do + zo) * Ho + yo) * Wo + xo) * lddy + 4 * q);
            const float w = wz * wy * wx;
            o.x += w * d.x; o.y += w * d.y; o.z += w * d.z; o.w += w * d.w;
          }
        }
      }
    }
    st4(dx + ((((long long)b * Di + zi) * Hi + yi) * Wi + xi) * lddx + 4 * q, o);
  }
}

// ---- separable form of the trilinear resize: one axis at a time --------------------------------------------
// A tensor viewed as [outer][n][inner4] float4s (inner4 = everything after the axis, channels included) is resized
// along n with the same align_corners index / weight arithmetic.  Trilinear interpolation is the product of three such
// passes; for the decoder's 2x up-sampling to 128^3 (mmvit4.py:269) three streaming passes (every element read once
// per pass, coalesced) replace 8 gathered loads per output in the forward and 64 per input in the gather-form
// backward: 0.83 -> ~0.5 ms and 1.54 -> ~0.5 ms for the 16-channel level.
__global__ void __launch_bounds__(THREADS) linear_axis_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y,
                                                                  unsigned n_in, unsigned n_out, unsigned inner4,
                                                                  unsigned total) {
  const float sc = ac_scale((int)n_in, (int)n_out);
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned row = idx / inner4, i = idx - row * inner4;
    const unsigned o = row / n_out, j = row - o * n_out;
    int i0, i1;
    float w;
    ac_src((int)j, sc, (int)n_in, i0, i1, w);
    const float4 a = x[((size_t)o * n_in + i0) * inner4 + i], b = x[((size_t)o * n_in + i1) * inner4 + i];
    const float u = 1.f - w;
    y[idx] = make_float4(u * a.x + w * b.x, u * a.y + w * b.y, u * a.z + w * b.z, u * a.w + w * b.w);
  }
}
// adjoint of the pass above in gather form: input index i sums its (at most MAXC) contributing outputs
__global__ void __launch_bounds__(THREADS) linear_axis_bwd_kernel(const float4* __restrict__ dy, float4* __restrict__ dx,
                                                                  unsigned n_in, unsigned n_out, unsigned inner4,
                                                                  unsigned total) {
  const float sc = ac_scale((int)n_in, (int)n_out);
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned row = idx / inner4, i = idx - row * inner4;
    const unsigned o = row / n_in, k = row - o * n_in;
    float w[MAXC];
    const int lo = ac_candidates((int)k, sc, (int)n_in, (int)n_out, w);
    float4 acc = f4(0.f);
    const float4* p = dy + ((size_t)o * n_out + lo) * inner4 + i;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (w[c] == 0.f) continue;
      const float4 d = p[(size_t)c * inner4];
      acc.x += w[c] * d.x; acc.y += w[c] * d.y; acc.z += w[c] * d.z; acc.w += w[c] * d.w;
    }
    dx[idx] = acc;
  }
}

// the same two passes for a long inner extent (the z and y axes): a block walks whole rows, so the index / weight
// arithmetic of a row is done once per thread and row instead of once per float4
__global__ void __launch_bounds__(THREADS) linear_axis_fwd_rows_kernel(const float4* __restrict__ x, float4* __restrict__ y,
                                                                       unsigned n_in, unsigned n_out, unsigned inner4,
                                                                       unsigned rows, unsigned split) {
  const float sc = ac_scale((int)n_in, (int)n_out);
  // a row is cut into `split` pieces so that small row counts still fill the GPU
  for (unsigned w = blockIdx.x; w < rows * split; w += gridDim.x) {
    const unsigned row = w / split, piece = w - row * split;
    const unsigned o = row / n_out, j = row - o * n_out;
    int i0, i1;
    float wt;
    ac_src((int)j, sc, (int)n_in, i0, i1, wt);
    const float u = 1.f - wt;
    const float4* a = x + ((size_t)o * n_in + i0) * inner4;
    const float4* b = x + ((size_t)o * n_in + i1) * inner4;
    float4* d = y + (size_t)row * inner4;
    const unsigned lo = (unsigned)(((unsigned long long)inner4 * piece) / split), hi = (unsigned)(((unsigned long long)inner4 * (piece + 1)) / split);
    for (unsigned i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const float4 p = a[i], q = b[i];
      d[i] = make_float4(u * p.x + wt * q.x, u * p.y + wt * q.y, u * p.z + wt * q.z, u * p.w + wt * q.w);
    }
  }
}
__global__ void __launch_bounds__(THREADS) linear_axis_bwd_rows_kernel(const float4* __restrict__ dy, float4* __restrict__ dx,
                                                                       unsigned n_in, unsigned n_out, unsigned inner4,
                                                                       unsigned rows, unsigned split) {
  const float sc = ac_scale((int)n_in, (int)n_out);
  for (unsigned w = blockIdx.x; w < rows * split; w += gridDim.x) {
    const unsigned row = w / split, piece = w - row * split;
    const unsigned o = row / n_in, k = row - o * n_in;
    float wt[MAXC];
    const int c0 = ac_candidates((int)k, sc, (int)n_in, (int)n_out, wt);
    const float4* p = dy + ((size_t)o * n_out + c0) * inner4;
    float4* d = dx + (size_t)row * inner4;
    const unsigned lo = (unsigned)(((unsigned long long)inner4 * piece) / split), hi = (unsigned)(((unsigned long long)inner4 * (piece + 1)) / split);
    for (unsigned i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      float4 acc = f4(0.f);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (wt[c] == 0.f) continue;
        const float4 v = p[(size_t)c * inner4 + i];
        acc.x += wt[c] * v.x; acc.y += wt[c] * v.y; acc.z += wt[c] * v.z; acc.w += wt[c] * v.w;
      }
      d[i] = acc;
    }
  }
}

// torch 'nearest': src = min(floor(dst * (in / out)), in - 1) in float arithmetic
__device__ __forceinline__ int nn_src(int dst, float scale, int in) {
  const int s = (int)floorf((float)dst * scale);
  return s < in - 1 ? s : in - 1;
}
__global__ void __launch_bounds__(THREADS) nearest_fwd_kernel(const float* __restrict__ x, long long ldx, float* y,
                                                              long long ldy, int C, int Di, int Hi, int Wi, int Do,
                                                              int Ho, int Wo, long long total) {
  const int Q = C / 4;
  const float sz = (float)Di / (float)Do, sy = (float)Hi / (float)Ho, sx = (float)Wi / (float)Wo;
  // one output row (b, zo, yo) per block iteration: the row's source row is resolved once, and the threads of the block
  // walk its Wo * Q float4s with 32-bit arithmetic (the flat 64-bit div / mod chain per element made this write-bound
  // broadcast run at 2.3 TB/s)
  const long long rows = total / ((long long)Wo * Q);
  const unsigned per_row = (unsigned)Wo * (unsigned)Q;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int yo = (int)(row % Ho);
    const long long t = row / Ho;
    const int zo = (int)(t % Do);
    const long long b = t / Do;
    const float* srow = x + (((long long)b * Di + nn_src(zo, sz, Di)) * Hi + nn_src(yo, sy, Hi)) * Wi * ldx;
    float* drow = y + row * Wo * ldy;
    for (unsigned e = threadIdx.x; e < per_row; e += blockDim.x) {
      const unsigned xo = e / (unsigned)Q, q = e - xo * (unsigned)Q;
      st4(drow + (long long)xo * ldy + 4 * q, ld4(srow + (long long)nn_src((int)xo, sx, Wi) * ldx + 4 * q));
    }
  }
}
__device__ __forceinline__ void nn_range(int i, float scale, int in, int out, int& lo, int& hi) {
  const float inv = 1.f / scale;
  lo = (int)floorf((float)i * inv) - 1;
  hi = (int)ceilf((float)(i + 1) * inv) + 1;
  lo = lo < 0 ? 0 : lo;
  hi = hi > out - 1 ? out - 1 : hi;
  while (lo <= hi && nn_src(lo, scale, in) != i) ++lo;
  while (hi >= lo && nn_src(hi, scale, in) != i) --hi;
}
// One warp per input voxel.  Lanes are (slot, quad): `slots` contributing output voxels are read side by side, each
// with its Q = C/4 quads by consecutive lanes, so a warp load covers slots x (C*4 contiguous bytes) - the first version
// (one warp per quad, lanes over voxels) fetched 16 bytes out of every 32-byte sector.  Partial sums of the slots are
// folded with shuffles.  C > 128 loops over groups of 32 quads.
__global__ void __launch_bounds__(THREADS) nearest_bwd_kernel(const float* __restrict__ dy, long long lddy, float* dx,
                                                              long long lddx, int C, int Di, int Hi, int Wi, int Do,
                                                              int Ho, int Wo, long long total_warps) {
  const int Q = C / 4, lane = threadIdx.x & 31;
  const float sz = (float)Di / (float)Do, sy = (float)Hi / (float)Ho, sx = (float)Wi / (float)Wo;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= total_warps) return;
  long long v = wid;
  const int xi = (int)(v % Wi); v /= Wi;
  const int yi = (int)(v % Hi); v /= Hi;
  const int zi = (int)(v % Di);
  const long long b = v / Di;
  int zl, zh, yl, yh, xl, xh;
  nn_range(zi, sz, Di, Do, zl, zh); nn_range(yi, sy, Hi, Ho, yl, yh); nn_range(xi, sx, Wi, Wo, xl, xh);
  const int nz = zh - zl + 1, ny = yh - yl + 1, nx = xh - xl + 1;
  const int n = (nz > 0 && ny > 0 && nx > 0) ? nz * ny * nx : 0;
  float* o = dx + ((((long long)b * Di + zi) * Hi + yi) * Wi + xi) * lddx;
  for (int q0 = 0; q0 < Q; q0 += 32) {
    const int qn = Q - q0 < 32 ? Q - q0 : 32;            // quads handled in this round
    const int slots = 32 / qn;                           // output voxels read side by side
    const int slot = lane / qn, q = q0 + lane - slot * qn;
    float4 acc = f4(0.f);
    if (slot < slots)
      for (int j = slot; j < n; j += slots) {
        const int xo = xl + j % nx, yo = yl + (j / nx) % ny, zo = zl + j / (nx * ny);
        const float4 d = ld4_stream(dy + ((((long long)b * Do + zo) * Ho + yo) * Wo + xo) * lddy + 4 * q);
        acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
      }
    for (int k = 1; k < slots; ++k) {                    // fold slot k onto slot 0 (lanes 0 .. qn-1)
      const int srcl = lane + k * qn;
      const float ax = __shfl_sync(0xffffffffu, acc.x, srcl & 31), ay = __shfl_sync(0xffffffffu, acc.y, srcl & 31),
                  az = __shfl_sync(0xffffffffu, acc.z, srcl & 31), aw = __shfl_sync(0xffffffffu, acc.w, srcl & 31);
      if (lane < qn) { acc.x += ax; acc.y += ay; acc.z += az; acc.w += aw; }
    }
    if (lane < qn) st4(o + 4 * q, acc);
  }
}

// ---- 1x1x1 convolution with C = 8 or 16 channels on both sides (d1_out / d2_out, mmvit4.py:231-236) -----------------
// 8 -> 8 channels at 128^3 is 64 FMAs per 64 bytes moved: HBM-bound by a wide margin, and the window-staging tensor-core
// kernel ran it at a third of the bandwidth (0.45 ms against a 0.17 ms floor).  One thread per voxel: the voxel's C
// inputs in registers, the C x C weights broadcast from shared memory, exact fp32 FMAs; the InstanceNorm statistics of
// the stored output are kept per thread and leave through one block reduction.  The data gradient is the same kernel
// with the transposed weight.
template <int C>
__global__ void __launch_bounds__(THREADS) conv1_small_kernel(const float* __restrict__ x, long long ldx,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              float* __restrict__ out, long long ldo, double* stats,
                                                              long long nvox, int relu, int transpose) {
  __shared__ __align__(16) float sw[C][C];              // sw[ci][co]
  __shared__ float sb[C];
  __shared__ double sred[2][C];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
    const int ci = i / C, co = i - ci * C;
    sw[ci][co] = transpose ? w[ci * C + co] : w[co * C + ci];     // w is [Cout][Cin]; transposed: out index = ci of w
  }
  if (threadIdx.x < C) { sb[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f; sred[0][threadIdx.x] = 0.0; sred[1][threadIdx.x] = 0.0; }
  __syncthreads();
  const int b = blockIdx.y;
  float s1[C], s2[C];
#pragma unroll
  for (int j = 0; j < C; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
    const float* xp = x + ((long long)b * nvox + v) * ldx;
    float xi[C], o[C];
#pragma unroll
    for (int j = 0; j < C; j += 4) { const float4 t = ld4(xp + j); xi[j] = t.x; xi[j + 1] = t.y; xi[j + 2] = t.z; xi[j + 3] = t.w; }
#pragma unroll
    for (int j = 0; j < C; ++j) o[j] = sb[j];
#pragma unroll
    for (int ci = 0; ci < C; ++ci) {
#pragma unroll
      for (int j = 0; j < C; j += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(&sw[ci][j]);
        o[j] = fmaf(xi[ci], wv.x, o[j]); o[j + 1] = fmaf(xi[ci], wv.y, o[j + 1]);
        o[j + 2] = fmaf(xi[ci], wv.z, o[j + 2]); o[j + 3] = fmaf(xi[ci], wv.w, o[j + 3]);
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < C; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    float* op = out + ((long long)b * nvox + v) * ldo;
#pragma unroll
    for (int j = 0; j < C; j += 4) st4(op + j, make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]));
#pragma unroll
    for (int j = 0; j < C; ++j) { s1[j] += o[j]; s2[j] = fmaf(o[j], o[j], s2[j]); }
  }
  if (stats) {
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const float a1 = warp_sum(s1[j]), a2 = warp_sum(s2[j]);
      if ((threadIdx.x & 31) == 0) { atomicAdd(&sred[0][j], (double)a1); atomicAdd(&sred[1][j], (double)a2); }
    }
    __syncthreads();
    if (threadIdx.x < C) {
      atomicAdd(stats + ((long long)b * C + threadIdx.x) * 2, sred[0][threadIdx.x]);
      atomicAdd(stats + ((long long)b * C + threadIdx.x) * 2 + 1, sred[1][threadIdx.x]);
    }
  }
}
// dW[co][ci] += sum over rows of g[row][co] * x[row][ci], C = 8: 64 partial sums per thread, folded through shuffles and
// shared memory into one atomic per element and block
__global__ void __launch_bounds__(THREADS) conv1_small_wgrad8_kernel(const float* __restrict__ x, long long ldx,
                                                                     const float* __restrict__ g, long long ldg,
                                                                     float* __restrict__ dW, long long rows) {
  constexpr int C = 8;
  __shared__ float sred[C * C];
  if (threadIdx.x < C * C) sred[threadIdx.x] = 0.f;
  __syncthreads();
  float acc[C][C];
#pragma unroll
  for (int a = 0; a < C; ++a)
#pragma unroll
    for (int c = 0; c < C; ++c) acc[a][c] = 0.f;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const float4 x0 = ld4(x + r * ldx), x1 = ld4(x + r * ldx + 4), g0 = ld4(g + r * ldg), g1 = ld4(g + r * ldg + 4);
    const float xv[C] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    const float gv[C] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int co = 0; co < C; ++co)
#pragma unroll
      for (int ci = 0; ci < C; ++ci) acc[co][ci] = fmaf(gv[co], xv[ci], acc[co][ci]);
  }
#pragma unroll
  for (int co = 0; co < C; ++co)
#pragma unroll
    for (int ci = 0; ci < C; ++ci) {
      const float t = warp_sum(acc[co][ci]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&sred[co * C + ci], t);
    }
  __syncthreads();
  if (threadIdx.x < C * C) atomicAdd(dW + threadIdx.x, sred[threadIdx.x]);
}

static int grid_for(long long nvox, int Q) {
  const int rows = ((THREADS / Q) * Q) / Q;
  long long blocks = (nvox + (long long)rows * 8 - 1) / ((long long)rows * 8);      // ~8 voxels per thread
  const long long cap = (long long)num_sms() * 16;
  blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
  return (int)blocks;
}

}  // namespace vol
}  // namespace corrif

using namespace corrif;
using namespace corrif::vol;

#define VOL_CHECK(p, ld, C, what)                                                                          \
  CORRIF_REQUIRE((p) != nullptr && ((uintptr_t)(p) % 16) == 0 && (C) > 0 && (C) % 4 == 0 && (ld) % 4 == 0 && \
                     (ld) >= (C) && (C) / 4 <= THREADS,                                                      \
                 what ": volume must be 16-byte aligned with C %% 4 == 0, ld %% 4 == 0, ld >= C")

extern "C" int corrif_instnorm_apply(float* x, int64_t ld, const double* stats, float* mean, float* rstd, int32_t B,
                                     int64_t nvox, int32_t C, float eps, void* stream) {
  VOL_CHECK(x, ld, C, "instnorm_apply");
  CORRIF_REQUIRE(stats && mean && rstd && B > 0 && nvox > 0, "instnorm_apply: null statistics / empty volume");
  dim3 grid(grid_for(nvox, C / 4), B);
  instnorm_apply_kernel<<<grid, THREADS, 0, (cudaStream_t)stream>>>(x, ld, stats, mean, rstd, nvox, C, eps);
  return launch_status("instnorm_apply");
}

extern "C" int corrif_instnorm_bwd_stats(const float* dy, int64_t lddy, const float* y, int64_t ldy, double* sums,
                                         int32_t B, int64_t nvox, int32_t C, void* stream) {
  VOL_CHECK(dy, lddy, C, "instnorm_bwd_stats(dy)");
  VOL_CHECK(y, ldy, C, "instnorm_bwd_stats(y)");
  CORRIF_REQUIRE(sums && B > 0 && nvox > 0, "instnorm_bwd_stats: null sums / empty volume");
  dim3 grid(grid_for(nvox, C / 4), B);
  instnorm_bwd_stats_kernel<<<grid, THREADS, 0, (cudaStream_t)stream>>>(dy, lddy, y, ldy, sums, nvox, C);
  return launch_status("instnorm_bwd_stats");
}

extern "C" int corrif_instnorm_relu_bwd_apply(const float* dy, int64_t lddy, const float* y, int64_t ldy,
                                              const float* mean, const float* rstd, const double* sums, float* g,
                                              int64_t ldg, float* dbias, int32_t B, int64_t nvox, int32_t C,
                                              int32_t relu, void* stream) {
  VOL_CHECK(dy, lddy, C, "instnorm_relu_bwd_apply(dy)");
  VOL_CHECK(y, ldy, C, "instnorm_relu_bwd_apply(y)");
  VOL_CHECK(g, ldg, C, "instnorm_relu_bwd_apply(g)");
  CORRIF_REQUIRE(mean && rstd && sums && B > 0 && nvox > 0, "instnorm_relu_bwd_apply: null statistics / empty volume");
  dim3 grid(grid_for(nvox, C / 4), B);
  instnorm_relu_bwd_apply_kernel<<<grid, THREADS, 0, (cudaStream_t)stream>>>(dy, lddy, y, ldy, mean, rstd, sums, g, ldg,
                                                                             dbias, nvox, C, relu);
  return launch_status("instnorm_relu_bwd_apply");
}

extern "C" int corrif_volume_colsum(const float* g, int64_t ldg, float* dbias, int64_t rows, int32_t C, void* stream) {
  VOL_CHECK(g, ldg, C, "volume_colsum");
  CORRIF_REQUIRE(dbias && rows > 0, "volume_colsum: null output / empty volume");
  colsum_kernel<<<grid_for(rows, C / 4), THREADS, 0, (cudaStream_t)stream>>>(g, ldg, dbias, rows, C);
  return launch_status("volume_colsum");
}

static int resize_check(const void* x, int64_t ldx, const void* y, int64_t ldy, int B, int C, int Di, int Hi, int Wi,
                        int Do, int Ho, int Wo) {
  VOL_CHECK(x, ldx, C, "resize(in)");
  VOL_CHECK(y, ldy, C, "resize(out)");
  CORRIF_REQUIRE(B > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "resize: empty volume");
  return 0;
}
static unsigned flat_grid(long long total) {
  long long blocks = (total + THREADS - 1) / THREADS;
  const long long cap = (long long)num_sms() * 32;
  return (unsigned)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

extern "C" int corrif_resize_trilinear_fwd(const float* x, int64_t ldx, float* y, int64_t ldy, int32_t B, int32_t C,
                                           int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho, int32_t Wo,
                                           void* stream) {
  int rc = resize_check(x, ldx, y, ldy, B, C, Di, Hi, Wi, Do, Ho, Wo);
  if (rc) return rc;
  const long long total = (long long)B * Do * Ho * Wo * (C / 4);
  trilinear_fwd_kernel<<<flat_grid(total), THREADS, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, C, Di, Hi, Wi, Do, Ho, Wo, total);
  return launch_status("resize_trilinear_fwd");
}
extern "C" int corrif_resize_trilinear_bwd(const float* dy, int64_t lddy, float* dx, int64_t lddx, int32_t B, int32_t C,
                                           int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho, int32_t Wo,
                                           void* stream) {
  int rc = resize_check(dx, lddx, dy, lddy, B, C, Di, Hi, Wi, Do, Ho, Wo);
  if (rc) return rc;
  const long long total = (long long)B * Di * Hi * Wi * (C / 4);
  // every input index has at most ceil(2 * (out-1)/(in-1)) + 1 contributing outputs per axis
  auto fits = [](int in, int out) { return in <= 1 ? out <= 8 : (2.0 * (out - 1) / (in - 1) + 1.0) <= 7.0; };
  const int small = fits(Di, Do) && fits(Hi, Ho) && fits(Wi, Wo);
  trilinear_bwd_kernel<<<flat_grid(total), THREADS, 0, (cudaStream_t)stream>>>(dy, lddy, dx, lddx, C, Di, Hi, Wi, Do, Ho, Wo, total, small);
  return launch_status("resize_trilinear_bwd");
}
extern "C" int corrif_conv1_small_fwd(const float* x, int64_t ldx, const float* w, const float* bias, float* out,
                                      int64_t ldo, double* stats, int32_t B, int64_t nvox, int32_t C, int32_t relu,
                                      int32_t transpose, void* stream) {
  CORRIF_REQUIRE(C == 8 || C == 16, "conv1_small_fwd: C must be 8 or 16 (got %d)", C);
  VOL_CHECK(x, ldx, C, "conv1_small_fwd(in)");
  VOL_CHECK(out, ldo, C, "conv1_small_fwd(out)");
  CORRIF_REQUIRE(w != nullptr && B > 0 && nvox > 0 && B <= 65535, "conv1_small_fwd: bad arguments");
  long long bx = (nvox + THREADS * 4 - 1) / (THREADS * 4);          // ~4 voxels per thread: the statistics amortise
  const long long cap = ((long long)num_sms() * 8 + B - 1) / B;
  bx = bx < 1 ? 1 : (bx > cap ? cap : bx);
  const dim3 grid((unsigned)bx, (unsigned)B);
  if (C == 8) conv1_small_kernel<8><<<grid, THREADS, 0, (cudaStream_t)stream>>>(x, ldx, w, bias, out, ldo, stats, nvox, relu, transpose);
  else conv1_small_kernel<16><<<grid, THREADS, 0, (cudaStream_t)stream>>>(x, ldx, w, bias, out, ldo, stats, nvox, relu, transpose);
  return launch_status("conv1_small_fwd");
}
extern "C" int corrif_conv1_small_wgrad(const float* x, int64_t ldx, const float* g, int64_t ldg, float* dW, int64_t rows,
                                        int32_t C, void* stream) {
  CORRIF_REQUIRE(C == 8, "conv1_small_wgrad: C must be 8 (got %d)", C);
  VOL_CHECK(x, ldx, C, "conv1_small_wgrad(x)");
  VOL_CHECK(g, ldg, C, "conv1_small_wgrad(g)");
  CORRIF_REQUIRE(dW != nullptr && rows > 0, "conv1_small_wgrad: bad arguments");
  long long bx = (rows + THREADS * 16 - 1) / (THREADS * 16);
  const long long cap = (long long)num_sms() * 4;
  bx = bx < 1 ? 1 : (bx > cap ? cap : bx);
  conv1_small_wgrad8_kernel<<<(unsigned)bx, THREADS, 0, (cudaStream_t)stream>>>(x, ldx, g, ldg, dW, rows);
  return launch_status("conv1_small_wgrad");
}

/* One axis of the trilinear resize on a contiguous tensor [outer][n][inner] (inner a multiple of 4 floats):
 * y[outer][n_out][inner] from x[outer][n_in][inner], and the adjoint (dx from dy; needs (n_out - 1) <= 3 (n_in - 1)). */
extern "C" int corrif_resize_linear_axis_fwd(const float* x, float* y, int64_t outer, int32_t n_in, int32_t n_out,
                                             int64_t inner, void* stream) {
  CORRIF_REQUIRE(x && y && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0, "resize_linear_axis_fwd: null / unaligned pointer");
  CORRIF_REQUIRE(outer > 0 && n_in > 0 && n_out > 0 && inner > 0 && inner % 4 == 0, "resize_linear_axis_fwd: bad shape");
  const long long total = (long long)outer * n_out * (inner / 4);
  CORRIF_REQUIRE(total < (1ll << 31) && (long long)outer * n_in * (inner / 4) < (1ll << 31), "resize_linear_axis_fwd: tensor too large");
  if (inner / 4 >= 4 * THREADS) {
    const long long rows = (long long)outer * n_out, cap = (long long)num_sms() * 16;
    const unsigned split = (unsigned)(rows >= cap ? 1 : (cap + rows - 1) / rows);
    linear_axis_fwd_rows_kernel<<<(unsigned)(rows * split < cap ? rows * split : cap), THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), (unsigned)n_in, (unsigned)n_out, (unsigned)(inner / 4),
        (unsigned)rows, split);
    return launch_status("resize_linear_axis_fwd");
  }
  linear_axis_fwd_kernel<<<flat_grid(total), THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), (unsigned)n_in, (unsigned)n_out, (unsigned)(inner / 4), (unsigned)total);
  return launch_status("resize_linear_axis_fwd");
}
extern "C" int corrif_resize_linear_axis_bwd(const float* dy, float* dx, int64_t outer, int32_t n_in, int32_t n_out,
                                             int64_t inner, void* stream) {
  CORRIF_REQUIRE(dx && dy && ((uintptr_t)dx % 16) == 0 && ((uintptr_t)dy % 16) == 0, "resize_linear_axis_bwd: null / unaligned pointer");
  CORRIF_REQUIRE(outer > 0 && n_in > 0 && n_out > 0 && inner > 0 && inner % 4 == 0, "resize_linear_axis_bwd: bad shape");
  CORRIF_REQUIRE(n_in > 1 ? (2.0 * (n_out - 1) / (n_in - 1) + 1.0) <= 7.0 : n_out <= 8,
                 "resize_linear_axis_bwd: more than 8 contributing outputs per input (use corrif_resize_trilinear_bwd)");
  const long long total = (long long)outer * n_in * (inner / 4);
  CORRIF_REQUIRE(total < (1ll << 31) && (long long)outer * n_out * (inner / 4) < (1ll << 31), "resize_linear_axis_bwd: tensor too large");
  if (inner / 4 >= 4 * THREADS) {
    const long long rows = (long long)outer * n_in, cap = (long long)num_sms() * 16;
    const unsigned split = (unsigned)(rows >= cap ? 1 : (cap + rows - 1) / rows);
    linear_axis_bwd_rows_kernel<<<(unsigned)(rows * split < cap ? rows * split : cap), THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx), (unsigned)n_in, (unsigned)n_out, (unsigned)(inner / 4),
        (unsigned)rows, split);
    return launch_status("resize_linear_axis_bwd");
  }
  linear_axis_bwd_kernel<<<flat_grid(total), THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx), (unsigned)n_in, (unsigned)n_out, (unsigned)(inner / 4), (unsigned)total);
  return launch_status("resize_linear_axis_bwd");
}
extern "C" int corrif_resize_nearest_fwd(const float* x, int64_t ldx, float* y, int64_t ldy, int32_t B, int32_t C,
                                         int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho, int32_t Wo,
                                         void* stream) {
  int rc = resize_check(x, ldx, y, ldy, B, C, Di, Hi, Wi, Do, Ho, Wo);
  if (rc) return rc;
  const long long total = (long long)B * Do * Ho * Wo * (C / 4);
  const long long rows = (long long)B * Do * Ho;
  const long long cap = (long long)num_sms() * 32;
  nearest_fwd_kernel<<<(unsigned)(rows < cap ? rows : cap), Wo * (C / 4) >= THREADS ? THREADS : 128, 0, (cudaStream_t)stream>>>(
      x, ldx, y, ldy, C, Di, Hi, Wi, Do, Ho, Wo, total);
  return launch_status("resize_nearest_fwd");
}
extern "C" int corrif_resize_nearest_bwd(const float* dy, int64_t lddy, float* dx, int64_t lddx, int32_t B, int32_t C,
                                         int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho, int32_t Wo,
                                         void* stream) {
  int rc = resize_check(dx, lddx, dy, lddy, B, C, Di, Hi, Wi, Do, Ho, Wo);
  if (rc) return rc;
  const long long warps = (long long)B * Di * Hi * Wi;
  const long long blocks = (warps * 32 + THREADS - 1) / THREADS;
  nearest_bwd_kernel<<<(unsigned)blocks, THREADS, 0, (cudaStream_t)stream>>>(dy, lddy, dx, lddx, C, Di, Hi, Wi, Do, Ho, Wo, warps);
  return launch_status("resize_nearest_bwd");
}

/* ---- train-mode BatchNorm (+ residual) (+ ReLU) on [rows, C] channels-last data (C <= 1024 per call) ---- */
extern "C" int corrif_batchnorm_fwd(const float* x, int64_t ldx, const double* stats, const float* gamma,
                                    const float* beta, const float* res, int64_t ldr, float* y, int64_t ldy, float* mean,
                                    float* var, float* rstd, int64_t rows, int32_t C, float eps, int32_t relu, void* stream) {
  VOL_CHECK(x, ldx, C, "batchnorm_fwd(x)");
  VOL_CHECK(y, ldy, C, "batchnorm_fwd(y)");
  CORRIF_REQUIRE(stats && gamma && beta && mean && var && rstd && rows > 0, "batchnorm_fwd: null pointer / empty");
  CORRIF_REQUIRE(res == nullptr || (((uintptr_t)res % 16) == 0 && ldr % 4 == 0 && ldr >= C), "batchnorm_fwd: residual unaligned");
  bn_apply_kernel<<<grid_for(rows, C / 4), THREADS, 0, (cudaStream_t)stream>>>(x, ldx, stats, gamma, beta, res, ldr, y, ldy,
                                                                              mean, var, rstd, rows, C, eps, relu);
  return launch_status("batchnorm_fwd");
}

extern "C" int corrif_batchnorm_bwd_stats(const float* dy, int64_t lddy, const float* y, int64_t ldy, const float* x,
                                          int64_t ldx, const float* mean, const float* rstd, double* sums, int64_t rows,
                                          int32_t C, int32_t relu, void* stream) {
  VOL_CHECK(dy, lddy, C, "batchnorm_bwd_stats(dy)");
  VOL_CHECK(x, ldx, C, "batchnorm_bwd_stats(x)");
  CORRIF_REQUIRE(mean && rstd && sums && rows > 0 && (!relu || y != nullptr), "batchnorm_bwd_stats: null pointer / empty");
  bn_bwd_stats_kernel<<<grid_for(rows, C / 4), THREADS, 0, (cudaStream_t)stream>>>(dy, lddy, y, ldy, x, ldx, mean, rstd, sums,
                                                                                  rows, C, relu);
  return launch_status("batchnorm_bwd_stats");
}

extern "C" int corrif_batchnorm_bwd_apply(const float* dy, int64_t lddy, const float* y, int64_t ldy, const float* x,
                                          int64_t ldx, const float* mean, const float* rstd, const float* gamma,
                                          const double* sums, float* dx, int64_t lddx, float* dres, int64_t lddr,
                                          float* dgamma, float* dbeta, int64_t rows, int32_t C, int32_t relu, void* stream) {
  VOL_CHECK(dy, lddy, C, "batchnorm_bwd_apply(dy)");
  VOL_CHECK(x, ldx, C, "batchnorm_bwd_apply(x)");
  VOL_CHECK(dx, lddx, C, "batchnorm_bwd_apply(dx)");
  CORRIF_REQUIRE(mean && rstd && gamma && sums && dgamma && dbeta && rows > 0 && (!relu || y != nullptr),
                 "batchnorm_bwd_apply: null pointer / empty");
  bn_bwd_apply_kernel<<<grid_for(rows, C / 4), THREADS, 0, (cudaStream_t)stream>>>(dy, lddy, y, ldy, x, ldx, mean, rstd, gamma,
                                                                                  sums, dx, lddx, dres, lddr, dgamma, dbeta,
                                                                                  rows, C, relu);
  return launch_status("batchnorm_bwd_apply");
}
