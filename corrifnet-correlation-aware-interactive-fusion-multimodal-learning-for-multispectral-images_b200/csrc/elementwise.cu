// HBM-bound kernels of the fusion block: LayerNorm(+pos), row softmax, dropout, transposes and the
// small reductions of the backward.  All accesses are 128-bit and coalesced; row reductions use
// warp shuffles (one warp per row, the row lives in registers).
#include "common.cuh"

namespace corrif {

// ============================================================================================
// batched transpose   in [batch, rows, cols] -> out [batch, cols, rows]
// ============================================================================================
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows,
                                 int cols, int round) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const float* src = in + b * (int64_t)rows * cols;
  float* dst = out + b * (int64_t)rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[(int64_t)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) {
      const float v = tile[threadIdx.x][i];
      dst[(int64_t)c * rows + r] = round ? round_tf32(v) : v;
    }
  }
}

// 64 x 64 tiles, 128-bit accesses on both sides (rows and cols multiples of 64: every transpose of the fusion
// block - 64 / 192 / 512).  The 32 x 32 scalar kernel above moved these 2-6 MB tensors at 0.5 TB/s (8.5 us each).
__global__ void __launch_bounds__(256)
transpose64_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols, int round) {
  __shared__ float tile[64][65];
  const int64_t b = blockIdx.z;
  const float* src = in + b * (int64_t)rows * cols;
  float* dst = out + b * (int64_t)rows * cols;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int q = threadIdx.x & 15, rr = threadIdx.x >> 4;          // 16 float4 per 64-float row, 16 rows per pass
  float4 v[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) v[p] = ld4_stream(src + (int64_t)(r0 + p * 16 + rr) * cols + c0 + q * 4);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float* t = &tile[p * 16 + rr][q * 4];
    t[0] = v[p].x; t[1] = v[p].y; t[2] = v[p].z; t[3] = v[p].w;
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int c = p * 16 + rr;                                      // output row (an input column)
    float4 o = make_float4(tile[q * 4 + 0][c], tile[q * 4 + 1][c], tile[q * 4 + 2][c], tile[q * 4 + 3][c]);
    if (round) o = round_tf32_4(o);
    st4(dst + (int64_t)(c0 + c) * rows + r0 + q * 4, o);
  }
}

// ============================================================================================
// LayerNorm (+ positional add), C == 512: one warp per row, 16 floats per lane
// ============================================================================================
constexpr int LN_C = 512;
constexpr int LN_V = LN_C / 128;  // float4 per lane

__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pos, int64_t pos_rows,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* __restrict__ x1_out, float* __restrict__ y, float* __restrict__ mean,
                     float* __restrict__ rstd, int64_t rows, int round) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * LN_C;
  float4 v[LN_V];
#pragma unroll
  for (int j = 0; j < LN_V; ++j) v[j] = ld4_stream(xr + lane * 4 + j * 128);
  if (pos != nullptr) {
    const float* pr = pos + (row % pos_rows) * LN_C;
#pragma unroll
    for (int j = 0; j < LN_V; ++j) {
      const float4 p = ld4(pr + lane * 4 + j * 128);
      v[j].x += p.x; v[j].y += p.y; v[j].z += p.z; v[j].w += p.w;
    }
    if (x1_out != nullptr) {
#pragma unroll
      for (int j = 0; j < LN_V; ++j) st4(x1_out + row * LN_C + lane * 4 + j * 128, v[j]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_V; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mu = warp_sum(s) * (1.0f / LN_C);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < LN_V; ++j) {
    const float a = v[j].x - mu, b = v[j].y - mu, c = v[j].z - mu, d = v[j].w - mu;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rs = rsqrtf(warp_sum(q) * (1.0f / LN_C) + 1e-5f);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
  for (int j = 0; j < LN_V; ++j) {
    const float4 g = ld4(gamma + lane * 4 + j * 128), b = ld4(beta + lane * 4 + j * 128);
    float4 o;
    o.x = (v[j].x - mu) * rs * g.x + b.x;
    o.y = (v[j].y - mu) * rs * g.y + b.y;
    o.z = (v[j].z - mu) * rs * g.z + b.z;
    o.w = (v[j].w - mu) * rs * g.w + b.w;
    if (round) o = round_tf32_4(o);
    st4(y + row * LN_C + lane * 4 + j * 128, o);
  }
}

// Backward: each warp walks rows with a grid stride, keeps its dgamma/dbeta partials in registers,
// the 8 warps of a block are combined through shared memory, one partial row per block goes to
// scratch and a second kernel folds the partials (deterministic, no atomics).
constexpr int LN_BWD_MAX_BLOCKS = 592;  // 4 per SM

// RES2 (a second residual input) is a template parameter: as a run-time pointer it cost the common
// instantiation its last free registers (127 + a 16-byte spill; ncu 177 -> 213 us over the step's 8 launches).
template <bool RES2>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x1,
                     const float* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ dres,
                     float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows,
                     float* __restrict__ dx_drop, uint32_t thresh, float keep_scale, uint64_t seed,
                     const uint64_t* seed_dev, uint32_t site_a, uint32_t site_b, int groups, int group_rows,
                     const float* __restrict__ dres2) {
  __shared__ __align__(16) float red[8][2][LN_C];
  uint64_t key_a = 0, key_b = 0;
  if (dx_drop != nullptr) {      // fused dropout of the outgoing gradient (backward of the next block's
    if (seed_dev != nullptr) seed += *seed_dev;   // proj_drop + PreNormDrop.dropout, mmvit4.py:314,339)
    key_a = dropout_key(seed, site_a);
    key_b = site_b != CORRIF_NO_SITE ? dropout_key(seed, site_b) : 0ull;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 g[LN_V], dg[LN_V], db[LN_V];
#pragma unroll
  for (int j = 0; j < LN_V; ++j) {
    g[j] = ld4(gamma + lane * 4 + j * 128);
    dg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
    const float mu = mean[row], rs = rstd[row];
    // groups > 0: rows are [batch][group][group_rows]; dx is written group-major, [group][batch][group_rows]
    // (the multimodal token gradient leaves in the layout its consumers want: no strided copy afterwards)
    int64_t orow = row;
    if (groups > 0) {
      const int64_t per_b = (int64_t)groups * group_rows, bb = row / per_b, rem = row - bb * per_b;
      const int64_t X = rem / group_rows, sidx = rem - X * group_rows;
      orow = (X * (rows / per_b) + bb) * group_rows + sidx;
    }
    float4 xh[LN_V], d[LN_V], rres[LN_V];
    float c1 = 0.f, c2 = 0.f;
    // all 12 loads of the row are issued before the first use (the residual gradient used to be fetched
    // after the two warp reductions: a second, dependent DRAM round trip per row)
#pragma unroll
    for (int j = 0; j < LN_V; ++j) xh[j] = ld4_stream(x1 + row * LN_C + lane * 4 + j * 128);
#pragma unroll
    for (int j = 0; j < LN_V; ++j) d[j] = ld4_stream(dy + row * LN_C + lane * 4 + j * 128);
    if (dres != nullptr) {
#pragma unroll
      for (int j = 0; j < LN_V; ++j) rres[j] = ld4_stream(dres + row * LN_C + lane * 4 + j * 128);
    }
#pragma unroll
    for (int j = 0; j < LN_V; ++j) {
      const float4 xv = xh[j];
      xh[j].x = (xv.x - mu) * rs; xh[j].y = (xv.y - mu) * rs;
      xh[j].z = (xv.z - mu) * rs; xh[j].w = (xv.w - mu) * rs;
      dg[j].x += d[j].x * xh[j].x; dg[j].y += d[j].y * xh[j].y;
      dg[j].z += d[j].z * xh[j].z; dg[j].w += d[j].w * xh[j].w;
      db[j].x += d[j].x; db[j].y += d[j].y; db[j].z += d[j].z; db[j].w += d[j].w;
      d[j].x *= g[j].x; d[j].y *= g[j].y; d[j].z *= g[j].z; d[j].w *= g[j].w;
      c1 += (d[j].x + d[j].y) + (d[j].z + d[j].w);
      c2 += (d[j].x * xh[j].x + d[j].y * xh[j].y) + (d[j].z * xh[j].z + d[j].w * xh[j].w);
    }
    c1 = warp_sum(c1) * (1.0f / LN_C);
    c2 = warp_sum(c2) * (1.0f / LN_C);
#pragma unroll
    for (int j = 0; j < LN_V; ++j) {
      float4 o;
      o.x = rs * (d[j].x - c1 - xh[j].x * c2);
      o.y = rs * (d[j].y - c1 - xh[j].y * c2);
      o.z = rs * (d[j].z - c1 - xh[j].z * c2);
      o.w = rs * (d[j].w - c1 - xh[j].w * c2);
      if (dres != nullptr) {
        const float4 r = rres[j];
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      if (RES2) {                  // a second incoming gradient of the same tensor (the skip path, mmvit4.py:505)
        const float4 r = ld4_stream(dres2 + row * LN_C + lane * 4 + j * 128);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      st4(dx + orow * LN_C + lane * 4 + j * 128, o);
      if (dx_drop != nullptr) {
        const uint64_t quad = (uint64_t)(row * LN_C + lane * 4 + j * 128) >> 2;
        uint32_t km = dropout_keepmask4(key_a, quad, thresh);
        if (site_b != CORRIF_NO_SITE) km &= dropout_keepmask4(key_b, quad, thresh);
        st4(dx_drop + row * LN_C + lane * 4 + j * 128,
            make_float4((km & 1u) ? o.x * keep_scale : 0.f, (km & 2u) ? o.y * keep_scale : 0.f,
                        (km & 4u) ? o.z * keep_scale : 0.f, (km & 8u) ? o.w * keep_scale : 0.f));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < LN_V; ++j) {
    st4(&red[warp][0][lane * 4 + j * 128], dg[j]);
    st4(&red[warp][1][lane * 4 + j * 128], db[j]);
  }
  __syncthreads();
  {   // block totals straight into dgamma / dbeta (256 threads x float4 = 2 x 512 columns)
    const int which = threadIdx.x >> 7, c = (threadIdx.x & 127) * 4;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float4 v = ld4(&red[w][which][c]);
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    float* dst = (which ? dbeta : dgamma) + c;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(dst), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
  }
}

// partial [nparts][width] -> out: columns [0,split) go to out0, [split,width) to out1.  16 threads
// share one column (strided over the partials, coalesced across columns), combined through smem.
__global__ void __launch_bounds__(512)
fold_partials_kernel(const float* __restrict__ partial, int nparts, int width,
                     float* __restrict__ out0, float* __restrict__ out1, int split, int accumulate) {
  __shared__ float red[16][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < width)
    for (int p = threadIdx.y; p < nparts; p += 16) s += partial[(int64_t)p * width + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < width) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 16; ++y) t += red[y][threadIdx.x];
    float* dst = c < split ? out0 + c : out1 + (c - split);
    *dst = accumulate ? *dst + t : t;
  }
}

// ============================================================================================
// row softmax (in place) with optional Philox dropout copy; one warp per row
// ============================================================================================
template <int NV>  // float4 per lane capacity; cols = 128 * nv, nv <= NV
__global__ void __launch_bounds__(256)
softmax_fwd_kernel(float* __restrict__ S, float* __restrict__ Pd, int64_t rows, int cols, int nv,
                   uint32_t thresh, float keep_scale, uint64_t seed, const uint64_t* seed_dev,
                   uint32_t site, int round) {
  if (seed_dev != nullptr) seed += *seed_dev;
  const uint64_t key = dropout_key(seed, site);
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* sr = S + row * cols;
  float4 v[NV];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (j < nv) {
      v[j] = ld4(sr + lane * 4 + j * 128);
      mx = fmaxf(mx, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (j < nv) {
      v[j].x = __expf(v[j].x - mx); v[j].y = __expf(v[j].y - mx);
      v[j].z = __expf(v[j].z - mx); v[j].w = __expf(v[j].w - mx);
      sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  const float inv = 1.0f / warp_sum(sum);
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (j < nv) {
      v[j].x *= inv; v[j].y *= inv; v[j].z *= inv; v[j].w *= inv;
      if (round) v[j] = round_tf32_4(v[j]);
      st4(sr + lane * 4 + j * 128, v[j]);
      if (Pd != nullptr) {
        float m[4];
        const uint64_t e = (uint64_t)row * cols + lane * 4 + j * 128;
        dropout_keep4(key, e >> 2, thresh, keep_scale, m);
        float4 pd = make_float4(v[j].x * m[0], v[j].y * m[1], v[j].z * m[2], v[j].w * m[3]);
        if (round) pd = round_tf32_4(pd);
        st4(Pd + row * cols + lane * 4 + j * 128, pd);
      }
    }
}

template <int NV>
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ P, float* __restrict__ dP, int64_t rows, int cols,
                   int nv, float scale, uint32_t thresh, float keep_scale, uint64_t seed,
                   const uint64_t* seed_dev, uint32_t site, int use_drop) {
  if (seed_dev != nullptr) seed += *seed_dev;
  const uint64_t key = dropout_key(seed, site);
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 p[NV], d[NV];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (j < nv) {
      p[j] = ld4_stream(P + row * cols + lane * 4 + j * 128);
      d[j] = ld4(dP + row * cols + lane * 4 + j * 128);
      if (use_drop) {
        float m[4];
        const uint64_t e = (uint64_t)row * cols + lane * 4 + j * 128;
        dropout_keep4(key, e >> 2, thresh, keep_scale, m);
        d[j].x *= m[0]; d[j].y *= m[1]; d[j].z *= m[2]; d[j].w *= m[3];
      }
      dot += (p[j].x * d[j].x + p[j].y * d[j].y) + (p[j].z * d[j].z + p[j].w * d[j].w);
    }
  dot = warp_sum(dot);
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (j < nv) {
      float4 o;
      o.x = p[j].x * (d[j].x - dot) * scale; o.y = p[j].y * (d[j].y - dot) * scale;
      o.z = p[j].z * (d[j].z - dot) * scale; o.w = p[j].w * (d[j].w - dot) * scale;
      st4(dP + row * cols + lane * 4 + j * 128, o);
    }
}

// ============================================================================================
// dropout / mask
// ============================================================================================
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t nquads,
                               uint32_t thresh, float keep_scale, uint64_t seed,
                               const uint64_t* seed_dev, uint32_t site, int mask_only) {
  if (seed_dev != nullptr) seed += *seed_dev;
  const uint64_t key = dropout_key(seed, site);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads;
       q += (int64_t)gridDim.x * blockDim.x) {
    float m[4];
    dropout_keep4(key, (uint64_t)q, thresh, mask_only ? 1.0f : keep_scale, m);
    float4 v = mask_only ? make_float4(1.f, 1.f, 1.f, 1.f) : ld4(x + q * 4);
    st4(out + q * 4, make_float4(v.x * m[0], v.y * m[1], v.z * m[2], v.w * m[3]));
  }
}

// dropout of a [rows, cols] gradient PLUS its column sums (the bias gradient of the layer it feeds, e.g.
// FeedForward's second Linear: mmvit4.py:354-355 backward) in one pass - the stand-alone colsum re-read
// the dropped tensor.  256 threads = 256 / (cols/4) rows per sweep, thread -> fixed column quad, 4 rows in
// flight; block totals go out as red.global.add.v4.  Same keep decisions as dropout_kernel (quad = flat / 4).
__global__ void __launch_bounds__(256)
dropout_colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t rows, int cq,
                      uint32_t thresh, float keep_scale, uint64_t seed, const uint64_t* seed_dev,
                      uint32_t site, float* __restrict__ colsum) {
  __shared__ __align__(16) float red[256][4];
  if (seed_dev != nullptr) seed += *seed_dev;
  const uint64_t key = dropout_key(seed, site);
  const int rpi = 256 / cq, c = threadIdx.x % cq, roff = threadIdx.x / cq;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t stride = (int64_t)gridDim.x * rpi;
  for (int64_t r0 = (int64_t)blockIdx.x * rpi + roff; r0 < rows; r0 += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = r0 + u * stride;
      v[u] = r < rows ? ld4_stream(x + (r * cq + c) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = r0 + u * stride;
      if (r < rows) {
        float m[4];
        dropout_keep4(key, (uint64_t)(r * cq + c), thresh, keep_scale, m);
        const float4 o = make_float4(v[u].x * m[0], v[u].y * m[1], v[u].z * m[2], v[u].w * m[3]);
        st4(out + (r * cq + c) * 4, o);
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
    }
  }
  st4(&red[threadIdx.x][0], acc);
  __syncthreads();
  if (threadIdx.x < cq) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int g = 0; g < rpi; ++g) {
      const float4 w = ld4(&red[g * cq + threadIdx.x][0]);
      t.x += w.x; t.y += w.y; t.z += w.z; t.w += w.w;
    }
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(colsum + threadIdx.x * 4), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
  }
}

__global__ void round_tf32_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t nquads) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads;
       q += (int64_t)gridDim.x * blockDim.x)
    st4(out + q * 4, round_tf32_4(ld4(x + q * 4)));
}

__global__ void dropout_add_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                   float* __restrict__ out, int64_t nquads, uint32_t thresh,
                                   float keep_scale, uint64_t seed, const uint64_t* seed_dev,
                                   uint32_t site_a, uint32_t site_b) {
  if (seed_dev != nullptr) seed += *seed_dev;
  const uint64_t key_a = dropout_key(seed, site_a), key_b = dropout_key(seed, site_b);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads;
       q += (int64_t)gridDim.x * blockDim.x) {
    float m[4], m2[4] = {1.f, 1.f, 1.f, 1.f};
    dropout_keep4(key_a, (uint64_t)q, thresh, keep_scale, m);
    if (site_b != CORRIF_NO_SITE) dropout_keep4(key_b, (uint64_t)q, thresh, keep_scale, m2);
    const float4 v = ld4(x + q * 4);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res != nullptr) r = ld4(res + q * 4);
    st4(out + q * 4, make_float4(v.x * m[0] * m2[0] + r.x, v.y * m[1] * m2[1] + r.y,
                                 v.z * m[2] * m2[2] + r.z, v.w * m[3] * m2[3] + r.w));
  }
}

// ============================================================================================
// reductions of the backward
// ============================================================================================
__global__ void __launch_bounds__(128)
colsum_partial_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int cols,
                      float* __restrict__ partial, int64_t x_bstride, int64_t out_bstride) {
  const int c = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (c >= cols) return;
  x += (int64_t)blockIdx.z * x_bstride;
  partial += (int64_t)blockIdx.z * out_bstride;
  const int64_t chunk = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = blockIdx.y * chunk, r1 = min(rows, r0 + chunk);
  // 8 independent row loads in flight per thread: with 2 the kernel was latency-bound at ~1.1 TB/s
  float4 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t r = r0;
  for (; r + 8 <= r1; r += 8) {
    float4 u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = ld4_stream(x + (r + i) * ld + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i].x += u[i].x; acc[i].y += u[i].y; acc[i].z += u[i].z; acc[i].w += u[i].w; }
  }
  for (; r < r1; ++r) {
    const float4 u = ld4_stream(x + r * ld + c);
    acc[0].x += u.x; acc[0].y += u.y; acc[0].z += u.z; acc[0].w += u.w;
  }
#pragma unroll
  for (int i = 1; i < 8; ++i) { acc[0].x += acc[i].x; acc[0].y += acc[i].y; acc[0].z += acc[i].z; acc[0].w += acc[i].w; }
  float* dst = partial + c;   // `partial` is the output vector: add this block's share
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(dst), "f"(acc[0].x), "f"(acc[0].y), "f"(acc[0].z), "f"(acc[0].w) : "memory");
}

__global__ void round_tf32_multi_kernel(const float* const* __restrict__ src, float* const* __restrict__ dst,
                                        const int64_t* __restrict__ n) {
  const float* x = src[blockIdx.y];
  float* o = dst[blockIdx.y];
  const int64_t cnt = n[blockIdx.y];
  const bool copy = cnt < 0;                      // negative count: plain copy (stacked biases)
  const int64_t nq = (copy ? -cnt : cnt) >> 2;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ld4(x + q * 4);
    st4(o + q * 4, copy ? v : round_tf32_4(v));
  }
}

__global__ void batchsum_kernel(const float* __restrict__ x, int64_t batch, int64_t stride,
                                int64_t nquads, float* __restrict__ out, int accumulate) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads;
       q += (int64_t)gridDim.x * blockDim.x) {
    float4 a = accumulate ? ld4(out + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t b = 0;
    for (; b + 8 <= batch; b += 8) {       // 8 independent loads in flight (one at a time ran at 1.5 TB/s)
      float4 u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = ld4_stream(x + (b + i) * stride + q * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) { a.x += u[i].x; a.y += u[i].y; a.z += u[i].z; a.w += u[i].w; }
    }
    for (; b < batch; ++b) {
      const float4 u = ld4_stream(x + b * stride + q * 4);
      a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
    }
    st4(out + q * 4, a);
  }
}

__global__ void add_rows_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b,
                                int64_t ldb, float* __restrict__ out, int64_t ldo, int64_t rows,
                                int cq) {
  const int64_t total = rows * cq;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cq;
    const int c = (int)(i % cq) * 4;
    const float4 u = ld4(a + r * lda + c), w = ld4(b + r * ldb + c);
    st4(out + r * ldo + c, make_float4(u.x + w.x, u.y + w.y, u.z + w.z, u.w + w.w));
  }
}

// ============================================================================================
// train-step tail: BCE-with-logits on probabilities (+grad), Adam
// ============================================================================================
__global__ void __launch_bounds__(256)
bce_probs_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n,
                 float grad_scale, double* __restrict__ loss_sum, float* __restrict__ dx) {
  __shared__ double wsum[8];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i], yv = y[i];
    // max(x,0) - x*y + log1p(exp(-|x|))
    acc += (double)(fmaxf(xv, 0.f) - xv * yv + log1pf(expf(-fabsf(xv))));
    if (dx != nullptr) dx[i] = (1.0f / (1.0f + expf(-xv)) - yv) * grad_scale;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += wsum[w];
    atomicAdd(loss_sum, s);
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2,
                            float eps, float grad_scale, float bc1, float bc2_sqrt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

static inline int grid_for(int64_t work, int threads) {
  int64_t b = (work + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace corrif

using namespace corrif;

extern "C" {

int corrif_round_tf32(const float* in, float* out, int64_t n, void* stream) {
  CORRIF_REQUIRE(in && out && n > 0 && n % 4 == 0, "round_tf32: n must be a positive multiple of 4");
  round_tf32_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n / 4);
  return launch_status("round_tf32");
}

int corrif_transpose(const float* in, float* out, int64_t batch, int32_t rows, int32_t cols,
                     int32_t round_tf32, void* stream) {
  CORRIF_REQUIRE(in && out && batch > 0 && rows > 0 && cols > 0, "transpose: bad arguments");
  CORRIF_REQUIRE(batch <= 65535, "transpose: batch %lld > 65535", (long long)batch);
  if (rows % 64 == 0 && cols % 64 == 0 && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0)) {
    dim3 grid64(cols / 64, rows / 64, (unsigned)batch);
    transpose64_kernel<<<grid64, 256, 0, (cudaStream_t)stream>>>(in, out, rows, cols, round_tf32);
    return launch_status("transpose");
  }
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, (unsigned)batch), block(32, 8);
  transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, out, rows, cols, round_tf32);
  return launch_status("transpose");
}

int corrif_layernorm_fwd(const float* x, const float* pos, int64_t pos_rows, const float* gamma,
                         const float* beta, float* x1_out, float* y, float* mean, float* rstd,
                         int64_t rows, int32_t C, int32_t round_tf32, void* stream) {
  CORRIF_REQUIRE(C == LN_C, "layernorm: C must be 512, got %d", C);
  CORRIF_REQUIRE(x && gamma && beta && y && mean && rstd && rows > 0, "layernorm_fwd: null/empty");
  CORRIF_REQUIRE(pos == nullptr || pos_rows > 0, "layernorm_fwd: pos_rows");
  layernorm_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      x, pos, pos_rows, gamma, beta, x1_out, y, mean, rstd, rows, round_tf32);
  return launch_status("layernorm_fwd");
}

int64_t corrif_layernorm_bwd_scratch_floats(int64_t rows, int32_t C) {
  (void)rows;
  return (int64_t)LN_BWD_MAX_BLOCKS * 2 * C;
}

int corrif_layernorm_bwd(const float* dy, const float* x1, const float* gamma, const float* mean,
                         const float* rstd, const float* dres, float* dx, float* dgamma,
                         float* dbeta, float* scratch, int64_t rows, int32_t C, int32_t accumulate,
                         float* dx_drop, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                         uint32_t site_a, uint32_t site_b, void* stream) {
  return corrif_layernorm_bwd_regroup(dy, x1, gamma, mean, rstd, dres, dx, dgamma, dbeta, scratch, rows, C, accumulate,
                                      dx_drop, p_drop, seed, seed_dev, site_a, site_b, 0, 0, nullptr, stream);
}

int corrif_layernorm_bwd_regroup(const float* dy, const float* x1, const float* gamma, const float* mean,
                                 const float* rstd, const float* dres, float* dx, float* dgamma,
                                 float* dbeta, float* scratch, int64_t rows, int32_t C, int32_t accumulate,
                                 float* dx_drop, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                                 uint32_t site_a, uint32_t site_b, int32_t groups, int32_t group_rows,
                                 const float* dres2, void* stream) {
  (void)scratch;
  CORRIF_REQUIRE(groups >= 0 && (groups == 0 || (group_rows > 0 && rows % ((int64_t)groups * group_rows) == 0 && dx_drop == nullptr && dx != dy && dx != dres)),
                 "layernorm_bwd: regrouping needs rows %% (groups * group_rows) == 0, no dx_drop and an out-of-place dx");
  CORRIF_REQUIRE(C == LN_C, "layernorm: C must be 512, got %d", C);
  CORRIF_REQUIRE(dy && x1 && gamma && mean && rstd && dx && dgamma && dbeta && rows > 0,
                 "layernorm_bwd: null/empty");
  CORRIF_REQUIRE(dx_drop == nullptr || (p_drop > 0.f && p_drop < 1.f), "layernorm_bwd: dx_drop needs 0 < p < 1");
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(dgamma, 0, LN_C * sizeof(float), (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(dbeta, 0, LN_C * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_last_error("layernorm_bwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  int blocks = (int)((rows + 7) / 8);
  const int cap = num_sms() * 2;     // 128 registers x 256 threads: two resident blocks per SM, one wave
  if (blocks > cap) blocks = cap;
  float ks = 1.0f;
  if (dx_drop) { ks = 1.0f / (1.0f - p_drop); if (site_b != CORRIF_NO_SITE) ks *= ks; }
  auto kern = dres2 != nullptr ? layernorm_bwd_kernel<true> : layernorm_bwd_kernel<false>;
  kern<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      dy, x1, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, dx_drop, dx_drop ? dropout_threshold(p_drop) : 0u,
      ks, seed, seed_dev, site_a, site_b, groups, group_rows, dres2);
  return launch_status("layernorm_bwd");
}

int corrif_softmax_fwd(float* S, float* Pdrop, int64_t rows, int32_t cols, float p_drop,
                       uint64_t seed, const uint64_t* seed_dev, uint32_t site, int32_t round_tf32,
                       void* stream) {
  CORRIF_REQUIRE(S && rows > 0 && cols > 0 && cols % 128 == 0 && cols <= 4096,
                 "softmax_fwd: cols must be a multiple of 128 and <= 4096 (got %d)", cols);
  CORRIF_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "softmax_fwd: p_drop");
  if (p_drop == 0.f) Pdrop = nullptr;
  const int nv = cols / 128;
  const uint32_t th = dropout_threshold(p_drop);
  const float ks = 1.0f / (1.0f - p_drop);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (nv <= 4) softmax_fwd_kernel<4><<<grid, 256, 0, st>>>(S, Pdrop, rows, cols, nv, th, ks, seed, seed_dev, site, round_tf32);
  else if (nv <= 16) softmax_fwd_kernel<16><<<grid, 256, 0, st>>>(S, Pdrop, rows, cols, nv, th, ks, seed, seed_dev, site, round_tf32);
  else softmax_fwd_kernel<32><<<grid, 256, 0, st>>>(S, Pdrop, rows, cols, nv, th, ks, seed, seed_dev, site, round_tf32);
  return launch_status("softmax_fwd");
}

int corrif_softmax_bwd(const float* P, float* dP, int64_t rows, int32_t cols, float scale,
                       float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t site,
                       void* stream) {
  CORRIF_REQUIRE(P && dP && rows > 0 && cols > 0 && cols % 128 == 0 && cols <= 4096,
                 "softmax_bwd: cols must be a multiple of 128 and <= 4096 (got %d)", cols);
  CORRIF_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "softmax_bwd: p_drop");
  const int nv = cols / 128;
  const uint32_t th = dropout_threshold(p_drop);
  const float ks = 1.0f / (1.0f - p_drop);
  const int ud = p_drop > 0.f;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (nv <= 4) softmax_bwd_kernel<4><<<grid, 256, 0, st>>>(P, dP, rows, cols, nv, scale, th, ks, seed, seed_dev, site, ud);
  else if (nv <= 16) softmax_bwd_kernel<16><<<grid, 256, 0, st>>>(P, dP, rows, cols, nv, scale, th, ks, seed, seed_dev, site, ud);
  else softmax_bwd_kernel<32><<<grid, 256, 0, st>>>(P, dP, rows, cols, nv, scale, th, ks, seed, seed_dev, site, ud);
  return launch_status("softmax_bwd");
}

int corrif_dropout(const float* x, float* out, int64_t n, float p, uint64_t seed,
                   const uint64_t* seed_dev, uint32_t site, void* stream) {
  CORRIF_REQUIRE(x && out && n > 0 && n % 4 == 0, "dropout: n must be a positive multiple of 4");
  CORRIF_REQUIRE(p >= 0.f && p < 1.f, "dropout: p");
  dropout_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      x, out, n / 4, dropout_threshold(p), 1.0f / (1.0f - p), seed, seed_dev, site, 0);
  return launch_status("dropout");
}

int corrif_dropout_colsum(const float* x, float* out, int64_t rows, int32_t cols, float p, uint64_t seed,
                          const uint64_t* seed_dev, uint32_t site, float* colsum, void* stream) {
  CORRIF_REQUIRE(x && out && colsum && rows > 0, "dropout_colsum: null/empty");
  CORRIF_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 1024 && 256 % (cols / 4) == 0,
                 "dropout_colsum: cols / 4 must divide 256 (got cols = %d)", cols);
  CORRIF_REQUIRE(p >= 0.f && p < 1.f, "dropout_colsum: p");
  CORRIF_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)colsum % 16 == 0),
                 "dropout_colsum: alignment");
  const int cq = cols / 4, rpi = 256 / cq;
  int64_t blocks = (rows + 4 * rpi - 1) / (4 * rpi);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  dropout_colsum_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      x, out, rows, cq, dropout_threshold(p), 1.0f / (1.0f - p), seed, seed_dev, site, colsum);
  return launch_status("dropout_colsum");
}

int corrif_dropout_mask(float* mask, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev,
                        uint32_t site, void* stream) {
  CORRIF_REQUIRE(mask && n > 0 && n % 4 == 0, "dropout_mask: n must be a positive multiple of 4");
  CORRIF_REQUIRE(p >= 0.f && p < 1.f, "dropout_mask: p");
  dropout_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      nullptr, mask, n / 4, dropout_threshold(p), 1.0f, seed, seed_dev, site, 1);
  return launch_status("dropout_mask");
}

int corrif_dropout_add(const float* x, const float* res, float* out, int64_t n, float p,
                       uint64_t seed, const uint64_t* seed_dev, uint32_t site_a, uint32_t site_b,
                       void* stream) {
  CORRIF_REQUIRE(x && out && n > 0 && n % 4 == 0, "dropout_add: n must be a positive multiple of 4");
  CORRIF_REQUIRE(p >= 0.f && p < 1.f, "dropout_add: p");
  dropout_add_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      x, res, out, n / 4, dropout_threshold(p), 1.0f / (1.0f - p), seed, seed_dev, site_a, site_b);
  return launch_status("dropout_add");
}

static int colsum_chunks(int64_t rows, int32_t cols) {
  const int bx = (cols / 4 + 127) / 128;
  int64_t want = ((int64_t)num_sms() * 8 + bx - 1) / bx;
  if (want > (rows + 15) / 16) want = (rows + 15) / 16;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return (int)want;
}

int64_t corrif_colsum_scratch_floats(int64_t rows, int32_t cols) {
  (void)rows; (void)cols;
  return 4;
}

int corrif_colsum_batched(const float* x, int64_t ld, int64_t rows, int32_t cols, float* out, int32_t batch,
                          int64_t x_bstride, int64_t out_bstride, int accumulate, void* stream) {
  CORRIF_REQUIRE(x && out && rows > 0 && cols > 0 && cols % 4 == 0 && ld % 4 == 0,
                 "colsum: cols and ld must be multiples of 4");
  CORRIF_REQUIRE(batch >= 1 && batch <= 65535 && x_bstride % 4 == 0 && out_bstride % 4 == 0,
                 "colsum: batch strides must be multiples of 4");
  if (!accumulate) {
    for (int b = 0; b < batch; ++b) {
      cudaError_t e = cudaMemsetAsync(out + (int64_t)b * out_bstride, 0, (size_t)cols * sizeof(float),
                                      (cudaStream_t)stream);
      if (e != cudaSuccess) { set_last_error("colsum: memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
  }
  int chunks = colsum_chunks(rows, cols);
  if (batch > 1) chunks = (chunks + batch - 1) / batch < 8 ? 8 : (chunks + batch - 1) / batch;
  dim3 grid((cols / 4 + 127) / 128, chunks, batch);
  colsum_partial_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, ld, rows, cols, out, x_bstride, out_bstride);
  return launch_status("colsum");
}

int corrif_colsum(const float* x, int64_t ld, int64_t rows, int32_t cols, float* out,
                  int accumulate, float* scratch, void* stream) {
  (void)scratch;
  return corrif_colsum_batched(x, ld, rows, cols, out, 1, 0, 0, accumulate, stream);
}

int corrif_round_tf32_multi(const float* const* src, float* const* dst, const int64_t* n, int32_t count,
                            void* stream) {
  CORRIF_REQUIRE(src && dst && n && count > 0 && count <= 65535, "round_tf32_multi: bad arguments");
  dim3 grid(num_sms() * 2 / (count < 8 ? count : 8) + 1, count);
  round_tf32_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  return launch_status("round_tf32_multi");
}

int corrif_batchsum(const float* x, int64_t batch, int64_t stride, int64_t n, float* out,
                    int accumulate, void* stream) {
  CORRIF_REQUIRE(x && out && batch > 0 && n > 0 && n % 4 == 0 && stride % 4 == 0,
                 "batchsum: n and stride must be multiples of 4");
  batchsum_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, batch, stride, n / 4,
                                                                          out, accumulate);
  return launch_status("batchsum");
}

int corrif_add_rows(const float* a, int64_t lda, const float* b, int64_t ldb, float* out,
                    int64_t ldo, int64_t rows, int32_t cols, void* stream) {
  CORRIF_REQUIRE(a && b && out && rows > 0 && cols > 0 && cols % 4 == 0 && lda % 4 == 0 &&
                     ldb % 4 == 0 && ldo % 4 == 0, "add_rows: cols/ld must be multiples of 4");
  add_rows_kernel<<<grid_for(rows * (cols / 4), 256), 256, 0, (cudaStream_t)stream>>>(
      a, lda, b, ldb, out, ldo, rows, cols / 4);
  return launch_status("add_rows");
}

int corrif_bce_probs_fwd_bwd(const float* x, const float* y, int64_t n, float grad_scale,
                             double* loss_sum, float* dx, void* stream) {
  CORRIF_REQUIRE(x && y && loss_sum && n > 0, "bce: null/empty");
  bce_probs_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, grad_scale,
                                                                       loss_sum, dx);
  return launch_status("bce_probs");
}

int corrif_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                     float beta1, float beta2, float eps, float grad_scale, int32_t step,
                     void* stream) {
  CORRIF_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adam: null/empty/step");
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
  adam_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2,
                                                                  eps, grad_scale, bc1, bc2s);
  return launch_status("adam");
}

}  // extern "C"
