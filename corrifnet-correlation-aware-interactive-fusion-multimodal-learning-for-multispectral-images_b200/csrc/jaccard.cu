// Jaccard family (F5_JACCARD2.py:4-36) and the integer confusion matrix.
//
// jaccard_sums: one streaming pass, three double accumulators per thread, warp-shuffle + one
// double atomicAdd per block.  For {0,1}-valued inputs every partial is an exact integer, so the
// result is independent of summation order and equals the reference's fp32 sums bit for bit
// (while counts stay below 2^24).
// jaccard_finish: resolves the `if y.sum(0)==0` inversion (F5_JACCARD2.py:12-14) on the device,
// removing the reference's host sync, and evaluates the three ratios in fp32 with the reference's
// operation order (round-to-nearest intrinsics so nothing is contracted or re-associated).
// confusion_counts: K x K histogram of uint8 (label, pred) pairs, privatised per warp in shared
// memory, merged into 64-bit global counters.
#include "common.cuh"

namespace corrif {

__global__ void __launch_bounds__(256)
jaccard_sums_kernel(const float* __restrict__ y, const float* __restrict__ yp, int64_t P,
                    double* __restrict__ sums) {
  __shared__ double red[8][3];
  double sy = 0.0, sp = 0.0, syp = 0.0;
  const int64_t nq = P / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
    const float4 a = ld4_stream(y + q * 4), b = ld4_stream(yp + q * 4);
    // fp32 partials of 4 elements are exact for {0,1} data and accurate to 1 ulp otherwise
    sy += (double)((a.x + a.y) + (a.z + a.w));
    sp += (double)((b.x + b.y) + (b.z + b.w));
    syp += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(P - nq * 4)) {   // ragged tail (P % 4 elements)
    const float a = y[nq * 4 + threadIdx.x], b = yp[nq * 4 + threadIdx.x];
    sy += a; sp += b; syp += (double)a * b;
  }
  sy = warp_sum(sy); sp = warp_sum(sp); syp = warp_sum(syp);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[warp][0] = sy; red[warp][1] = sp; red[warp][2] = syp; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(sums + threadIdx.x, s);
  }
  if (blockIdx.x == 0 && threadIdx.x == 3) atomicAdd(sums + 3, (double)P);
}

// ------------------------------------------------------------------------------------------------
// Train-step tail in ONE pass (F4_TRAIN.py:58-71): BCE-with-logits of ALL B*CH*P output elements
// (loss sum in fp64, optional gradient) and, on channel 0 only, the three Jaccard sums that
// F4_TRAIN.py:70 feeds to Jaccard2 - the reference reads outputs and masks ~8 times for this and
// synchronises the host twice.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
loss_jaccard_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t nquads, int64_t Pq,
                    int CH, float grad_scale, double* __restrict__ loss_sum, float* __restrict__ dx,
                    double* __restrict__ sums, double pixels) {
  __shared__ double red[8][4];
  double ls = 0.0, sy = 0.0, sp = 0.0, syp = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += stride) {
    const float4 xv = ld4_stream(x + q * 4), yv = ld4_stream(y + q * 4);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
    float g[4], l4 = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      l4 += fmaxf(xs[e], 0.f) - xs[e] * ys[e] + log1pf(expf(-fabsf(xs[e])));   // as bce_probs_kernel
      g[e] = (1.0f / (1.0f + expf(-xs[e])) - ys[e]) * grad_scale;
    }
    ls += (double)l4;
    if (dx != nullptr) st4(dx + q * 4, make_float4(g[0], g[1], g[2], g[3]));
    if ((q / Pq) % CH == 0) {                       // channel 0: the plane the metric is taken on
      sy += (double)((ys[0] + ys[1]) + (ys[2] + ys[3]));
      sp += (double)((xs[0] + xs[1]) + (xs[2] + xs[3]));
      syp += (double)ys[0] * xs[0] + (double)ys[1] * xs[1] + (double)ys[2] * xs[2] + (double)ys[3] * xs[3];
    }
  }
  ls = warp_sum(ls); sy = warp_sum(sy); sp = warp_sum(sp); syp = warp_sum(syp);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[warp][0] = ls; red[warp][1] = sy; red[warp][2] = sp; red[warp][3] = syp; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(threadIdx.x == 0 ? loss_sum : sums + (threadIdx.x - 1), t);
  }
  if (blockIdx.x == 0 && threadIdx.x == 4) atomicAdd(sums + 3, pixels);   // number of metric pixels
}

__device__ __forceinline__ float jac_ratio(float tp, float fp, float fn, float eps) {
  // (TP+eps) / (TP+FP+FN+eps), left to right as F5_JACCARD2.py:19
  return __fdiv_rn(__fadd_rn(tp, eps), __fadd_rn(__fadd_rn(__fadd_rn(tp, fp), fn), eps));
}

__global__ void jaccard_finish_kernel(const double* __restrict__ sums, float eps,
                                      float* __restrict__ out) {
  const double sy = sums[0], sp = sums[1], syp = sums[2], P = sums[3];
  // plain Jaccard (F5_JACCARD2.py:4-9)
  const float tp = (float)syp, fp = (float)(sy - syp), fn = (float)(sp - syp);
  out[0] = jac_ratio(tp, fp, fn, eps);
  // Jaccard2 / JaccardAndF1: invert when the mask is empty (:12-14, :23-25)
  float tp2 = tp, fp2 = fp, fn2 = fn;
  if (sy == 0.0) {
    tp2 = (float)(P - sy - sp + syp);   // sum (1-y_pred)(1-y)
    fp2 = (float)(sp - syp);            // sum y_pred (1-y)      ("FP" = (1-y_pred')*y')
    fn2 = (float)(sy - syp);            // sum y (1-y_pred)      ("FN" = (1-y')*y_pred')
  }
  out[1] = jac_ratio(tp2, fp2, fn2, eps);
  const float recall = __fdiv_rn(tp2, __fadd_rn(__fadd_rn(tp2, fn2), eps));          // :33
  const float prec = __fdiv_rn(tp2, __fadd_rn(__fadd_rn(tp2, fp2), eps));            // :34
  out[2] = __fdiv_rn(__fmul_rn(2.0f, __fmul_rn(recall, prec)),
                     __fadd_rn(__fadd_rn(recall, prec), eps));                        // :35
}

constexpr int CM_MAX_K = 16;
constexpr int CM_WARPS = 8;

__global__ void __launch_bounds__(CM_WARPS * 32)
confusion_kernel(const uint8_t* __restrict__ label, const uint8_t* __restrict__ pred, int64_t P,
                 int K, unsigned long long* __restrict__ counts) {
  __shared__ unsigned int hist[CM_WARPS][CM_MAX_K * CM_MAX_K];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < CM_WARPS * CM_MAX_K * CM_MAX_K; i += blockDim.x)
    (&hist[0][0])[i] = 0u;
  __syncthreads();
  unsigned int* h = hist[warp];
  const int64_t nvec = P / 16;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint4* l16 = reinterpret_cast<const uint4*>(label);
  const uint4* p16 = reinterpret_cast<const uint4*>(pred);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 a = l16[i], b = p16[i];
    const unsigned int aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned int l = (aw[w] >> (8 * k)) & 0xffu, p = (bw[w] >> (8 * k)) & 0xffu;
        if (l < (unsigned)K && p < (unsigned)K) atomicAdd(&h[l * K + p], 1u);
      }
  }
  if (blockIdx.x == 0) {   // ragged tail
    for (int64_t i = nvec * 16 + threadIdx.x; i < P; i += blockDim.x) {
      const unsigned int l = label[i], p = pred[i];
      if (l < (unsigned)K && p < (unsigned)K) atomicAdd(&h[l * K + p], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) {
    unsigned long long s = 0;
#pragma unroll
    for (int w = 0; w < CM_WARPS; ++w) s += hist[w][i];
    if (s) atomicAdd(counts + i, s);
  }
}

}  // namespace corrif

using namespace corrif;

extern "C" {

int corrif_jaccard_sums(const float* y, const float* y_pred, int64_t P, double* sums,
                        void* stream) {
  CORRIF_REQUIRE(y && y_pred && sums && P > 0, "jaccard_sums: null/empty");
  CORRIF_REQUIRE(((uintptr_t)y % 16 == 0) && ((uintptr_t)y_pred % 16 == 0),
                 "jaccard_sums: inputs must be 16-byte aligned");
  int64_t blocks = (P / 4 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  jaccard_sums_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, y_pred, P, sums);
  return launch_status("jaccard_sums");
}

int corrif_loss_jaccard_fused(const float* x, const float* y, int64_t B, int32_t CH, int64_t P,
                              float grad_scale, double* loss_sum, float* dx, double* sums, void* stream) {
  CORRIF_REQUIRE(x && y && loss_sum && sums && B > 0 && CH > 0 && P > 0, "loss_jaccard: null/empty");
  CORRIF_REQUIRE(P % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) &&
                     (dx == nullptr || (uintptr_t)dx % 16 == 0),
                 "loss_jaccard: the plane size must be a multiple of 4 and the tensors 16-byte aligned");
  const int64_t nquads = B * CH * P / 4;
  int64_t blocks = (nquads + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  loss_jaccard_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, nquads, P / 4, CH, grad_scale,
                                                                          loss_sum, dx, sums, (double)B * (double)P);
  return launch_status("loss_jaccard");
}

int corrif_jaccard_finish(const double* sums, float epsilon, float* out3, void* stream) {
  CORRIF_REQUIRE(sums && out3, "jaccard_finish: null");
  jaccard_finish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sums, epsilon, out3);
  return launch_status("jaccard_finish");
}

int corrif_confusion_counts(const uint8_t* label, const uint8_t* pred, int64_t P,
                            int32_t num_classes, unsigned long long* counts, void* stream) {
  CORRIF_REQUIRE(label && pred && counts && P > 0, "confusion_counts: null/empty");
  CORRIF_REQUIRE(num_classes >= 1 && num_classes <= CM_MAX_K, "confusion_counts: 1 <= K <= 16");
  CORRIF_REQUIRE(((uintptr_t)label % 16 == 0) && ((uintptr_t)pred % 16 == 0),
                 "confusion_counts: inputs must be 16-byte aligned");
  int64_t blocks = (P / 16 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  confusion_kernel<<<(unsigned)blocks, CM_WARPS * 32, 0, (cudaStream_t)stream>>>(label, pred, P,
                                                                                num_classes, counts);
  return launch_status("confusion_counts");
}

}  // extern "C"
