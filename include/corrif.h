/*
 * corrif.h - C ABI of libcorrif_b200.so: the sm_100a kernels of CorrIFNet's fusion hot path.
 *
 * The reference (iremulku/CorrIFNet) has no FFI or operator-plugin interface: its seam is
 * Python.  Each entry point below therefore names the reference lines it
 * replaces; the Python host layer (corrif_b200.ops / dropin/) binds them with ctypes and
 * registers them as torch custom ops, which is what a maintainer of the reference would call
 * (INTEGRATION.md).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 data unless stated otherwise; the caller owns all
 *     memory (no allocation, no global state besides a per-thread last-error string);
 *   - `stream` is a cudaStream_t passed as void*; all work is stream ordered, no host sync;
 *   - return 0 on success, a negative CORRIF_E* for argument errors, or a positive cudaError_t;
 *   - "tokens" are row-major [rows, C] fp32 with token s = d*64 + h*8 + w (mmvit4.py:458-461).
 */
#ifndef CORRIF_H_
#define CORRIF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CORRIF_ABI_VERSION 3

#define CORRIF_EINVAL (-1)  /* bad argument (shape, alignment, null pointer) */
#define CORRIF_EARCH  (-2)  /* device is not sm_100 */
#define CORRIF_EDRIVER (-3) /* driver entry point (cuTensorMapEncodeTiled) unavailable */

int corrif_abi_version(void);
/* Last error message of the calling thread ("" if none). */
const char* corrif_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x). */
int corrif_check_device(void);

/* ------------------------------------------------------------------------------------------
 * GEMM  D = epilogue(alpha * A . B^T)   A: logical [M,K], B: logical [N,K], D: [M,N] row-major.
 * Replaces every nn.Linear / 1x1x1 Conv3d / bmm on the path: mmvit4.py:307,309,312,313 (attention),
 * :351,354 (FeedForward), :398-402 (encode convs), :417-419 (qkv convs), :426 (decode conv) and
 * their autograd backward (dgrad, wgrad).
 *
 * Operand storage:  *_mn_major == 0 : element (r,k) at ptr[r*ld + k]   (K-major, "row . row")
 *                   *_mn_major == 1 : element (r,k) at ptr[k*ld + r]   (MN-major, transposed)
 * Batching: grid of batch_outer x batch_inner problems; problem (bo,bi) uses
 *   A + bo*a_bo + bi*a_bi (element offsets), likewise B and D, aux, residual.
 * split_k > 1 splits the contraction across CTAs and REQUIRES epilogue CORRIF_EPI_ATOMIC_ADD
 *   (D += alpha*A.B^T with red.global.add; used for weight gradients, which also gives gradient
 *   accumulation across micro-batches).
 * precision: CORRIF_GEMM_TF32 = tcgen05.mma kind::tf32 (fp32 operands read as TF32, fp32
 *   accumulate in TMEM); CORRIF_GEMM_FP32 = CUDA-core FFMA (exact-fp32 checking mode).
 * ------------------------------------------------------------------------------------------ */
enum {
  CORRIF_EPI_STORE = 0,       /* D = v                                     */
  CORRIF_EPI_BIAS = 1,        /* D = v + bias[n]                            */
  CORRIF_EPI_BIAS_GELU = 2,   /* aux = v + bias[n]; D = gelu_erf(aux)   (mmvit4.py:345,351-352) */
  CORRIF_EPI_BIAS_RESIDUAL = 3, /* D = v + bias[n] + residual[m,n]       (mmvit4.py:322)        */
  CORRIF_EPI_MUL_DGELU = 4,   /* D = v * gelu_erf'(aux[m,n])            (backward of :345)     */
  CORRIF_EPI_ATOMIC_ADD = 5   /* D += v  (atomic)                                            */
};
enum { CORRIF_GEMM_TF32 = 0, CORRIF_GEMM_FP32 = 1 };
/* flags: round D to TF32 (nearest) at store time.  The tensor core truncates fp32 operands to TF32;
 * an output that is only consumed by further GEMMs is rounded once here instead. */
#define CORRIF_GEMM_ROUND_TF32 1

typedef struct corrif_gemm_desc {
  const float* A; const float* B; float* D;
  const float* bias;       /* [N] or NULL                                      */
  const float* residual;   /* [M,N] with ldr, or NULL                          */
  float* aux;              /* [M,N] with ldaux: written by BIAS_GELU, read by MUL_DGELU */
  int64_t lda, ldb, ldd, ldr, ldaux;
  int32_t M, N, K;
  int32_t a_mn_major, b_mn_major;
  int32_t batch_outer, batch_inner;         /* >= 1 */
  int64_t a_bo, a_bi, b_bo, b_bi, d_bo, d_bi; /* element offsets per batch index (D offsets also
                                                 apply to residual and aux)   */
  int32_t split_k;                          /* >= 1 */
  int32_t epilogue;
  int32_t precision;
  float alpha;
  int32_t flags;
  /* Dropout fused into the epilogue (drop_p > 0; not with ATOMIC_ADD): the value that the epilogue
   * would add the residual to / store is first multiplied by keep(site_a) * keep(site_b) / (1-p)^k,
   * k = number of sites != CORRIF_NO_SITE, with the element index m*N + n (requires ldd == N):
   *   BIAS_RESIDUAL: D = (v + bias) * keeps + residual   (proj_drop + PreNormDrop.dropout + Residual,
   *                                                       mmvit4.py:314,339,322; FeedForward :355,322)
   *   BIAS_GELU    : D = gelu(v + bias) * keep            (:352-353)
   *   MUL_DGELU    : D = v * gelu'(aux) * keep            (backward of :352-353)
   * Same decisions as corrif_dropout with the same (seed, seed_dev, site). */
  float drop_p;
  uint32_t drop_site_a, drop_site_b;
  uint32_t drop_site_bo;   /* batched problems: problem (bo, bi) draws from sites site_a/_b + bo * drop_site_bo */
  uint64_t drop_seed;
  const uint64_t* drop_seed_dev;
  int64_t bias_bo;         /* batched problems: problem (bo, bi) adds bias + bo * bias_bo (element offset) */
} corrif_gemm_desc;

int corrif_gemm(const corrif_gemm_desc* desc, void* stream);
/* sizeof(corrif_gemm_desc) as compiled into the library (binding sanity check). */
int corrif_sizeof_gemm_desc(void);

/* ------------------------------------------------------------------------------------------
 * Layout: batched 2-D transpose  in [batch, rows, cols] -> out [batch, cols, rows].
 * Replaces .permute(0,2,3,4,1).contiguous() / .permute(0,4,1,2,3).contiguous()
 * (mmvit4.py:459-461, 474-475, 499-501, 511-513, 526-528).
 * ------------------------------------------------------------------------------------------ */
int corrif_transpose(const float* in, float* out, int64_t batch, int32_t rows, int32_t cols,
                     int32_t round_tf32, void* stream);
/* out = round-to-nearest-TF32(in): makes GEMM-ready copies of weights (in place allowed).
 * _multi: `count` tensors in one launch; src/dst/n are DEVICE arrays of pointers / element counts;
 *         a NEGATIVE count -n copies n elements unrounded (stacked bias copies for batched GEMMs). */
int corrif_round_tf32(const float* in, float* out, int64_t n, void* stream);
int corrif_round_tf32_multi(const float* const* src, float* const* dst, const int64_t* n, int32_t count,
                            void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (C == 512, eps 1e-5, affine) with the positional add fused:
 *   x1 = x + pos[row % pos_rows]      (mmvit4.py:385; pos == NULL: x1 = x, x1_out may be NULL)
 *   y  = LN(x1) * gamma + beta         (mmvit4.py:327-339)
 * mean / rstd [rows] are saved for the backward.  round_tf32 != 0 rounds y (not x1) to TF32.
 * ------------------------------------------------------------------------------------------ */
int corrif_layernorm_fwd(const float* x, const float* pos, int64_t pos_rows, const float* gamma,
                         const float* beta, float* x1_out, float* y, float* mean, float* rstd,
                         int64_t rows, int32_t C, int32_t round_tf32, void* stream);
/* dx = LN'(dy) (+ dres if not NULL).  dgamma/dbeta [C] are overwritten, or added to when
 * accumulate != 0 (block totals are added with red.global.add: the summation order is not fixed).
 * `scratch` is unused (kept for source compatibility).  When dx_drop != NULL the kernel also writes
 * dx_drop = dx * keep(site_a) * keep(site_b) / (1-p)^k - the backward of the dropouts that followed
 * this tensor in the forward (mmvit4.py:314,339) - with the decisions of corrif_dropout_add. */
int64_t corrif_layernorm_bwd_scratch_floats(int64_t rows, int32_t C);
int corrif_layernorm_bwd(const float* dy, const float* x1, const float* gamma, const float* mean,
                         const float* rstd, const float* dres, float* dx, float* dgamma,
                         float* dbeta, float* scratch, int64_t rows, int32_t C, int32_t accumulate,
                         float* dx_drop, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                         uint32_t site_a, uint32_t site_b, void* stream);

/* Same, with the rows of dx regrouped: the input rows are [batch][group][group_rows] (the multimodal token
 * set: group = modality / fused token set, mmvit4.py:515-522) and dx is written [group][batch][group_rows],
 * the layout the consumers of the token gradient use - the strided copy the host did before disappears.
 * groups == 0: no regrouping; regrouping needs dx_drop == NULL and dx distinct from dy / dres.
 * dres2 (may be NULL): a second incoming gradient of the same tensor, added like dres (the skip path of
 * mmvit4.py:505 - the host's separate add pass disappears). */
int corrif_layernorm_bwd_regroup(const float* dy, const float* x1, const float* gamma, const float* mean,
                                 const float* rstd, const float* dres, float* dx, float* dgamma,
                                 float* dbeta, float* scratch, int64_t rows, int32_t C, int32_t accumulate,
                                 float* dx_drop, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                                 uint32_t site_a, uint32_t site_b, int32_t groups, int32_t group_rows,
                                 const float* dres2, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row softmax in place, for the materialised attention path: P = softmax(S) over `cols`
 * (mmvit4.py:310; the 0.125 scale of :309 is folded into the producing GEMM's alpha), with the
 * optional attention dropout of :311: when p_drop > 0 the dropped-and-rescaled probabilities
 * (counter RNG keyed by seed/site/element index, scaled 1/(1-p)) are written to Pdrop, while S keeps
 * the un-dropped P that the backward needs.
 * Backward in place on dP:  dS = P * (dP*keep/(1-p) - sum_j(dP*keep/(1-p)*P)) * scale.
 * ------------------------------------------------------------------------------------------ */
int corrif_softmax_fwd(float* S, float* Pdrop, int64_t rows, int32_t cols, float p_drop,
                       uint64_t seed, const uint64_t* seed_dev, uint32_t site, int32_t round_tf32,
                       void* stream);
int corrif_softmax_bwd(const float* P, float* dP, int64_t rows, int32_t cols, float scale,
                       float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t site,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-head self-attention (SelfAttention.forward, mmvit4.py:307-312, and its autograd
 * backward): per (batch, head)  O = dropout(softmax(Q K^T * scale)) V  with Q, K, V the strided
 * column blocks [0,C), [C,2C), [2C,3C) of qkv [B*N, 3C] (head h = columns h*D..h*D+D of a block),
 * O / dO [B*N, C], dqkv [B*N, 3C].  tcgen05 (TF32 operands, fp32 accumulate in TMEM); the N x N
 * probabilities never reach HBM.  D must be 64, N a multiple of 128.
 *   lse      [B*H, N]        log2-domain log-sum-exp, written by fwd, read by bwd
 *   maskbits [B*H, N, N/32]  dropout keep bits (bit kv%32 of word kv/32), written by fwd when
 *                            p_drop > 0 (counter RNG keyed by seed/site/element index as corrif_dropout),
 *                            read by bwd; may be NULL when p_drop == 0
 *   delta    [B*H, N]        scratch of bwd (rowsum(dO*O))
 * Several independent attention modules of the same shape (the three intra-modal transformers) run
 * as ONE launch over their stacked buffers: with group_batches = G > 0, batch b belongs to module
 * b / G and draws its dropout decisions from site + (b / G) * group_site_stride with the element
 * index of batch b % G - exactly what G-batch launches per module would draw.  0 = one module.
 * ------------------------------------------------------------------------------------------ */
int corrif_attention_fwd(const float* qkv, float* O, float* lse, uint32_t* maskbits, int32_t B,
                         int32_t N, int32_t H, int32_t D, float scale, float p_drop, uint64_t seed,
                         const uint64_t* seed_dev, uint32_t site, int32_t group_batches,
                         uint32_t group_site_stride, int32_t round_tf32, void* stream);
/* The forward's dropout decisions as a stand-alone pass, and the forward that consumes them: the keep bits
 * depend on (seed, site, element) only, so a host can produce them on a second stream while the GEMMs that
 * precede the attention run, and keep the counter hash out of the forward's issue-bound softmax loop.
 * corrif_attention_keepbits writes exactly the words corrif_attention_fwd would store for the same arguments;
 * corrif_attention_fwd_premasked == corrif_attention_fwd with those words read instead of regenerated. */
int corrif_attention_keepbits(uint32_t* maskbits, int32_t B, int32_t N, int32_t H, float p_drop, uint64_t seed,
                              const uint64_t* seed_dev, uint32_t site, int32_t group_batches,
                              uint32_t group_site_stride, int32_t max_blocks /* 0 = fill the GPU */, void* stream);
int corrif_attention_fwd_premasked(const float* qkv, float* O, float* lse, const uint32_t* maskbits,
                                   int32_t B, int32_t N, int32_t H, int32_t D, float scale,
                                   float p_drop, int32_t round_tf32, void* stream);
int corrif_attention_bwd(const float* qkv, const float* O, const float* dO, const float* lse,
                         const uint32_t* maskbits, float* delta, float* dqkv, int32_t B, int32_t N,
                         int32_t H, int32_t D, float scale, float p_drop, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dropout  out = x * keep(seed, site, index) / (1-p)   (nn.Dropout sites mmvit4.py:311,314,339,
 * 353,355).  Counter based (SplitMix64 hash of seed, site and element index, 16-bit draws; p is realised as
 * round(p*65536)/65536): the same call on a gradient is the backward.
 * In place allowed.  corrif_dropout_mask writes the 0/1 keep mask itself (for tests).
 * corrif_dropout_add: out = x * keep_a * keep_b / (1-p)^2 + res  - the two stacked dropouts of the
 *   attention branch (proj_drop :314 then PreNormDrop.dropout :339) and the residual add (:322) in
 *   one pass; site_b == CORRIF_NO_SITE applies one mask only (FeedForward output, :355 + :322).
 * Effective seed = seed + (seed_dev ? *seed_dev : 0): a device-resident step counter lets a
 *   captured CUDA graph draw fresh masks on every replay.
 * ------------------------------------------------------------------------------------------ */
#define CORRIF_NO_SITE 0xFFFFFFFFu
int corrif_dropout(const float* x, float* out, int64_t n, float p, uint64_t seed,
                   const uint64_t* seed_dev, uint32_t site, void* stream);
/* out = dropout(x) over a [rows, cols] matrix and colsum[c] += sum_r out[r, c] in the same pass (the bias
 * gradient of the Linear the dropped gradient feeds: backward of mmvit4.py:354-355).  Keep decisions are
 * those of corrif_dropout on the flat tensor.  cols / 4 must divide 256; colsum is ACCUMULATED into. */
int corrif_dropout_colsum(const float* x, float* out, int64_t rows, int32_t cols, float p, uint64_t seed,
                          const uint64_t* seed_dev, uint32_t site, float* colsum, void* stream);
int corrif_dropout_mask(float* mask, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev,
                        uint32_t site, void* stream);
int corrif_dropout_add(const float* x, const float* res, float* out, int64_t n, float p,
                       uint64_t seed, const uint64_t* seed_dev, uint32_t site_a, uint32_t site_b,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Reductions used by the backward:
 *   colsum:   out[c] (+)= sum_r x[r*ld + c]          bias gradients
 *   batchsum: out[i] (+)= sum_b x[b*stride + i]      positional-embedding gradients
 *   add:      out[i] = a[i] + b[i]
 * colsum is one launch: per-block partial column sums are added to `out` with red.global.add (the
 * summation order across blocks is not fixed); accumulate == 0 zeroes `out` first.  `scratch` is
 * unused (kept for ABI stability; corrif_colsum_scratch_floats returns a small constant).
 * ------------------------------------------------------------------------------------------ */
int64_t corrif_colsum_scratch_floats(int64_t rows, int32_t cols);
int corrif_colsum(const float* x, int64_t ld, int64_t rows, int32_t cols, float* out,
                  int accumulate, float* scratch, void* stream);
/* `batch` independent column sums in one launch: problem b reads x + b*x_bstride and adds into
 * out + b*out_bstride (the three intra-modal branches' bias gradients). */
int corrif_colsum_batched(const float* x, int64_t ld, int64_t rows, int32_t cols, float* out, int32_t batch,
                          int64_t x_bstride, int64_t out_bstride, int accumulate, void* stream);
int corrif_batchsum(const float* x, int64_t batch, int64_t stride, int64_t n, float* out,
                    int accumulate, void* stream);
/* out[r*ldo + c] = a[r*lda + c] + b[r*ldb + c]   (rows x cols, cols % 4 == 0) */
int corrif_add_rows(const float* a, int64_t lda, const float* b, int64_t ldb, float* out,
                    int64_t ldo, int64_t rows, int32_t cols, void* stream);

/* ------------------------------------------------------------------------------------------
 * Inter-modal correlation (InterFormer), mmvit4.py:481-507, including the batch-mixing .view of
 * :485 and the skip add of :505-507, written straight into the
 * multimodal token buffer (which removes the cats of :515-521):
 *   qkv   [M, B, S, 3C]  q = cols [0,C), k = [C,2C), v = [2C,3C)   (outputs of the qkv_* convs)
 *   skip  [M, B, S, C]   pre-transformer tokens (mmvit4.py:462)
 *   tokens[B, (M+1)*S, C] rows X*S+s receive skip_X + sum_i A_X[m,b] * v_i[b'],
 *                          (m,b) = divmod(M*b'+i, B), A_X[:,b] = softmax_m(q_X[b]*k_m[b]/sqrt(M)).
 * Backward: g = dL/dtokens (same layout) -> dqkv [M,B,S,3C] (overwritten).  dskip_X = g_X.
 * M = 2..6 modalities (the reference hard-wires 3, mmvit4.py:15; BASELINE.json configs[4] runs 6).
 * ------------------------------------------------------------------------------------------ */
int corrif_inter_corr_fwd(const float* qkv, const float* skip, float* tokens, int32_t M,
                          int32_t B, int32_t S, int32_t C, void* stream);
int corrif_inter_corr_bwd(const float* qkv, const float* g_tokens, float* dqkv, int32_t M,
                          int32_t B, int32_t S, int32_t C, void* stream);
/* g_group_major != 0: g_tokens is [M+1][B][S][C] (see corrif_layernorm_bwd_regroup) instead of [B][(M+1)S][C] */
int corrif_inter_corr_bwd_layout(const float* qkv, const float* g_tokens, float* dqkv, int32_t M,
                                 int32_t B, int32_t S, int32_t C, int32_t g_group_major, void* stream);

/* ------------------------------------------------------------------------------------------
 * Jaccard family, F5_JACCARD2.py:4-36 / F5_JACCARD.py:4-9.
 * corrif_jaccard_sums: one pass over y, y_pred [P] fp32 producing, in double, sums[0..3] =
 *   sum(y), sum(y_pred), sum(y*y_pred), P.  From these (all exact integers for {0,1} inputs)
 *   TP = s2, "FP" = s0 - s2, "FN" = s1 - s2, and for the empty-mask inversion of :12-14
 *   TP' = P - s1, "FP'" = s1, "FN'" = 0 - resolved on device by corrif_jaccard_finish, which
 *   writes out[0] = Jaccard, out[1] = Jaccard2, out[2] = JaccardAndF1 (fp32 arithmetic as :19,
 *   :33-35) without a host sync.  sums must be zeroed by the caller (or accumulate over calls).
 * corrif_confusion_counts: K x K integer confusion matrix (rows label, cols pred) of uint8 class
 *   maps, shared-memory privatised; counts [K*K] uint64 accumulate (caller zeroes).
 * ------------------------------------------------------------------------------------------ */
int corrif_jaccard_sums(const float* y, const float* y_pred, int64_t P, double* sums,
                        void* stream);
/* Train-step tail in one pass (F4_TRAIN.py:58-71): x = model outputs [B, CH, P] (sigmoid probabilities
 * fed to BCEWithLogitsLoss as the reference does), y = masks.  Adds sum_i bce(x_i, y_i) over ALL
 * elements to *loss_sum (fp64), writes dx = (sigmoid(x) - y) * grad_scale when dx != NULL, and adds the
 * Jaccard sums of channel 0 (sum y, sum x, sum x*y, B*P) to sums[0..3] - feed those to
 * corrif_jaccard_finish.  loss_sum and sums must be zeroed by the caller.  P % 4 == 0. */
int corrif_loss_jaccard_fused(const float* x, const float* y, int64_t B, int32_t CH, int64_t P,
                              float grad_scale, double* loss_sum, float* dx, double* sums, void* stream);
int corrif_jaccard_finish(const double* sums, float epsilon, float* out3, void* stream);
int corrif_confusion_counts(const uint8_t* label, const uint8_t* pred, int64_t P,
                            int32_t num_classes, unsigned long long* counts, void* stream);

/* ------------------------------------------------------------------------------------------
 * Train-step tail, F4_TRAIN.py:58-62: BCEWithLogitsLoss applied to the sigmoid output, and Adam.
 *   bce_probs_fwd_bwd: loss_sum (double, accumulates; caller zeroes) += sum softplus(x) - x*y;
 *                      dx = (sigmoid(x) - y) * grad_scale      (grad_scale = 1/numel)
 *   adam_step: torch.optim.Adam defaults (F2_MAIN.py:168-169), in place on p, m, v.
 * ------------------------------------------------------------------------------------------ */
int corrif_bce_probs_fwd_bwd(const float* x, const float* y, int64_t n, float grad_scale,
                             double* loss_sum, float* dx, void* stream);
int corrif_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                     float beta1, float beta2, float eps, float grad_scale, int32_t step,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * Channels-last 3-D volume operators: the blocks either side of the fusion hot path (SURVEY.md section 8f rows
 * N1 / N2).  A "volume" is [B, D, H, W, C] fp32 with channels contiguous and `ld` floats between consecutive
 * voxels (ld >= C: a channel slice of a wider buffer is a valid volume, so the reference's torch.cat calls
 * (mmvit4.py:77, 272-287) are just two producers writing into one buffer, or one consumer reading several
 * sources).  C and ld are multiples of 4, pointers 16-byte aligned.
 *
 * corrif_conv3d_fwd: out = act(conv(cat(src...), W) + bias), stride 1, "same" padding, kernel 1x1x1 or 3x3x3,
 *   as warp-level TF32 tensor-core MMAs over an input window staged once in shared memory.  Replaces nn.Conv3d
 *   (+ F.pad replicate) + ReLU of general_conv3d_prenorm (mmvit4.py:29-45), EarlyFusionBlock (:64-81) and the
 *   bias-only 1x1x1 convs (:231, :264, adapt1-5/conv6 :157-164).  If `stats` is given, sum and sum of squares of
 *   the stored output are accumulated per (sample, channel): the InstanceNorm3d statistics (mmvit4.py:24) come
 *   out of the convolution's epilogue instead of a second reduction pass.
 *   The SAME kernel computes the data gradient: run it on d(pre-activation) with weights packed with
 *   `transpose_flip` = 1 and zero padding; for replicate padding corrif_conv3d_dgrad_border then adds what the
 *   clamped border taps contribute (replication_pad3d_backward without a padded tensor).
 * corrif_conv3d_wgrad: dW[co][ci][tap] += sum_voxels x[voxel + tap][ci] * g[voxel][co]  (persistent CTAs keep the
 *   partial weight gradient in registers across tiles; one atomic add per element and CTA at the end).
 * ------------------------------------------------------------------------------------------ */
typedef struct corrif_vol_src {
  const float* p;
  int32_t C;
  int32_t reserved;
  int64_t ld;
} corrif_vol_src;

/* CORRIF_PAD_REPLICATE_ADJOINT (corrif_conv3d_tc_fwd only): the adjoint of replicate padding - what the data gradient
 * of a replicate-padded convolution applies at the volume border (a clamped tap uses the mirrored weight). */
enum { CORRIF_PAD_ZEROS = 0, CORRIF_PAD_REPLICATE = 1, CORRIF_PAD_REPLICATE_ADJOINT = 2 };

typedef struct corrif_conv3d_desc {
  corrif_vol_src src[3];      /* inputs, concatenated along channels in this order      */
  int32_t nsrc;
  int32_t B, D, H, W;
  int32_t Cin, Cout;          /* Cin = sum of src[i].C; both multiples of 8             */
  int32_t ksize;              /* 1 or 3                                                 */
  int32_t pad_mode;           /* CORRIF_PAD_* (ignored for ksize 1)                     */
  int32_t relu;               /* apply max(.,0) after the bias                          */
  const float* wpk;           /* weights packed by corrif_conv3d_pack_weights           */
  const float* bias;          /* [Cout] or NULL                                         */
  float* out;                 /* [B,D,H,W,Cout] with voxel stride ldo                   */
  int64_t ldo;
  double* stats;              /* [B][Cout][2] += (sum, sum of squares) of out, or NULL  */
} corrif_conv3d_desc;

int corrif_sizeof_conv3d_desc(void);
/* floats needed for the packed form of a [Cout, Cin, k, k, k] weight (0 on bad arguments) */
int64_t corrif_conv3d_pack_floats(int32_t Cin, int32_t Cout, int32_t ksize);
/* w: torch layout [Cout][Cin][k][k][k].  transpose_flip = 0: forward operand.  transpose_flip = 1: the operand of
 * the data gradient (in/out channels swapped, taps mirrored); Cin / Cout are still those of the forward conv. */
int corrif_conv3d_pack_weights(const float* w, float* wpk, int32_t Cin, int32_t Cout, int32_t ksize,
                               int32_t transpose_flip, void* stream);
int corrif_conv3d_fwd(const corrif_conv3d_desc* desc, void* stream);
/* desc: src / geometry / ksize / pad_mode as in the forward (out, wpk, bias, stats ignored); g = d(pre-activation)
 * [B,D,H,W,Cout] with stride ldg; dW: torch layout, accumulated. */
int corrif_conv3d_wgrad(const corrif_conv3d_desc* desc, const float* g, int64_t ldg, float* dW, void* stream);
/* The same 3x3x3 convolution as a tcgen05 "line convolution" (csrc/conv3d_tc.cu): a row of 128 voxels along x is the
 * M dimension of the MMA, the three x-taps are folded into N (= 3 * Cout), every input line is staged once by TMA and
 * feeds the TMEM accumulators of the nine output lines it touches; the x-tap sum, bias, ReLU and the InstanceNorm
 * statistics run in the epilogue.  Replaces, for the shapes it supports, corrif_conv3d_fwd (forward: mmvit4.py:29-45,
 * 222-292) and corrif_conv3d_fwd + corrif_conv3d_dgrad_border (data gradient: describe the gradient as the source,
 * Cin / Cout swapped, pad_mode CORRIF_PAD_REPLICATE_ADJOINT for a replicate-padded forward, weights packed with
 * transpose_flip = 1).  Supported: ksize 3, W in {16, 32, 64, 128} with B a multiple of 128 / W, source channel counts
 * that are multiples of 8 (a concatenation runs as uniform 8 / 16 / 32-channel pieces), Cout a multiple of 8 (processed
 * in chunks of 32, 16 or 8 output channels), and the packed weights of one chunk resident in shared memory beside
 * three line buffers (corrif_conv3d_tc_supported returns 1).  desc->wpk must come from corrif_conv3d_tc_pack_weights
 * for a descriptor with the same channel split; `w` there is always the forward weight [Cout_fwd][Cin_fwd][3][3][3]. */
int corrif_conv3d_tc_supported(const corrif_conv3d_desc* desc);
int64_t corrif_conv3d_tc_pack_floats(const corrif_conv3d_desc* desc);
int corrif_conv3d_tc_pack_weights(const corrif_conv3d_desc* desc, const float* w, float* wpk, int32_t transpose_flip,
                                  void* stream);
int corrif_conv3d_tc_fwd(const corrif_conv3d_desc* desc, void* stream);
/* The weight gradient of the big 3x3x3 layers on tcgen05 (csrc/conv3d_wgrad_tc.cu): transposer warps rewrite the input
 * and gradient lines as K(= voxel)-major swizzled tiles, one warp issues M = 128 ((4 input lines) x (32 channels))
 * x N = 8 * Cout MMAs whose three accumulators (z-taps) stay in TMEM for the whole kernel.  Same contract as
 * corrif_conv3d_wgrad (dW accumulated, torch layout).  Supported: ksize 3, W a multiple of 64, H even, Cout 8 or 16,
 * Cin <= 64 (32 channels per pass), pad_mode zeros / replicate. */
int corrif_conv3d_wgrad_tc_supported(const corrif_conv3d_desc* desc);
int corrif_conv3d_wgrad_tc(const corrif_conv3d_desc* desc, const float* g, int64_t ldg, float* dW, void* stream);
/* dx[B,D,H,W,Cin] (stride ldx) += the replicate-padding part of the 3x3x3 data gradient; w_taps_major is the weight
 * transposed to [27][Cout][Cin] (weight.permute(2,3,4,0,1)), so that threads of consecutive ci read consecutive words */
int corrif_conv3d_dgrad_border(const float* g, int64_t ldg, const float* w_taps_major, float* dx, int64_t ldx, int32_t B,
                               int32_t D, int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* stream);

/* InstanceNorm3d (affine=False, eps, biased variance; mmvit4.py:24) after ReLU, in place on a volume:
 * y = (r - mean) * rstd with the statistics the convolution accumulated; also writes mean / rstd [B][C] for the
 * backward.  Backward of ReLU -> InstanceNorm given dy and the saved OUTPUT y (the ReLU mask is recovered from y:
 * r > 0  <=>  y > (0 - mean) * rstd, evaluated with the forward's arithmetic):
 *   sums[b][c] = (sum dy, sum dy*y)                                              (..._bwd_stats)
 *   g = [r > 0] * rstd * (dy - sum_dy / n - y * sum_dyy / n);  dbias[c] += sum g   (..._bwd_apply) */
int corrif_instnorm_apply(float* x, int64_t ld, const double* stats, float* mean, float* rstd, int32_t B,
                          int64_t nvox, int32_t C, float eps, void* stream);
int corrif_instnorm_bwd_stats(const float* dy, int64_t lddy, const float* y, int64_t ldy, double* sums, int32_t B,
                              int64_t nvox, int32_t C, void* stream);
int corrif_instnorm_relu_bwd_apply(const float* dy, int64_t lddy, const float* y, int64_t ldy, const float* mean,
                                   const float* rstd, const double* sums, float* g, int64_t ldg, float* dbias,
                                   int32_t B, int64_t nvox, int32_t C, int32_t relu, void* stream);
/* dbias[c] += sum over voxels of g (bias gradient of a convolution without norm) */
int corrif_volume_colsum(const float* g, int64_t ldg, float* dbias, int64_t rows, int32_t C, void* stream);

/* Train-mode BatchNorm3d (+ residual add) (+ ReLU) on channels-last data [rows, C] (rows = B*D*H*W, C <= 1024 per call,
 * wider tensors in channel chunks): the encoders' Bottleneck3D (mmvit4.py:196-212: conv -> BN -> ReLU and
 * conv -> BN -> += identity -> ReLU) as one normalise pass forward and a statistics + apply pair backward, instead of
 * cuDNN batch-norm + separate add / ReLU / threshold_backward kernels.
 *   fwd:  stats[c] = (sum x, sum x^2) (from corrif_instnorm_bwd_stats(x, x) with B = 1);
 *         y = relu?((x - mean) * rstd * gamma + beta (+ res)); mean / var (biased) / rstd [C] are written.
 *   bwd:  g = dy * [y > 0]; sums[c] = (sum g, sum g * xhat); dx = gamma * rstd * (g - sum_g/n - xhat * sum_gx/n);
 *         dres = g (optional); dbeta = sum g; dgamma = sum g * xhat. */
int corrif_batchnorm_fwd(const float* x, int64_t ldx, const double* stats, const float* gamma, const float* beta,
                         const float* res, int64_t ldr, float* y, int64_t ldy, float* mean, float* var, float* rstd,
                         int64_t rows, int32_t C, float eps, int32_t relu, void* stream);
int corrif_batchnorm_bwd_stats(const float* dy, int64_t lddy, const float* y, int64_t ldy, const float* x, int64_t ldx,
                               const float* mean, const float* rstd, double* sums, int64_t rows, int32_t C, int32_t relu,
                               void* stream);
int corrif_batchnorm_bwd_apply(const float* dy, int64_t lddy, const float* y, int64_t ldy, const float* x, int64_t ldx,
                               const float* mean, const float* rstd, const float* gamma, const double* sums, float* dx,
                               int64_t lddx, float* dres, int64_t lddr, float* dgamma, float* dbeta, int64_t rows, int32_t C,
                               int32_t relu, void* stream);

/* Trilinear resize with align_corners=True (nn.Upsample / F.interpolate, mmvit4.py:187-191, 260, 269-285) and
 * nearest resize (F.interpolate default, mmvit4.py:271-286) on channels-last volumes, forward and backward
 * (backward in gather form: every input voxel sums its few contributing output voxels - no atomics; `dx` is
 * overwritten). */
int corrif_resize_trilinear_fwd(const float* x, int64_t ldx, float* y, int64_t ldy, int32_t B, int32_t C,
                                int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho, int32_t Wo, void* stream);
int corrif_resize_trilinear_bwd(const float* dy, int64_t lddy, float* dx, int64_t lddx, int32_t B, int32_t C,
                                int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho, int32_t Wo, void* stream);
/* 1x1x1 convolution with the same small channel count C (8 or 16) on both sides - the decoder's d1_out / d2_out blocks
 * (mmvit4.py:231-236, 8 -> 8 channels at 128^3): HBM-bound, so one thread per voxel with exact fp32 FMAs instead of
 * the window-staging tensor-core kernel.  out = act(W x + bias) with W = w [Cout][Cin] (transpose = 0) or w^T (the
 * data gradient); `stats` as in corrif_conv3d_fwd.  corrif_conv1_small_wgrad: dW[co][ci] += sum g[row][co] x[row][ci]
 * (C = 8). */
int corrif_conv1_small_fwd(const float* x, int64_t ldx, const float* w, const float* bias, float* out, int64_t ldo,
                           double* stats, int32_t B, int64_t nvox, int32_t C, int32_t relu, int32_t transpose,
                           void* stream);
int corrif_conv1_small_wgrad(const float* x, int64_t ldx, const float* g, int64_t ldg, float* dW, int64_t rows, int32_t C,
                             void* stream);
/* One axis of the trilinear resize (same align_corners arithmetic) on a contiguous tensor [outer][n][inner], inner a
 * multiple of 4 floats: the separable form of nn.Upsample(scale_factor=2, trilinear, align_corners=True)
 * (mmvit4.py:269) - three streaming passes instead of 8 gathered loads per output (forward) / 64 per input (backward).
 * The adjoint needs (n_out - 1) <= 3 (n_in - 1). */
int corrif_resize_linear_axis_fwd(const float* x, float* y, int64_t outer, int32_t n_in, int32_t n_out, int64_t inner,
                                  void* stream);
int corrif_resize_linear_axis_bwd(const float* dy, float* dx, int64_t outer, int32_t n_in, int32_t n_out, int64_t inner,
                                  void* stream);
int corrif_resize_nearest_fwd(const float* x, int64_t ldx, float* y, int64_t ldy, int32_t B, int32_t C,
                              int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho, int32_t Wo, void* stream);
int corrif_resize_nearest_bwd(const float* dy, int64_t lddy, float* dx, int64_t lddx, int32_t B, int32_t C,
                              int32_t Di, int32_t Hi, int32_t Wi, int32_t Do, int32_t Ho, int32_t Wo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CORRIF_H_ */
