"""Tensor-level wrappers over the C ABI (include/corrif.h): torch owns memory and streams, the
library does the work.  Every function launches on ``torch.cuda.current_stream()`` and never syncs.

No CPU path: tensors must be CUDA fp32 contiguous, otherwise ``CorrifError``/``ValueError``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple, Union

import torch

from . import _lib as L
from ._lib import (EPI_ATOMIC_ADD, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_MUL_DGELU,  # noqa: F401
                   EPI_STORE, GEMM_FP32, GEMM_ROUND_TF32, GEMM_TF32, NO_SITE, CorrifError, GemmDesc)

TensorOrView = Union[torch.Tensor, Tuple[torch.Tensor, int]]  # (tensor, element offset)

_launches = 0          # number of library kernels enqueued (bench.py reports it as gpu_launches)


def launch_count() -> int:
    return _launches


def _count(n: int = 1):
    global _launches
    _launches += n


# ---- optional per-launch device timing (bench.py's roofline leg) ---------------------------------
_prof = None     # list of (kernel class, algorithmic work, start event, end event) while profiling


class profile:
    """``with ops.profile() as rec:`` brackets every library launch with CUDA events on the launching
    stream.  ``rec.summary()`` (after a sync) returns {class: (launches, ms, work)} where work is
    algorithmic FLOPs for GEMMs and algorithmic bytes for the HBM-bound kernels."""

    def __enter__(self):
        global _prof
        _prof = []
        self.records = _prof
        return self

    def __exit__(self, *exc):
        global _prof
        _prof = None

    def details(self):
        """[(class, detail string, ms, work)] per launch, in launch order."""
        torch.cuda.synchronize()
        return [(r[0], r[4] if len(r) > 4 else "", r[2].elapsed_time(r[3]), r[1]) for r in self.records]

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for cls, work, e0, e1, *_ in self.records:
            n, ms, w = out.get(cls, (0, 0.0, 0.0))
            out[cls] = (n + 1, ms + e0.elapsed_time(e1), w + work)
        return out


class _Rec:
    __slots__ = ("cls", "work", "e0", "detail")

    def __init__(self, cls, work, detail=""):
        self.cls, self.work, self.detail = cls, work, detail

    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e0.record()

    def __exit__(self, *exc):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        if _prof is not None:
            _prof.append((self.cls, self.work, self.e0, e1, self.detail))


class _NoRec:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NOREC = _NoRec()


def _rec(cls: str, work: float, detail: str = ""):
    return _NOREC if _prof is None else _Rec(cls, work, detail)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[TensorOrView], dtype=torch.float32) -> Optional[int]:
    if t is None:
        return None
    off = 0
    if isinstance(t, tuple):
        t, off = t
    if not t.is_cuda:
        raise ValueError("corrif ops need CUDA tensors (there is no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError("expected %s, got %s" % (dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("corrif ops need contiguous tensors")
    return t.data_ptr() + off * t.element_size()


def lib():
    return L.load()


def check_device():
    L.check(lib().corrif_check_device(), "corrif_check_device")


# ------------------------------------------------------------------------------------------------
def gemm(A: TensorOrView, B: TensorOrView, D: TensorOrView, *, M: int, N: int, K: int, lda: int,
         ldb: int, ldd: int, a_mn: bool = False, b_mn: bool = False, bias=None, residual=None,
         ldr: int = 0, aux=None, ldaux: int = 0, batch=(1, 1), a_step=(0, 0), b_step=(0, 0),
         d_step=(0, 0), split_k: int = 1, epilogue: int = EPI_STORE, precision: int = GEMM_TF32,
         alpha: float = 1.0, round_out: bool = False, tag: str = "", drop_p: float = 0.0,
         drop_sites=(NO_SITE, NO_SITE), drop_seed: int = 0, drop_seed_dev=None, bias_step: int = 0,
         drop_site_step: int = 0):
    """D = epilogue(alpha * A . B^T); see corrif_gemm in include/corrif.h for the layout rules."""
    g = GemmDesc()
    g.A, g.B, g.D = _ptr(A), _ptr(B), _ptr(D)
    g.bias, g.residual, g.aux = _ptr(bias), _ptr(residual), _ptr(aux)
    g.lda, g.ldb, g.ldd, g.ldr, g.ldaux = lda, ldb, ldd, ldr, ldaux
    g.M, g.N, g.K = M, N, K
    g.a_mn_major, g.b_mn_major = int(a_mn), int(b_mn)
    g.batch_outer, g.batch_inner = batch
    g.a_bo, g.a_bi = a_step
    g.b_bo, g.b_bi = b_step
    g.d_bo, g.d_bi = d_step
    g.split_k, g.epilogue, g.precision, g.alpha = split_k, epilogue, precision, alpha
    g.flags = GEMM_ROUND_TF32 if round_out else 0
    if drop_p > 0:
        g.drop_p, g.drop_site_a, g.drop_site_b = drop_p, drop_sites[0], drop_sites[1]
        g.drop_seed, g.drop_seed_dev = drop_seed, _seed_dev(drop_seed_dev)
        g.drop_site_bo = drop_site_step
    g.bias_bo = bias_step
    cls = ("gemm_tf32" if precision == GEMM_TF32 else "gemm_fp32") + ("/" + tag if tag else "")
    det = "" if _prof is None else "M%d N%d K%d a%d b%d epi%d split%d z%d" % (
        M, N, K, int(a_mn), int(b_mn), epilogue, split_k, batch[0] * batch[1])
    with _rec(cls, 2.0 * M * N * K * batch[0] * batch[1], det):
        L.check(lib().corrif_gemm(C.byref(g), _stream()), "corrif_gemm")
    _count()


def transpose(x: TensorOrView, out: TensorOrView, batch: int, rows: int, cols: int, round_out=False):
    with _rec('transpose', 8.0 * batch * rows * cols):
        L.check(lib().corrif_transpose(_ptr(x), _ptr(out), batch, rows, cols, int(round_out), _stream()),
                "corrif_transpose")
    _count()


def round_tf32(x: TensorOrView, out: TensorOrView, n: int):
    with _rec('round_tf32', 8.0 * n):
        L.check(lib().corrif_round_tf32(_ptr(x), _ptr(out), n, _stream()), "corrif_round_tf32")
    _count()


def round_tf32_multi(src_ptrs, dst_ptrs, counts, count: int, total_elems: int):
    """src_ptrs/dst_ptrs/counts: int64 device tensors holding pointers / element counts."""
    with _rec('round_tf32', 8.0 * total_elems):
        L.check(lib().corrif_round_tf32_multi(_ptr(src_ptrs, torch.int64), _ptr(dst_ptrs, torch.int64),
                                              _ptr(counts, torch.int64), count, _stream()),
                "corrif_round_tf32_multi")
    _count()


def layernorm_fwd(x, pos, pos_rows, gamma, beta, x1_out, y, mean, rstd, rows, C_=512, round_out=False):
    with _rec('layernorm_fwd', (12.0 if pos is not None else 8.0) * rows * C_):
        L.check(lib().corrif_layernorm_fwd(_ptr(x), _ptr(pos), pos_rows, _ptr(gamma), _ptr(beta),
                                           _ptr(x1_out), _ptr(y), _ptr(mean), _ptr(rstd), rows, C_,
                                           int(round_out), _stream()), "corrif_layernorm_fwd")
    _count()


def layernorm_bwd_scratch_floats(rows, C_=512) -> int:
    return int(lib().corrif_layernorm_bwd_scratch_floats(rows, C_))


def layernorm_bwd(dy, x1, gamma, mean, rstd, dres, dx, dgamma, dbeta, scratch, rows, C_=512,
                  accumulate=False, dx_drop=None, p=0.0, seed=0, seed_dev=None, site_a=NO_SITE, site_b=NO_SITE,
                  groups=0, group_rows=0, dres2=None):
    """groups > 0: rows are [batch][group][group_rows] and dx is written [group][batch][group_rows];
    dres2: a second incoming gradient added like dres."""
    nbytes = (16.0 if dres is not None else 12.0) + (4.0 if dx_drop is not None else 0.0) + (4.0 if dres2 is not None else 0.0)
    with _rec('layernorm_bwd', nbytes * rows * C_):
        L.check(lib().corrif_layernorm_bwd_regroup(_ptr(dy), _ptr(x1), _ptr(gamma), _ptr(mean), _ptr(rstd),
                                                   _ptr(dres), _ptr(dx), _ptr(dgamma), _ptr(dbeta),
                                                   _ptr(scratch), rows, C_, int(accumulate), _ptr(dx_drop), p, seed,
                                                   _seed_dev(seed_dev), site_a, site_b, groups, group_rows, _ptr(dres2),
                                                   _stream()),
                    "corrif_layernorm_bwd_regroup")
    _count(1)


def _seed_dev(seed_dev):
    return None if seed_dev is None else _ptr(seed_dev, torch.int64)


def softmax_fwd(S, Pdrop, rows, cols, p=0.0, seed=0, seed_dev=None, site=0, round_out=False):
    with _rec('softmax_fwd', (8.0 if Pdrop is None else 12.0) * rows * cols):
        L.check(lib().corrif_softmax_fwd(_ptr(S), _ptr(Pdrop), rows, cols, p, seed, _seed_dev(seed_dev),
                                         site, int(round_out), _stream()), "corrif_softmax_fwd")
    _count()


def softmax_bwd(P, dP, rows, cols, scale, p=0.0, seed=0, seed_dev=None, site=0):
    with _rec('softmax_bwd', 12.0 * rows * cols):
        L.check(lib().corrif_softmax_bwd(_ptr(P), _ptr(dP), rows, cols, scale, p, seed,
                                         _seed_dev(seed_dev), site, _stream()), "corrif_softmax_bwd")
    _count()


def attention_fwd(qkv, O, lse, maskbits, B, N, H=8, D=64, scale=0.125, p=0.0, seed=0, seed_dev=None,
                  site=0, round_out=False, group_batches=0, group_site_stride=0):
    with _rec("attn_fwd", 4.0 * B * H * N * N * D, "B%d N%d" % (B, N)):
        L.check(lib().corrif_attention_fwd(_ptr(qkv), _ptr(O), _ptr(lse), _ptr(maskbits, torch.int32), B, N,
                                           H, D, scale, p, seed, _seed_dev(seed_dev), site, group_batches,
                                           group_site_stride, int(round_out), _stream()), "corrif_attention_fwd")
    _count()


def attention_keepbits(maskbits, B, N, H, p, seed, site, seed_dev=None, group_batches=0, group_site_stride=0,
                       max_blocks=0):
    """The dropout keep bits attention_fwd would store, as a stand-alone pass (data-independent)."""
    with _rec("attn_keepbits", 0.125 * B * H * N * N, "B%d N%d" % (B, N)):
        L.check(lib().corrif_attention_keepbits(_ptr(maskbits, torch.int32), B, N, H, p, seed, _seed_dev(seed_dev), site,
                                                group_batches, group_site_stride, max_blocks, _stream()),
                "corrif_attention_keepbits")
    _count()


def attention_fwd_premasked(qkv, O, lse, maskbits, B, N, H=8, D=64, scale=0.125, p=0.1, round_out=False):
    with _rec("attn_fwd", 4.0 * B * H * N * N * D, "B%d N%d premasked" % (B, N)):
        L.check(lib().corrif_attention_fwd_premasked(_ptr(qkv), _ptr(O), _ptr(lse), _ptr(maskbits, torch.int32), B, N,
                                                     H, D, scale, p, int(round_out), _stream()),
                "corrif_attention_fwd_premasked")
    _count()


def attention_bwd(qkv, O, dO, lse, maskbits, delta, dqkv, B, N, H=8, D=64, scale=0.125, p=0.0):
    # algorithmic FLOPs of the backward: dV, dP, dQ, dK = 4 products (recomputing S is not counted)
    with _rec("attn_bwd", 8.0 * B * H * N * N * D, "B%d N%d" % (B, N)):
        L.check(lib().corrif_attention_bwd(_ptr(qkv), _ptr(O), _ptr(dO), _ptr(lse),
                                           _ptr(maskbits, torch.int32), _ptr(delta), _ptr(dqkv), B, N, H,
                                           D, scale, p, _stream()), "corrif_attention_bwd")
    _count(3)


def dropout(x, out, n, p, seed, site, seed_dev=None):
    with _rec('dropout', 8.0 * n):
        L.check(lib().corrif_dropout(_ptr(x), _ptr(out), n, p, seed, _seed_dev(seed_dev), site, _stream()),
                "corrif_dropout")
    _count()


def dropout_colsum(x, out, rows, cols, p, seed, site, colsum, seed_dev=None):
    """out = dropout(x) ([rows, cols]) and colsum += column sums of out, one pass."""
    with _rec('dropout_colsum', 8.0 * rows * cols):
        L.check(lib().corrif_dropout_colsum(_ptr(x), _ptr(out), rows, cols, p, seed, _seed_dev(seed_dev), site,
                                            _ptr(colsum), _stream()), "corrif_dropout_colsum")
    _count()


def dropout_mask(mask, n, p, seed, site, seed_dev=None):
    with _rec('dropout', 4.0 * n):
        L.check(lib().corrif_dropout_mask(_ptr(mask), n, p, seed, _seed_dev(seed_dev), site, _stream()),
                "corrif_dropout_mask")
    _count()


def dropout_add(x, res, out, n, p, seed, site_a, site_b=NO_SITE, seed_dev=None):
    with _rec('dropout', (8.0 if res is None else 12.0) * n):
        L.check(lib().corrif_dropout_add(_ptr(x), _ptr(res), _ptr(out), n, p, seed, _seed_dev(seed_dev),
                                         site_a, site_b, _stream()), "corrif_dropout_add")
    _count()


def colsum_scratch_floats(rows, cols) -> int:
    return int(lib().corrif_colsum_scratch_floats(rows, cols))


def colsum(x, ld, rows, cols, out, scratch, accumulate=False):
    with _rec('colsum', 4.0 * rows * cols):
        L.check(lib().corrif_colsum(_ptr(x), ld, rows, cols, _ptr(out), int(accumulate), _ptr(scratch),
                                    _stream()), "corrif_colsum")
    _count(1)


def colsum_batched(x, ld, rows, cols, out, batch, x_bstride, out_bstride, accumulate=False):
    with _rec('colsum', 4.0 * rows * cols * batch):
        L.check(lib().corrif_colsum_batched(_ptr(x), ld, rows, cols, _ptr(out), batch, x_bstride, out_bstride,
                                            int(accumulate), _stream()), "corrif_colsum_batched")
    _count(1)


def batchsum(x, batch, stride, n, out, accumulate=False):
    with _rec('batchsum', 4.0 * (batch + 1) * n):
        L.check(lib().corrif_batchsum(_ptr(x), batch, stride, n, _ptr(out), int(accumulate), _stream()),
                "corrif_batchsum")
    _count()


def add_rows(a, lda, b, ldb, out, ldo, rows, cols):
    with _rec('add_rows', 12.0 * rows * cols):
        L.check(lib().corrif_add_rows(_ptr(a), lda, _ptr(b), ldb, _ptr(out), ldo, rows, cols, _stream()),
                "corrif_add_rows")
    _count()


def inter_corr_fwd(qkv, skip, tokens, M, B, S, C_):
    with _rec('inter_corr_fwd', 4.0 * (3 * M + 2 * M) * B * S * C_):
        L.check(lib().corrif_inter_corr_fwd(_ptr(qkv), _ptr(skip), _ptr(tokens), M, B, S, C_, _stream()),
                "corrif_inter_corr_fwd")
    _count()


def inter_corr_bwd(qkv, g_tokens, dqkv, M, B, S, C_, g_group_major=False):
    """g_tokens [B][(M+1)S][C], or [M+1][B][S][C] with g_group_major."""
    with _rec('inter_corr_bwd', 4.0 * (3 * M + M + 3 * M) * B * S * C_):
        L.check(lib().corrif_inter_corr_bwd_layout(_ptr(qkv), _ptr(g_tokens), _ptr(dqkv), M, B, S, C_,
                                                   int(g_group_major), _stream()), "corrif_inter_corr_bwd_layout")
    _count()


def jaccard_sums(y, y_pred, P, sums):
    with _rec('jaccard', 8.0 * P):
        L.check(lib().corrif_jaccard_sums(_ptr(y), _ptr(y_pred), P, _ptr(sums, torch.float64), _stream()),
                "corrif_jaccard_sums")
    _count()


def loss_jaccard_fused(x, y, B, CH, P, grad_scale, loss_sum, dx, sums):
    with _rec('loss_jaccard', (12.0 if dx is not None else 8.0) * B * CH * P):
        L.check(lib().corrif_loss_jaccard_fused(_ptr(x), _ptr(y), B, CH, P, grad_scale, _ptr(loss_sum, torch.float64),
                                                _ptr(dx), _ptr(sums, torch.float64), _stream()),
                "corrif_loss_jaccard_fused")
    _count()


def jaccard_finish(sums, epsilon, out3):
    L.check(lib().corrif_jaccard_finish(_ptr(sums, torch.float64), epsilon, _ptr(out3), _stream()),
            "corrif_jaccard_finish")
    _count()


def confusion_counts(label, pred, P, num_classes, counts):
    with _rec('confusion', 2.0 * P):
        L.check(lib().corrif_confusion_counts(_ptr(label, torch.uint8), _ptr(pred, torch.uint8), P,
                                              num_classes, _ptr(counts, torch.int64), _stream()),
                "corrif_confusion_counts")
    _count()


def bce_probs_fwd_bwd(x, y, n, grad_scale, loss_sum, dx):
    L.check(lib().corrif_bce_probs_fwd_bwd(_ptr(x), _ptr(y), n, grad_scale,
                                           _ptr(loss_sum, torch.float64), _ptr(dx), _stream()),
            "corrif_bce_probs_fwd_bwd")
    _count()


def adam_step(p, g, m, v, n, lr, beta1, beta2, eps, grad_scale, step):
    L.check(lib().corrif_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), n, lr, beta1, beta2, eps,
                                   grad_scale, step, _stream()), "corrif_adam_step")
    _count()
