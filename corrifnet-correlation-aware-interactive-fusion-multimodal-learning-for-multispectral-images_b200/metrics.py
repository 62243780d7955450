"""Jaccard family on the GPU (reference F5_JACCARD2.py:4-36, F5_JACCARD.py:4-9).

Same signatures and return shape as the reference (``[1]`` float32 tensor for ``[P,1]`` inputs), but
one streaming pass instead of ~6 eager tensor passes, and the ``if y.sum(0)==0`` branch of
F5_JACCARD2.py:12 is resolved on the device, so the call never synchronises the host.
Inputs must be CUDA float32; there is no CPU path.
"""
from __future__ import annotations

import torch

from . import ops


def _prep(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError("corrif_b200 Jaccard needs CUDA tensors (no CPU fallback)")
    if t.dim() == 2 and t.shape[1] != 1:
        raise ValueError("expected [P,1] (or [P]) inputs as in F4_TRAIN.py:70, got %s" % (tuple(t.shape),))
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous().view(-1)


def jaccard_all(y: torch.Tensor, y_pred: torch.Tensor, epsilon: float = 1e-8):
    """Returns (out3, sums): out3 = [Jaccard, Jaccard2, JaccardAndF1] float32[3];
    sums = [sum y, sum y_pred, sum y*y_pred, P] float64[4]."""
    yv, pv = _prep(y), _prep(y_pred)
    if yv.numel() != pv.numel() or yv.numel() == 0:
        raise ValueError("y and y_pred must be non-empty and of equal size")
    sums = torch.zeros(4, dtype=torch.float64, device=yv.device)
    out = torch.empty(3, dtype=torch.float32, device=yv.device)
    ops.jaccard_sums(yv, pv, yv.numel(), sums)
    ops.jaccard_finish(sums, float(epsilon), out)
    return out, sums


def loss_and_jaccard(outputs: torch.Tensor, masks: torch.Tensor, want_grad: bool = True, epsilon: float = 1e-8):
    """The train-step tail of F4_TRAIN.py:58-71 in one pass over ``outputs`` / ``masks`` [B, CH, 1, H, W]:
    returns (loss, d loss / d outputs or None, Jaccard2 of channel 0 as float32[1], pixels).
    loss = nn.BCEWithLogitsLoss()(outputs, masks) (mean over all elements); no host synchronisation."""
    if not (outputs.is_cuda and masks.is_cuda):
        raise ValueError("corrif_b200 loss/metric tail needs CUDA tensors (no CPU fallback)")
    x, y = outputs.detach().float().contiguous(), masks.detach().float().contiguous()
    if x.shape != y.shape or x.dim() < 3:
        raise ValueError("outputs and masks must have the same [B, CH, ...] shape")
    B, CH = x.shape[0], x.shape[1]
    P = x[0, 0].numel()
    loss_sum = torch.zeros(1, dtype=torch.float64, device=x.device)
    sums = torch.zeros(4, dtype=torch.float64, device=x.device)
    dx = torch.empty_like(x) if want_grad else None
    out3 = torch.empty(3, dtype=torch.float32, device=x.device)
    ops.loss_jaccard_fused(x, y, B, CH, P, 1.0 / x.numel(), loss_sum, dx, sums)
    ops.jaccard_finish(sums, float(epsilon), out3)
    return (loss_sum / x.numel()).float().reshape(()), dx, out3[1:2], B * P


def Jaccard(y, y_pred, epsilon=1e-8):
    return jaccard_all(y, y_pred, epsilon)[0][0:1]


def Jaccard2(y, y_pred, epsilon=1e-8):
    return jaccard_all(y, y_pred, epsilon)[0][1:2]


def JaccardAndF1(y, y_pred, epsilon=1e-8):
    return jaccard_all(y, y_pred, epsilon)[0][2:3]


def confusion_matrix(label: torch.Tensor, pred: torch.Tensor, num_classes: int) -> torch.Tensor:
    """K x K int64 confusion matrix (rows = label, cols = pred) of uint8 class maps."""
    if not (label.is_cuda and pred.is_cuda):
        raise ValueError("confusion_matrix needs CUDA tensors (no CPU fallback)")
    lv = label.detach().to(torch.uint8).contiguous().view(-1)
    pv = pred.detach().to(torch.uint8).contiguous().view(-1)
    if lv.numel() != pv.numel() or lv.numel() == 0:
        raise ValueError("label and pred must be non-empty and of equal size")
    counts = torch.zeros(num_classes * num_classes, dtype=torch.int64, device=lv.device)
    ops.confusion_counts(lv, pv, lv.numel(), num_classes, counts)
    return counts.view(num_classes, num_classes)


def per_class_iou(cm: torch.Tensor, epsilon: float = 1e-8) -> torch.Tensor:
    """Per-class Jaccard2-style IoU from an integer confusion matrix (counts are exact)."""
    tp = cm.diag().double()
    fp = cm.sum(1).double() - tp
    fn = cm.sum(0).double() - tp
    return ((tp + epsilon) / (tp + fp + fn + epsilon)).float()
