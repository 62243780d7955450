"""CPU oracle for the CorrIFNet fusion hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference's hot path so the CUDA
kernels can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package never does (it raises when the CUDA library is missing).

Pinning.  The reference ships no golden vectors or tests, and its
arithmetic lives in a third-party dependency that is not vendored: PyTorch (unpinned by the
reference; torch 2.11.0 in this image).  The oracle is therefore pinned by outputs of the
*unmodified reference run live* in the build container: ``tests/golden/make_golden.py``
imports ``/root/reference/mmvit4.py`` / ``F5_JACCARD2.py`` and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function here against
those files.

Two layers:
  * torch-functional restatement (any dtype, autograd supplies the backward) of the whole
    block, following /root/reference/mmvit4.py:295-388, 398-426, 456-529;
  * plain numpy closed forms for the pieces that have integer/index structure: the
    ``inter_attn`` batch-mixing correlation (forward and hand-derived backward) and the
    Jaccard sums (/root/reference/F5_JACCARD2.py:4-36).
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch
import torch.nn.functional as F

MODALITIES = ("RGB", "NIR", "SWIR")
DIM = 512          # transformer_basic_dims, mmvit4.py:11
HEADS = 8          # num_heads, mmvit4.py:13
MLP = 512          # mlp_dim, mmvit4.py:12
PATCH = 8          # patch_size, mmvit4.py:16
TOKENS = PATCH ** 3
ENC_CH = 64        # basic_dims*8, mmvit4.py:398
DEC_CH = 192       # basic_dims*8*3, mmvit4.py:426


# --------------------------------------------------------------------------------------
# parameter inventory (state_dict keys of the block, mmvit4.py:398-426)
# --------------------------------------------------------------------------------------
def transformer_param_shapes(prefix: str, dim: int = DIM, mlp: int = MLP):
    a = f"{prefix}.cross_attention_list.0.fn"
    f = f"{prefix}.cross_ffn_list.0.fn"
    return {
        f"{a}.norm.weight": (dim,), f"{a}.norm.bias": (dim,),
        f"{a}.fn.qkv.weight": (3 * dim, dim),
        f"{a}.fn.proj.weight": (dim, dim), f"{a}.fn.proj.bias": (dim,),
        f"{f}.norm.weight": (dim,), f"{f}.norm.bias": (dim,),
        f"{f}.fn.net.0.weight": (mlp, dim), f"{f}.fn.net.0.bias": (mlp,),
        f"{f}.fn.net.3.weight": (dim, mlp), f"{f}.fn.net.3.bias": (dim,),
    }


def param_shapes(modalities=MODALITIES, dim: int = DIM, tokens: int = TOKENS):
    """Ordered {state_dict key: shape} of every parameter on the hot path."""
    nm = len(modalities)
    shapes = {}
    for m in modalities:
        shapes[f"{m}_encode_conv.weight"] = (dim, ENC_CH, 1, 1, 1)
        shapes[f"{m}_encode_conv.bias"] = (dim,)
    shapes["fused6_encode_conv.weight"] = (dim, ENC_CH * nm, 1, 1, 1)
    shapes["fused6_encode_conv.bias"] = (dim,)
    for m in modalities:
        shapes[f"{m}_pos"] = (1, tokens, dim)
    shapes["fused6_pos"] = (1, tokens, dim)
    for m in modalities:
        shapes.update(transformer_param_shapes(f"{m}_transformer", dim))
    for m in modalities:
        shapes[f"qkv_{m}.weight"] = (3 * dim, dim, 1, 1, 1)
        shapes[f"qkv_{m}.bias"] = (3 * dim,)
    shapes.update(transformer_param_shapes("multimodal_transformer", dim))
    shapes["multimodal_decode_conv.weight"] = (ENC_CH * nm, dim * (nm + 1), 1, 1, 1)
    shapes["multimodal_decode_conv.bias"] = (ENC_CH * nm,)
    return shapes


def make_params(seed: int, modalities=MODALITIES, dtype=torch.float32):
    """Deterministic synthetic weights, reproducible on any machine (numpy PCG64, keyed by
    parameter name so the result does not depend on iteration order).  Magnitudes follow
    the reference's initialisers (kaiming-normal convs mmvit4.py:437-439, nn.Linear default,
    LayerNorm ~ identity) except the positional embeddings, which the reference zero-inits
    (mmvit4.py:408-411) and which are given N(0, 0.02) so a missing ``+pos`` is visible."""
    out = {}
    for name, shape in param_shapes(modalities).items():
        rng = np.random.default_rng([int(seed), zlib.crc32(name.encode())])
        if name.endswith("_pos"):
            a = rng.standard_normal(shape) * 0.02
        elif ".norm.weight" in name:
            a = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith(".bias"):
            a = 0.05 * rng.standard_normal(shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            a = rng.standard_normal(shape) * math.sqrt(1.0 / fan_in)
        out[name] = torch.from_numpy(a.astype(np.float64)).to(dtype)
    return out


def make_full_model_state(seed: int, inventory: dict, dtype=torch.float32):
    """Deterministic values for EVERY entry of MMVit4.state_dict() (``inventory`` = key -> shape, the
    committed tests/golden/mmvit4_state_dict_inventory.json).  Hot-path keys reuse make_params."""
    hot = make_params(seed, dtype=dtype)
    out = {}
    for name, shape in inventory.items():
        if name in hot:
            out[name] = hot[name]
            continue
        rng = np.random.default_rng([int(seed), zlib.crc32(name.encode())])
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros((), dtype=torch.int64)
            continue
        if name.endswith("running_var"):
            a = 1.0 + 0.1 * rng.random(shape)
        elif name.endswith("running_mean"):
            a = 0.05 * rng.standard_normal(shape)
        elif ("bn" in name.split(".")[-2] or "downsample.1" in name) and name.endswith(".weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith(".bias"):
            a = 0.05 * rng.standard_normal(shape)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
            a = rng.standard_normal(shape) * math.sqrt(2.0 / fan_in)
        out[name] = torch.from_numpy(np.asarray(a, dtype=np.float64)).to(dtype)
    return out


def make_inputs(seed: int, batch: int, modalities=MODALITIES, dtype=torch.float32):
    """x6 per modality [B,64,8,8,8], fused_x6 [B,192,8,8,8] and the upstream gradient
    [B,192,8,8,8] (SURVEY.md section 8d synthetic inputs)."""
    rng = np.random.default_rng([int(seed), int(batch), 77])
    nm = len(modalities)
    x6 = [torch.from_numpy(rng.standard_normal((batch, ENC_CH, PATCH, PATCH, PATCH))).to(dtype)
          for _ in modalities]
    fused = torch.from_numpy(rng.standard_normal((batch, ENC_CH * nm, PATCH, PATCH, PATCH))).to(dtype)
    gout = torch.from_numpy(rng.standard_normal((batch, ENC_CH * nm, PATCH, PATCH, PATCH))).to(dtype)
    return x6, fused, gout


# --------------------------------------------------------------------------------------
# torch-functional restatement
# --------------------------------------------------------------------------------------
def _drop(x, masks, key):
    """Dropout with an explicit, pre-scaled keep mask (``None`` = p=0, the parity setting)."""
    if masks is None or key not in masks:
        return x
    return x * masks[key]


def self_attention(x, p, prefix, masks=None, heads=HEADS):
    """SelfAttention.forward, mmvit4.py:305-315 (qkv has no bias, scale = head_dim**-0.5)."""
    B, N, C = x.shape
    d = C // heads
    qkv = F.linear(x, p[f"{prefix}.qkv.weight"])                       # :307
    qkv = qkv.reshape(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]                                   # :308
    attn = (q @ k.transpose(-2, -1)) * (d ** -0.5)                     # :309
    attn = attn.softmax(dim=-1)                                        # :310
    attn = _drop(attn, masks, f"{prefix}.attn_drop")                   # :311
    y = (attn @ v).transpose(1, 2).reshape(B, N, C)                    # :312
    y = F.linear(y, p[f"{prefix}.proj.weight"], p[f"{prefix}.proj.bias"])  # :313
    return _drop(y, masks, f"{prefix}.proj_drop")                      # :314


def transformer(x, pos, p, prefix, masks=None):
    """Transformer.forward with depth 1, mmvit4.py:383-388: pos add, Residual(PreNormDrop(
    attention)) (:317-322, :332-339), Residual(PreNorm(FeedForward)) (:324-330, :347-358)."""
    a = f"{prefix}.cross_attention_list.0.fn"
    f = f"{prefix}.cross_ffn_list.0.fn"
    C = x.shape[-1]
    x = x + pos                                                                     # :385
    h = F.layer_norm(x, (C,), p[f"{a}.norm.weight"], p[f"{a}.norm.bias"], 1e-5)     # :339
    y = self_attention(h, p, f"{a}.fn", masks)
    y = _drop(y, masks, f"{a}.dropout")                                             # :339
    x = y + x                                                                       # :322
    h = F.layer_norm(x, (C,), p[f"{f}.norm.weight"], p[f"{f}.norm.bias"], 1e-5)     # :330
    u = F.linear(h, p[f"{f}.fn.net.0.weight"], p[f"{f}.fn.net.0.bias"])
    u = F.gelu(u)                                                                   # :345 exact erf
    u = _drop(u, masks, f"{f}.fn.net.2")
    u = F.linear(u, p[f"{f}.fn.net.3.weight"], p[f"{f}.fn.net.3.bias"])
    u = _drop(u, masks, f"{f}.fn.net.4")
    return u + x                                                                    # :322


def conv1x1_tokens(x, w, b):
    """1x1x1 Conv3d followed by permute(0,2,3,4,1).view(B,-1,C): mmvit4.py:458-461, 510-513.
    x [B,Cin,D,H,W] -> tokens [B, D*H*W, Cout]."""
    B, Cin = x.shape[:2]
    t = x.reshape(B, Cin, -1).transpose(1, 2)            # [B,S,Cin]
    return F.linear(t, w.reshape(w.shape[0], Cin), b)


def inter_attn(q, ks, vs):
    """inter_attn closure, mmvit4.py:481-487, on token-major [B,S,C] tensors.

    The reference works on [B,C,8,8,8]; every op is element-wise in (c,s) so the layout does
    not matter, only the (modality, batch) re-interpretation of the ``.view`` at :485 does.
    We reproduce that view literally by moving C,S into one trailing axis."""
    B = q.shape[0]
    n = len(ks)
    scores = [(q * k).reshape(1, -1) for k in ks]                       # :482-483
    attn = torch.softmax(torch.cat(scores, dim=0) / math.sqrt(n), dim=0)  # :484 Softmax(dim=0)
    attn = attn.view(B, n, *q.shape[1:])                                # :485 the batch-mixing view
    return sum(attn[:, i] * v for i, v in enumerate(vs))               # :486-487


def fusion_block(params, x6, fused_x6, masks=None, modalities=MODALITIES):
    """The hot path, mmvit4.py:456-529.  x6: list of [B,64,8,8,8]; fused_x6 [B,192,8,8,8].
    Returns x6_inter [B,192,8,8,8]."""
    p = params
    B = fused_x6.shape[0]
    C = p[f"{modalities[0]}_pos"].shape[-1]
    trans, skip = [], []
    for m, x in zip(modalities, x6):                                     # tokenize :457-466
        tok = conv1x1_tokens(x, p[f"{m}_encode_conv.weight"], p[f"{m}_encode_conv.bias"])
        skip.append(tok)
        trans.append(transformer(tok, p[f"{m}_pos"], p, f"{m}_transformer", masks))
    qs, ks, vs = [], [], []
    for m, t in zip(modalities, trans):                                  # qkv convs :469-479
        w = p[f"qkv_{m}.weight"]
        y = F.linear(t, w.reshape(w.shape[0], C), p[f"qkv_{m}.bias"])
        q, k, v = y.chunk(3, dim=-1)
        qs.append(q), ks.append(k), vs.append(v)
    fused_tokens = [skip[i] + inter_attn(qs[i], ks, vs) for i in range(len(modalities))]  # :489-507
    f6 = conv1x1_tokens(fused_x6, p["fused6_encode_conv.weight"], p["fused6_encode_conv.bias"])
    tokens = torch.cat(fused_tokens + [f6], dim=1)                       # :515-521
    pos = torch.cat([p[f"{m}_pos"] for m in modalities] + [p["fused6_pos"]], dim=1)
    mm = transformer(tokens, pos, p, "multimodal_transformer", masks)    # :519-522
    nt = len(modalities) + 1
    S = mm.shape[1] // nt
    g = mm.reshape(B * S, nt * C)                                        # :526 view(B,8,8,8,2048)
    w = p["multimodal_decode_conv.weight"]
    y = F.linear(g, w.reshape(w.shape[0], nt * C), p["multimodal_decode_conv.bias"])
    return y.reshape(B, S, -1).transpose(1, 2).reshape(B, -1, PATCH, PATCH, PATCH)  # :527-529


def random_masks(batch: int, p: float, dtype=torch.float32, modalities=MODALITIES):
    """Fresh pre-scaled Bernoulli keep masks for all 20 dropout sites (what the 20 nn.Dropout modules
    of the reference draw in train mode, mmvit4.py:311,314,339,353,355); used by the CPU timing legs so
    the baseline pays the same RNG cost as the reference does."""
    masks = {}
    names = [f"{m}_transformer" for m in modalities] + ["multimodal_transformer"]
    for t, name in enumerate(names):
        N = TOKENS if t < len(modalities) else TOKENS * (len(modalities) + 1)
        a, f = f"{name}.cross_attention_list.0.fn", f"{name}.cross_ffn_list.0.fn"
        for key, shape in ((f"{a}.fn.attn_drop", (batch, HEADS, N, N)), (f"{a}.fn.proj_drop", (batch, N, DIM)),
                           (f"{a}.dropout", (batch, N, DIM)), (f"{f}.fn.net.2", (batch, N, MLP)),
                           (f"{f}.fn.net.4", (batch, N, DIM))):
            masks[key] = torch.empty(shape, dtype=dtype).bernoulli_(1 - p).div_(1 - p)
    return masks


def fusion_block_fwd_bwd(params, x6, fused_x6, gout, masks=None, dtype=torch.float64):
    """Forward + autograd backward in ``dtype``.  Returns (out, {name: grad}) where the grad
    dict holds every parameter plus ``x6.<i>`` and ``fused_x6``."""
    p = {k: v.detach().to(dtype).requires_grad_(True) for k, v in params.items()}
    xs = [x.detach().to(dtype).requires_grad_(True) for x in x6]
    fx = fused_x6.detach().to(dtype).requires_grad_(True)
    mk = None if masks is None else {k: v.to(dtype) for k, v in masks.items()}
    out = fusion_block(p, xs, fx, mk)
    out.backward(gout.to(dtype))
    grads = {k: v.grad for k, v in p.items()}
    for i, x in enumerate(xs):
        grads[f"x6.{i}"] = x.grad
    grads["fused_x6"] = fx.grad
    return out.detach(), grads


# --------------------------------------------------------------------------------------
# numpy closed forms: inter-modal correlation with the batch-mixing quirk
# --------------------------------------------------------------------------------------
def inter_corr_fwd_np(q, k, v, skip=None):
    """q,k,v: [M,B,S,C] (modality-major stacks).  Returns out [M,B,S,C] where
    out[X,b'] = skip[X,b'] + sum_i A_X[m,b] * v[i,b'],  (m,b) = divmod(M*b'+i, B),
    A_X[:,b] = softmax_m(q[X,b]*k[m,b]/sqrt(M))          (mmvit4.py:481-487)."""
    M, B = q.shape[:2]
    out = np.zeros_like(q) if skip is None else skip.copy()
    for X in range(M):
        s = q[X][None] * k / math.sqrt(M)                # [M(m),B,S,C]
        s = s - s.max(axis=0, keepdims=True)
        e = np.exp(s)
        A = e / e.sum(axis=0, keepdims=True)
        for bp in range(B):
            for i in range(M):
                m, b = divmod(M * bp + i, B)
                out[X, bp] += A[m, b] * v[i, bp]
    return out


def inter_corr_bwd_np(q, k, v, g):
    """Hand-derived backward of ``inter_corr_fwd_np`` (after the table).
    g [M,B,S,C] = dL/dout.  Returns dq, dk, dv (dskip = g)."""
    M, B = q.shape[:2]
    r = 1.0 / math.sqrt(M)
    dq, dk, dv = np.zeros_like(q), np.zeros_like(k), np.zeros_like(v)
    for X in range(M):
        s = q[X][None] * k * r
        s = s - s.max(axis=0, keepdims=True)
        e = np.exp(s)
        A = e / e.sum(axis=0, keepdims=True)             # [m,b,...]
        dA = np.zeros_like(A)
        for bp in range(B):
            for i in range(M):
                m, b = divmod(M * bp + i, B)
                dv[i, bp] += A[m, b] * g[X, bp]
                dA[m, b] = g[X, bp] * v[i, bp]
        ds = A * (dA - (A * dA).sum(axis=0, keepdims=True)) * r
        dq[X] = (ds * k).sum(axis=0)
        dk += ds * q[X][None]
    return dq, dk, dv


# --------------------------------------------------------------------------------------
# numpy: Jaccard family (F5_JACCARD2.py:4-36, F5_JACCARD.py:4-9)
# --------------------------------------------------------------------------------------
def jaccard_sums_np(y, y_pred, invert_if_empty: bool):
    """The three float32 sums of F5_JACCARD2.py:16-18 and whether :12-14 inverted.  Sums are
    accumulated in float64 then rounded to float32; for {0,1} inputs with every count below
    2**24 this equals the reference's fp32 result exactly."""
    y = np.asarray(y, dtype=np.float32).reshape(-1)
    yp = np.asarray(y_pred, dtype=np.float32).reshape(-1)
    inverted = False
    if invert_if_empty and float(y.astype(np.float64).sum()) == 0.0:     # :12
        y, yp, inverted = 1 - y, 1 - yp, True                            # :13-14
    tp = np.float32((yp.astype(np.float64) * y).sum())                   # :16
    fp = np.float32(((1 - yp).astype(np.float64) * y).sum())             # :17
    fn = np.float32(((1 - y).astype(np.float64) * yp).sum())             # :18
    return tp, fp, fn, inverted


def jaccard_np(y, y_pred, epsilon=1e-8):
    tp, fp, fn, _ = jaccard_sums_np(y, y_pred, False)
    eps = np.float32(epsilon)
    return np.float32((tp + eps) / (tp + fp + fn + eps))                 # F5_JACCARD2.py:8


def jaccard2_np(y, y_pred, epsilon=1e-8):
    tp, fp, fn, _ = jaccard_sums_np(y, y_pred, True)
    eps = np.float32(epsilon)
    return np.float32((tp + eps) / (tp + fp + fn + eps))                 # F5_JACCARD2.py:19


def jaccard_and_f1_np(y, y_pred, epsilon=1e-8):
    tp, fp, fn, _ = jaccard_sums_np(y, y_pred, True)
    eps = np.float32(epsilon)
    recall = np.float32(tp / (tp + fn + eps))                            # :33
    prec = np.float32(tp / (tp + fp + eps))                              # :34
    return np.float32(np.float32(2) * (recall * prec) / (recall + prec + eps))  # :35


def confusion_counts_np(label, pred, num_classes):
    """Integer K x K confusion matrix (rows = label, cols = pred) of uint8 class maps, and the
    per-class (TP, FP, FN) in the reference's naming (F5_JACCARD2.py:16-18: "FP" = label
    positive & pred negative, "FN" = label negative & pred positive)."""
    label = np.asarray(label).reshape(-1).astype(np.int64)
    pred = np.asarray(pred).reshape(-1).astype(np.int64)
    cm = np.bincount(label * num_classes + pred, minlength=num_classes * num_classes)
    cm = cm.reshape(num_classes, num_classes)
    tp = np.diag(cm)
    fp = cm.sum(axis=1) - tp
    fn = cm.sum(axis=0) - tp
    return cm, tp, fp, fn


# --------------------------------------------------------------------------------------
# train-step pieces (F4_TRAIN.py:52-71): loss on the already-sigmoided output, Adam
# --------------------------------------------------------------------------------------
def bce_with_logits_on_probs(probs, masks):
    """nn.BCEWithLogitsLoss()(outputs, masks) where outputs are sigmoid probabilities
    (mmvit4.py:291, F4_TRAIN.py:58-60): mean over all elements of softplus(x) - x*y."""
    return F.binary_cross_entropy_with_logits(probs, masks)


def adam_step_np(p, g, m, v, step, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (F2_MAIN.py:168-169), one step, float64 numpy."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mhat = m / (1 - b1 ** step)
    vhat = v / (1 - b2 ** step)
    return p - lr * mhat / (np.sqrt(vhat) + eps), m, v


# --------------------------------------------------------------------------------------------------
# Next rows of the scope table: restatements pinned by tests/golden/next_rows.npz,
# which tests/golden/make_golden.py generates from the UNMODIFIED reference classes.  No kernels yet.
# --------------------------------------------------------------------------------------------------
def instance_norm3d(x, eps: float = 1e-5):
    """nn.InstanceNorm3d defaults (affine=False, no running stats), mmvit4.py:24: per (sample, channel)
    normalisation over the spatial volume with the biased variance."""
    dims = tuple(range(2, x.dim()))
    mu = x.mean(dim=dims, keepdim=True)
    var = x.var(dim=dims, unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def early_fusion_block(xs, weight, bias):
    """EarlyFusionBlock.forward (mmvit4.py:76-81): cat over channels -> 1x1x1 Conv3d -> ReLU -> InstanceNorm3d.
    xs: three [B, c, D, H, W]; weight [3c, 3c, 1, 1, 1]; bias [3c]."""
    x = torch.cat(list(xs), dim=1)
    y = torch.einsum("bkdhw,nk->bndhw", x, weight.reshape(weight.shape[0], weight.shape[1])) + bias.view(1, -1, 1, 1, 1)
    return instance_norm3d(torch.relu(y))


def general_conv3d_prenorm(x, weight, bias, k_size: int = 3, pad_type: str = "replicate"):
    """general_conv3d_prenorm.forward (mmvit4.py:29-45) as the decoder uses it (mmvit4.py:225-237): Conv3d with
    stride 1, padding k//2 of `pad_type`, bias -> ReLU -> InstanceNorm3d."""
    pad = k_size // 2
    if pad:
        x = torch.nn.functional.pad(x, (pad,) * 6, mode=pad_type if pad_type != "zeros" else "constant")
    return instance_norm3d(torch.relu(torch.nn.functional.conv3d(x, weight, bias)))



# --------------------------------------------------------------------------------------------------
# Whole model (MMVit4.forward, mmvit4.py:441-532) as a functional restatement keyed by the reference's state_dict
# names: the CPU leg of the full-train-step measurement (bench.py cpu_baseline / --impl reference) and the
# checker of the drop-in model.  Pinned by tests/golden/mmvit4_full_small.npz (the unmodified reference, fp64).
# Train-mode semantics: BatchNorm3d normalises with the statistics of the local batch (mmvit4.py:121,132-136; the
# running buffers do not influence train-mode outputs and are not updated here).
# --------------------------------------------------------------------------------------------------
def _bn_train(x, p, prefix, eps=1e-5):
    return F.batch_norm(x, None, None, p[f"{prefix}.weight"], p[f"{prefix}.bias"], training=True, eps=eps)


def _bottleneck3d(x, p, prefix, stride):
    """Bottleneck3D.forward (mmvit4.py:196-212); ResNet-50 v1.5 places the stride on the 3x3 conv, inflated to
    (1,3,3) kernels (inflate_conv with time_dim=1, mmvit4.py:83-111, 128-150)."""
    out = torch.relu(_bn_train(F.conv3d(x, p[f"{prefix}.conv1.weight"]), p, f"{prefix}.bn1"))
    out = F.conv3d(out, p[f"{prefix}.conv2.weight"], stride=(1, stride, stride), padding=(0, 1, 1))
    out = torch.relu(_bn_train(out, p, f"{prefix}.bn2"))
    out = _bn_train(F.conv3d(out, p[f"{prefix}.conv3.weight"]), p, f"{prefix}.bn3")
    identity = x
    if f"{prefix}.downsample.0.weight" in p:
        identity = _bn_train(F.conv3d(x, p[f"{prefix}.downsample.0.weight"], stride=(1, stride, stride)),
                             p, f"{prefix}.downsample.1")
    return torch.relu(out + identity)


def encoder(x, p, prefix):
    """Encoder.forward (mmvit4.py:170-194).  x [B,1,3,H,W] -> (x1..x5 adapted, x6 [B,64,8,8,8])."""
    y = F.conv3d(x, p[f"{prefix}.e1_c1.weight"], stride=(1, 2, 2), padding=(1, 3, 3))
    y = _bn_train(torch.relu(y), p, f"{prefix}.e1_bn")                                     # ReLU before BN (:173)
    y = F.max_pool3d(y, kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
    feats = [y]
    for stage, blocks, stride in (("e2", 3, 1), ("e3", 4, 2), ("e4", 6, 2), ("e5", 3, 2)):
        for b in range(blocks):
            y = _bottleneck3d(y, p, f"{prefix}.{stage}.{b}", stride if b == 0 else 1)
        feats.append(y)
    lv = [F.conv3d(f, p[f"{prefix}.adapt{i}.weight"], p[f"{prefix}.adapt{i}.bias"]) for i, f in enumerate(feats, 1)]
    pooled = [F.interpolate(t, size=(8, 8, 8), mode="trilinear", align_corners=True) for t in lv]   # :187-191
    x6 = F.conv3d(torch.cat(pooled, dim=1), p[f"{prefix}.conv6.weight"], p[f"{prefix}.conv6.bias"])
    return (*lv, x6)


def _prenorm(x, p, prefix, k=3, pad_type="replicate"):
    return general_conv3d_prenorm(x, p[f"{prefix}.conv.weight"], p[f"{prefix}.conv.bias"], k, pad_type)


def _rfm(x, p, prefix):
    """fusion_prenorm (mmvit4.py:47-56): 1x1x1, 3x3x3 (zero padding: the class default), 1x1x1."""
    x = _prenorm(x, p, f"{prefix}.fusion_layer.0", 1)
    x = _prenorm(x, p, f"{prefix}.fusion_layer.1", 3, "zeros")
    return _prenorm(x, p, f"{prefix}.fusion_layer.2", 1)


def decoder_fuse(x1, x2, x3, x4, x5, p, prefix="decoder_fuse"):
    """Decoder_fuse.forward (mmvit4.py:266-292)."""
    up2 = lambda t: F.interpolate(t, scale_factor=2, mode="trilinear", align_corners=True)   # noqa: E731
    y = _rfm(x5, p, f"{prefix}.RFM5")
    y = F.conv3d(y, p[f"{prefix}.RFM5_reduce.weight"], p[f"{prefix}.RFM5_reduce.bias"])
    for lvl, skip, cube in ((4, x4, 16), (3, x3, 32), (2, x2, 64), (1, x1, 128)):
        y = _prenorm(up2(y), p, f"{prefix}.d{lvl}_c1")
        s = F.interpolate(_rfm(skip, p, f"{prefix}.RFM{lvl}"), (cube, cube, cube))          # nearest (:271)
        y = _prenorm(torch.cat((s, y), dim=1), p, f"{prefix}.d{lvl}_c2")
        y = _prenorm(y, p, f"{prefix}.d{lvl}_out", 1)
    y = F.interpolate(y, size=(1, 224, 224), mode="trilinear", align_corners=True)          # :263, 288
    return torch.sigmoid(F.conv3d(y, p[f"{prefix}.final_conv.weight"], p[f"{prefix}.final_conv.bias"]))


def full_model(p, x, masks=None, compute_unused=True):
    """MMVit4.forward (mmvit4.py:441-532).  p: the 1140-entry state_dict (or its parameters); x [B,3,3,H,W] ->
    [B,3,1,224,224] sigmoid probabilities.  ``compute_unused`` also evaluates fusion5, whose result the reference
    computes and drops (:453)."""
    feats = [encoder(x[:, i:i + 1], p, f"{m}_encoder") for i, m in enumerate(MODALITIES)]
    fused = []
    for lv in range(6):
        if lv == 4 and not compute_unused:
            fused.append(None)
            continue
        fused.append(early_fusion_block([f[lv] for f in feats], p[f"fusion{lv + 1}.conv.weight"],
                                        p[f"fusion{lv + 1}.conv.bias"]))
    x6_inter = fusion_block(p, [f[5] for f in feats], fused[5], masks)
    return decoder_fuse(fused[0], fused[1], fused[2], fused[3], x6_inter, p)


def train_step_cpu(state, x, masks, lr=1e-4, dropout_p=0.1):
    """One F4_TRAIN.py:52-71 step body on the CPU (forward, BCE-with-logits on the probabilities, backward, Adam,
    Jaccard2 of channel 0); ``state`` = {name: leaf tensor} is updated in place.  Returns (loss, jaccard2)."""
    params = [v for v in state.values() if v.is_floating_point() and v.requires_grad]
    if not hasattr(train_step_cpu, "_opt") or train_step_cpu._opt[0] is not state:
        train_step_cpu._opt = (state, torch.optim.Adam(params, lr))
    opt = train_step_cpu._opt[1]
    opt.zero_grad()
    dm = random_masks(x.shape[0], dropout_p) if dropout_p > 0 else None
    out = full_model(state, x, dm)
    loss = bce_with_logits_on_probs(out, masks)
    loss.backward()
    opt.step()
    load = masks.shape[0] * 224 * 224
    jac = jaccard2_np(masks[:, 0].reshape(load, 1).numpy(), out.detach()[:, 0].reshape(load, 1).numpy())
    return float(loss), float(jac)
