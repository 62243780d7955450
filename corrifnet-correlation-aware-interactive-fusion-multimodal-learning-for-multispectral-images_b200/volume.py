"""Channels-last 3-D volume operators of the blocks either side of the fusion hot path (SURVEY.md section 8f rows
N1 / N2): ``conv_block`` = Conv3d (1x1x1 or 3x3x3, replicate / zero padding) -> ReLU -> InstanceNorm3d, i.e. the
reference's ``general_conv3d_prenorm`` (mmvit4.py:29-45) and ``EarlyFusionBlock`` (:64-81, with three sources), the
bias-only 1x1x1 convolutions, and the trilinear / nearest resizes of the decoder (:260-288), each with a hand-written
backward, as torch.autograd Functions over libcorrif_b200 kernels (include/corrif.h, "volume operators").  Per shape
the convolution runs on the tcgen05 line-convolution kernels (csrc/conv3d_tc.cu, conv3d_wgrad_tc.cu), the per-voxel
8-channel kernel or the warp-level kernels; the weight gradient of a 3x3x3 block is an autograd node of its own on a side
stream (_WgradLater).

A volume is a tensor [B, D, H, W, C] whose channel stride is 1 and whose voxel stride ``ld`` is uniform - a
contiguous tensor or a channel slice of one.  ``torch.cat`` along channels never happens: a convolution takes up to
three sources and concatenates them inside its loader, and its data gradient is one buffer whose channel slices are
returned as the sources' gradients.

What is saved for the backward of a block: its sources (alive anyway: they are the previous blocks' outputs), its
OUTPUT y (alive anyway: the next block's source) and 2 x B x C statistics.  The ReLU mask is recovered from y
(r > 0 <=> y > (0 - mean) * rstd, the forward's own arithmetic), so no pre-activation or post-ReLU tensor is kept:
stock PyTorch keeps the padded input, the conv output and the ReLU output of every block.

No CPU path: CUDA fp32 tensors only.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import ops
import os

from ._lib import PAD_REPLICATE, PAD_REPLICATE_ADJOINT, PAD_ZEROS  # noqa: F401

EPS = 1e-5          # nn.InstanceNorm3d default (mmvit4.py:24)
_TRUNC_COMP = 1.0 + 3.52e-4      # mean of the tensor core's TF32 truncation of an fp32 operand, inverted


def _ld(t: torch.Tensor) -> int:
    """Voxel stride of a channels-last volume view, or raise."""
    if t.dim() != 5 or not t.is_cuda or t.dtype != torch.float32:
        raise ValueError("volume ops need CUDA fp32 tensors of shape [B, D, H, W, C] (no CPU fallback)")
    B, D, H, W, Cc = t.shape
    st = t.stride()
    inner = {3: 1, 2: W, 1: H * W, 0: D * H * W}              # voxels spanned by one step along each axis
    ld = Cc
    for ax in (3, 2, 1, 0):
        if t.shape[ax] > 1:
            ld = st[ax] // inner[ax]
            break
    ok = (st[4] == 1 or Cc == 1) and all(t.shape[ax] == 1 or st[ax] == inner[ax] * ld for ax in (3, 2, 1, 0))
    if not ok or ld < Cc or ld % 4 or Cc % 4 or t.data_ptr() % 16:
        raise ValueError("not a channels-last volume view: shape %s strides %s" % (tuple(t.shape), st))
    return ld


def as_volume(t: torch.Tensor) -> torch.Tensor:
    """Return ``t`` if it already is a valid volume view, else a contiguous copy."""
    try:
        _ld(t)
        return t
    except ValueError:
        return t.contiguous()


def to_channels_last(x: torch.Tensor) -> torch.Tensor:
    """[B, C, D, H, W] (any strides) -> volume [B, D, H, W, C]; free when x is in channels_last_3d memory format."""
    return as_volume(x.permute(0, 2, 3, 4, 1))


def to_channels_first(v: torch.Tensor) -> torch.Tensor:
    """volume [B, D, H, W, C] -> logical [B, C, D, H, W] view (channels_last_3d strides; no copy)."""
    return v.permute(0, 4, 1, 2, 3)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _merge_adjacent(srcs):
    """[(data_ptr, C, ld)] of the sources, with consecutive channel slices of ONE buffer fused into one entry: when two
    producers wrote side by side into a wider buffer (conv_block(out=...), resize_nearest(out=...)) the consumer reads
    cat(24, 8) as 128-byte rows of one source instead of four 32-byte pieces per voxel."""
    B, D, H, W = srcs[0].shape[:4]
    out = []
    for s in srcs:
        if tuple(s.shape[:4]) != (B, D, H, W):
            raise ValueError("sources of a convolution must share [B, D, H, W]")
        ptr, C_, ld = s.data_ptr(), s.shape[4], _ld(s)
        if out and out[-1][2] == ld and ptr == out[-1][0] + 4 * out[-1][1] and out[-1][1] + C_ <= ld:
            out[-1] = (out[-1][0], out[-1][1] + C_, ld)
        else:
            out.append((ptr, C_, ld))
    return out


def _desc(srcs: Sequence[torch.Tensor], Cout: int, ksize: int, pad_mode: int) -> L.Conv3dDesc:
    d = L.Conv3dDesc()
    B, D, H, W = srcs[0].shape[:4]
    cin = 0
    srcs = _merge_adjacent(srcs)
    for i, (ptr, C_, ld) in enumerate(srcs):
        d.src[i].p, d.src[i].C, d.src[i].ld = ptr, C_, ld
        cin += C_
    d.nsrc, d.B, d.D, d.H, d.W, d.Cin, d.Cout = len(srcs), B, D, H, W, cin, Cout
    d.ksize, d.pad_mode = ksize, pad_mode
    return d


# Packed operands are cached per weight tensor and re-used until the weight changes: a step of 8 micro-batches packs
# each weight twice (forward + mirrored form) instead of 16 times.  "Changed" = another storage address, a bumped
# autograd version counter (every in-place torch update, load_state_dict), or a bumped epoch (FlatAdam and
# broadcast_module change weights behind autograd's version counters and call bump_weight_epoch(); so must any caller
# that writes through ``param.data``).  Entries are tied to the tensor OBJECT by a weak reference.
_PACK_CACHE: dict = {}
_WEIGHT_EPOCH = 0


def bump_weight_epoch() -> None:
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1
    if len(_PACK_CACHE) > 4096:
        _PACK_CACHE.clear()


def _cached_pack(weight: torch.Tensor, key_extra, make) -> torch.Tensor:
    if torch.cuda.is_current_stream_capturing():
        return make()                                         # inside a CUDA graph the packing must replay with the graph
    key = (id(weight),) + tuple(key_extra)
    stamp = (weight.data_ptr(), weight._version, _WEIGHT_EPOCH)
    hit = _PACK_CACHE.get(key)
    # the entry must belong to THIS tensor object (ids and addresses are recycled once a tensor dies)
    if hit is not None and hit[0]() is weight and hit[1] == stamp:
        return hit[2]
    wpk = make()
    try:
        _PACK_CACHE[key] = (weakref.ref(weight), stamp, wpk)
    except TypeError:
        pass
    return wpk


def pack_weights(weight: torch.Tensor, transpose_flip: bool = False) -> torch.Tensor:
    """[Cout, Cin, k, k, k] -> the warp-level kernels' fragment-ordered TF32 operand (forward, or data-gradient form)."""
    return _cached_pack(weight, (bool(transpose_flip),), lambda: _pack_weights(weight, transpose_flip))


def tc_enabled() -> bool:
    """The tcgen05 line convolution serves the shapes it supports unless CORRIF_CONV_TC=0 (A/B switch)."""
    return os.environ.get("CORRIF_CONV_TC", "1") != "0"


def tc_supported(d: L.Conv3dDesc) -> bool:
    return d.ksize == 3 and tc_enabled() and bool(ops.lib().corrif_conv3d_tc_supported(C.byref(d)))


def pack_weights_tc(weight: torch.Tensor, d: L.Conv3dDesc, transpose_flip: bool = False) -> torch.Tensor:
    """Forward weight [Cout, Cin, 3, 3, 3] -> the shared-memory image of the tcgen05 line convolution described by
    ``d`` (swizzled [3*Cout x Cin] tiles per (dz, dy) tap; depends on how the input channels split into sources)."""
    chans = tuple(d.src[i].C for i in range(d.nsrc))

    def make():
        n = ops.lib().corrif_conv3d_tc_pack_floats(C.byref(d))
        if n <= 0:
            raise ValueError("conv3d_tc: shape not supported")
        w = weight.detach().contiguous()
        wpk = torch.empty(n, device=weight.device, dtype=torch.float32)
        with ops._rec("conv_pack", 8.0 * w.numel()):
            L.check(ops.lib().corrif_conv3d_tc_pack_weights(C.byref(d), w.data_ptr(), wpk.data_ptr(), int(transpose_flip),
                                                            _stream()), "conv3d_tc_pack_weights")
        ops._count()
        return wpk

    return _cached_pack(weight, ("tc", bool(transpose_flip), chans, d.Cout), make)


def _pack_weights(weight: torch.Tensor, transpose_flip: bool) -> torch.Tensor:
    Cout, Cin, k = weight.shape[0], weight.shape[1], weight.shape[2]
    n = ops.lib().corrif_conv3d_pack_floats(Cin, Cout, k)
    if n <= 0:
        raise ValueError("conv3d: Cin (%d) and Cout (%d) must be multiples of 8, kernel 1 or 3" % (Cin, Cout))
    w = weight.detach().contiguous()
    wpk = torch.zeros(n, device=weight.device, dtype=torch.float32)
    with ops._rec("conv_pack", 8.0 * w.numel()):
        L.check(ops.lib().corrif_conv3d_pack_weights(w.data_ptr(), wpk.data_ptr(), Cin, Cout, k, int(transpose_flip),
                                                     _stream()), "conv3d_pack_weights")
    ops._count()
    return wpk


def conv3d_forward(srcs, wpk, bias, Cout, ksize, pad_mode, relu, out, stats=None):
    d = _desc(srcs, Cout, ksize, pad_mode)
    d.relu, d.wpk, d.bias = int(relu), wpk.data_ptr(), (bias.data_ptr() if bias is not None else None)
    d.out, d.ldo, d.stats = out.data_ptr(), _ld(out), (stats.data_ptr() if stats is not None else None)
    nvox = d.B * d.D * d.H * d.W
    taps = 27 if ksize == 3 else 1
    # algorithmic work: 2 * voxels * taps * Cin * Cout FLOP; bytes = read every source once + write the output once
    with ops._rec("conv3d_fwd", 2.0 * nvox * taps * d.Cin * Cout, "k%d %dx%dx%dx%d %d->%d bytes=%d" % (
            ksize, d.B, d.D, d.H, d.W, d.Cin, Cout, 4 * nvox * (d.Cin + Cout))):
        L.check(ops.lib().corrif_conv3d_fwd(C.byref(d), _stream()), "conv3d_fwd")
    ops._count()


def _small_pointwise(srcs, Cout, ksize) -> bool:
    """1x1x1 with 8 channels on both sides (d1_out at 128^3): the HBM-bound per-voxel kernel (the 16-channel
    instantiation exists in the library but measured slower than the tensor-core kernel: 0.69 vs 0.14 ms at 64^3)."""
    return (ksize == 1 and len(srcs) == 1 and srcs[0].shape[4] == Cout and Cout == 8
            and os.environ.get("CORRIF_CONV1_SMALL", "1") != "0" and srcs[0].shape[0] <= 65535)


def conv1_small(x, weight, bias, out, stats, relu, transpose):
    B, D, H, W, Cc = x.shape
    nvox = D * H * W
    w = weight.detach().contiguous()
    with ops._rec("conv3d_dgrad" if transpose else "conv3d_fwd", 2.0 * B * nvox * Cc * Cc,
                  "k1 %dx%dx%dx%d %d->%d small bytes=%d" % (B, D, H, W, Cc, Cc, 8 * B * nvox * Cc)):
        L.check(ops.lib().corrif_conv1_small_fwd(x.data_ptr(), _ld(x), w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                                 out.data_ptr(), _ld(out), stats.data_ptr() if stats is not None else None,
                                                 B, nvox, Cc, int(relu), int(transpose), _stream()), "conv1_small_fwd")
    ops._count()


def conv3d_forward_auto(srcs, weight, bias, Cout, ksize, pad_mode, relu, out, stats=None):
    """The convolution forward on the tcgen05 line kernel where it applies, else on the warp-level kernel."""
    if _small_pointwise(srcs, Cout, ksize):
        return conv1_small(srcs[0], weight, bias, out, stats, relu, transpose=False)
    d = _desc(srcs, Cout, ksize, pad_mode)
    if not tc_supported(d):
        return conv3d_forward(srcs, pack_weights(weight), bias, Cout, ksize, pad_mode, relu, out, stats)
    wpk = pack_weights_tc(weight, d)
    d.relu, d.wpk, d.bias = int(relu), wpk.data_ptr(), (bias.data_ptr() if bias is not None else None)
    d.out, d.ldo, d.stats = out.data_ptr(), _ld(out), (stats.data_ptr() if stats is not None else None)
    nvox = d.B * d.D * d.H * d.W
    with ops._rec("conv3d_fwd", 2.0 * nvox * 27 * d.Cin * Cout, "k3 %dx%dx%dx%d %d->%d tc bytes=%d" % (
            d.B, d.D, d.H, d.W, d.Cin, Cout, 4 * nvox * (d.Cin + Cout))):
        L.check(ops.lib().corrif_conv3d_tc_fwd(C.byref(d), _stream()), "conv3d_tc_fwd")
    ops._count()


def conv3d_wgrad(srcs, g, dW, ksize, pad_mode):
    Cout = g.shape[4]
    if _small_pointwise(srcs, Cout, ksize) and Cout == 8:
        x = srcs[0]
        rows = x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3]
        with ops._rec("conv3d_wgrad", 2.0 * rows * 64, "k1 %dx%dx%dx%d 8->8 small" % tuple(x.shape[:4])):
            L.check(ops.lib().corrif_conv1_small_wgrad(x.data_ptr(), _ld(x), g.data_ptr(), _ld(g), dW.data_ptr(), rows, 8,
                                                       _stream()), "conv1_small_wgrad")
        ops._count()
        return
    d = _desc(srcs, Cout, ksize, pad_mode)
    nvox = d.B * d.D * d.H * d.W
    taps = 27 if ksize == 3 else 1
    if (ksize == 3 and os.environ.get("CORRIF_WGRAD_TC", "1") != "0"
            and ops.lib().corrif_conv3d_wgrad_tc_supported(C.byref(d))):
        with ops._rec("conv3d_wgrad", 2.0 * nvox * 27 * d.Cin * Cout, "k3 %dx%dx%dx%d %d->%d tc" % (
                d.B, d.D, d.H, d.W, d.Cin, Cout)):
            L.check(ops.lib().corrif_conv3d_wgrad_tc(C.byref(d), g.data_ptr(), _ld(g), dW.data_ptr(), _stream()),
                    "conv3d_wgrad_tc")
        ops._count()
        return
    with ops._rec("conv3d_wgrad", 2.0 * nvox * taps * d.Cin * Cout, "k%d %dx%dx%dx%d %d->%d" % (
            ksize, d.B, d.D, d.H, d.W, d.Cin, Cout)):
        L.check(ops.lib().corrif_conv3d_wgrad(C.byref(d), g.data_ptr(), _ld(g), dW.data_ptr(), _stream()), "conv3d_wgrad")
    ops._count()


def conv3d_dgrad(g, weight, Cin, ksize, pad_mode, dx):
    """dx [B,D,H,W,Cin] = d(cat of the sources) from g = d(pre-activation)."""
    Cout = g.shape[4]
    if _small_pointwise([g], Cin, ksize):
        return conv1_small(g, weight, None, dx, None, relu=False, transpose=True)
    if ksize == 3:
        # tcgen05 line kernel: the adjoint of replicate padding is part of the same pass (no border kernel)
        dt = _desc([g], Cin, ksize, PAD_REPLICATE_ADJOINT if pad_mode == PAD_REPLICATE else PAD_ZEROS)
        if tc_supported(dt):
            wpk_t = pack_weights_tc(weight, dt, transpose_flip=True)
            dt.relu, dt.wpk, dt.bias, dt.out, dt.ldo, dt.stats = 0, wpk_t.data_ptr(), None, dx.data_ptr(), _ld(dx), None
            nvox = dt.B * dt.D * dt.H * dt.W
            with ops._rec("conv3d_dgrad", 2.0 * nvox * 27 * Cin * Cout, "k3 %dx%dx%dx%d %d->%d tc bytes=%d" % (
                    dt.B, dt.D, dt.H, dt.W, Cout, Cin, 4 * nvox * (Cin + Cout))):
                L.check(ops.lib().corrif_conv3d_tc_fwd(C.byref(dt), _stream()), "conv3d_tc_dgrad")
            ops._count()
            return
    wpk_t = pack_weights(weight, transpose_flip=True)
    d = _desc([g], Cin, ksize, PAD_ZEROS)
    d.relu, d.wpk, d.bias, d.out, d.ldo, d.stats = 0, wpk_t.data_ptr(), None, dx.data_ptr(), _ld(dx), None
    nvox = d.B * d.D * d.H * d.W
    taps = 27 if ksize == 3 else 1
    with ops._rec("conv3d_dgrad", 2.0 * nvox * taps * Cin * Cout, "k%d %dx%dx%dx%d %d->%d" % (
            ksize, d.B, d.D, d.H, d.W, Cout, Cin)):
        L.check(ops.lib().corrif_conv3d_fwd(C.byref(d), _stream()), "conv3d_dgrad")
    ops._count()
    if ksize == 3 and pad_mode == PAD_REPLICATE:
        w = weight.detach().permute(2, 3, 4, 0, 1).contiguous()          # [27][Cout][Cin]: coalesced over ci
        with ops._rec("conv3d_dgrad_border", 0.0):
            L.check(ops.lib().corrif_conv3d_dgrad_border(g.data_ptr(), _ld(g), w.data_ptr(), dx.data_ptr(), _ld(dx),
                                                         d.B, d.D, d.H, d.W, Cin, Cout, _stream()), "conv3d_dgrad_border")
        ops._count()


class _ConvBlock(torch.autograd.Function):
    """conv (+bias) [-> ReLU] [-> InstanceNorm3d] over the channel concatenation of 1..3 volumes."""

    @staticmethod
    def forward(ctx, weight, bias, ksize, pad_mode, relu, norm, out_holder, *srcs):
        if relu and not norm:
            raise ValueError("conv_block: ReLU without InstanceNorm does not occur in the model and has no backward")
        srcs = [as_volume(s) for s in srcs]
        Cout = weight.shape[0]
        B, D, H, W = srcs[0].shape[:4]
        dev = weight.device
        # out_holder: (tensor,) = write the block's output into this channel slice of a wider buffer (its consumer then
        # reads the buffer as ONE source: the reference's torch.cat, mmvit4.py:272, without a copy and without narrow rows)
        out = (out_holder[0] if out_holder is not None and out_holder[0] is not None
               else torch.empty(B, D, H, W, Cout, device=dev, dtype=torch.float32))
        if tuple(out.shape) != (B, D, H, W, Cout):
            raise ValueError("conv_block: out has shape %s, expected %s" % (tuple(out.shape), (B, D, H, W, Cout)))
        stats = torch.zeros(B, Cout, 2, device=dev, dtype=torch.float64) if norm else None
        b = bias.detach().contiguous() if bias is not None else None
        conv3d_forward_auto(srcs, weight, b, Cout, ksize, pad_mode, relu, out, stats)
        mean = rstd = None
        if norm:
            mean = torch.empty(B, Cout, device=dev, dtype=torch.float32)
            rstd = torch.empty(B, Cout, device=dev, dtype=torch.float32)
            nvox = D * H * W
            with ops._rec("instnorm_apply", 8.0 * B * nvox * Cout):
                L.check(ops.lib().corrif_instnorm_apply(out.data_ptr(), _ld(out), stats.data_ptr(), mean.data_ptr(),
                                                        rstd.data_ptr(), B, nvox, Cout, EPS, _stream()), "instnorm_apply")
            ops._count()
        ctx.cfg = (ksize, pad_mode, relu, norm, bias is not None, [s.shape[4] for s in srcs])
        ctx.wgrad_holder = out_holder[1] if out_holder is not None and len(out_holder) > 1 else None
        ctx.save_for_backward(weight, out if norm else None, mean, rstd, *srcs)
        return out

    @staticmethod
    def backward(ctx, dy):
        ksize, pad_mode, relu, norm, has_bias, chans = ctx.cfg
        weight, y, mean, rstd, *srcs = ctx.saved_tensors
        dy = as_volume(dy)
        B, D, H, W, Cout = dy.shape
        nvox = D * H * W
        dev = dy.device
        dbias = torch.zeros(Cout, device=dev, dtype=torch.float32) if has_bias else None
        if norm:
            sums = torch.zeros(B, Cout, 2, device=dev, dtype=torch.float64)
            with ops._rec("instnorm_bwd_stats", 8.0 * B * nvox * Cout):
                L.check(ops.lib().corrif_instnorm_bwd_stats(dy.data_ptr(), _ld(dy), y.data_ptr(), _ld(y), sums.data_ptr(),
                                                            B, nvox, Cout, _stream()), "instnorm_bwd_stats")
            g = torch.empty(B, D, H, W, Cout, device=dev, dtype=torch.float32)
            with ops._rec("instnorm_bwd_apply", 12.0 * B * nvox * Cout):
                L.check(ops.lib().corrif_instnorm_relu_bwd_apply(
                    dy.data_ptr(), _ld(dy), y.data_ptr(), _ld(y), mean.data_ptr(), rstd.data_ptr(), sums.data_ptr(),
                    g.data_ptr(), Cout, dbias.data_ptr() if has_bias else None, B, nvox, Cout, int(relu), _stream()),
                    "instnorm_relu_bwd_apply")
            ops._count(2)
        else:
            g = dy
            if has_bias:
                with ops._rec("volume_colsum", 4.0 * B * nvox * Cout):
                    L.check(ops.lib().corrif_volume_colsum(g.data_ptr(), _ld(g), dbias.data_ptr(), B * nvox, Cout,
                                                           _stream()), "volume_colsum")
                ops._count()
        dW = None
        if ctx.needs_input_grad[0]:
            dW = torch.zeros_like(weight, memory_format=torch.contiguous_format)
            conv3d_wgrad(srcs, g, dW, ksize, pad_mode)
        elif ctx.wgrad_holder is not None:
            # the weight gradient runs in its own autograd node on a side stream (_WgradLater): hand it the operands
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            ctx.wgrad_holder.update(srcs=srcs, g=g, ksize=ksize, pad_mode=pad_mode, ready=ev,
                                    main=torch.cuda.current_stream())
        dsrcs: List[Optional[torch.Tensor]] = [None] * len(srcs)
        if any(ctx.needs_input_grad[7:]):
            cin = sum(chans)
            dx = torch.empty(B, D, H, W, cin, device=dev, dtype=torch.float32)
            conv3d_dgrad(g, weight, cin, ksize, pad_mode, dx)
            off = 0
            for i, c in enumerate(chans):
                if ctx.needs_input_grad[7 + i]:
                    dsrcs[i] = dx[..., off:off + c]
                off += c
        return (dW, dbias, None, None, None, None, None, *dsrcs)


class _WgradLater(torch.autograd.Function):
    """The weight gradient of a conv block as an autograd node of its own whose forward "ran" on a side stream, so the
    engine executes its backward there (and joins that stream at the end of the backward pass, also under graph
    capture).  The block's own backward (data path) stashes g = d(pre-activation) in ``holder`` and goes on; the
    HBM-bound passes of the NEXT block's backward then overlap this tensor-bound kernel."""

    @staticmethod
    def forward(ctx, weight, holder):
        ctx.holder = holder
        ctx.save_for_backward(weight)
        return weight.new_empty(0)

    @staticmethod
    def backward(ctx, _token_grad):
        (weight,) = ctx.saved_tensors
        h = ctx.holder
        if "g" not in h:                                   # the block's backward did not run (its output was unused)
            return None, None
        side = torch.cuda.current_stream()
        side.wait_event(h["ready"])
        for t in (h["g"], *h["srcs"]):
            t.record_stream(side)
        dW = torch.zeros_like(weight, memory_format=torch.contiguous_format)
        conv3d_wgrad(h["srcs"], h["g"], dW, h["ksize"], h["pad_mode"])
        dW.record_stream(h["main"])
        h.clear()
        return dW, None


class _Tap(torch.autograd.Function):
    """out -> out, with a second (empty) input whose gradient edge makes the engine run _WgradLater."""

    @staticmethod
    def forward(ctx, out, token):
        return out.view_as(out)

    @staticmethod
    def backward(ctx, g):
        return g, g.new_empty(0)


_WGRAD_STREAMS: dict = {}


def _wgrad_stream(dev: torch.device):
    """Side stream of the deferred weight gradients (CORRIF_WGRAD_STREAM=0 keeps them inside the block's backward)."""
    if os.environ.get("CORRIF_WGRAD_STREAM", "1") == "0":
        return None
    key = (dev.type, dev.index)
    if key not in _WGRAD_STREAMS:
        _WGRAD_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _WGRAD_STREAMS[key]


def conv_block(srcs: Sequence[torch.Tensor], weight: torch.Tensor, bias: Optional[torch.Tensor], ksize: int,
               pad_mode: int = PAD_ZEROS, relu: bool = True, norm: bool = True,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """general_conv3d_prenorm / EarlyFusionBlock (relu=norm=True) or a plain biased convolution (relu=norm=False)
    over cat(srcs, channels).  srcs and the result are channels-last volumes [B, D, H, W, C]."""
    if not 1 <= len(srcs) <= 3:
        raise ValueError("conv_block takes 1..3 sources")
    # the block's own backward must run (it produces g): some other input has to carry a gradient
    chain = (bias is not None and bias.requires_grad) or any(s.requires_grad for s in srcs)
    side = _wgrad_stream(weight.device) if (ksize == 3 and weight.is_cuda and weight.requires_grad and chain
                                            and torch.is_grad_enabled()) else None
    if side is None:
        return _ConvBlock.apply(weight, bias, ksize, pad_mode, relu, norm, (out,) if out is not None else None, *srcs)
    holder: dict = {}
    with torch.cuda.stream(side):                          # created FIRST: the engine then runs it after the block's backward
        token = _WgradLater.apply(weight, holder)
    y = _ConvBlock.apply(weight.detach(), bias, ksize, pad_mode, relu, norm, (out, holder), *srcs)
    return _Tap.apply(y, token)


class _PointwiseGemm(torch.autograd.Function):
    """Bias-only 1x1x1 convolution with many channels and few voxels (the encoders' adapt1-5 / conv6,
    mmvit4.py:157-164, and RFM5_reduce, :231) as ONE tcgen05 TF32 GEMM [voxels x Cin] . [Cout x Cin]^T + bias - the
    kernel family of the fusion block (corrif_gemm), which at these shapes (K up to 2048, 1.5 k-98 k rows) is the
    right tool; the window-staging convolution kernel would run a handful of CTAs through 64 channel passes.
    ``channels_first``: x is a contiguous [B, Cin, D, H, W] tensor (cuDNN's output); the NCDHW -> channels-last
    transpose then doubles as the TF32 round-to-nearest pass of the A operand.  Returns a volume [B, D, H, W, Cout]."""

    @staticmethod
    def forward(ctx, x, weight, bias, channels_first):
        Cout, Cin = weight.shape[0], weight.shape[1]
        dev = weight.device
        if channels_first:
            x = x.contiguous()
            B, _, D, H, W = x.shape
            S = D * H * W
            xt = torch.empty(B * S, Cin, device=dev, dtype=torch.float32)
            ops.transpose(x, xt, B, Cin, S, round_out=True)
        else:
            # channels-last input: the GEMM reads it in place.  The tensor core truncates it to TF32; the mean of that
            # truncation (-3.52e-4, see csrc/conv3d_tc.cu) is folded into the rounded weight copy instead of spending a
            # rounding pass (one read + one write of the activation, forward and again for dy backward)
            x = x.contiguous()
            B, D, H, W, _ = x.shape
            S = D * H * W
            xt = x.view(B * S, Cin)
        comp = not channels_first
        wsrc = weight.detach().contiguous()
        if comp:
            wsrc = wsrc * _TRUNC_COMP
        wr = torch.empty(Cout, Cin, device=dev, dtype=torch.float32)
        ops.round_tf32(wsrc, wr, wr.numel())
        out = torch.empty(B, D, H, W, Cout, device=dev, dtype=torch.float32)
        ops.gemm(xt, wr, out, M=B * S, N=Cout, K=Cin, lda=Cin, ldb=Cin, ldd=Cout, bias=bias.detach().contiguous(),
                 epilogue=ops.EPI_BIAS, tag="pointwise")
        ctx.save_for_backward(xt, wr)
        ctx.cfg = (channels_first, (B, D, H, W), comp)
        return out

    @staticmethod
    def backward(ctx, dy):
        xt, wr = ctx.saved_tensors
        channels_first, (B, D, H, W), comp = ctx.cfg
        Cout, Cin = wr.shape
        R = B * D * H * W
        dev = dy.device
        dyc = dy.contiguous()
        if comp:
            dyr = dyc.view(R, Cout)                       # truncated by the tensor core; compensated below
        else:
            dyr = torch.empty(R, Cout, device=dev, dtype=torch.float32)
            ops.round_tf32(dyc, dyr, dyr.numel())
        dbias = torch.zeros(Cout, device=dev, dtype=torch.float32)
        with ops._rec("volume_colsum", 4.0 * R * Cout):
            L.check(ops.lib().corrif_volume_colsum(dyc.data_ptr(), Cout, dbias.data_ptr(), R, Cout, _stream()), "volume_colsum")
        ops._count()
        dW = torch.zeros(Cout, Cin, 1, 1, 1, device=dev, dtype=torch.float32)
        split = max(1, min(64, R // 2048)) if R % 32 == 0 else 1
        # both operands truncated (comp): (1 + 3.52e-4)^2 restores the mean; wr already carries one factor for dy below
        ops.gemm(dyr, xt, dW, M=Cout, N=Cin, K=R, lda=Cout, ldb=Cin, ldd=Cin, a_mn=True, b_mn=True, split_k=split,
                 epilogue=ops.EPI_ATOMIC_ADD, alpha=_TRUNC_COMP * _TRUNC_COMP if comp else 1.0, tag="pointwise_wgrad")
        dx = None
        if ctx.needs_input_grad[0]:
            dxt = torch.empty(R, Cin, device=dev, dtype=torch.float32)
            ops.gemm(dyr, wr, dxt, M=R, N=Cin, K=Cout, lda=Cout, ldb=Cin, ldd=Cin, b_mn=True, tag="pointwise_dgrad")
            if channels_first:
                dx = torch.empty(B, Cin, D, H, W, device=dev, dtype=torch.float32)
                ops.transpose(dxt, dx, B, D * H * W, Cin)
            else:
                dx = dxt.view(B, D, H, W, Cin)
        return dx, dW, dbias, None


def pointwise_conv(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, channels_first: bool = False) -> torch.Tensor:
    """nn.Conv3d(kernel_size=1) with bias, no activation / norm, as a tcgen05 GEMM; see _PointwiseGemm."""
    if weight.shape[0] % 4 or weight.shape[1] % 4:
        raise ValueError("pointwise_conv: channel counts must be multiples of 4")
    if channels_first and x.dim() == 5 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last_3d):
        x, channels_first = x.permute(0, 2, 3, 4, 1), False          # cuDNN already produced the volume layout
    return _PointwiseGemm.apply(x, weight, bias, channels_first)


class _BatchNormReLU(torch.autograd.Function):
    """Train-mode BatchNorm3d (+ residual add) (+ ReLU) on a channels-last volume: statistics pass + one normalise pass
    forward, statistics + apply backward (include/corrif.h, corrif_batchnorm_*).  Returns (y, mean, biased var)."""
    CH = 1024                                            # channels per launch

    @staticmethod
    def forward(ctx, x, gamma, beta, res, relu, eps):
        x = as_volume(x)
        res = as_volume(res) if res is not None else None
        B, D, H, W, Cc = x.shape
        rows, ldx, dev = B * D * H * W, _ld(x), x.device
        y = torch.empty(B, D, H, W, Cc, device=dev, dtype=torch.float32)
        stats = torch.zeros(Cc, 2, device=dev, dtype=torch.float64)
        mean, var, rstd = (torch.empty(Cc, device=dev, dtype=torch.float32) for _ in range(3))
        g, bt = gamma.detach().contiguous(), beta.detach().contiguous()
        lib = ops.lib()
        for c0 in range(0, Cc, _BatchNormReLU.CH):
            c = min(_BatchNormReLU.CH, Cc - c0)
            o = 4 * c0
            with ops._rec("batchnorm_stats", 4.0 * rows * c):
                L.check(lib.corrif_instnorm_bwd_stats(x.data_ptr() + o, ldx, x.data_ptr() + o, ldx, stats.data_ptr() + 16 * c0,
                                                      1, rows, c, _stream()), "batchnorm stats")
            with ops._rec("batchnorm_fwd", (12.0 if res is not None else 8.0) * rows * c):
                L.check(lib.corrif_batchnorm_fwd(x.data_ptr() + o, ldx, stats.data_ptr() + 16 * c0, g.data_ptr() + o,
                                                 bt.data_ptr() + o, (res.data_ptr() + o) if res is not None else None,
                                                 _ld(res) if res is not None else 0, y.data_ptr() + o, Cc,
                                                 mean.data_ptr() + o, var.data_ptr() + o, rstd.data_ptr() + o, rows, c,
                                                 float(eps), int(relu), _stream()), "batchnorm_fwd")
            ops._count(2)
        ctx.save_for_backward(x, y if relu else None, mean, rstd, g)
        ctx.cfg = (bool(relu), res is not None)
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, dy, _dmean, _dvar):
        x, y, mean, rstd, g = ctx.saved_tensors
        relu, has_res = ctx.cfg
        dy = as_volume(dy)
        B, D, H, W, Cc = x.shape
        rows, dev = B * D * H * W, x.device
        sums = torch.zeros(Cc, 2, device=dev, dtype=torch.float64)
        dx = torch.empty(B, D, H, W, Cc, device=dev, dtype=torch.float32)
        dres = torch.empty(B, D, H, W, Cc, device=dev, dtype=torch.float32) if has_res else None
        dgamma, dbeta = torch.empty(Cc, device=dev), torch.empty(Cc, device=dev)
        lib = ops.lib()
        ldx, lddy = _ld(x), _ld(dy)
        for c0 in range(0, Cc, _BatchNormReLU.CH):
            c = min(_BatchNormReLU.CH, Cc - c0)
            o = 4 * c0
            yp = (y.data_ptr() + o) if relu else None
            with ops._rec("batchnorm_bwd_stats", (12.0 if relu else 8.0) * rows * c):
                L.check(lib.corrif_batchnorm_bwd_stats(dy.data_ptr() + o, lddy, yp, Cc, x.data_ptr() + o, ldx, mean.data_ptr() + o,
                                                       rstd.data_ptr() + o, sums.data_ptr() + 16 * c0, rows, c, int(relu),
                                                       _stream()), "batchnorm_bwd_stats")
            with ops._rec("batchnorm_bwd_apply", ((16.0 if relu else 12.0) + (4.0 if has_res else 0.0)) * rows * c):
                L.check(lib.corrif_batchnorm_bwd_apply(dy.data_ptr() + o, lddy, yp, Cc, x.data_ptr() + o, ldx, mean.data_ptr() + o,
                                                       rstd.data_ptr() + o, g.data_ptr() + o, sums.data_ptr() + 16 * c0,
                                                       dx.data_ptr() + o, Cc, (dres.data_ptr() + o) if has_res else None, Cc,
                                                       dgamma.data_ptr() + o, dbeta.data_ptr() + o, rows, c, int(relu),
                                                       _stream()), "batchnorm_bwd_apply")
            ops._count(2)
        return dx, dgamma, dbeta, dres, None, None


def batchnorm_relu(x: torch.Tensor, bn: torch.nn.BatchNorm3d, residual: Optional[torch.Tensor] = None,
                   relu: bool = True) -> torch.Tensor:
    """y = relu?(BatchNorm3d(x) (+ residual)) in TRAIN mode on channels-last volumes [B, D, H, W, C], updating the
    module's running statistics like nn.BatchNorm3d does (momentum, unbiased running variance)."""
    y, mean, var = _BatchNormReLU.apply(x, bn.weight, bn.bias, residual, relu, bn.eps)
    if bn.track_running_stats and bn.running_mean is not None:
        with torch.no_grad():
            n = x.numel() // x.shape[-1]
            m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked.item() + 1)
            bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
            bn.running_var.mul_(1 - m).add_(var, alpha=m * n / max(n - 1, 1))
            bn.num_batches_tracked += 1
    return y


def _separable(shape_in, size) -> bool:
    """Trilinear resizes that grow the volume (the decoder's 2x up-sampling, mmvit4.py:269) run as three streaming
    one-axis passes; the small / strongly down-sampling ones keep the single gather kernel."""
    B, Di, Hi, Wi, Cc = shape_in
    Do, Ho, Wo = size
    if os.environ.get("CORRIF_RESIZE_SEPARABLE", "1") == "0" or B * Do * Ho * Wo * Cc < (1 << 22) or B * Do * Ho * Wo * Cc >= (1 << 33):
        return False
    return all(o >= i and (i > 1 and (o - 1) <= 3 * (i - 1) or i == o) for i, o in ((Di, Do), (Hi, Ho), (Wi, Wo)))


def _linear_axes(t: torch.Tensor, sizes_in, sizes_out, Cc: int, backward: bool) -> torch.Tensor:
    """Apply the one-axis passes (forward: x, y, z; adjoint: z, y, x) to a contiguous [B, D, H, W, C] tensor."""
    B = t.shape[0]
    dims = list(sizes_out if backward else sizes_in)          # current D, H, W
    lib = ops.lib()
    order = (0, 1, 2) if backward else (2, 1, 0)
    for ax in order:
        n_from, n_to = (sizes_out[ax], sizes_in[ax]) if backward else (sizes_in[ax], sizes_out[ax])
        if n_from == n_to:
            continue
        outer = B
        for a in range(ax):
            outer *= dims[a]
        inner = Cc
        for a in range(ax + 1, 3):
            inner *= dims[a]
        new_dims = list(dims)
        new_dims[ax] = n_to
        out = torch.empty(B, *new_dims, Cc, device=t.device, dtype=torch.float32)
        tag = "resize_trilinear_" + ("bwd" if backward else "fwd")
        with ops._rec(tag, 4.0 * (t.numel() + out.numel()), "axis%d %dx%dx%dx%d->%d C%d" % (ax, B, *dims, n_to, Cc)):
            if backward:
                L.check(lib.corrif_resize_linear_axis_bwd(t.data_ptr(), out.data_ptr(), outer, n_to, n_from, inner, _stream()),
                        "resize_linear_axis_bwd")
            else:
                L.check(lib.corrif_resize_linear_axis_fwd(t.data_ptr(), out.data_ptr(), outer, n_from, n_to, inner, _stream()),
                        "resize_linear_axis_fwd")
        ops._count()
        t, dims = out, new_dims
    return t


class _Resize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size, mode, out_holder=None):
        x = as_volume(x)
        B, Di, Hi, Wi, Cc = x.shape
        Do, Ho, Wo = size
        if mode == "trilinear" and out_holder is None and _separable(x.shape, size):
            ctx.cfg = (mode, (Di, Hi, Wi), size)
            return _linear_axes(x.contiguous(), (Di, Hi, Wi), size, Cc, backward=False)
        y = out_holder[0] if out_holder is not None else torch.empty(B, Do, Ho, Wo, Cc, device=x.device, dtype=torch.float32)
        if tuple(y.shape) != (B, Do, Ho, Wo, Cc):
            raise ValueError("resize: out has shape %s, expected %s" % (tuple(y.shape), (B, Do, Ho, Wo, Cc)))
        fn = ops.lib().corrif_resize_trilinear_fwd if mode == "trilinear" else ops.lib().corrif_resize_nearest_fwd
        with ops._rec("resize_" + mode + "_fwd", 4.0 * Cc * B * (Di * Hi * Wi + Do * Ho * Wo),
                      "%dx%dx%dx%d->%dx%dx%d C%d" % (B, Di, Hi, Wi, Do, Ho, Wo, Cc)):
            L.check(fn(x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), B, Cc, Di, Hi, Wi, Do, Ho, Wo, _stream()), "resize_fwd")
        ops._count()
        ctx.cfg = (mode, (Di, Hi, Wi), size)
        return y

    @staticmethod
    def backward(ctx, dy):
        mode, (Di, Hi, Wi), (Do, Ho, Wo) = ctx.cfg
        dy = as_volume(dy)
        B, Cc = dy.shape[0], dy.shape[4]
        if mode == "trilinear" and _separable((B, Di, Hi, Wi, Cc), (Do, Ho, Wo)):
            return _linear_axes(dy.contiguous(), (Di, Hi, Wi), (Do, Ho, Wo), Cc, backward=True), None, None, None
        dx = torch.empty(B, Di, Hi, Wi, Cc, device=dy.device, dtype=torch.float32)
        fn = ops.lib().corrif_resize_trilinear_bwd if mode == "trilinear" else ops.lib().corrif_resize_nearest_bwd
        with ops._rec("resize_" + mode + "_bwd", 4.0 * Cc * B * (Di * Hi * Wi + Do * Ho * Wo),
                      "%dx%dx%dx%d->%dx%dx%d C%d" % (B, Di, Hi, Wi, Do, Ho, Wo, Cc)):
            L.check(fn(dy.data_ptr(), _ld(dy), dx.data_ptr(), Cc, B, Cc, Di, Hi, Wi, Do, Ho, Wo, _stream()), "resize_bwd")
        ops._count()
        return dx, None, None, None


def resize_trilinear(x: torch.Tensor, size: Tuple[int, int, int]) -> torch.Tensor:
    """F.interpolate(mode='trilinear', align_corners=True) / nn.Upsample on a channels-last volume."""
    return _Resize.apply(x, tuple(int(s) for s in size), "trilinear")


def resize_nearest(x: torch.Tensor, size: Tuple[int, int, int], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """F.interpolate(x, size) (default nearest, mmvit4.py:271) on a channels-last volume; ``out``: write into this
    channel slice of a wider buffer."""
    return _Resize.apply(x, tuple(int(s) for s in size), "nearest", (out,) if out is not None else None)
