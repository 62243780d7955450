"""Drop-in for the reference's ``mmvit4`` module: ``MMVit4(num_cls=1)`` with the same ``forward(x)``
contract ([B,3,3,H,W] -> [B,3,1,224,224] sigmoid probabilities) and the SAME 1140 ``state_dict`` keys
and shapes, so ``iremmodel{i}.pt`` / ``Finaliremmodel{i}.pt`` checkpoints load strictly in both
directions (reference F4_TRAIN.py:84-86,180).

What differs is where the work runs.  The fusion hot path (reference mmvit4.py:456-529) is one call to
``torch.ops.corrif.fusion_block``; the six EarlyFusionBlocks (:64-81, 449-454) and the whole decoder
(:29-56, 222-292) run on the channels-last volume kernels of corrif_b200.volume (conv -> ReLU ->
InstanceNorm as one fused block with a hand-written backward, trilinear / nearest resizes) - all
hand-written sm_100a kernels behind the C ABI (include/corrif.h).  Only the modality encoders'
ResNet-50 trunks (carried on stock PyTorch / cuDNN) and the final 8 -> 3
channel 1x1x1 convolution + sigmoid on the 224^2 output remain ATen calls.  The sub-modules are
re-stated here from the architecture description, table-driven, because the class has to own their
parameters under the reference's names.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from corrif_b200 import fusion as _fusion  # noqa: E402
from corrif_b200 import module as _module  # noqa: E402  (registers torch.ops.corrif.*)
from corrif_b200 import volume as _V  # noqa: E402

basic_dims = 8
transformer_basic_dims = 512
mlp_dim = 512
num_heads = 8
depth = 1
num_modals = 3
patch_size = 8

_MODS = ("RGB", "NIR", "SWIR")
# Measured switches.  CORRIF_FUSED_BN=1 runs the encoders' train-mode BatchNorm (+ residual)
# (+ ReLU) on the fused channels-last kernels: 1.5 ms less GPU time per micro-batch but 636 more Python-level launches,
# and the step is host-bound there (93.4 vs 97.9 imgs/s) - off until the encoder step is captured in a CUDA graph.
# CORRIF_ENCODER_NCDHW=1 keeps the encoder trunks in PyTorch's default memory format instead of channels_last_3d.
_FUSED_BN = os.environ.get("CORRIF_FUSED_BN") == "1"
_ENC_CL3D = os.environ.get("CORRIF_ENCODER_NCDHW") != "1"
_ENC_STREAMS = os.environ.get("CORRIF_ENCODER_STREAMS", "1") != "0"      # the three encoders on three streams


# ----------------------------------------------------------------------------------------------
# encoder: ResNet-50 topology inflated to (1,k,k) 3-D convolutions (reference mmvit4.py:83-212)
# ----------------------------------------------------------------------------------------------
def _conv2d_as_3d(cin, cout, k, stride, pad, depth_k=1):
    return nn.Conv3d(cin, cout, (depth_k, k, k), stride=(1, stride, stride),
                     padding=(depth_k // 2, pad, pad), bias=False)


class Bottleneck3D(nn.Module):
    def __init__(self, cin, planes, stride, project):
        super().__init__()
        self.conv1, self.bn1 = _conv2d_as_3d(cin, planes, 1, 1, 0), nn.BatchNorm3d(planes)
        self.conv2, self.bn2 = _conv2d_as_3d(planes, planes, 3, stride, 1), nn.BatchNorm3d(planes)
        self.conv3, self.bn3 = _conv2d_as_3d(planes, 4 * planes, 1, 1, 0), nn.BatchNorm3d(4 * planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if project:
            self.downsample = nn.Sequential(_conv2d_as_3d(cin, 4 * planes, 1, stride, 0),
                                            nn.BatchNorm3d(4 * planes))

    @staticmethod
    def _bn(bn, x, residual=None, relu=True):
        """BatchNorm (+ residual) (+ ReLU): in training on the fused channels-last kernels (one normalise pass instead of
        cuDNN batch-norm + add + ReLU), in eval mode (running statistics) on stock PyTorch."""
        if not (bn.training and x.is_cuda and _FUSED_BN):
            if bn.training and bn.track_running_stats and bn.momentum is not None and getattr(bn, "_deferred_count", False):
                # same arithmetic as nn.BatchNorm3d.forward; the num_batches_tracked increment of all the encoder's
                # BatchNorms is ONE multi-tensor add at the end of Encoder.forward instead of 53 one-element kernels
                y = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps)
            else:
                y = bn(x)
            if residual is not None:
                y = y + residual
            return F.relu(y) if relu else y
        r = None if residual is None else _V.to_channels_last(residual)
        return _V.to_channels_first(_V.batchnorm_relu(_V.to_channels_last(x), bn, r, relu))

    def forward(self, x):
        y = self._bn(self.bn1, self.conv1(x))
        y = self._bn(self.bn2, self.conv2(y))
        identity = x if self.downsample is None else self._bn(self.downsample[1], self.downsample[0](x), relu=False)
        return self._bn(self.bn3, self.conv3(y), residual=identity)


def _stage(cin, planes, blocks, stride):
    layers = [Bottleneck3D(cin, planes, stride, True)]
    layers += [Bottleneck3D(4 * planes, planes, 1, False) for _ in range(blocks - 1)]
    return nn.Sequential(*layers)


class Encoder(nn.Module):
    """One modality encoder.  Returns the five adapted pyramid levels and x6 [B,64,8,8,8].  The ResNet-50 trunk's
    convolutions and max-pool are stock PyTorch / cuDNN in channels_last_3d memory format, i.e.
    directly on the volume layout; its train-mode BatchNorm (+ residual) (+ ReLU) and the tail run on corrif_b200.volume
    kernels."""

    def __init__(self, inflate_time=3):
        super().__init__()
        self.e1_c1 = _conv2d_as_3d(1, 64, 7, 2, 3, depth_k=inflate_time)
        self.e1_bn = nn.BatchNorm3d(64)
        self.e1_relu = nn.ReLU(inplace=True)
        self.e1_mp = nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        self.e2 = _stage(64, 64, 3, 1)
        self.e3 = _stage(256, 128, 4, 2)
        self.e4 = _stage(512, 256, 6, 2)
        self.e5 = _stage(1024, 512, 3, 2)
        widths = (basic_dims, basic_dims * 2, basic_dims * 4, basic_dims * 8, basic_dims * 8)
        self.conv6 = nn.Conv3d(sum(widths), basic_dims * 8, kernel_size=1)
        for i, (cin, cout) in enumerate(zip((64, 256, 512, 1024, 2048), widths), start=1):
            setattr(self, f"adapt{i}", nn.Conv3d(cin, cout, kernel_size=1))

    def forward(self, x):
        bns = [m for m in self.modules() if isinstance(m, nn.BatchNorm3d)]
        defer = self.training and x.is_cuda and not _FUSED_BN
        for m in bns:
            m._deferred_count = defer
        if defer:
            torch._foreach_add_([m.num_batches_tracked for m in bns if m.track_running_stats and m.momentum is not None], 1)
        f1 = self.e1_mp(Bottleneck3D._bn(self.e1_bn, self.e1_relu(self.e1_c1(x)), relu=False))   # ReLU BEFORE BN, as the reference
        f2 = self.e2(f1)
        f3 = self.e3(f2)
        f4 = self.e4(f3)
        f5 = self.e5(f4)
        # encoder tail (reference mmvit4.py:181-193) on the channels-last volume kernels: adapt1-5 are bias-only
        # 1x1x1 convolutions, the five trilinear resizes to 8^3 and conv6 over their concatenation follow
        lv = []
        for i, f in enumerate((f1, f2, f3, f4, f5), start=1):
            a = getattr(self, f"adapt{i}")
            lv.append(_V.pointwise_conv(f, a.weight, a.bias, channels_first=True))         # tcgen05 GEMM
        pooled = torch.cat([_V.resize_trilinear(t, (8, 8, 8)) for t in lv], dim=-1)       # [B,8,8,8,184]
        x6 = _V.pointwise_conv(pooled, self.conv6.weight, self.conv6.bias)
        # same shapes as the reference's outputs ([B,C,D,H,W]); the memory underneath stays channels-last
        return tuple(_V.to_channels_first(t) for t in (*lv, x6))


class EarlyFusionBlock(nn.Module):
    """cat(3 modalities) -> 1x1x1 conv -> ReLU -> InstanceNorm3d (reference mmvit4.py:64-81) as ONE fused block on
    channels-last volumes: the three sources are concatenated by the convolution's loader.  Takes the encoders'
    [B,c,D,H,W] maps, returns a volume [B,D,H,W,3c]."""

    def __init__(self, in_channels):
        super().__init__()
        c = num_modals * in_channels
        self.conv = nn.Conv3d(c, c, kernel_size=1)          # parameter holder: keys conv.weight / conv.bias
        self.norm = nn.InstanceNorm3d(c)                    # no parameters, no buffers (affine=False)

    def forward(self, a, b, c):
        srcs = [_V.to_channels_last(t) for t in (a, b, c)]
        return _V.conv_block(srcs, self.conv.weight, self.conv.bias, 1, _V.PAD_ZEROS)


# ----------------------------------------------------------------------------------------------
# decoder (reference mmvit4.py:29-56, 222-292) on channels-last volumes [B,D,H,W,C]
# ----------------------------------------------------------------------------------------------
class general_conv3d_prenorm(nn.Module):
    """Conv3d (k=1 or 3, stride 1, 'same' padding of pad_type) -> ReLU -> InstanceNorm3d (reference mmvit4.py:29-45)
    as one fused block: padding resolved by the convolution's loader, norm statistics from its epilogue.
    ``forward(*sources)`` convolves the channel concatenation of its sources (the reference's torch.cat, :272)."""

    def __init__(self, in_ch, out_ch, k_size=3, stride=1, padding=1, pad_type="zeros"):
        super().__init__()
        if stride != 1 or padding != k_size // 2 or k_size not in (1, 3):
            raise ValueError("general_conv3d_prenorm: the model only uses stride 1, 'same' padding, k in {1,3}")
        self.conv = nn.Conv3d(in_ch, out_ch, k_size, stride=stride, padding=padding,
                              padding_mode=pad_type, bias=True)
        self.norm = nn.InstanceNorm3d(out_ch)
        self.k, self.pad_mode = k_size, (_V.PAD_REPLICATE if pad_type == "replicate" else _V.PAD_ZEROS)

    def forward(self, *xs, out=None):
        return _V.conv_block(xs, self.conv.weight, self.conv.bias, self.k, self.pad_mode, out=out)


class fusion_prenorm(nn.Module):
    def __init__(self, in_channel):
        super().__init__()
        self.fusion_layer = nn.Sequential(
            general_conv3d_prenorm(in_channel, in_channel, k_size=1, padding=0),
            general_conv3d_prenorm(in_channel, in_channel, k_size=3, padding=1),
            general_conv3d_prenorm(in_channel, in_channel, k_size=1, padding=0))

    def forward(self, x):
        return self.fusion_layer(x)


class Decoder_fuse(nn.Module):
    # (level, skip channels, in channels of *_c1, out channels, cube size the skip is resized to)
    _LEVELS = ((4, 192, 128, 64, 16), (3, 96, 64, 32, 32), (2, 48, 32, 16, 64), (1, 24, 16, 8, 128))

    def __init__(self, num_cls=1):
        super().__init__()
        rep = dict(pad_type="replicate")
        for lvl, skip, cin, cout, _ in self._LEVELS:
            c1_out = cin if lvl == 4 else cout
            setattr(self, f"d{lvl}_c1", general_conv3d_prenorm(cin, c1_out, **rep))
            setattr(self, f"d{lvl}_c2", general_conv3d_prenorm(skip + c1_out, cout, **rep))
            setattr(self, f"d{lvl}_out", general_conv3d_prenorm(cout, cout, k_size=1, padding=0, **rep))
            setattr(self, f"RFM{lvl}", fusion_prenorm(skip))
        for name, cin in (("seg_d4", 64), ("seg_d3", 64), ("seg_d2", 32), ("seg_d1", 16), ("seg_layer", 8)):
            setattr(self, name, nn.Conv3d(cin, num_cls, kernel_size=1))          # unused in forward
        self.RFM5 = fusion_prenorm(192)
        self.RFM5_reduce = nn.Conv3d(192, 128, kernel_size=1)
        self.final_conv = nn.Conv3d(8, 3, kernel_size=1)

    def forward(self, x1, x2, x3, x4, x5):
        """x1..x4: early-fusion volumes [B,D,H,W,C]; x5: x6_inter [B,8,8,8,192] -> sigmoid probs [B,3,1,224,224]."""
        # The skip branches (RFM_l -> nearest resize, :271) depend only on the early-fusion maps: they run on a side
        # stream, level 4 first, beside the serial up-path; an event per level hands each result over.  The 1.6 GB
        # nearest up-sampling of level 1 (bandwidth-bound) then overlaps the small-grid convolutions of levels 4-2.
        skips = (x4, x3, x2, x1)
        side_ok = _ENC_STREAMS and x5.is_cuda
        ready = [None] * 4
        s_list = [None] * 4
        if side_ok:
            cur = torch.cuda.current_stream()
            if getattr(self, "_side", None) is None or self._side.device != x5.device:
                self._side = torch.cuda.Stream(device=x5.device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                for i, ((lvl, _, _, _, cube), skip) in enumerate(zip(self._LEVELS, skips)):
                    s_list[i] = _V.resize_nearest(getattr(self, f"RFM{lvl}")(skip), (cube, cube, cube))
                    ready[i] = torch.cuda.Event()
                    ready[i].record(self._side)
        y = self.RFM5(x5)
        y = _V.pointwise_conv(y, self.RFM5_reduce.weight, self.RFM5_reduce.bias)           # tcgen05 GEMM
        for i, ((lvl, _, _, _, cube), skip) in enumerate(zip(self._LEVELS, skips)):
            up = _V.resize_trilinear(y, tuple(2 * d for d in y.shape[1:4]))                 # self.up2 (:269)
            # Measured and not used: letting d*_c1 and the nearest resize write side by side into ONE buffer
            # (conv_block(out=...), resize_nearest(out=...)) so that d*_c2 reads a single 128-byte-row source makes its
            # forward 0.4 ms faster (1.11 -> 0.71 ms at 128^3) but the in-place InstanceNorm passes over the 8-of-32
            # channel slice lose as much (0.92 -> 1.30 ms apply, 1.54 -> 1.78 ms backward statistics).
            y = getattr(self, f"d{lvl}_c1")(up)
            if side_ok:
                cur.wait_event(ready[i])
                s = s_list[i]
                s.record_stream(cur)
            else:
                s = _V.resize_nearest(getattr(self, f"RFM{lvl}")(skip), (cube, cube, cube))     # F.interpolate (:271)
            y = getattr(self, f"d{lvl}_out")(getattr(self, f"d{lvl}_c2")(s, y))             # cat((s, y)) (:272)
        # nn.Upsample(size=(1,224,224), trilinear, align_corners=True) (:263, 288): with ONE output slice along
        # depth the source index is 0 for every output voxel, i.e. only depth slice 0 of the 128^3 volume is read
        top = _V.resize_trilinear(y[:, 0:1], (1, 224, 224))
        return torch.sigmoid(self.final_conv(_V.to_channels_first(top)))


# ----------------------------------------------------------------------------------------------
class MMVit4(nn.Module):
    def __init__(self, num_cls=1, dropout_rate=0.1, precision="tf32"):
        super().__init__()
        C, E = transformer_basic_dims, basic_dims * 8
        self.dropout_rate, self.precision = dropout_rate, precision
        self._step, self.base_seed, self.device_seed = 0, _module.default_base_seed(), False
        for m in _MODS:
            setattr(self, f"{m}_encoder", Encoder().to(memory_format=torch.channels_last_3d) if _ENC_CL3D else Encoder())
        for m in _MODS:
            setattr(self, f"{m}_encode_conv", nn.Conv3d(E, C, 1))
        self.fused6_encode_conv = nn.Conv3d(E * 3, C, 1)
        for m in _MODS:
            setattr(self, f"{m}_decode_conv", nn.Conv3d(C, E, 1))                 # unused in forward
        for m in _MODS + ("fused6",):
            setattr(self, f"{m}_pos", nn.Parameter(torch.zeros(1, patch_size ** 3, C)))
        # transformer parameters live under the reference's dotted names
        for prefix in [f"{m}_transformer" for m in _MODS] + ["multimodal_transformer"]:
            for key in _fusion.transformer_keys(prefix).values():
                shape = _module.fusion_param_shapes()[key]
                p = nn.Parameter(torch.empty(shape))
                if key.endswith("norm.weight"):
                    nn.init.ones_(p)
                elif key.endswith(".bias"):
                    nn.init.zeros_(p)
                else:
                    nn.init.kaiming_uniform_(p, a=5 ** 0.5)
                _module._attach(self, key, p)
        for m in _MODS:
            setattr(self, f"qkv_{m}", nn.Conv3d(C, C * 3, 1))
        self.multimodal_decode_conv = nn.Conv3d(C * 4, E * 3, 1)
        self.decoder_fuse = Decoder_fuse(num_cls=num_cls)
        for i, c in enumerate((1, 2, 4, 8, 8, 8), start=1):
            setattr(self, f"fusion{i}", EarlyFusionBlock(basic_dims * c))
        for mod in self.modules():
            if isinstance(mod, nn.Conv3d):
                nn.init.kaiming_normal_(mod.weight)
        self._fusion_names = _fusion.param_names()

    def fusion_parameters(self):
        named = dict(self.named_parameters())
        return [named[n] for n in self._fusion_names]

    def forward(self, x):
        fmt = torch.channels_last_3d if _ENC_CL3D else torch.contiguous_format
        if _ENC_STREAMS and x.is_cuda:
            # The three modality encoders are independent: on three streams their small late-stage kernels (8 x 8 and
            # 16 x 16 maps: a dozen CTAs each) overlap instead of leaving most of the 148 SMs idle.  Autograd runs each
            # encoder's backward on its forward stream, so the backward overlaps the same way; under TrainStep's CUDA
            # graphs the fork / join becomes parallel branches of the graph.
            cur = torch.cuda.current_stream()
            if getattr(self, "_enc_streams", None) is None or self._enc_streams[0].device != x.device:
                self._enc_streams = [torch.cuda.Stream(device=x.device) for _ in _MODS[1:]]
            feats = [None] * len(_MODS)
            for i, m in enumerate(_MODS):
                st = cur if i == 0 else self._enc_streams[i - 1]
                if i:
                    st.wait_stream(cur)
                with torch.cuda.stream(st):
                    feats[i] = getattr(self, f"{m}_encoder")(x[:, i:i + 1].contiguous(memory_format=fmt))
            for st in self._enc_streams:
                cur.wait_stream(st)
            for fs in feats[1:]:
                for t in fs:
                    t.record_stream(cur)
        else:
            feats = [getattr(self, f"{m}_encoder")(x[:, i:i + 1].contiguous(memory_format=fmt)) for i, m in enumerate(_MODS)]
        if _ENC_STREAMS and x.is_cuda:
            # fusion1-4 feed only the decoder's skip branches: they run on the decoder's side stream (which goes on with
            # RFM_l -> nearest resize there), beside fusion6 -> transformer fusion block on this one
            cur = torch.cuda.current_stream()
            dec = self.decoder_fuse
            if getattr(dec, "_side", None) is None or dec._side.device != x.device:
                dec._side = torch.cuda.Stream(device=x.device)
            dec._side.wait_stream(cur)
            with torch.cuda.stream(dec._side):
                fused_x1, fused_x2, fused_x3, fused_x4 = (getattr(self, f"fusion{lv + 1}")(*(f[lv] for f in feats))
                                                          for lv in (0, 1, 2, 3))
            fused_x6 = self.fusion6(*(f[5] for f in feats))
        else:
            fused = [getattr(self, f"fusion{lv + 1}")(*(f[lv] for f in feats)) for lv in (0, 1, 2, 3, 5)]
            fused_x1, fused_x2, fused_x3, fused_x4, fused_x6 = fused          # fusion5's output is unused
        p = self.dropout_rate if self.training else 0.0
        self._step += 1
        # device_seed (set by TrainStep when it captures the model in CUDA graphs): the dropout seed is a device-resident
        # counter that advances inside the graph, encoded as -1 - base for the operator
        seed = (-1 - (self.base_seed & 0x3FFFFFFFFFFF)) if self.device_seed else self.base_seed + self._step
        x6_inter = torch.ops.corrif.fusion_block(feats[0][5], feats[1][5], feats[2][5], _V.to_channels_first(fused_x6),
                                                 self.fusion_parameters(), p, seed, self.precision)
        return self.decoder_fuse(fused_x1, fused_x2, fused_x3, fused_x4, _V.to_channels_last(x6_inter))
