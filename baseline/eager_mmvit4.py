"""Stock-PyTorch eager CorrIFNet on the GPU: "the existing Blackwell path to beat" (SURVEY.md section 2.3,
BASELINE.md): every op an ATen call into cuBLAS / cuDNN, no corrif_b200 kernel anywhere.

``EagerFusionBlock`` restates mmvit4.py:295-388, 398-426, 456-529 with nn.Linear / nn.Conv3d / nn.LayerNorm /
nn.Dropout modules under the reference's attribute names, so a drop-in (or reference) ``state_dict`` loads into it
strictly for the hot-path subset; ``EagerMMVit4`` is the whole model - encoders, early-fusion blocks, fusion block, decoder -
on stock PyTorch under the same state_dict keys.  Baseline only: bench.py times
it beside the kernels, tests use it as a same-device comparison; the product never imports it.
"""
import math
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200", "dropin")
for _p in (ROOT, DROPIN):
    if _p not in sys.path:
        sys.path.insert(0, _p)

MODS = ("RGB", "NIR", "SWIR")
DIM, HEADS, MLP, PATCH, ENC = 512, 8, 512, 8, 64


class _Attention(nn.Module):
    def __init__(self, p):
        super().__init__()
        self.qkv, self.proj = nn.Linear(DIM, 3 * DIM, bias=False), nn.Linear(DIM, DIM)
        self.attn_drop, self.proj_drop = nn.Dropout(p), nn.Dropout(p)

    def forward(self, x):
        B, N, C = x.shape
        q, k, v = self.qkv(x).reshape(B, N, 3, HEADS, C // HEADS).permute(2, 0, 3, 1, 4)
        a = self.attn_drop(((q @ k.transpose(-2, -1)) * (C // HEADS) ** -0.5).softmax(dim=-1))
        return self.proj_drop(self.proj((a @ v).transpose(1, 2).reshape(B, N, C)))


class _Wrap(nn.Module):          # Residual(PreNorm[Drop](fn)) with the reference's nesting: .fn.norm / .fn.fn / .fn.dropout
    class _Pre(nn.Module):
        def __init__(self, fn, p):
            super().__init__()
            self.norm, self.fn = nn.LayerNorm(DIM), fn
            self.dropout = nn.Dropout(p) if p is not None else None

        def forward(self, x):
            y = self.fn(self.norm(x))
            return y if self.dropout is None else self.dropout(y)

    def __init__(self, fn, p):
        super().__init__()
        self.fn = _Wrap._Pre(fn, p)

    def forward(self, x):
        return self.fn(x) + x


class _FeedForward(nn.Module):
    def __init__(self, p):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(DIM, MLP), nn.GELU(), nn.Dropout(p), nn.Linear(MLP, DIM), nn.Dropout(p))

    def forward(self, x):
        return self.net(x)


class _Transformer(nn.Module):
    def __init__(self, p):
        super().__init__()
        self.cross_attention_list = nn.ModuleList([_Wrap(_Attention(p), p)])
        self.cross_ffn_list = nn.ModuleList([_Wrap(_FeedForward(p), None)])

    def forward(self, x, pos):
        return self.cross_ffn_list[0](self.cross_attention_list[0](x + pos))


class EagerFusionBlock(nn.Module):
    """forward(x6 list of [B,64,8,8,8], fused_x6 [B,192,8,8,8]) -> x6_inter [B,192,8,8,8]."""

    def __init__(self, dropout_rate=0.1):
        super().__init__()
        for m in MODS:
            setattr(self, f"{m}_encode_conv", nn.Conv3d(ENC, DIM, 1))
            setattr(self, f"{m}_pos", nn.Parameter(torch.zeros(1, PATCH ** 3, DIM)))
            setattr(self, f"{m}_transformer", _Transformer(dropout_rate))
            setattr(self, f"qkv_{m}", nn.Conv3d(DIM, 3 * DIM, 1))
        self.fused6_encode_conv = nn.Conv3d(3 * ENC, DIM, 1)
        self.fused6_pos = nn.Parameter(torch.zeros(1, PATCH ** 3, DIM))
        self.multimodal_transformer = _Transformer(dropout_rate)
        self.multimodal_decode_conv = nn.Conv3d(4 * DIM, 3 * ENC, 1)

    def forward(self, x6, fused_x6):
        B = fused_x6.shape[0]
        tok = lambda conv, x: conv(x).permute(0, 2, 3, 4, 1).contiguous().view(B, -1, DIM)      # noqa: E731
        vol = lambda t: t.view(B, PATCH, PATCH, PATCH, -1).permute(0, 4, 1, 2, 3).contiguous()  # noqa: E731
        skip = [tok(getattr(self, f"{m}_encode_conv"), x) for m, x in zip(MODS, x6)]
        trans = [getattr(self, f"{m}_transformer")(s, getattr(self, f"{m}_pos")) for m, s in zip(MODS, skip)]
        qkv = [getattr(self, f"qkv_{m}")(vol(t)).chunk(3, dim=1) for m, t in zip(MODS, trans)]
        ks, vs = [t[1] for t in qkv], [t[2] for t in qkv]

        def inter(q):                                    # mmvit4.py:481-487 incl. the batch-mixing view
            s = torch.cat([(q * k).reshape(1, -1) for k in ks], dim=0)
            a = torch.softmax(s / math.sqrt(3), dim=0).view(B, 3 * DIM, PATCH, PATCH, PATCH)
            return sum(a[:, i * DIM:(i + 1) * DIM] * v for i, v in enumerate(vs))
        fused = [s + inter(t[0]).permute(0, 2, 3, 4, 1).reshape(B, -1, DIM) for s, t in zip(skip, qkv)]
        tokens = torch.cat(fused + [tok(self.fused6_encode_conv, fused_x6)], dim=1)
        pos = torch.cat([getattr(self, f"{m}_pos") for m in MODS] + [self.fused6_pos], dim=1)
        mm = self.multimodal_transformer(tokens, pos)
        return self.multimodal_decode_conv(
            mm.view(B, PATCH, PATCH, PATCH, 4 * DIM).permute(0, 4, 1, 2, 3).contiguous())


def _conv(cin, cout, k, stride, pad, depth_k=1):    # torchvision conv2d inflated to (depth_k, k, k) (mmvit4.py:83-111)
    return nn.Conv3d(cin, cout, (depth_k, k, k), stride=(1, stride, stride), padding=(depth_k // 2, pad, pad), bias=False)


class _Bottleneck(nn.Module):                      # Bottleneck3D (mmvit4.py:196-212), ResNet-50 v1.5 strides
    def __init__(self, cin, planes, stride, project):
        super().__init__()
        self.conv1, self.bn1 = _conv(cin, planes, 1, 1, 0), nn.BatchNorm3d(planes)
        self.conv2, self.bn2 = _conv(planes, planes, 3, stride, 1), nn.BatchNorm3d(planes)
        self.conv3, self.bn3 = _conv(planes, 4 * planes, 1, 1, 0), nn.BatchNorm3d(4 * planes)
        self.downsample = nn.Sequential(_conv(cin, 4 * planes, 1, stride, 0), nn.BatchNorm3d(4 * planes)) if project else None

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class _Encoder(nn.Module):                         # Encoder (mmvit4.py:113-194), all stock PyTorch
    def __init__(self):
        super().__init__()
        self.e1_c1, self.e1_bn = _conv(1, 64, 7, 2, 3, depth_k=3), nn.BatchNorm3d(64)
        self.e1_mp = nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        for name, cin, planes, blocks, stride in (("e2", 64, 64, 3, 1), ("e3", 256, 128, 4, 2), ("e4", 512, 256, 6, 2),
                                                  ("e5", 1024, 512, 3, 2)):
            setattr(self, name, nn.Sequential(_Bottleneck(cin, planes, stride, True),
                                              *[_Bottleneck(4 * planes, planes, 1, False) for _ in range(blocks - 1)]))
        self.conv6 = nn.Conv3d(184, 64, 1)
        for i, (cin, cout) in enumerate(zip((64, 256, 512, 1024, 2048), (8, 16, 32, 64, 64)), start=1):
            setattr(self, f"adapt{i}", nn.Conv3d(cin, cout, 1))

    def forward(self, x):
        f1 = self.e1_mp(self.e1_bn(F.relu(self.e1_c1(x))))
        f2 = self.e2(f1); f3 = self.e3(f2); f4 = self.e4(f3); f5 = self.e5(f4)       # noqa: E702
        lv = [getattr(self, f"adapt{i}")(f) for i, f in enumerate((f1, f2, f3, f4, f5), start=1)]
        pooled = [F.interpolate(t, size=(8, 8, 8), mode="trilinear", align_corners=True) for t in lv]
        return (*lv, self.conv6(torch.cat(pooled, dim=1)))


class _ConvReluNorm(nn.Module):                    # general_conv3d_prenorm (mmvit4.py:29-45): conv -> ReLU -> IN
    def __init__(self, cin, cout, k=3, pad_type="zeros"):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, k, padding=k // 2, padding_mode=pad_type)
        self.norm = nn.InstanceNorm3d(cout)

    def forward(self, x):
        return self.norm(F.relu(self.conv(x)))


class _RFM(nn.Module):                             # fusion_prenorm (mmvit4.py:47-56)
    def __init__(self, c):
        super().__init__()
        self.fusion_layer = nn.Sequential(_ConvReluNorm(c, c, 1), _ConvReluNorm(c, c, 3), _ConvReluNorm(c, c, 1))

    def forward(self, x):
        return self.fusion_layer(x)


class _EarlyFusion(nn.Module):                     # EarlyFusionBlock (mmvit4.py:64-81)
    def __init__(self, c):
        super().__init__()
        self.conv, self.norm = nn.Conv3d(3 * c, 3 * c, 1), nn.InstanceNorm3d(3 * c)

    def forward(self, a, b, c):
        return self.norm(F.relu(self.conv(torch.cat([a, b, c], dim=1))))


class _Decoder(nn.Module):                         # Decoder_fuse (mmvit4.py:222-292)
    LEVELS = ((4, 192, 128, 64, 16), (3, 96, 64, 32, 32), (2, 48, 32, 16, 64), (1, 24, 16, 8, 128))

    def __init__(self, num_cls=1):
        super().__init__()
        for lvl, skip, cin, cout, _ in self.LEVELS:
            c1 = cin if lvl == 4 else cout
            setattr(self, f"d{lvl}_c1", _ConvReluNorm(cin, c1, 3, "replicate"))
            setattr(self, f"d{lvl}_c2", _ConvReluNorm(skip + c1, cout, 3, "replicate"))
            setattr(self, f"d{lvl}_out", _ConvReluNorm(cout, cout, 1, "replicate"))
            setattr(self, f"RFM{lvl}", _RFM(skip))
        for name, cin in (("seg_d4", 64), ("seg_d3", 64), ("seg_d2", 32), ("seg_d1", 16), ("seg_layer", 8)):
            setattr(self, name, nn.Conv3d(cin, num_cls, 1))
        self.RFM5, self.RFM5_reduce, self.final_conv = _RFM(192), nn.Conv3d(192, 128, 1), nn.Conv3d(8, 3, 1)

    def forward(self, x1, x2, x3, x4, x5):
        y = self.RFM5_reduce(self.RFM5(x5))
        for (lvl, _, _, _, cube), skip in zip(self.LEVELS, (x4, x3, x2, x1)):
            y = getattr(self, f"d{lvl}_c1")(F.interpolate(y, scale_factor=2, mode="trilinear", align_corners=True))
            s = F.interpolate(getattr(self, f"RFM{lvl}")(skip), (cube, cube, cube))
            y = getattr(self, f"d{lvl}_out")(getattr(self, f"d{lvl}_c2")(torch.cat((s, y), dim=1)))
        y = F.interpolate(y, size=(1, 224, 224), mode="trilinear", align_corners=True)
        return torch.sigmoid(self.final_conv(y))


class EagerMMVit4(EagerFusionBlock):
    """Whole CorrIFNet on stock PyTorch with the reference's 1140 state_dict keys."""

    def __init__(self, num_cls=1, dropout_rate=0.1):
        super().__init__(dropout_rate)
        for m in MODS:
            setattr(self, f"{m}_encoder", _Encoder())
            setattr(self, f"{m}_decode_conv", nn.Conv3d(DIM, ENC, 1))                  # unused in forward
        for i, c in enumerate((1, 2, 4, 8, 8, 8), start=1):
            setattr(self, f"fusion{i}", _EarlyFusion(8 * c))
        self.decoder_fuse = _Decoder(num_cls)
        for mod in self.modules():
            if isinstance(mod, nn.Conv3d):
                nn.init.kaiming_normal_(mod.weight)

    def forward(self, x):
        feats = [getattr(self, f"{m}_encoder")(x[:, i:i + 1]) for i, m in enumerate(MODS)]
        fused = [getattr(self, f"fusion{lv + 1}")(*(f[lv] for f in feats)) for lv in range(6)]   # fusion5 too (:453)
        x6_inter = EagerFusionBlock.forward(self, [f[5] for f in feats], fused[5])
        return self.decoder_fuse(fused[0], fused[1], fused[2], fused[3], x6_inter)
