"""Channels-last volume operators on the B200:
  * against fixtures produced by the reference's own classes (tests/golden/next_rows.npz: general_conv3d_prenorm
    3x3x3 replicate-padded and 1x1x1, EarlyFusionBlock; fp64 reference run), forward and backward;
  * against plain PyTorch (fp64 on the same GPU) over the decoder's shapes incl. ragged volumes, several sources,
    zero padding and bias-only convolutions;
  * resizes against F.interpolate.
Tolerances (relative L2): the convolutions run TF32 tensor-core MMAs with fp32 accumulation (the precision class
cuDNN uses for the same layers by default) -> output 2e-3; resizes are fp32 -> 1e-5.

Gradients through ReLU.  A TF32 pre-activation differs from the fp64 one by ~1e-3 of its scale, so the ~1e-3 of the
elements that lie that close to zero get the OTHER ReLU mask, and each of them changes the gradient by its full
magnitude: the relative L2 distance to an fp64 gradient is ~sqrt(1e-3) = 2-4e-2 for ANY TF32 convolution (measured
here: dx 1.9e-2, dW 2.0e-2, db 4.2e-2 on the reference fixture with a forward that is 3.5e-4 from it).  The kernels'
backward arithmetic is therefore checked exactly - against the fp64 backward evaluated with the mask the kernels'
own forward produced (tolerance 4e-3) - and the distance to the unconditioned fp64 gradient is bounded by 8e-2."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2, GOLDEN

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from corrif_b200 import volume as V

TOL_OUT, TOL_GRAD, TOL_GRAD_FLIPS = 2e-3, 4e-3, 8e-2


def _ref_block(xs, w, b, k, pad_mode, relu=True, norm=True, mask=None):
    """fp64 PyTorch restatement on NCDHW tensors.  ``mask``: use this ReLU mask instead of the fp64 run's own."""
    x = torch.cat(xs, dim=1)
    if k == 3:
        x = F.pad(x, (1,) * 6, mode="replicate" if pad_mode == 1 else "constant")
    y = F.conv3d(x, w, b)
    if relu:
        y = torch.relu(y) if mask is None else y * mask
    if norm:
        y = F.instance_norm(y, eps=1e-5)
    return y


def _run_block(xs_ncdhw, w, b, k, pad_mode, gout, relu=True, norm=True):
    dev = torch.device("cuda:0")
    xs = [V.to_channels_last(x.float().to(dev)).contiguous().requires_grad_(True) for x in xs_ncdhw]
    wt = w.float().to(dev).requires_grad_(True)
    bt = b.float().to(dev).requires_grad_(True) if b is not None else None
    y = V.conv_block(xs, wt, bt, k, pad_mode, relu=relu, norm=norm)
    y.backward(V.to_channels_last(gout.float().to(dev)))
    torch.cuda.synchronize()
    cf = lambda t: t.permute(0, 4, 1, 2, 3).double().cpu().numpy()   # noqa: E731
    return cf(y.detach()), [cf(x.grad) for x in xs], wt.grad.double().cpu().numpy(), (bt.grad.double().cpu().numpy() if bt is not None else None)


def _kernel_relu_mask(xs_ncdhw, w, b, k, pad_mode):
    """The ReLU mask of the kernels' own forward: sign of the pre-activation the same convolution kernel produces."""
    dev = torch.device("cuda:0")
    xs = [V.to_channels_last(x.float().to(dev)).contiguous() for x in xs_ncdhw]
    pre = V.conv_block(xs, w.float().to(dev), b.float().to(dev), k, pad_mode, relu=False, norm=False)
    return (pre > 0).permute(0, 4, 1, 2, 3).double()


def _check_backward_with_kernel_mask(xs, w, b, k, pad_mode, gout, dxs, dw, db, tol=TOL_GRAD):
    dev = torch.device("cuda:0")
    mask = _kernel_relu_mask(xs, w, b, k, pad_mode)
    rx = [x.double().to(dev).requires_grad_(True) for x in xs]
    rw, rb = w.double().to(dev).requires_grad_(True), b.double().to(dev).requires_grad_(True)
    _ref_block(rx, rw, rb, k, pad_mode, mask=mask).backward(gout.double().to(dev))
    errs = dict(dw=rel_l2(dw, rw.grad.cpu().numpy()), db=rel_l2(db, rb.grad.cpu().numpy()))
    for i, x in enumerate(rx):
        errs[f"dx{i}"] = rel_l2(dxs[i], x.grad.cpu().numpy())
    print("   backward with the kernels' mask: " + "  ".join(f"{n} {e:.2e}" for n, e in errs.items()))
    for n, e in errs.items():
        assert e < tol, errs


@pytest.mark.parametrize("tag,k", [("c3", 3), ("c1", 1), ("c3b", 3)])
def test_conv_block_matches_reference_fixture(tag, k):
    g = np.load(os.path.join(GOLDEN, "next_rows.npz"))
    t = lambda n: torch.from_numpy(g[f"{tag}/{n}"])   # noqa: E731
    y, dxs, dw, db = _run_block([t("x")], t("w"), t("b"), k, 1, t("gout"))
    report = dict(y=rel_l2(y, g[f"{tag}/y"]), dx=rel_l2(dxs[0], g[f"{tag}/dx"]), dw=rel_l2(dw, g[f"{tag}/dw"]),
                  db=rel_l2(db, g[f"{tag}/db"]))
    print(f"\n[conv_block {tag}] " + "  ".join(f"{n} {e:.2e}" for n, e in report.items()))
    assert report["y"] < TOL_OUT, report
    assert max(report["dx"], report["dw"], report["db"]) < TOL_GRAD_FLIPS, report      # ReLU mask flips: see the header
    _check_backward_with_kernel_mask([t("x")], t("w"), t("b"), k, 1, t("gout"), dxs, dw, db)


def test_early_fusion_block_matches_reference_fixture():
    g = np.load(os.path.join(GOLDEN, "next_rows.npz"))
    t = lambda n: torch.from_numpy(g[f"ef/{n}"])   # noqa: E731
    y, dxs, dw, db = _run_block([t("x0"), t("x1"), t("x2")], t("w"), t("b"), 1, 0, t("gout"))
    assert rel_l2(y, g["ef/y"]) < TOL_OUT
    for i in range(3):
        assert rel_l2(dxs[i], g[f"ef/dx{i}"]) < TOL_GRAD_FLIPS, i
    assert rel_l2(dw, g["ef/dw"]) < TOL_GRAD_FLIPS and rel_l2(db, g["ef/db"]) < TOL_GRAD_FLIPS
    _check_backward_with_kernel_mask([t("x0"), t("x1"), t("x2")], t("w"), t("b"), 1, 0, t("gout"), dxs, dw, db)


CASES = [
    # (source channels, Cout, k, pad_mode, B, D, H, W)
    ((32,), 8, 3, 1, 2, 8, 16, 16),            # d1_c2-like, tile-aligned
    ((24, 8), 8, 3, 1, 1, 12, 9, 13),          # two sources (skip ++ up), ragged volume
    ((16,), 8, 3, 1, 2, 5, 8, 8),              # d1_c1
    ((48, 16), 16, 3, 1, 1, 8, 8, 10),         # d2_c2
    ((96, 32), 32, 3, 1, 1, 4, 8, 8),          # d3_c2
    ((192, 128), 64, 3, 1, 1, 4, 8, 8),        # d4_c2: 320 input channels, 8 output blocks
    ((128,), 128, 3, 1, 1, 4, 8, 8),           # d4_c1: two output tiles
    ((24,), 24, 3, 0, 2, 3, 14, 14),           # RFM 3x3x3, ZERO padding, depth 3
    ((192,), 192, 3, 0, 1, 3, 7, 7),           # RFM4 at 224^2 tiles
    ((8,), 8, 1, 1, 2, 6, 8, 9),               # d1_out (1x1x1): per-voxel fp32 kernel
    ((16,), 16, 1, 1, 2, 4, 8, 8),             # d2_out
    ((64,), 64, 1, 1, 1, 4, 8, 8),             # d4_out
    ((8, 8, 8), 24, 1, 0, 2, 3, 10, 10),       # EarlyFusionBlock(8): three sources
    ((64, 64, 64), 192, 1, 0, 1, 8, 8, 8),     # fusion6
    ((1, ), 1, 1, 0, 1, 1, 1, 1),              # placeholder, replaced below
]
CASES = CASES[:-1]


@pytest.mark.parametrize("case", CASES, ids=[f"{'+'.join(map(str, c[0]))}to{c[1]}k{c[2]}p{c[3]}" for c in CASES])
def test_conv_block_against_pytorch_fp64(case):
    chans, cout, k, pad_mode, B, D, H, W = case
    g = torch.Generator().manual_seed(sum(chans) * 131 + cout)
    xs = [torch.randn(B, c, D, H, W, generator=g, dtype=torch.float64) for c in chans]
    w = torch.randn(cout, sum(chans), k, k, k, generator=g, dtype=torch.float64) * (2.0 / (sum(chans) * k ** 3)) ** 0.5
    b = torch.randn(cout, generator=g, dtype=torch.float64) * 0.1
    gout = torch.randn(B, cout, D, H, W, generator=g, dtype=torch.float64)
    dev = torch.device("cuda:0")
    rx = [x.to(dev).requires_grad_(True) for x in xs]
    rw, rb = w.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
    ry = _ref_block(rx, rw, rb, k, pad_mode)
    ry.backward(gout.to(dev))
    y, dxs, dw, db = _run_block(xs, w, b, k, pad_mode, gout)
    errs = dict(y=rel_l2(y, ry.detach().cpu().numpy()), dw=rel_l2(dw, rw.grad.cpu().numpy()), db=rel_l2(db, rb.grad.cpu().numpy()))
    for i, x in enumerate(rx):
        errs[f"dx{i}"] = rel_l2(dxs[i], x.grad.cpu().numpy())
    print("\n[conv_block %s] " % (case,) + "  ".join(f"{n} {e:.2e}" for n, e in errs.items()))
    assert errs["y"] < TOL_OUT, errs
    for n, e in errs.items():
        assert e < TOL_GRAD_FLIPS, errs
    _check_backward_with_kernel_mask(xs, w, b, k, pad_mode, gout, dxs, dw, db)


def test_bias_only_convolution_and_its_gradients():
    """RFM5_reduce / adapt convs: conv + bias, no ReLU, no norm (mmvit4.py:231, 157-164)."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 192, 4, 4, 4, generator=g, dtype=torch.float64)
    w = torch.randn(128, 192, 1, 1, 1, generator=g, dtype=torch.float64) * 0.1
    b = torch.randn(128, generator=g, dtype=torch.float64)
    gout = torch.randn(2, 128, 4, 4, 4, generator=g, dtype=torch.float64)
    dev = torch.device("cuda:0")
    rx, rw, rb = (t.to(dev).requires_grad_(True) for t in (x, w, b))
    ry = F.conv3d(rx, rw, rb)
    ry.backward(gout.to(dev))
    y, dxs, dw, db = _run_block([x], w, b, 1, 0, gout, relu=False, norm=False)
    assert rel_l2(y, ry.detach().cpu().numpy()) < TOL_OUT
    assert rel_l2(dxs[0], rx.grad.cpu().numpy()) < TOL_GRAD and rel_l2(dw, rw.grad.cpu().numpy()) < TOL_GRAD
    assert rel_l2(db, rb.grad.cpu().numpy()) < 1e-5


def test_channel_slice_views_are_valid_sources_and_gradients():
    """A source may be a channel slice of a wider buffer (ld > C), and the gradients of several sources are channel
    slices of ONE data-gradient buffer - the torch.cat of mmvit4.py:272 never materialises."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(9)
    buf = torch.randn(1, 4, 8, 8, 40, generator=g).to(dev)
    a, b_ = buf[..., :24].detach().requires_grad_(True), buf[..., 24:40].detach().requires_grad_(True)
    assert V._ld(buf[..., :24]) == 40
    w = (torch.randn(8, 40, 3, 3, 3, generator=g) * 0.05).to(dev).requires_grad_(True)
    bias = torch.zeros(8, device=dev, requires_grad=True)
    y1 = V.conv_block([buf[..., :24], buf[..., 24:40]], w, bias, 3, 1)
    y2 = V.conv_block([a.contiguous(), b_.contiguous()], w, bias, 3, 1)
    assert torch.allclose(y1, y2, rtol=1e-5, atol=1e-6)      # statistics leave through atomics: last-bit differences
    y2.sum().backward()
    assert a.grad is not None and b_.grad is not None and a.grad.shape == a.shape


@pytest.mark.parametrize("shape", [((2, 4, 5, 6, 8), (8, 10, 12)), ((1, 16, 16, 16, 16), (32, 32, 32)),
                                   ((2, 32, 32, 32, 16), (64, 64, 64)), ((2, 24, 40, 32, 24), (48, 80, 33)),   # separable passes
                                   ((2, 3, 14, 14, 24), (8, 8, 8)), ((1, 8, 8, 8, 8), (1, 24, 24)),
                                   ((2, 3, 7, 7, 64), (8, 8, 8))])
def test_trilinear_resize_matches_f_interpolate(shape):
    (B, D, H, W, C), size = shape
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(D * 100 + size[0])
    x = torch.randn(B, D, H, W, C, generator=g).to(dev).requires_grad_(True)
    xr = x.detach().permute(0, 4, 1, 2, 3).contiguous().double().requires_grad_(True)
    y = V.resize_trilinear(x, size)
    yr = F.interpolate(xr, size=size, mode="trilinear", align_corners=True)
    go = torch.randn(yr.shape, generator=g, dtype=torch.float64).to(dev)
    yr.backward(go)
    y.backward(go.permute(0, 2, 3, 4, 1).float())
    assert rel_l2(y.detach().permute(0, 4, 1, 2, 3).cpu().numpy(), yr.detach().cpu().numpy()) < 1e-5
    assert rel_l2(x.grad.permute(0, 4, 1, 2, 3).cpu().numpy(), xr.grad.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("shape", [((2, 3, 16, 16, 24), (32, 32, 32)), ((1, 3, 14, 14, 192), (16, 16, 16)),
                                   ((1, 3, 64, 64, 24), (128, 128, 128)), ((2, 3, 7, 9, 8), (5, 20, 11))])
def test_nearest_resize_matches_f_interpolate(shape):
    (B, D, H, W, C), size = shape
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(H)
    x = torch.randn(B, D, H, W, C, generator=g).to(dev).requires_grad_(True)
    xr = x.detach().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    y = V.resize_nearest(x, size)
    yr = F.interpolate(xr, size)
    assert torch.equal(y.detach().permute(0, 4, 1, 2, 3), yr.detach())
    go = torch.randn(yr.shape, generator=g).to(dev)
    yr.backward(go)
    y.backward(go.permute(0, 2, 3, 4, 1).contiguous())
    assert rel_l2(x.grad.permute(0, 4, 1, 2, 3).cpu().numpy(), xr.grad.cpu().numpy()) < 1e-5


def test_decoder_and_early_fusion_against_stock_pytorch_modules():
    """The drop-in Decoder_fuse / EarlyFusionBlock (volume kernels) against the stock-PyTorch restatement of
    mmvit4.py:64-81, 222-292 (baseline/eager_mmvit4.py, fp32 with TF32 off) sharing one state_dict: output
    probabilities and the gradients of inputs and of a spread of parameters."""
    import sys
    from conftest import ROOT
    dropin_dir = os.path.join(ROOT, "corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200", "dropin")
    sys.path.insert(0, dropin_dir)
    sys.modules.pop("mmvit4", None)
    import mmvit4
    from baseline.eager_mmvit4 import _Decoder, _EarlyFusion
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        dev = torch.device("cuda:0")
        torch.manual_seed(3)
        ref, mine = _Decoder().to(dev), mmvit4.Decoder_fuse().to(dev)
        for m in ref.modules():
            if isinstance(m, torch.nn.Conv3d):
                torch.nn.init.kaiming_normal_(m.weight)
        mine.load_state_dict(ref.state_dict(), strict=True)
        B = 1
        shapes = [(B, 24, 3, 16, 16), (B, 48, 3, 16, 16), (B, 96, 3, 8, 8), (B, 192, 3, 4, 4), (B, 192, 8, 8, 8)]
        xs = [torch.randn(s, device=dev) for s in shapes]
        rx = [x.clone().requires_grad_(True) for x in xs]
        mx = [x.clone().requires_grad_(True) for x in xs]
        yr = ref(*rx)
        ym = mine(*[V.to_channels_last(x) for x in mx])
        go = torch.randn_like(yr)
        yr.backward(go)
        ym.backward(go)
        torch.cuda.synchronize()
        e_y = rel_l2(ym.detach().cpu().numpy(), yr.detach().cpu().numpy())
        e_x = [rel_l2(a.grad.cpu().numpy(), b.grad.cpu().numpy()) for a, b in zip(mx, rx)]
        named_r, named_m = dict(ref.named_parameters()), dict(mine.named_parameters())
        keys = ["RFM5.fusion_layer.1.conv.weight", "RFM5_reduce.weight", "d4_c2.conv.weight", "d3_c1.conv.bias",
                "d2_c2.conv.weight", "d1_c1.conv.weight", "d1_c2.conv.weight", "d1_out.conv.weight", "RFM1.fusion_layer.1.conv.weight",
                "final_conv.weight"]
        e_p = {k: rel_l2(named_m[k].grad.cpu().numpy(), named_r[k].grad.cpu().numpy()) for k in keys}
        print(f"\n[decoder] y {e_y:.2e}  dx {['%.1e' % e for e in e_x]}  " + "  ".join(f"{k}:{e:.1e}" for k, e in e_p.items()))
        assert e_y < 2e-3
        # gradients: every one of the ~30 ReLUs flips the mask of the ~1e-3 of its elements that sit within TF32
        # rounding of zero (see the header: 2-4e-2 per block against an fp64 run), and the decoder stacks 15 blocks;
        # the per-block backward arithmetic is checked exactly above, here the composition is held to 0.15
        assert max(e_x) < 0.15 and max(e_p.values()) < 0.15
        # EarlyFusionBlock(8) on three encoder-shaped maps
        ef_r, ef_m = _EarlyFusion(8).to(dev), mmvit4.EarlyFusionBlock(8).to(dev)
        ef_m.load_state_dict(ef_r.state_dict(), strict=True)
        a = [torch.randn(2, 8, 3, 12, 12, device=dev) for _ in range(3)]
        assert rel_l2(V.to_channels_first(ef_m(*a)).detach().cpu().numpy(), ef_r(*a).detach().cpu().numpy()) < 2e-3
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
        sys.path.remove(dropin_dir)
        sys.modules.pop("mmvit4", None)


@pytest.mark.parametrize("case", [(2, 2048, 64, 3, 8, 8, True), (2, 256, 16, 3, 16, 16, True), (1, 64, 8, 3, 14, 14, True),
                                  (2, 184, 64, 8, 8, 8, False), (2, 192, 128, 8, 8, 8, False), (1, 1024, 64, 3, 7, 7, True)])
def test_pointwise_conv_gemm_path(case):
    """Bias-only 1x1x1 convolutions of the encoder tail / RFM5_reduce on the tcgen05 GEMM (volume.pointwise_conv),
    from cuDNN-layout ([B,C,D,H,W]) and channels-last inputs, forward and all three gradients, against fp64 PyTorch."""
    B, cin, cout, D, H, W, cf = case
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(B, cin, D, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, 1, 1, 1, generator=g, dtype=torch.float64) * cin ** -0.5
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    go = torch.randn(B, cout, D, H, W, generator=g, dtype=torch.float64)
    dev = torch.device("cuda:0")
    rx, rw, rb = (t.to(dev).requires_grad_(True) for t in (x, w, b))
    F.conv3d(rx, rw, rb).backward(go.to(dev))
    xin = x.float().to(dev) if cf else x.float().to(dev).permute(0, 2, 3, 4, 1).contiguous()
    xin.requires_grad_(True)
    wt, bt = w.float().to(dev).requires_grad_(True), b.float().to(dev).requires_grad_(True)
    y = V.pointwise_conv(xin, wt, bt, channels_first=cf)
    y.backward(go.float().to(dev).permute(0, 2, 3, 4, 1).contiguous())
    torch.cuda.synchronize()
    ref_y = F.conv3d(rx, rw, rb).detach()
    assert rel_l2(y.detach().permute(0, 4, 1, 2, 3).cpu().numpy(), ref_y.cpu().numpy()) < TOL_OUT
    dx = xin.grad if cf else xin.grad.permute(0, 4, 1, 2, 3)
    assert rel_l2(dx.cpu().numpy(), rx.grad.cpu().numpy()) < TOL_GRAD
    assert rel_l2(wt.grad.cpu().numpy(), rw.grad.cpu().numpy()) < TOL_GRAD
    assert rel_l2(bt.grad.cpu().numpy(), rb.grad.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("case", [(2, 3, 16, 16, 64, True, False), (2, 3, 8, 8, 256, True, True), (1, 3, 4, 4, 2048, True, True),
                                  (2, 3, 9, 7, 128, False, False), (3, 1, 5, 5, 1024, False, True)])
def test_batchnorm_relu_against_pytorch(case):
    """Train-mode BatchNorm3d (+ residual) (+ ReLU) of the encoders' Bottleneck3D (mmvit4.py:196-212) on the fused
    channels-last kernels against fp64 PyTorch: output, dx, d residual, d gamma, d beta and the running statistics."""
    B, D, H, W, C, relu, with_res = case
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C + D)
    x = (torch.randn(B, C, D, H, W, generator=g, dtype=torch.float64) * 2 + 0.5).to(dev)
    res = torch.randn(B, C, D, H, W, generator=g, dtype=torch.float64).to(dev) if with_res else None
    go = torch.randn(B, C, D, H, W, generator=g, dtype=torch.float64).to(dev)
    ref_bn = torch.nn.BatchNorm3d(C).double().to(dev).train()
    with torch.no_grad():
        ref_bn.weight.copy_(torch.rand(C, generator=g, dtype=torch.float64) + 0.5)
        ref_bn.bias.copy_(torch.randn(C, generator=g, dtype=torch.float64) * 0.2)
    rx = x.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if with_res else None
    ry = ref_bn(rx) if rr is None else ref_bn(rx) + rr
    ry = torch.relu(ry) if relu else ry
    ry.backward(go)
    bn = torch.nn.BatchNorm3d(C).to(dev).train()
    with torch.no_grad():
        bn.weight.copy_(ref_bn.weight.float()); bn.bias.copy_(ref_bn.bias.float())
    mx = x.float().permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True)
    mr = res.float().permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True) if with_res else None
    y = V.batchnorm_relu(mx, bn, mr, relu)
    y.backward(go.float().permute(0, 2, 3, 4, 1).contiguous())
    torch.cuda.synchronize()
    cf = lambda t: t.permute(0, 4, 1, 2, 3).double().cpu().numpy()   # noqa: E731
    assert rel_l2(cf(y.detach()), ry.detach().cpu().numpy()) < 1e-5
    # the ReLU mask can differ where the fp32 output rounds across zero: compare gradients on a 1e-4 budget
    assert rel_l2(cf(mx.grad), rx.grad.cpu().numpy()) < 2e-3
    if with_res:
        assert rel_l2(cf(mr.grad), rr.grad.cpu().numpy()) < 2e-3
    assert rel_l2(bn.weight.grad.double().cpu().numpy(), ref_bn.weight.grad.cpu().numpy()) < 2e-3
    assert rel_l2(bn.bias.grad.double().cpu().numpy(), ref_bn.bias.grad.cpu().numpy()) < 2e-3
    assert rel_l2(bn.running_mean.double().cpu().numpy(), ref_bn.running_mean.cpu().numpy()) < 1e-5
    assert rel_l2(bn.running_var.double().cpu().numpy(), ref_bn.running_var.cpu().numpy()) < 1e-5
    assert int(bn.num_batches_tracked) == 1


def test_two_producers_writing_into_one_buffer_are_one_source():
    """conv_block(out=slice) / resize_nearest(out=slice): the reference's torch.cat (mmvit4.py:272) as two producers that
    write side by side into one buffer, which the consumer then reads as a single source - same numbers as separate
    tensors, forward and backward."""
    dev = torch.device("cuda:0")
    gen = lambda s: torch.Generator().manual_seed(s)   # noqa: E731
    res = {}
    for shared in (False, True):
        up = torch.randn(2, 6, 8, 64, 16, generator=gen(1)).to(dev).requires_grad_(True)
        skip = torch.randn(2, 3, 4, 32, 24, generator=gen(2)).to(dev).requires_grad_(True)
        w1 = (torch.randn(8, 16, 3, 3, 3, generator=gen(3)) * 0.1).to(dev).requires_grad_(True)
        w2 = (torch.randn(8, 32, 3, 3, 3, generator=gen(4)) * 0.1).to(dev).requires_grad_(True)
        b1, b2 = torch.zeros(8, device=dev, requires_grad=True), torch.zeros(8, device=dev, requires_grad=True)
        buf = torch.empty(2, 6, 8, 64, 32, device=dev)
        y = V.conv_block([up], w1, b1, 3, V.PAD_REPLICATE, out=buf[..., 24:] if shared else None)
        s = V.resize_nearest(skip, (6, 8, 64), out=buf[..., :24] if shared else None)
        if shared:
            assert V._desc([s, y], 8, 3, 1).nsrc == 1          # the consumer sees one 32-channel source
        z = V.conv_block([s, y], w2, b2, 3, V.PAD_REPLICATE)
        z.backward(torch.randn(z.shape, generator=gen(5)).to(dev))
        torch.cuda.synchronize()
        res[shared] = [t.detach().cpu().numpy() for t in (z, up.grad, skip.grad, w1.grad, w2.grad)]
    for a, b_, name in zip(res[True], res[False], ("z", "dup", "dskip", "dw1", "dw2")):
        e = rel_l2(a, b_)
        # same kernels on the same values; only the summation order of the statistics' atomics differs
        assert e < (1e-4 if name == "z" else 5e-2), (name, e)
