"""The tcgen05 convolution kernels on the B200 - the line convolution (csrc/conv3d_tc.cu, corrif_conv3d_tc_fwd):
forward (replicate / zero padding, 1..3 sources, bias, ReLU, InstanceNorm statistics) and data gradient (incl. the
adjoint of replicate padding in one pass), and the weight gradient (csrc/conv3d_wgrad_tc.cu) - against fp64 PyTorch on
the same inputs, for every tile geometry the kernel has: line widths 128 / 64 /
32 / 16 (1 / 2 / 4 / 8 samples per MMA tile), 32- / 64- / 128-byte swizzled channel chunks, one to four lines per
strip, several output-channel chunks, ragged strips and depths down to one plane.
Tolerance (relative L2): TF32 operands (activations truncated by the tensor core with the mean compensated in the
packed weights, weights rounded to nearest), fp32 accumulation -> 1e-3, as for the warp-level kernels."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from corrif_b200 import volume as V

TOL = 1e-3

CASES = [
    # (source channels, Cout, pad_mode, B, D, H, W)
    ((32,), 8, 1, 1, 5, 8, 128),             # d1_c1-like: one 128-byte chunk, 4 lines per strip
    ((16,), 8, 1, 1, 3, 6, 128),             # 64-byte swizzle, ragged last strip
    ((8,), 32, 1, 1, 4, 5, 128),             # data-gradient shape of 32 -> 8: 32-byte swizzle, N = 96, 1 line per strip
    ((16, 16), 8, 1, 2, 4, 4, 64),           # two sources (skip ++ up), two samples per tile
    ((64,), 16, 1, 2, 6, 8, 64),             # d2_c2-like: two 128-byte chunks, N = 48
    ((16,), 64, 1, 2, 3, 4, 64),             # two output-channel chunks
    ((32,), 16, 0, 4, 3, 32, 32),            # ZERO padding, four samples per tile
    ((8, 8, 8), 8, 0, 8, 2, 16, 16),         # three sources, eight samples per tile, zero padding
    ((32,), 32, 1, 8, 1, 16, 16),            # a single plane (both z taps clamp onto it)
    ((32,), 8, 1, 1, 40, 4, 128),            # several z ranges per strip
    ((24, 8), 8, 1, 1, 4, 6, 128),           # d1_c2: cat(skip 24, up 8) as four 8-channel chunks
    ((48, 16), 16, 1, 2, 3, 8, 64),          # d2_c2: four 16-channel chunks
    ((32,), 128, 1, 8, 4, 32, 32),           # four output-channel chunks whose weights are staged one chunk at a time
    ((24,), 24, 0, 2, 3, 64, 64),            # RFM1 3x3x3: three chunks of 8 output channels, zero padding, depth 3
    ((64,), 320, 1, 8, 5, 16, 16),           # data-gradient shape of d4_c2: 20 chunks of 16 output channels
]
IDS = [f"{'+'.join(map(str, c[0]))}to{c[1]}p{c[2]}_{c[3]}x{c[4]}x{c[5]}x{c[6]}" for c in CASES]


def _desc(xs, cout, pad):
    return V._desc(xs, cout, 3, pad)


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_line_convolution_forward_and_statistics(case):
    chans, cout, pad, B, D, H, W = case
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(sum(chans) * 7 + cout + W)
    xs = [torch.randn(B, D, H, W, c, generator=g).to(dev) for c in chans]
    cin = sum(chans)
    w = (torch.randn(cout, cin, 3, 3, 3, generator=g) * (2.0 / (27 * cin)) ** 0.5).to(dev)
    b = (torch.randn(cout, generator=g) * 0.1).to(dev)
    assert V.tc_supported(_desc(xs, cout, pad)), "case must run on the tcgen05 kernel"
    for relu in (False, True):
        out = torch.full((B, D, H, W, cout), float("nan"), device=dev)
        stats = torch.zeros(B, cout, 2, device=dev, dtype=torch.float64)
        V.conv3d_forward_auto(xs, w, b, cout, 3, pad, relu, out, stats)
        torch.cuda.synchronize()
        x64 = torch.cat([x.double() for x in xs], dim=4).permute(0, 4, 1, 2, 3)
        ref = F.conv3d(F.pad(x64, (1,) * 6, mode="replicate" if pad == 1 else "constant"), w.double(), b.double())
        if relu:
            ref = torch.relu(ref)
        ref = ref.permute(0, 2, 3, 4, 1)
        assert torch.isfinite(out).all()
        e = rel_l2(out.cpu().numpy(), ref.cpu().numpy())
        s_ref = torch.stack([ref.sum(dim=(1, 2, 3)), (ref * ref).sum(dim=(1, 2, 3))], dim=2)
        e_s = rel_l2(stats.cpu().numpy(), s_ref.cpu().numpy())
        print(f"\n[conv_tc fwd {case} relu={relu}] out {e:.2e} stats {e_s:.2e}")
        assert e < TOL, (case, relu, e)
        assert e_s < 2e-3, (case, relu, e_s)       # sums cancel: a looser bound than the element-wise one


# forward shapes the line kernel does not take (weights too large) but whose data gradient it does
DGRAD_ONLY = [
    ((128,), 32, 1, 8, 4, 32, 32),           # d3_c2: dX has 128 channels = four chunks, weights staged per chunk
    ((64,), 32, 1, 4, 6, 32, 32),            # d3_c1
]


@pytest.mark.parametrize("case", CASES + DGRAD_ONLY, ids=IDS + [f"dgrad_only{i}" for i in range(len(DGRAD_ONLY))])
def test_line_convolution_data_gradient(case):
    """dX of the convolution whose forward is ``case`` (the kernel runs with the channel roles swapped)."""
    chans, cout, pad, B, D, H, W = case
    cin = sum(chans)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(cin * 11 + cout + H)
    gy = torch.randn(B, D, H, W, cout, generator=g).to(dev)
    w = (torch.randn(cout, cin, 3, 3, 3, generator=g) * (2.0 / (27 * cin)) ** 0.5).to(dev)
    dt = V._desc([gy], cin, 3, V.PAD_REPLICATE_ADJOINT if pad == 1 else V.PAD_ZEROS)
    if not V.tc_supported(dt):
        pytest.skip("gradient shape served by the warp-level kernel")
    dx = torch.full((B, D, H, W, cin), float("nan"), device=dev)
    V.conv3d_dgrad(gy, w, cin, 3, pad, dx)
    torch.cuda.synchronize()
    x64 = torch.zeros(B, cin, D, H, W, device=dev, dtype=torch.float64, requires_grad=True)
    y = F.conv3d(F.pad(x64, (1,) * 6, mode="replicate" if pad == 1 else "constant"), w.double())
    y.backward(gy.double().permute(0, 4, 1, 2, 3))
    e = rel_l2(dx.cpu().numpy(), x64.grad.permute(0, 2, 3, 4, 1).cpu().numpy())
    print(f"\n[conv_tc dgrad {case}] dx {e:.2e}")
    assert torch.isfinite(dx).all()
    assert e < TOL, (case, e)


def test_line_convolution_reads_channel_slices_and_writes_into_a_wider_buffer():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    buf = torch.randn(2, 4, 8, 64, 48, generator=g).to(dev)
    xs = [buf[..., 8:24], buf[..., 32:48]]
    w = (torch.randn(16, 32, 3, 3, 3, generator=g) * 0.05).to(dev)
    wide = torch.zeros(2, 4, 8, 64, 40, device=dev)
    out = wide[..., 8:24]
    assert V.tc_supported(_desc(xs, 16, 1))
    V.conv3d_forward_auto(xs, w, None, 16, 3, 1, False, out, None)
    torch.cuda.synchronize()
    cat = torch.cat([x.double() for x in xs], dim=4)
    ref = F.conv3d(F.pad(cat.permute(0, 4, 1, 2, 3), (1,) * 6, mode="replicate"), w.double()).permute(0, 2, 3, 4, 1)
    assert rel_l2(out.cpu().numpy(), ref.cpu().numpy()) < TOL
    assert float(wide[..., :8].abs().max()) == 0.0 and float(wide[..., 24:].abs().max()) == 0.0


def test_conv_block_on_the_line_kernel_matches_the_warp_level_kernel(monkeypatch):
    """The whole block (conv -> ReLU -> InstanceNorm, forward and backward) with the tcgen05 kernel on and off."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("CORRIF_CONV_TC", flag)
        xs = [torch.randn(2, 6, 8, 64, 16, generator=torch.Generator().manual_seed(5)).to(dev).requires_grad_(True),
              torch.randn(2, 6, 8, 64, 16, generator=torch.Generator().manual_seed(6)).to(dev).requires_grad_(True)]
        w = (torch.randn(8, 32, 3, 3, 3, generator=torch.Generator().manual_seed(7)) * 0.08).to(dev).requires_grad_(True)
        b = torch.zeros(8, device=dev, requires_grad=True)
        y = V.conv_block(xs, w, b, 3, V.PAD_REPLICATE)
        y.backward(torch.randn(y.shape, generator=torch.Generator().manual_seed(8)).to(dev))
        torch.cuda.synchronize()
        res[flag] = [t.detach().cpu().numpy() for t in (y, xs[0].grad, xs[1].grad, w.grad, b.grad)]
    for a, b_, name in zip(res["1"], res["0"], ("y", "dx0", "dx1", "dw", "db")):
        e = rel_l2(a, b_)
        print(f"[conv_block tc vs warp-level] {name} {e:.2e}")
        # two TF32 evaluations of the same block: ReLU mask flips at the 1e-3 level separate the gradients
        assert e < (2e-3 if name == "y" else 8e-2), (name, e)


WGRAD_CASES = [
    # (source channels, Cout, pad_mode, B, D, H, W)
    ((32,), 8, 1, 1, 5, 8, 128),             # d1_c1-like, two 64-voxel halves per line
    ((24, 8), 8, 1, 1, 3, 6, 128),           # d1_c2: cat(skip 24, up 8)
    ((16,), 8, 1, 2, 4, 4, 64),              # 16 input channels: half of the MMA rows are padding
    ((32,), 16, 1, 2, 6, 8, 64),             # d2_c1: 16 output channels (N = 128)
    ((48, 16), 16, 1, 1, 4, 6, 64),          # d2_c2: 64 input channels = two passes of 32
    ((32,), 8, 0, 1, 3, 4, 64),              # zero padding
    ((32,), 8, 1, 1, 1, 2, 64),              # a single plane, a single line pair
    ((32,), 8, 1, 2, 37, 32, 64),            # several z ranges, many items per CTA (ring wrap-around across items)
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[f"{'+'.join(map(str, c[0]))}to{c[1]}p{c[2]}_{c[3]}x{c[4]}x{c[5]}x{c[6]}" for c in WGRAD_CASES])
def test_tensor_core_weight_gradient(case):
    """corrif_conv3d_wgrad_tc (transposing producers + tcgen05 MMAs over voxels) against the fp64 weight gradient."""
    chans, cout, pad, B, D, H, W = case
    cin = sum(chans)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(cin * 13 + cout + D)
    xs = [torch.randn(B, D, H, W, c, generator=g).to(dev) for c in chans]
    gy = torch.randn(B, D, H, W, cout, generator=g).to(dev)
    d = V._desc(xs, cout, 3, pad)
    from corrif_b200 import ops
    assert ops.lib().corrif_conv3d_wgrad_tc_supported(d), "case must run on the tcgen05 kernel"
    dW = torch.zeros(cout, cin, 3, 3, 3, device=dev)
    V.conv3d_wgrad(xs, gy, dW, 3, pad)
    torch.cuda.synchronize()
    x64 = torch.cat([x.double() for x in xs], dim=4).permute(0, 4, 1, 2, 3)
    w64 = torch.zeros(cout, cin, 3, 3, 3, device=dev, dtype=torch.float64, requires_grad=True)
    y = F.conv3d(F.pad(x64, (1,) * 6, mode="replicate" if pad == 1 else "constant"), w64)
    y.backward(gy.double().permute(0, 4, 1, 2, 3))
    e = rel_l2(dW.cpu().numpy(), w64.grad.cpu().numpy())
    print(f"\n[conv_tc wgrad {case}] dW {e:.2e}")
    assert torch.isfinite(dW).all()
    assert e < TOL, (case, e)
