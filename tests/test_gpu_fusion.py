"""Whole fusion block (forward + backward) on the B200 against fixtures produced by the unmodified
reference (tests/golden/make_golden.py) and against the CPU oracle.

Tolerances (relative L2 per tensor, stated per the north-star):
  precision="fp32" (CUDA-core checking mode): out 1e-5, input grads 1e-5, parameter grads 2e-5
  precision="tf32" (tcgen05 hot path):        out 1e-3, input grads 3e-3, parameter grads 6e-3
Measured on B200 (round 1): fp32 mode 9.6e-7 / 8.8e-7 / 1.5e-6; tf32 mode 5.5e-4 / 1.2e-3 / 2.4e-3
(worst tensor: LayerNorm-1 weight and qkv weight gradients of the intra transformers).  The TF32
numbers agree with SURVEY.md section 7: rounding operands to 10 mantissa bits gives ~6e-4 on the
output and ~1e-3 on gradients; the north-star's 1e-3 holds for the output, gradients carry the
per-tensor tolerance stated here.
"""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2, GOLDEN

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from corrif_b200 import fusion, module
    from oracle import corrif_oracle as O

TOL = {"fp32": dict(out=1e-5, xgrad=1e-5, pgrad=2e-5), "tf32": dict(out=1e-3, xgrad=3e-3, pgrad=6e-3)}


def _sample_idx(n, k=2048):
    return np.arange(n) if n <= k else np.linspace(0, n - 1, k).astype(np.int64)


def _run_engine(batch, precision, seed=1234):
    dev = torch.device("cuda:0")
    params = {k: v.to(dev).contiguous() for k, v in O.make_params(seed).items()}
    x6, fused, gout = O.make_inputs(seed, batch)
    eng = fusion.FusionBlockEngine(params, dropout_p=0.0, precision=precision)
    out = eng.forward([x.to(dev) for x in x6], fused.to(dev)).clone()
    dx6, dfused, grads = eng.backward(gout.to(dev))
    torch.cuda.synchronize()
    return out, dx6, dfused, grads


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("batch", [1, 2, 3])
def test_fusion_block_matches_reference_fixture(batch, precision):
    g = np.load(os.path.join(GOLDEN, f"fusion_block_b{batch}.npz"))
    out, dx6, dfused, grads = _run_engine(batch, precision)
    tol = TOL[precision]
    report = {"out": rel_l2(out.cpu().numpy(), g["out"])}
    for i in range(3):
        report[f"x6.{i}"] = rel_l2(dx6[i].cpu().numpy(), g[f"grad/x6.{i}"])
    report["fused_x6"] = rel_l2(dfused.cpu().numpy(), g["grad/fused_x6"])
    worst = {}
    for k, v in grads.items():
        gk = v.reshape(-1).cpu().numpy()
        worst[k] = rel_l2(gk[_sample_idx(gk.size)], g[f"gsample/{k}"])
    top = sorted(worst.items(), key=lambda kv: -kv[1])[:5]
    print(f"\n[fusion B={batch} {precision}] out {report['out']:.2e}  xgrads "
          f"{max(report[f'x6.{i}'] for i in range(3)):.2e}/{report['fused_x6']:.2e}  worst pgrads "
          + ", ".join(f"{k}:{e:.2e}" for k, e in top))
    assert report["out"] < tol["out"], report
    for i in range(3):
        assert report[f"x6.{i}"] < tol["xgrad"], report
    assert report["fused_x6"] < tol["xgrad"], report
    for k, e in worst.items():
        assert e < tol["pgrad"], (k, e)


def test_fusion_block_batch8_vs_cpu_oracle():
    """B = 8 (the DP micro-batch, and B > num_modals for the batch-mixing quirk) against the fp64
    CPU oracle, forward and input gradients."""
    batch, seed = 8, 4321
    params = O.make_params(seed)
    x6, fused, gout = O.make_inputs(seed, batch)
    ref_out, ref_g = O.fusion_block_fwd_bwd(params, x6, fused, gout, dtype=torch.float64)
    out, dx6, dfused, grads = _run_engine(batch, "tf32", seed=seed)
    assert rel_l2(out.cpu().numpy(), ref_out.numpy()) < TOL["tf32"]["out"]
    for i in range(3):
        assert rel_l2(dx6[i].cpu().numpy(), ref_g[f"x6.{i}"].numpy()) < TOL["tf32"]["xgrad"]
    assert rel_l2(dfused.cpu().numpy(), ref_g["fused_x6"].numpy()) < TOL["tf32"]["xgrad"]
    for k in ("RGB_pos", "multimodal_decode_conv.weight", "qkv_NIR.weight",
              "multimodal_transformer.cross_attention_list.0.fn.fn.qkv.weight",
              "SWIR_transformer.cross_ffn_list.0.fn.fn.net.0.weight", "fused6_encode_conv.bias"):
        assert rel_l2(grads[k].cpu().numpy(), ref_g[k].numpy()) < TOL["tf32"]["pgrad"], k


def test_custom_op_autograd_path():
    """torch.ops.corrif.fusion_block through autograd on an nn.Module with reference key names."""
    dev = torch.device("cuda:0")
    blk = module.CorrIFusionBlock(precision="fp32").to(dev)
    sd = {k: v for k, v in O.make_params(1234).items()}
    blk.load_state_dict(sd, strict=True)
    assert set(blk.state_dict().keys()) == set(O.param_shapes().keys())
    blk.eval()                                   # dropout off: deterministic parity
    g = np.load(os.path.join(GOLDEN, "fusion_block_b2.npz"))
    x6, fused, gout = O.make_inputs(1234, 2)
    xs = [x.to(dev).requires_grad_(True) for x in x6]
    fx = fused.to(dev).requires_grad_(True)
    out = blk(xs, fx)
    out.backward(gout.to(dev))
    assert rel_l2(out.detach().cpu().numpy(), g["out"]) < TOL["fp32"]["out"]
    assert rel_l2(xs[1].grad.cpu().numpy(), g["grad/x6.1"]) < TOL["fp32"]["xgrad"]
    assert rel_l2(fx.grad.cpu().numpy(), g["grad/fused_x6"]) < TOL["fp32"]["xgrad"]
    named = dict(blk.named_parameters())
    k = "NIR_transformer.cross_attention_list.0.fn.fn.proj.weight"
    gk = named[k].grad.reshape(-1).cpu().numpy()
    assert rel_l2(gk[_sample_idx(gk.size)], g[f"gsample/{k}"]) < TOL["fp32"]["pgrad"]


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_dropout_train_mode_matches_oracle_with_exported_masks(precision):
    """Train-mode dropout (p=0.1, 20 sites): export the Philox keep-masks the kernels use, feed them
    to the CPU oracle as explicit masks, compare forward and input gradients."""
    from corrif_b200 import ops
    dev = torch.device("cuda:0")
    batch, seed, p = 2, 777, 0.1
    params = O.make_params(seed)
    x6, fused, gout = O.make_inputs(seed, batch)
    eng = fusion.FusionBlockEngine({k: v.to(dev) for k, v in params.items()}, dropout_p=p, precision=precision)
    eng.seed = 99
    out = eng.forward([x.to(dev) for x in x6], fused.to(dev)).clone()
    dx6, dfused, grads = eng.backward(gout.to(dev))

    def mask(site, shape):
        n = int(np.prod(shape))
        m = torch.empty(n, device=dev)
        ops.dropout_mask(m, n, p, 99, site)
        return (m.view(*shape) / (1 - p)).double().cpu()

    masks = {}
    names = [f"{m}_transformer" for m in O.MODALITIES] + ["multimodal_transformer"]
    for t, name in enumerate(names):
        N = 512 if t < 3 else 2048
        a, f = f"{name}.cross_attention_list.0.fn", f"{name}.cross_ffn_list.0.fn"
        masks[f"{a}.fn.attn_drop"] = mask(t * 8 + 0, (batch, 8, N, N))
        masks[f"{a}.fn.proj_drop"] = mask(t * 8 + 1, (batch, N, 512))
        masks[f"{a}.dropout"] = mask(t * 8 + 2, (batch, N, 512))
        masks[f"{f}.fn.net.2"] = mask(t * 8 + 3, (batch, N, 512))
        masks[f"{f}.fn.net.4"] = mask(t * 8 + 4, (batch, N, 512))
    ref_out, ref_g = O.fusion_block_fwd_bwd(params, x6, fused, gout, masks=masks, dtype=torch.float64)
    tol = TOL[precision]
    assert rel_l2(out.cpu().numpy(), ref_out.numpy()) < tol["out"]
    for i in range(3):
        assert rel_l2(dx6[i].cpu().numpy(), ref_g[f"x6.{i}"].numpy()) < tol["xgrad"]
    assert rel_l2(dfused.cpu().numpy(), ref_g["fused_x6"].numpy()) < tol["xgrad"]
    for k in ("multimodal_transformer.cross_ffn_list.0.fn.fn.net.3.weight",
              "multimodal_transformer.cross_attention_list.0.fn.fn.qkv.weight",
              "RGB_transformer.cross_attention_list.0.fn.fn.qkv.weight"):
        assert rel_l2(grads[k].cpu().numpy(), ref_g[k].numpy()) < tol["pgrad"], k


@pytest.mark.gpu
def test_cuda_graph_replay_matches_stream_launches():
    """use_graphs=True: forward/backward captured after two eager calls and replayed must give exactly
    the eager results for the same seed (dropout on: the seed travels through seed_dev)."""
    dev = torch.device("cuda:0")
    params = {k: v.to(dev) for k, v in O.make_params(3).items()}
    x6, fused, gout = O.make_inputs(3, 2)
    x6 = [t.to(dev) for t in x6]
    fused, gout = fused.to(dev), gout.to(dev)
    ref = fusion.FusionBlockEngine(params, dropout_p=0.1, precision="tf32")
    eng = fusion.FusionBlockEngine(params, dropout_p=0.1, precision="tf32", use_graphs=True)
    flat, grads = eng.new_grad_buffers()
    for step in range(5):                      # steps 3.. are graph replays
        ref.seed = 100 + step
        r_out = ref.forward(x6, fused).clone()
        r_dx6, r_df, r_g = ref.backward(gout)
        eng.set_seed(100 + step)
        flat.zero_()
        out = eng.forward(x6, fused)
        dx6, df, _ = eng.backward(gout, grads)
        torch.cuda.synchronize()
        assert torch.equal(out, r_out), step
        assert torch.equal(dx6, r_dx6) and torch.equal(df, r_df), step
        worst = max(float((grads[n] - r_g[n]).abs().max() / (r_g[n].abs().max() + 1e-30)) for n in grads)
        assert worst < 1e-5, (step, worst)     # weight gradients: reduce-add order is not fixed
    assert sum("graph" in e for e in eng._graphs.values()) == 2


@pytest.mark.gpu
def test_cuda_graph_replay_with_moving_inputs_switches_to_engine_staging():
    """Inputs at a new address every call (the full model: x6 / fused_x6 come out of the encoders wherever
    the allocator put them): after MAX_POINTER_KEYED address sets the engine copies into its own staging
    buffers and replays ONE graph per direction; results equal stream launches for the same seed."""
    dev = torch.device("cuda:0")
    params = {k: v.to(dev) for k, v in O.make_params(4).items()}
    ref = fusion.FusionBlockEngine(params, dropout_p=0.1, precision="tf32")
    eng = fusion.FusionBlockEngine(params, dropout_p=0.1, precision="tf32", use_graphs=True)
    flat, grads = eng.new_grad_buffers()
    keep = []
    for step in range(8):
        x6, fused, gout = O.make_inputs(10 + step, 2)
        x6 = [t.to(dev) for t in x6]
        fused, gout = fused.to(dev), gout.to(dev)
        keep.append((x6, fused, gout))             # keep them alive: every step sees fresh addresses
        ref.seed = 200 + step
        r_out = ref.forward(x6, fused).clone()
        r_dx6, r_df, r_g = ref.backward(gout)
        eng.set_seed(200 + step)
        flat.zero_()
        out = eng.forward(x6, fused)
        dx6, df, _ = eng.backward(gout, grads)
        torch.cuda.synchronize()
        assert torch.equal(out, r_out), step
        assert torch.equal(dx6, r_dx6) and torch.equal(df, r_df), step
    assert eng._graph_mode == {("fwd", 2): "static", ("bwd", 2): "static"}
    assert sum("graph" in e for e in eng._graphs.values()) == 2


def test_full_size_batch16_tf32_path_against_fp32_checking_mode_and_linearity():
    """BASELINE.json configs[1] size (batch 16), where the CPU oracle no longer finishes in seconds:
    (a) the tcgen05 TF32 path against the library's own fp32 checking mode (itself pinned to the reference
        fixtures at B = 1..3 and to the oracle at B = 8) - output and input gradients;
    (b) train mode, fixed seed: the backward is linear in the upstream gradient,
        bwd(2 g1 - 3 g2) == 2 bwd(g1) - 3 bwd(g2), and the forward is bit-reproducible."""
    batch, seed = 16, 99
    a = _run_engine(batch, "tf32", seed=seed)
    b = _run_engine(batch, "fp32", seed=seed)
    tol = TOL["tf32"]
    assert rel_l2(a[0].cpu().numpy(), b[0].cpu().numpy()) < tol["out"]
    assert rel_l2(a[1].cpu().numpy(), b[1].cpu().numpy()) < tol["xgrad"]
    assert rel_l2(a[2].cpu().numpy(), b[2].cpu().numpy()) < tol["xgrad"]
    for k in ("multimodal_transformer.cross_attention_list.0.fn.fn.qkv.weight", "qkv_RGB.weight", "NIR_pos",
              "multimodal_decode_conv.bias"):
        assert rel_l2(a[3][k].cpu().numpy(), b[3][k].cpu().numpy()) < tol["pgrad"], k
    del a, b
    dev = torch.device("cuda:0")
    params = {k: v.to(dev).contiguous() for k, v in O.make_params(seed).items()}
    x6, fused, g1 = O.make_inputs(seed, batch)
    g2 = O.make_inputs(seed + 1, batch)[2]
    x6, fused, g1, g2 = [t.to(dev) for t in x6], fused.to(dev), g1.to(dev), g2.to(dev)
    eng = fusion.FusionBlockEngine(params, dropout_p=0.1, precision="tf32")
    res = []
    for g in (g1, g2, 2.0 * g1 - 3.0 * g2):
        eng.seed = 5
        out = eng.forward(x6, fused).clone()
        dx6, df, _ = eng.backward(g)
        res.append((out, dx6.clone(), df.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][0], res[2][0])
    for i in (1, 2):
        lin = 2.0 * res[0][i] - 3.0 * res[1][i]
        assert rel_l2(res[2][i].cpu().numpy(), lin.cpu().numpy()) < 2e-3, i



def test_headline_batch16_tf32_against_fp64_oracle_per_tensor_table():
    """BASELINE.json configs[1] itself (batch 16, the size bench.py quotes the fusion block on) against the fp64 CPU
    oracle: output, all four input gradients and EVERY parameter gradient.  Writes the per-tensor relative-L2 table
    to gpurun_out/parity_b16_table.txt (committed as profiles/r02_parity_b16_table.txt; DESIGN.md section 2 marks
    the tensors above the north-star's flat 1e-3).  Tolerances are the per-tensor TF32 ones of this file's header."""
    batch, seed = 16, 2025
    params = O.make_params(seed)
    x6, fused, gout = O.make_inputs(seed, batch)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_out, ref_g = O.fusion_block_fwd_bwd(params, x6, fused, gout, dtype=torch.float64)
    out, dx6, dfused, grads = _run_engine(batch, "tf32", seed=seed)
    rows = [("out x6_inter", rel_l2(out.cpu().numpy(), ref_out.numpy()), TOL["tf32"]["out"])]
    for i in range(3):
        rows.append((f"d x6.{i}", rel_l2(dx6[i].cpu().numpy(), ref_g[f"x6.{i}"].numpy()), TOL["tf32"]["xgrad"]))
    rows.append(("d fused_x6", rel_l2(dfused.cpu().numpy(), ref_g["fused_x6"].numpy()), TOL["tf32"]["xgrad"]))
    for k in fusion.param_names():
        rows.append((f"d {k}", rel_l2(grads[k].cpu().numpy(), ref_g[k].numpy()), TOL["tf32"]["pgrad"]))
    lines = ["# fusion block, batch 16, tf32 hot path vs fp64 CPU oracle (relative L2 per tensor); * = above 1e-3",
             "%-72s %10s %8s" % ("tensor", "rel_l2", "tol")]
    lines += ["%-72s %10.3e %8.0e %s" % (n, e, t, "*" if e > 1e-3 else "") for n, e, t in rows]
    above = [n for n, e, _ in rows if e > 1e-3]
    lines.append("# %d of %d tensors above 1e-3; worst %s" % (len(above), len(rows), max(rows, key=lambda r: r[1])[:2]))
    os.makedirs(os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out", "parity_b16_table.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n" + "\n".join(lines[-12:]))
    for n, e, t in rows:
        assert e < t, (n, e, t)


def test_backward_of_a_stale_forward_raises():
    """Two forwards, then the backward of each: the second (most recent) works, the first raises instead of silently
    differentiating the wrong saved activations."""
    dev = torch.device("cuda:0")
    blk = module.CorrIFusionBlock(precision="tf32").to(dev).eval()
    blk.load_state_dict(O.make_params(7), strict=True)
    ins = []
    for s in (1, 2):
        x6, fused, gout = O.make_inputs(s, 2)
        ins.append(([x.to(dev).requires_grad_(True) for x in x6], fused.to(dev).requires_grad_(True), gout.to(dev)))
    out_a = blk(ins[0][0], ins[0][1])
    out_b = blk(ins[1][0], ins[1][1])
    out_b.backward(ins[1][2])                                   # most recent forward: fine
    assert ins[1][1].grad is not None
    with pytest.raises(module.StaleForwardError):
        out_a.backward(ins[0][2])
    out_c = blk(ins[0][0], ins[0][1])                           # a fresh forward can be differentiated again
    out_c.backward(ins[0][2])
    ref_out, ref_g = O.fusion_block_fwd_bwd(O.make_params(7), *O.make_inputs(1, 2), dtype=torch.float64)
    assert rel_l2(ins[0][1].grad.cpu().numpy(), ref_g["fused_x6"].numpy()) < TOL["tf32"]["xgrad"]


def test_kernels_against_stock_pytorch_on_the_same_gpu():
    """The plain-PyTorch fp32 restatement of the block (baseline/eager_mmvit4.EagerFusionBlock, cuBLAS/cuDNN with TF32
    off) on the same device: the comparison bench.py's eager_b200 arm times."""
    from baseline.eager_mmvit4 import EagerFusionBlock
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        dev = torch.device("cuda:0")
        ref = EagerFusionBlock(0.0).to(dev)
        ref.load_state_dict(O.make_params(1234), strict=True)
        x6, fused, gout = O.make_inputs(1234, 4)
        xs = [x.to(dev).requires_grad_(True) for x in x6]
        fx = fused.to(dev).requires_grad_(True)
        y = ref(xs, fx)
        y.backward(gout.to(dev))
        out, dx6, dfused, grads = _run_engine(4, "tf32")
        assert rel_l2(out.cpu().numpy(), y.detach().cpu().numpy()) < TOL["tf32"]["out"]
        assert rel_l2(dfused.cpu().numpy(), fx.grad.cpu().numpy()) < TOL["tf32"]["xgrad"]
        named = dict(ref.named_parameters())
        for k in ("multimodal_decode_conv.weight", "qkv_RGB.bias", "NIR_pos",
                  "multimodal_transformer.cross_ffn_list.0.fn.fn.net.3.weight"):
            assert rel_l2(grads[k].cpu().numpy(), named[k].grad.cpu().numpy()) < TOL["tf32"]["pgrad"], k
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved


def test_six_modalities_3584_token_block_against_oracle():
    """BASELINE.json configs[4] groundwork: every 3-band group of the 20-band cube as its own modality - six modalities,
    a 7 x 512 = 3584-token multimodal transformer.  The reference hard-wires three (mmvit4.py:15, 394-396), so the
    checker is the parametrised restatement (oracle.fusion_block(modalities=...)), which IS the reference at M = 3
    (tests/test_oracle_golden.py pins it to the reference fixtures); engine and restatement share nothing."""
    mods = ("RGB", "NIR", "SWIR", "G4", "G5", "G6")
    seed, batch = 606, 1
    params = O.make_params(seed, mods)
    x6, fused, gout = O.make_inputs(seed, batch, mods)
    assert fused.shape[1] == 64 * 6 and params["multimodal_decode_conv.weight"].shape[:2] == (384, 3584)
    torch.set_num_threads(os.cpu_count() or 1)
    p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    xs = [x.double().requires_grad_(True) for x in x6]
    fx = fused.double().requires_grad_(True)
    ref = O.fusion_block(p64, xs, fx, None, mods)
    ref.backward(gout.double())
    dev = torch.device("cuda:0")
    eng = fusion.FusionBlockEngine({k: v.to(dev).contiguous() for k, v in params.items()}, dropout_p=0.0,
                                   precision="tf32", modalities=mods)
    out = eng.forward([x.to(dev) for x in x6], fused.to(dev)).clone()
    dx6, dfused, grads = eng.backward(gout.to(dev))
    torch.cuda.synchronize()
    assert out.shape == (batch, 384, 8, 8, 8)
    tol = TOL["tf32"]
    assert rel_l2(out.cpu().numpy(), ref.detach().numpy()) < tol["out"]
    assert rel_l2(dfused.cpu().numpy(), fx.grad.numpy()) < tol["xgrad"]
    for i in range(6):
        assert rel_l2(dx6[i].cpu().numpy(), xs[i].grad.numpy()) < tol["xgrad"], i
    for k in ("G6_pos", "qkv_G5.weight", "G4_transformer.cross_attention_list.0.fn.fn.qkv.weight", "fused6_encode_conv.weight",
              "multimodal_transformer.cross_attention_list.0.fn.fn.proj.weight", "multimodal_decode_conv.weight", "RGB_encode_conv.bias"):
        assert rel_l2(grads[k].cpu().numpy(), p64[k].grad.numpy()) < tol["pgrad"], k
