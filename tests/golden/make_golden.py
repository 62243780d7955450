"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (``python tests/golden/make_golden.py``); the GPU box has no
/root/reference and only ever reads the committed .npz files.  Recipe = SURVEY.md appendix A:
  * ``mmvit4.resnet50`` patched to ``weights=None`` (no network; lossless);
  * the three encoders replaced by stubs that return a preset x6, the decoder by a stub that
    returns its fifth argument, so ``MMVit4.forward`` executes its own lines 449-529;
  * every nn.Dropout p=0, model kept in ``.train()``;
  * weights/inputs from ``oracle.corrif_oracle.make_params / make_inputs`` (numpy PCG64).
What is stored: the block output, d/d(x6), d/d(fused_x6) in full for small batches, and for the
10.3 M parameter gradients a per-tensor L2 norm plus a strided sample (keeps files small).
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import torchvision  # noqa: E402
import mmvit4 as ref_mmvit4  # noqa: E402  (the reference)
import F5_JACCARD2 as ref_jac  # noqa: E402
from oracle import corrif_oracle as O  # noqa: E402

ref_mmvit4.resnet50 = lambda pretrained=True: torchvision.models.resnet50(weights=None)

GRAD_SAMPLE = 2048


class _StubEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.x6 = None

    def forward(self, x):
        B = self.x6.shape[0]
        ph = [torch.zeros(B, c, 1, 2, 2, dtype=self.x6.dtype) for c in (8, 16, 32, 64, 64)]
        return (*ph, self.x6)


class _StubDecoder(nn.Module):
    def forward(self, x1, x2, x3, x4, x5):
        return x5


class _StubFusion6(nn.Module):
    """fusion6 (EarlyFusionBlock) sits before the hot path; replace it by a stub that returns the
    preset fused_x6 so the block's second input is controlled."""
    def __init__(self):
        super().__init__()
        self.fused = None

    def forward(self, a, b, c):
        return self.fused


def build_reference(seed, dtype):
    torch.manual_seed(0)
    m = ref_mmvit4.MMVit4(num_cls=1)
    m.RGB_encoder, m.NIR_encoder, m.SWIR_encoder = _StubEncoder(), _StubEncoder(), _StubEncoder()
    m.decoder_fuse = _StubDecoder()
    m.fusion6 = _StubFusion6()
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    params = O.make_params(seed)
    sd = m.state_dict()
    for k, v in params.items():
        assert k in sd and tuple(sd[k].shape) == tuple(v.shape), k
    m.load_state_dict(params, strict=False)
    m = m.to(dtype).train()
    return m, params


def sample_idx(n):
    if n <= GRAD_SAMPLE:
        return np.arange(n)
    return np.linspace(0, n - 1, GRAD_SAMPLE).astype(np.int64)


def run_block(seed, batch, dtype):
    m, params = build_reference(seed, dtype)
    x6, fused, gout = O.make_inputs(seed, batch)
    xs = [x.to(dtype).requires_grad_(True) for x in x6]
    fx = fused.to(dtype).requires_grad_(True)
    m.RGB_encoder.x6, m.NIR_encoder.x6, m.SWIR_encoder.x6 = xs
    m.fusion6.fused = fx
    out = m(torch.zeros(batch, 3, 1, 1, 1, dtype=dtype))
    out.backward(gout.to(dtype))
    rec = {"out": out.detach().numpy()}
    for i, x in enumerate(xs):
        rec[f"grad/x6.{i}"] = x.grad.numpy()
    rec["grad/fused_x6"] = fx.grad.numpy()
    named = dict(m.named_parameters())
    for k in params:
        g = named[k].grad.detach().reshape(-1).numpy()
        rec[f"gnorm/{k}"] = np.array(np.linalg.norm(g.astype(np.float64)))
        rec[f"gsample/{k}"] = g[sample_idx(g.size)]
    return rec


def make_block_golden():
    for batch in (1, 2, 3):
        rec64 = run_block(seed=1234, batch=batch, dtype=torch.float64)
        rec32 = run_block(seed=1234, batch=batch, dtype=torch.float32)
        out = {}
        for k, v in rec64.items():
            out[k] = v.astype(np.float32) if v.ndim else v
        # fp32-vs-fp64 spread of the reference itself: the floor any fp32 kernel is judged against
        out["ref_fp32_relerr_out"] = np.array(
            np.linalg.norm(rec32["out"] - rec64["out"]) / np.linalg.norm(rec64["out"]))
        np.savez_compressed(os.path.join(HERE, f"fusion_block_b{batch}.npz"), **out)
        print("fusion block golden B=%d  |out|=%.4f  ref fp32 relerr=%.2e" % (
            batch, np.linalg.norm(rec64["out"]), out["ref_fp32_relerr_out"]))


def make_inter_attn_golden():
    """The inter_attn closure cannot be imported (it is local to forward); run the reference's own
    lines by feeding identity-like weights is not possible either, so capture it through hooks:
    q,k,v are outputs of qkv_* convs, and x6_*_ are recovered from the multimodal transformer's
    input minus the skip tokens."""
    rec = {}
    for batch in (1, 2, 3, 4):
        m, params = build_reference(99, torch.float64)
        x6, fused, _ = O.make_inputs(99, batch)
        m.RGB_encoder.x6, m.NIR_encoder.x6, m.SWIR_encoder.x6 = [x.double() for x in x6]
        m.fusion6.fused = fused.double()
        cap = {}
        hooks = []
        for name in ("qkv_RGB", "qkv_NIR", "qkv_SWIR"):
            hooks.append(getattr(m, name).register_forward_hook(
                lambda mod, i, o, name=name: cap.__setitem__(name, o.detach())))
        for name in ("RGB_encode_conv", "NIR_encode_conv", "SWIR_encode_conv"):
            hooks.append(getattr(m, name).register_forward_hook(
                lambda mod, i, o, name=name: cap.__setitem__(name, o.detach())))
        hooks.append(m.multimodal_transformer.register_forward_hook(
            lambda mod, i, o: cap.__setitem__("mm_in", i[0].detach())))
        with torch.no_grad():
            m(torch.zeros(batch, 3, 1, 1, 1, dtype=torch.float64))
        for h in hooks:
            h.remove()
        # token-major [B,S,C] views
        def tok(t):
            return t.reshape(t.shape[0], t.shape[1], -1).transpose(1, 2)
        qkv = [tok(cap[n]) for n in ("qkv_RGB", "qkv_NIR", "qkv_SWIR")]
        q = np.stack([t[..., :512].numpy() for t in qkv])
        k = np.stack([t[..., 512:1024].numpy() for t in qkv])
        v = np.stack([t[..., 1024:].numpy() for t in qkv])
        skip = np.stack([tok(cap[n]).numpy() for n in ("RGB_encode_conv", "NIR_encode_conv", "SWIR_encode_conv")])
        fusedtok = cap["mm_in"][:, :1536].reshape(batch, 3, 512, 512).permute(1, 0, 2, 3).numpy()
        sl = (slice(None), slice(None), slice(0, 16), slice(0, 32))      # keep it small
        rec[f"b{batch}/q"] = q[sl].astype(np.float32)
        rec[f"b{batch}/k"] = k[sl].astype(np.float32)
        rec[f"b{batch}/v"] = v[sl].astype(np.float32)
        rec[f"b{batch}/skip"] = skip[sl].astype(np.float32)
        rec[f"b{batch}/out"] = fusedtok[sl].astype(np.float32)
        # exactness of the closed form on the full tensors, in fp64
        mine = O.inter_corr_fwd_np(q, k, v, skip)
        err = np.abs(mine - fusedtok).max()
        print("inter_attn closed form vs reference  B=%d  max abs err %.3e" % (batch, err))
        assert err < 1e-12
    np.savez_compressed(os.path.join(HERE, "inter_attn.npz"), **rec)


def make_jaccard_golden():
    rng = np.random.default_rng(5)
    rec = {}
    cases = {}
    n = 4 * 64 * 64
    label = rng.integers(0, 10, n)
    label[label == 7] = 3                                     # class 7 empty -> inversion branch
    pred = np.where(rng.random(n) < 0.7, label, rng.integers(0, 10, n))
    for c in range(10):
        cases[f"hard_c{c}"] = ((label == c).astype(np.float32), (pred == c).astype(np.float32))
    y = (rng.random(n) < 0.3).astype(np.float32)
    cases["soft"] = (y, rng.random(n).astype(np.float32))     # sigmoid-like probabilities
    cases["soft_empty"] = (np.zeros(n, np.float32), rng.random(n).astype(np.float32))
    cases["all_ones"] = (np.ones(n, np.float32), np.ones(n, np.float32))
    cases["tiny"] = (np.array([1, 0, 1], np.float32), np.array([1, 1, 0], np.float32))
    rec["label"] = label.astype(np.uint8)
    rec["pred"] = pred.astype(np.uint8)
    for name, (y, yp) in cases.items():
        ty, tp_ = torch.from_numpy(y).reshape(-1, 1), torch.from_numpy(yp).reshape(-1, 1)
        rec[f"{name}/y"] = y
        rec[f"{name}/y_pred"] = yp
        rec[f"{name}/Jaccard"] = ref_jac.Jaccard(ty, tp_).numpy()
        rec[f"{name}/Jaccard2"] = ref_jac.Jaccard2(ty, tp_).numpy()
        rec[f"{name}/JaccardAndF1"] = ref_jac.JaccardAndF1(ty, tp_).numpy()
        # the three sums as the reference forms them (fp32 torch reductions)
        yy, pp = (1 - ty, 1 - tp_) if ty.sum(0) == 0 else (ty, tp_)
        rec[f"{name}/sums2"] = np.array([(pp * yy).sum(0).item(), ((1 - pp) * yy).sum(0).item(),
                                         ((1 - yy) * pp).sum(0).item()], np.float32)
    np.savez_compressed(os.path.join(HERE, "jaccard.npz"), **rec)
    print("jaccard golden: %d cases" % len(cases))


def make_loss_golden():
    """BCEWithLogitsLoss on sigmoid output (F4_TRAIN.py:58-60) + its gradient wrt the probs."""
    rng = np.random.default_rng(11)
    probs = torch.from_numpy(rng.random((2, 3, 1, 32, 32)).astype(np.float32)).requires_grad_(True)
    masks = torch.from_numpy((rng.random((2, 1, 1, 32, 32)) < 0.3).astype(np.float32)).repeat(1, 3, 1, 1, 1)
    loss = nn.BCEWithLogitsLoss()(probs, masks)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, "bce_loss.npz"), probs=probs.detach().numpy(),
                        masks=masks.numpy(), loss=loss.detach().numpy(), grad=probs.grad.numpy())
    print("bce golden loss=%.6f" % loss.item())


def make_next_rows_golden():
    """SURVEY.md 8f rows N2 / N1, straight from the reference classes (fp64 run, stored as fp32):
    EarlyFusionBlock at the fusion6 shape (mmvit4.py:64-81, 449-454) and the decoder's
    general_conv3d_prenorm blocks, 3x3x3 replicate-padded and 1x1x1 (mmvit4.py:29-45, 225-237), forward and
    backward.  Parity anchors for the kernels of the next round."""
    g = torch.Generator().manual_seed(77)
    rec = {}
    # ---- N2: EarlyFusionBlock(in_channels=64) on three [B,64,8,8,8] bottlenecks
    blk = ref_mmvit4.EarlyFusionBlock(64).double()
    with torch.no_grad():
        blk.conv.weight.copy_(torch.randn(blk.conv.weight.shape, generator=g, dtype=torch.float64) * 0.05)
        blk.conv.bias.copy_(torch.randn(blk.conv.bias.shape, generator=g, dtype=torch.float64) * 0.1)
    xs = [torch.randn(2, 64, 8, 8, 8, generator=g, dtype=torch.float64).requires_grad_(True) for _ in range(3)]
    gout = torch.randn(2, 192, 8, 8, 8, generator=g, dtype=torch.float64)
    y = blk(*xs)
    y.backward(gout)
    rec["ef/w"], rec["ef/b"] = blk.conv.weight.detach().numpy(), blk.conv.bias.detach().numpy()
    for i, x in enumerate(xs):
        rec[f"ef/x{i}"], rec[f"ef/dx{i}"] = x.detach().numpy(), x.grad.numpy()
    rec["ef/gout"], rec["ef/y"] = gout.numpy(), y.detach().numpy()
    rec["ef/dw"], rec["ef/db"] = blk.conv.weight.grad.numpy(), blk.conv.bias.grad.numpy()
    # ---- N1: decoder blocks, small volumes with the decoder's channel counts (d1_c2: 32 -> 8, d1_out: 8 -> 8 1x1x1)
    for tag, cin, cout, k, vol in (("c3", 32, 8, 3, 12), ("c1", 8, 8, 1, 12), ("c3b", 16, 16, 3, 6)):
        m = ref_mmvit4.general_conv3d_prenorm(cin, cout, k_size=k, padding=k // 2, pad_type="replicate").double()
        with torch.no_grad():
            m.conv.weight.copy_(torch.randn(m.conv.weight.shape, generator=g, dtype=torch.float64) * 0.1)
            m.conv.bias.copy_(torch.randn(m.conv.bias.shape, generator=g, dtype=torch.float64) * 0.1)
        x = torch.randn(2, cin, vol, vol, vol, generator=g, dtype=torch.float64).requires_grad_(True)
        go = torch.randn(2, cout, vol, vol, vol, generator=g, dtype=torch.float64)
        y = m(x)
        y.backward(go)
        rec[f"{tag}/w"], rec[f"{tag}/b"] = m.conv.weight.detach().numpy(), m.conv.bias.detach().numpy()
        rec[f"{tag}/x"], rec[f"{tag}/gout"], rec[f"{tag}/y"] = x.detach().numpy(), go.numpy(), y.detach().numpy()
        rec[f"{tag}/dx"], rec[f"{tag}/dw"], rec[f"{tag}/db"] = x.grad.numpy(), m.conv.weight.grad.numpy(), m.conv.bias.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "next_rows.npz"), **{k: v.astype(np.float32) for k, v in rec.items()})
    print("next-rows golden: %d arrays" % len(rec))


def make_state_dict_inventory():
    """Key -> shape of the reference MMVit4.state_dict() (1140 entries): the drop-in must match it."""
    import json
    torch.manual_seed(0)
    m = ref_mmvit4.MMVit4(num_cls=1)
    inv = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(HERE, "mmvit4_state_dict_inventory.json"), "w") as f:
        json.dump(inv, f, indent=0, sort_keys=True)
    n_param = sum(p.numel() for p in m.parameters())
    print("state_dict inventory: %d entries, %d parameters" % (len(inv), n_param))


ADAM_LR = 1e-4          # lrFile.txt:1 of the reference's logged run

FULL_GRAD_KEYS = (
    "RGB_encoder.e1_c1.weight", "NIR_encoder.e3.1.conv2.weight", "SWIR_encoder.conv6.weight",
    "SWIR_encoder.e5.2.bn3.weight", "fusion1.conv.weight", "fusion6.conv.bias", "RGB_encode_conv.weight",
    "NIR_pos", "multimodal_transformer.cross_attention_list.0.fn.fn.qkv.weight",
    "RGB_transformer.cross_ffn_list.0.fn.fn.net.0.weight", "qkv_SWIR.weight",
    "multimodal_decode_conv.weight", "decoder_fuse.RFM5.fusion_layer.1.conv.weight",
    "decoder_fuse.d4_c2.conv.weight", "decoder_fuse.d1_c2.conv.weight", "decoder_fuse.final_conv.weight",
)


def _full_model_run(dtype):
    import json
    inv = json.load(open(os.path.join(HERE, "mmvit4_state_dict_inventory.json")))
    torch.manual_seed(0)
    m = ref_mmvit4.MMVit4(num_cls=1)
    m.load_state_dict(O.make_full_model_state(2024, inv), strict=True)
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    m = m.to(dtype).train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 3, 64, 64, generator=g)
    masks = (torch.rand(2, 1, 1, 224, 224, generator=g) < 0.3).float().repeat(1, 3, 1, 1, 1)
    y = m(x.to(dtype))
    loss = nn.BCEWithLogitsLoss()(y, masks.to(dtype))
    loss.backward()
    named = dict(m.named_parameters())
    grads = {k: named[k].grad.reshape(-1).double().numpy() for k in FULL_GRAD_KEYS}
    nograd = sorted(k for k, p in named.items() if p.grad is None)
    # optim.step() of F4_TRAIN.py:62 with the optimizer F2_MAIN.py:168-169 builds (Adam, default betas / eps)
    before = {k: named[k].detach().clone() for k in FULL_GRAD_KEYS}
    torch.optim.Adam(m.parameters(), ADAM_LR).step()
    deltas = {k: (named[k].detach() - before[k]).reshape(-1).double().numpy() for k in FULL_GRAD_KEYS}
    return x, masks, y.detach(), loss.item(), grads, nograd, deltas


def make_full_model_golden():
    """One full MMVit4 train-step forward/backward (F4_TRAIN.py:57-61) on a tiny batch, dropout off.
    Ground truth = the reference run in fp64; the reference's own fp32 run is compared against it and
    the per-tensor spread stored, because this decoder is ill-conditioned in fp32 (its gradients move
    by 1-4e-2 between fp32 and fp64) and a kernel cannot be held to more than the reference itself
    delivers.  Weights/inputs come from seeds (oracle.make_full_model_state); nothing large is stored."""
    x, masks, y64, loss64, g64, nograd, d64 = _full_model_run(torch.float64)
    _, _, y32, loss32, g32, _, d32 = _full_model_run(torch.float32)
    rec = {"x": x.numpy(), "masks": masks[:, :1].numpy(), "y": y64.float().numpy(), "loss": np.array(loss64),
           "ref_fp32_relerr/y": np.array(float((y32.double() - y64).norm() / y64.norm())),
           "ref_fp32_loss": np.array(loss32)}
    for k in FULL_GRAD_KEYS:
        rec[f"gnorm/{k}"] = np.array(np.linalg.norm(g64[k]))
        rec[f"gsample/{k}"] = g64[k][sample_idx(g64[k].size)].astype(np.float32)
        rec[f"ref_fp32_relerr/{k}"] = np.array(np.linalg.norm(g32[k] - g64[k]) / np.linalg.norm(g64[k]))
        # Adam-updated weights: the first step moves every weight by ~lr * sign(g); stored as the update itself
        rec[f"adam_delta_sample/{k}"] = d64[k][sample_idx(d64[k].size)].astype(np.float32)
        rec[f"adam_delta_ref_fp32_relerr/{k}"] = np.array(np.linalg.norm(d32[k] - d64[k]) / np.linalg.norm(d64[k]))
    rec["adam_lr"] = np.array(ADAM_LR)
    rec["nograd"] = np.array(nograd)
    jac = ref_jac.Jaccard2(masks[:, 0].reshape(-1, 1), y64.float()[:, 0].reshape(-1, 1))
    rec["jaccard2"] = jac.numpy()
    np.savez_compressed(os.path.join(HERE, "mmvit4_full_small.npz"), **rec)
    print("full model golden (fp64): loss %.6f  no-grad tensors %d  jaccard2 %.6f  worst ref fp32 grad spread %.2e" % (
        loss64, len(nograd), jac.item(), max(float(rec[f"ref_fp32_relerr/{k}"]) for k in FULL_GRAD_KEYS)))


def make_input_pipeline_golden():
    """the reference's get_images4 (F8_IMAGES4.py:11-95) and CrossVal (F6_CROSSVAL.py:5-37) run
    unmodified.  get_images4 hard-codes a Windows path; on Linux that is just a relative directory name, so the
    synthetic .mat tree is written under a scratch working directory.  os.listdir is pinned to sorted order for
    the call (the reference takes whatever order the file system returns)."""
    import tempfile
    import F6_CROSSVAL as ref_cv
    import F8_IMAGES4 as ref_im
    from synth_dstl import write_synthetic_dstl, TRIND, N_TILES
    rec = {}
    cwd = os.getcwd()
    real_listdir = os.listdir
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            write_synthetic_dstl(os.path.join(tmp, "C:/Users/Public/Server/data/DSTL"))
            os.listdir = lambda p=".": sorted(real_listdir(p))
            images, masks, r, g, b = ref_im.get_images4(N_TILES, 1, 5, None, TRIND, None, "x")
        finally:
            os.listdir = real_listdir
            os.chdir(cwd)
    rec["im/means_rgb"] = np.array([r, g, b], np.float32)
    rec["im/shape"], rec["im/mask_shape"] = np.array(images.shape), np.array(masks.shape)
    rec["im/sum64"] = np.array(images.double().sum().item())
    rec["im/sample"] = images.reshape(-1)[::9973].numpy()
    rec["im/band_means"] = images.double().mean(dim=(0, 3, 4)).numpy()
    rec["im/mask_sum"] = np.array(masks.sum().item())
    rec["im/mask_sample"] = masks.reshape(-1)[::9973].numpy()
    os.chdir("/root/reference")                    # randInd5985.txt lives beside the scripts
    try:
        for fno in (1, 2, 5):
            ts, tr, vl = ref_cv.CrossVal(5985, fno, 5)
            rec[f"cv/f{fno}/ts"], rec[f"cv/f{fno}/tr"], rec[f"cv/f{fno}/vl"] = (np.asarray(a, np.int32) for a in (ts, tr, vl))
        rec["cv/perm"] = np.asarray([int(x) for x in open("randInd5985.txt")], np.int32)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "input_pipeline.npz"), **rec)
    print("input pipeline golden: images", tuple(images.shape), "means", r, g, b)


if __name__ == "__main__":
    # python tests/golden/make_golden.py [--block] [--next-rows] [--input-pipeline] [--full-model]   (no flag: all)
    torch.set_num_threads(os.cpu_count())
    sys.path.insert(0, HERE)
    want = set(a for a in sys.argv[1:] if a.startswith("--"))
    every = not want
    if every or "--block" in want:
        make_jaccard_golden()
        make_loss_golden()
        make_inter_attn_golden()
        make_block_golden()
    if every or "--next-rows" in want:
        make_next_rows_golden()
    if every or "--input-pipeline" in want:
        make_input_pipeline_golden()
    if every or "--full-model" in want:
        make_state_dict_inventory()
        make_full_model_golden()
