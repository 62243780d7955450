"""Pin the CPU oracle against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import corrif_oracle as O
from conftest import rel_l2, GOLDEN


def _sample_idx(n, k=2048):
    return np.arange(n) if n <= k else np.linspace(0, n - 1, k).astype(np.int64)


@pytest.mark.parametrize("batch", [1, 2, 3])
def test_fusion_block_fwd_bwd_matches_reference(batch):
    g = np.load(os.path.join(GOLDEN, f"fusion_block_b{batch}.npz"))
    params = O.make_params(1234)
    x6, fused, gout = O.make_inputs(1234, batch)
    out, grads = O.fusion_block_fwd_bwd(params, x6, fused, gout, dtype=torch.float64)
    # fixtures are fp64 results rounded to fp32 -> 1e-6 is the storage precision
    assert rel_l2(out.numpy(), g["out"]) < 1e-6
    for i in range(3):
        assert rel_l2(grads[f"x6.{i}"].numpy(), g[f"grad/x6.{i}"]) < 1e-6
    assert rel_l2(grads["fused_x6"].numpy(), g["grad/fused_x6"]) < 1e-6
    for k in params:
        gk = grads[k].reshape(-1).numpy()
        assert abs(np.linalg.norm(gk) - float(g[f"gnorm/{k}"])) <= 1e-6 * float(g[f"gnorm/{k}"]) + 1e-12, k
        assert rel_l2(gk[_sample_idx(gk.size)], g[f"gsample/{k}"]) < 1e-5, k


def test_param_inventory_counts():
    shapes = O.param_shapes()
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert n == 10_310_336
    assert len(shapes) == 3 * 2 + 2 + 4 + 4 * 11 + 3 * 2 + 2


@pytest.mark.parametrize("batch", [1, 2, 3, 4])
def test_inter_corr_closed_form_matches_reference(batch):
    g = np.load(os.path.join(GOLDEN, "inter_attn.npz"))
    q, k, v, skip, ref = (g[f"b{batch}/{n}"].astype(np.float64) for n in ("q", "k", "v", "skip", "out"))
    out = O.inter_corr_fwd_np(q, k, v, skip)
    assert np.abs(out - ref).max() < 2e-6       # fixtures stored as fp32


@pytest.mark.parametrize("batch", [1, 2, 3, 5, 8])
def test_inter_corr_backward_matches_autograd(batch):
    rng = np.random.default_rng(batch)
    shp = (3, batch, 4, 8)
    q, k, v, gu = (rng.standard_normal(shp) for _ in range(4))
    tq, tk, tv = (torch.tensor(a, requires_grad=True) for a in (q, k, v))
    outs = [O.inter_attn(tq[i], list(tk), list(tv)) for i in range(3)]
    torch.stack(outs).backward(torch.tensor(gu))
    dq, dk, dv = O.inter_corr_bwd_np(q, k, v, gu)
    assert np.abs(dq - tq.grad.numpy()).max() < 1e-12
    assert np.abs(dk - tk.grad.numpy()).max() < 1e-12
    assert np.abs(dv - tv.grad.numpy()).max() < 1e-12
    fwd = O.inter_corr_fwd_np(q, k, v)
    assert np.abs(fwd - torch.stack(outs).detach().numpy()).max() < 1e-12


def test_jaccard_matches_reference_bit_exact():
    g = np.load(os.path.join(GOLDEN, "jaccard.npz"))
    names = sorted({k.split("/")[0] for k in g.files if "/" in k})
    assert len(names) == 14
    for n in names:
        y, yp = g[f"{n}/y"], g[f"{n}/y_pred"]
        hard = n.startswith("hard") or n in ("all_ones", "tiny")
        j, j2, f1 = O.jaccard_np(y, yp), O.jaccard2_np(y, yp), O.jaccard_and_f1_np(y, yp)
        if hard:  # integer-valued sums: bit-exact
            assert j == g[f"{n}/Jaccard"][0], n
            assert j2 == g[f"{n}/Jaccard2"][0], n
            assert f1 == g[f"{n}/JaccardAndF1"][0], n
            tp, fp, fn, _ = O.jaccard_sums_np(y, yp, True)
            assert [tp, fp, fn] == list(g[f"{n}/sums2"]), n
        else:     # soft inputs: fp32 summation order differs -> 1e-6 relative
            assert abs(j - g[f"{n}/Jaccard"][0]) <= 1e-6 * abs(g[f"{n}/Jaccard"][0]) + 1e-12, n
            assert abs(j2 - g[f"{n}/Jaccard2"][0]) <= 1e-6 * abs(g[f"{n}/Jaccard2"][0]), n
            assert abs(f1 - g[f"{n}/JaccardAndF1"][0]) <= 1e-6 * abs(g[f"{n}/JaccardAndF1"][0]), n


def test_confusion_counts_equal_reference_sums():
    g = np.load(os.path.join(GOLDEN, "jaccard.npz"))
    cm, tp, fp, fn = O.confusion_counts_np(g["label"], g["pred"], 10)
    assert cm.sum() == g["label"].size
    for c in range(10):
        y = g[f"hard_c{c}/y"]
        if y.sum() == 0:       # empty class: the reference inverts (F5_JACCARD2.py:12-14)
            n = y.size
            s = g[f"hard_c{c}/sums2"]
            # inverted (y'=1, p'=1-p): TP' = #(p=0), "FP'" = #(p=1), "FN'" = 0
            assert s[0] == n - g[f"hard_c{c}/y_pred"].sum() and s[2] == 0
            assert s[1] == g[f"hard_c{c}/y_pred"].sum()
        else:
            assert list(g[f"hard_c{c}/sums2"]) == [tp[c], fp[c], fn[c]]


def test_bce_on_probs_matches_reference():
    g = np.load(os.path.join(GOLDEN, "bce_loss.npz"))
    p = torch.from_numpy(g["probs"]).requires_grad_(True)
    loss = O.bce_with_logits_on_probs(p, torch.from_numpy(g["masks"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    assert np.abs(p.grad.numpy() - g["grad"]).max() < 1e-9


def test_next_rows_restatements_match_reference():
    """SURVEY.md 8f rows N2 (EarlyFusionBlock) and N1 (decoder conv -> ReLU -> InstanceNorm blocks): the oracle
    restatements against fixtures produced by the unmodified reference classes, forward and autograd backward
    (fp64 oracle vs the fp32-stored fp64 reference run)."""
    g = np.load(os.path.join(GOLDEN, "next_rows.npz"))
    t = lambda k: torch.from_numpy(g[k]).double()  # noqa: E731
    xs = [t(f"ef/x{i}").requires_grad_(True) for i in range(3)]
    w, b = t("ef/w").requires_grad_(True), t("ef/b").requires_grad_(True)
    y = O.early_fusion_block(xs, w, b)
    y.backward(t("ef/gout"))
    assert rel_l2(y.detach().numpy(), g["ef/y"]) < 1e-6
    for i in range(3):
        assert rel_l2(xs[i].grad.numpy(), g[f"ef/dx{i}"]) < 1e-5
    assert rel_l2(w.grad.numpy(), g["ef/dw"]) < 1e-5 and rel_l2(b.grad.numpy(), g["ef/db"]) < 1e-5
    for tag, k in (("c3", 3), ("c1", 1), ("c3b", 3)):
        x = t(f"{tag}/x").requires_grad_(True)
        w, b = t(f"{tag}/w").requires_grad_(True), t(f"{tag}/b").requires_grad_(True)
        y = O.general_conv3d_prenorm(x, w, b, k_size=k)
        y.backward(t(f"{tag}/gout"))
        assert rel_l2(y.detach().numpy(), g[f"{tag}/y"]) < 1e-6, tag
        assert rel_l2(x.grad.numpy(), g[f"{tag}/dx"]) < 1e-5, tag
        assert rel_l2(w.grad.numpy(), g[f"{tag}/dw"]) < 1e-5 and rel_l2(b.grad.numpy(), g[f"{tag}/db"]) < 1e-5, tag



def test_full_model_restatement_matches_reference_fixture():
    """oracle.full_model (encoders + early fusion + fusion block + decoder, mmvit4.py:441-532) against the
    unmodified reference's fp64 run stored in mmvit4_full_small.npz: output, loss, the gradient-less set and the
    gradients of 16 tensors spread over the model."""
    import json
    g = np.load(os.path.join(GOLDEN, "mmvit4_full_small.npz"))
    inv = json.load(open(os.path.join(GOLDEN, "mmvit4_state_dict_inventory.json")))
    state = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v)
             for k, v in O.make_full_model_state(2024, inv).items()}
    x = torch.from_numpy(g["x"]).double()
    masks = torch.from_numpy(g["masks"]).double().repeat(1, 3, 1, 1, 1)
    y = O.full_model(state, x)
    loss = O.bce_with_logits_on_probs(y, masks)
    loss.backward()
    assert rel_l2(y.detach().numpy(), g["y"]) < 1e-6
    assert abs(loss.item() - float(g["loss"])) < 1e-9
    params = {k for k, v in state.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    assert sorted(k for k in params if state[k].grad is None) == sorted(g["nograd"].tolist())
    for key in [k[6:] for k in g.files if k.startswith("gnorm/")]:
        gk = state[key].grad.reshape(-1).numpy()
        idx = np.arange(gk.size) if gk.size <= 2048 else np.linspace(0, gk.size - 1, 2048).astype(np.int64)
        assert rel_l2(gk[idx], g[f"gsample/{key}"]) < 1e-5, key
        assert abs(np.linalg.norm(gk) / float(g[f"gnorm/{key}"]) - 1) < 1e-6, key
