"""Parity of every libcorrif_b200 kernel against the CPU oracle / torch fp64, on a real B200.
All calls go through the C ABI (corrif_b200.ops -> ctypes -> libcorrif_b200.so)."""
import math
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2, GOLDEN

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from corrif_b200 import ops
    from oracle import corrif_oracle as O


def dev():
    return torch.device("cuda:0")


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev())


# ---------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------
def _gemm_case(M, N, K, a_mn, b_mn, precision, epilogue=None, split_k=1, seed=0, alpha=1.0):
    epilogue = ops.EPI_STORE if epilogue is None else epilogue
    A = _rand(M, K, seed=seed)
    Bm = _rand(N, K, seed=seed + 1)
    Am = A.t().contiguous() if a_mn else A
    Bs = Bm.t().contiguous() if b_mn else Bm
    bias = _rand(N, seed=seed + 2)
    res = _rand(M, N, seed=seed + 3)
    aux = _rand(M, N, seed=seed + 4)
    D0 = _rand(M, N, seed=seed + 5)
    D = D0.clone()
    ops.gemm(Am, Bs, D, M=M, N=N, K=K, lda=M if a_mn else K, ldb=N if b_mn else K, ldd=N, a_mn=a_mn,
             b_mn=b_mn, bias=bias, residual=res, ldr=N, aux=aux, ldaux=N, split_k=split_k,
             epilogue=epilogue, precision=precision, alpha=alpha)
    torch.cuda.synchronize()
    v = alpha * (A.double() @ Bm.double().t())
    if epilogue == ops.EPI_STORE:
        ref = v
    elif epilogue == ops.EPI_BIAS:
        ref = v + bias.double()
    elif epilogue == ops.EPI_BIAS_GELU:
        u = v + bias.double()
        ref = torch.nn.functional.gelu(u)
        assert rel_l2(aux.cpu().numpy(), u.cpu().numpy()) < (2e-3 if precision == ops.GEMM_TF32 else 1e-5)
    elif epilogue == ops.EPI_BIAS_RESIDUAL:
        ref = v + bias.double() + res.double()
    elif epilogue == ops.EPI_MUL_DGELU:
        u = aux.double().clone().requires_grad_(True)
        torch.nn.functional.gelu(u).sum().backward()
        ref = v * u.grad
    else:
        ref = D0.double() + v
    return D.double().cpu().numpy(), ref.cpu().numpy()


GEMM_SHAPES = [(128, 128, 32), (128, 128, 64), (256, 128, 512), (1024, 512, 64), (1024, 1536, 512),
               (1024, 192, 2048), (1024, 64, 512), (200, 192, 96), (512, 512, 512)]


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_tf32_layouts(M, N, K, a_mn, b_mn):
    if a_mn and M % 4:
        pytest.skip("MN-major A needs lda % 4 == 0")
    got, ref = _gemm_case(M, N, K, a_mn, b_mn, ops.GEMM_TF32)
    err = rel_l2(got, ref)
    # TF32 operands (10-bit mantissa): tolerance 2e-3 relative L2 on N(0,1) operands
    assert err < 2e-3, f"tf32 gemm M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn}: rel err {err:.3e}"


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 192, 96), (1024, 512, 512)])
def test_gemm_fp32_exact_mode(M, N, K, a_mn, b_mn):
    if a_mn and M % 4:
        pytest.skip("MN-major A needs lda % 4 == 0")
    got, ref = _gemm_case(M, N, K, a_mn, b_mn, ops.GEMM_FP32)
    assert rel_l2(got, ref) < 2e-6


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
@pytest.mark.parametrize("epi", ["BIAS", "BIAS_GELU", "BIAS_RESIDUAL", "MUL_DGELU", "ATOMIC_ADD"])
def test_gemm_epilogues(epi, precision):
    prec = ops.GEMM_TF32 if precision == "tf32" else ops.GEMM_FP32
    e = getattr(ops, "EPI_" + epi)
    got, ref = _gemm_case(512, 512, 512, False, False, prec, epilogue=e, alpha=0.5)
    assert rel_l2(got, ref) < (2e-3 if precision == "tf32" else 2e-6)


@pytest.mark.parametrize("epi", ["STORE", "BIAS", "BIAS_GELU", "BIAS_RESIDUAL", "MUL_DGELU", "ATOMIC_ADD"])
@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(2304, 640, 256, False, False), (1000, 260, 96, False, True),
                                              (19200, 512, 64, False, False), (192, 2048, 512, True, True)])
def test_gemm_pair_tiles_epilogues(M, N, K, a_mn, b_mn, epi):
    """CTA-pair (cta_group::2) kernel with the TMA epilogue: several 256 x 256 tiles per cluster (the
    residual / aux prefetch crosses tile boundaries), ragged M and N edges, every epilogue mode."""
    got, ref = _gemm_case(M, N, K, a_mn, b_mn, ops.GEMM_TF32, epilogue=getattr(ops, "EPI_" + epi), seed=11)
    assert rel_l2(got, ref) < 2e-3


@pytest.mark.parametrize("split_k", [2, 7, 16])
def test_gemm_split_k_wgrad_shape(split_k):
    # dW[512,512] += dY[4096,512]^T . X[4096,512]: both operands MN-major, atomics
    got, ref = _gemm_case(512, 512, 4096, True, True, ops.GEMM_TF32, epilogue=ops.EPI_ATOMIC_ADD,
                          split_k=split_k)
    assert rel_l2(got, ref) < 2e-3


def test_gemm_batched_attention_views():
    """QK^T, PV, and the four backward products on strided per-(batch, head) views of a qkv buffer."""
    B, H, N, d, Cc = 2, 8, 256, 64, 512
    qkv = _rand(B * N, 3 * Cc, seed=3)
    P = torch.empty(B * H, N, N, device=dev())
    ops.gemm(qkv, (qkv, Cc), P, M=N, N=N, K=d, lda=3 * Cc, ldb=3 * Cc, ldd=N, batch=(B, H),
             a_step=(N * 3 * Cc, d), b_step=(N * 3 * Cc, d), d_step=(H * N * N, N * N), alpha=0.125)
    q, k, v = (qkv.view(B, N, 3, H, d).permute(2, 0, 3, 1, 4)[i].double() for i in range(3))
    ref = 0.125 * q @ k.transpose(-1, -2)
    assert rel_l2(P.view(B, H, N, N).cpu().numpy(), ref.cpu().numpy()) < 2e-3
    # O = P V (V read MN-major in place)
    Pn = torch.softmax(ref, -1).float().contiguous().view(B * H, N, N)
    Obuf = torch.zeros(B * N, Cc, device=dev())
    ops.gemm(Pn, (qkv, 2 * Cc), Obuf, M=N, N=d, K=N, lda=N, ldb=3 * Cc, ldd=Cc, b_mn=True,
             batch=(B, H), a_step=(H * N * N, N * N), b_step=(N * 3 * Cc, d), d_step=(N * Cc, d))
    refO = (Pn.view(B, H, N, N).double() @ v).transpose(1, 2).reshape(B * N, Cc)
    assert rel_l2(Obuf.cpu().numpy(), refO.cpu().numpy()) < 2e-3
    # dV = P^T dO (both MN-major), written into the v slot of dqkv
    dO = _rand(B * N, Cc, seed=9)
    dqkv = torch.zeros(B * N, 3 * Cc, device=dev())
    ops.gemm(Pn, dO, (dqkv, 2 * Cc), M=N, N=d, K=N, lda=N, ldb=Cc, ldd=3 * Cc, a_mn=True, b_mn=True,
             batch=(B, H), a_step=(H * N * N, N * N), b_step=(N * Cc, d), d_step=(N * 3 * Cc, d))
    dOh = dO.view(B, N, H, d).permute(0, 2, 1, 3).double()
    refdV = Pn.view(B, H, N, N).double().transpose(-1, -2) @ dOh
    got = dqkv.view(B, N, 3, H, d)[:, :, 2].permute(0, 2, 1, 3)
    assert rel_l2(got.cpu().numpy(), refdV.cpu().numpy()) < 2e-3


# ---------------------------------------------------------------------------------------------
# element-wise / reductions
# ---------------------------------------------------------------------------------------------
def test_transpose():
    x = _rand(3, 64, 512)
    out = torch.empty(3, 512, 64, device=dev())
    ops.transpose(x, out, 3, 64, 512)
    assert torch.equal(out, x.transpose(1, 2).contiguous())
    x = _rand(2, 37, 50)
    out = torch.empty(2, 50, 37, device=dev())
    ops.transpose(x, out, 2, 37, 50)
    assert torch.equal(out, x.transpose(1, 2).contiguous())


@pytest.mark.parametrize("rows,with_pos", [(8, True), (1000, True), (4099, False)])
def test_layernorm_fwd_bwd(rows, with_pos):
    Cc = 512
    x = _rand(rows, Cc, seed=1)
    pos = _rand(40, Cc, seed=2) if with_pos else None
    g, b = 1 + 0.1 * _rand(Cc, seed=3), 0.1 * _rand(Cc, seed=4)
    x1, y = torch.empty_like(x), torch.empty_like(x)
    mean, rstd = torch.empty(rows, device=dev()), torch.empty(rows, device=dev())
    ops.layernorm_fwd(x, pos, 40, g, b, x1 if with_pos else None, y, mean, rstd, rows)
    xin = x.double()
    if with_pos:
        xin = xin + pos.double()[torch.arange(rows, device=dev()) % 40]
        assert rel_l2(x1.cpu().numpy(), xin.cpu().numpy()) < 1e-7
    xin = xin.detach().requires_grad_(True)
    gd, bd = g.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xin, (Cc,), gd, bd, 1e-5)
    assert rel_l2(y.cpu().numpy(), ref.detach().cpu().numpy()) < 2e-6
    dy, dres = _rand(rows, Cc, seed=5), _rand(rows, Cc, seed=6)
    ref.backward(dy.double())
    dx, dg, db = torch.empty_like(x), torch.empty(Cc, device=dev()), torch.empty(Cc, device=dev())
    scratch = torch.empty(ops.layernorm_bwd_scratch_floats(rows), device=dev())
    ops.layernorm_bwd(dy, x1 if with_pos else x, g, mean, rstd, dres, dx, dg, db, scratch, rows)
    assert rel_l2(dx.cpu().numpy(), (xin.grad + dres.double()).cpu().numpy()) < 5e-6
    assert rel_l2(dg.cpu().numpy(), gd.grad.cpu().numpy()) < 1e-5
    assert rel_l2(db.cpu().numpy(), bd.grad.cpu().numpy()) < 1e-5


def test_layernorm_bwd_regrouped_output_and_inter_corr_layout():
    """LayerNorm-backward writing dx group-major ([batch][group][rows] -> [group][batch][rows]) equals the
    plain kernel followed by the permutation, bit for bit; inter_corr_bwd reads either layout."""
    Cc, B, G, S = 512, 3, 4, 64
    rows = B * G * S
    x, dy, dres = _rand(rows, Cc, seed=1), _rand(rows, Cc, seed=2), _rand(rows, Cc, seed=3)
    g, b = 1 + 0.1 * _rand(Cc, seed=4), 0.1 * _rand(Cc, seed=5)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(rows, device=dev()), torch.empty(rows, device=dev())
    ops.layernorm_fwd(x, None, 1, g, b, None, y, mean, rstd, rows)
    scratch = torch.empty(ops.layernorm_bwd_scratch_floats(rows), device=dev())
    dx0, dx1 = torch.empty_like(x), torch.empty_like(x)
    dg, db = torch.empty(Cc, device=dev()), torch.empty(Cc, device=dev())
    ops.layernorm_bwd(dy, x, g, mean, rstd, dres, dx0, dg, db, scratch, rows)
    ops.layernorm_bwd(dy, x, g, mean, rstd, dres, dx1, dg, db, scratch, rows, groups=G, group_rows=S)
    assert torch.equal(dx1.view(G, B, S, Cc), dx0.view(B, G, S, Cc).transpose(0, 1))
    # inter_corr_bwd: [B][(M+1)S][C] vs [M+1][B][S][C] gradient layouts
    M, S2, C2 = 3, 16, 64
    qkv = _rand(M, B, S2, 3 * C2, seed=6)
    gt = _rand(B, (M + 1) * S2, C2, seed=7)
    d0, d1 = torch.empty_like(qkv), torch.empty_like(qkv)
    ops.inter_corr_bwd(qkv, gt, d0, M, B, S2, C2)
    ops.inter_corr_bwd(qkv, gt.view(B, M + 1, S2, C2).transpose(0, 1).contiguous(), d1, M, B, S2, C2, g_group_major=True)
    assert torch.equal(d0, d1)


@pytest.mark.parametrize("cols", [512, 2048, 3584])
def test_softmax_fwd_bwd(cols):
    rows = 300
    s = _rand(rows, cols, seed=1, scale=3.0)
    ref_in = s.double().requires_grad_(True)
    ref = torch.softmax(ref_in, -1)
    P = s.clone()
    ops.softmax_fwd(P, None, rows, cols)
    assert rel_l2(P.cpu().numpy(), ref.detach().cpu().numpy()) < 2e-6
    dP = _rand(rows, cols, seed=2)
    ref.backward(dP.double())
    d = dP.clone()
    ops.softmax_bwd(P, d, rows, cols, 0.125)
    assert rel_l2(d.cpu().numpy(), (0.125 * ref_in.grad).cpu().numpy()) < 1e-5


def test_dropout_is_counter_based_and_unbiased():
    n, p = 1 << 22, 0.1
    m1, m2 = torch.empty(n, device=dev()), torch.empty(n, device=dev())
    ops.dropout_mask(m1, n, p, seed=123, site=3)
    ops.dropout_mask(m2, n, p, seed=123, site=3)
    assert torch.equal(m1, m2)
    assert set(m1.unique().tolist()) <= {0.0, 1.0}
    keep = m1.mean().item()
    assert abs(keep - (1 - p)) < 4 * math.sqrt(p * (1 - p) / n) + 1e-4
    ops.dropout_mask(m2, n, p, seed=123, site=4)
    agree = (m1 == m2).float().mean().item()
    assert abs(agree - (0.81 + 0.01)) < 5e-3            # independent sites
    sd = torch.tensor([5], dtype=torch.int64, device=dev())
    ops.dropout_mask(m2, n, p, seed=118, site=3, seed_dev=sd)   # 118 + 5 == 123
    assert torch.equal(m1, m2)
    x = _rand(n, seed=7)
    y = torch.empty_like(x)
    ops.dropout(x, y, n, p, seed=123, site=3)
    assert torch.allclose(y, x * m1 / (1 - p), rtol=1e-6, atol=0)
    ops.dropout_mask(m2, n, p, seed=123, site=9)
    r = _rand(n, seed=8)
    ops.dropout_add(x, r, y, n, p, 123, 3, 9)
    assert torch.allclose(y, x * m1 * m2 / (1 - p) ** 2 + r, rtol=1e-5, atol=1e-6)


def test_dropout_colsum_fused():
    """dropout + column sums in one pass == dropout kernel followed by a column sum (same keep decisions)."""
    for rows, cols in ((1000, 512), (8192, 512), (37, 64)):
        x = _rand(rows, cols, seed=rows)
        ref, out = torch.empty_like(x), torch.empty_like(x)
        ops.dropout(x, ref, rows * cols, 0.1, seed=11, site=5)
        cs = torch.ones(cols, device=dev())
        ops.dropout_colsum(x, out, rows, cols, 0.1, 11, 5, cs)
        assert torch.equal(out, ref)
        assert rel_l2(cs.cpu().numpy(), (1.0 + ref.double().sum(0)).cpu().numpy()) < 1e-6


def test_softmax_dropout_consistency():
    rows, cols, p = 64, 512, 0.1
    s = _rand(rows, cols, seed=1)
    P, Pd = s.clone(), torch.empty_like(s)
    ops.softmax_fwd(P, Pd, rows, cols, p, seed=77, site=8)
    m = torch.empty(rows * cols, device=dev())
    ops.dropout_mask(m, rows * cols, p, seed=77, site=8)
    m = m.view(rows, cols)
    assert torch.allclose(Pd, P * m / (1 - p), rtol=1e-6, atol=0)
    # backward with explicit mask vs torch
    x = s.double().requires_grad_(True)
    (torch.softmax(x, -1) * m.double() / (1 - p)).backward(torch.ones_like(x) * _rand(rows, cols, seed=2).double())
    d = _rand(rows, cols, seed=2).clone()
    ops.softmax_bwd(P, d, rows, cols, 1.0, p, seed=77, site=8)
    assert rel_l2(d.cpu().numpy(), x.grad.cpu().numpy()) < 1e-5


def test_colsum_batchsum_add_rows():
    x = _rand(5000, 1536, seed=1)
    out = torch.zeros(1536, device=dev())
    scratch = torch.empty(ops.colsum_scratch_floats(5000, 1536), device=dev())
    ops.colsum(x, 1536, 5000, 1536, out, scratch)
    assert rel_l2(out.cpu().numpy(), x.double().sum(0).cpu().numpy()) < 1e-6
    ops.colsum(x, 1536, 5000, 1536, out, scratch, accumulate=True)
    assert rel_l2(out.cpu().numpy(), 2 * x.double().sum(0).cpu().numpy()) < 1e-6
    # strided view: columns [512,1024) of a wider matrix
    out2 = torch.zeros(512, device=dev())
    ops.colsum((x, 512), 1536, 5000, 512, out2, scratch)
    assert rel_l2(out2.cpu().numpy(), x[:, 512:1024].double().sum(0).cpu().numpy()) < 1e-6
    y = _rand(7, 512 * 512, seed=2)
    o = torch.empty(512 * 512, device=dev())
    ops.batchsum(y, 7, 512 * 512, 512 * 512, o)
    assert rel_l2(o.cpu().numpy(), y.double().sum(0).cpu().numpy()) < 1e-6
    a, b = _rand(100, 512, seed=3), _rand(100, 512, seed=4)
    c = torch.empty_like(a)
    ops.add_rows(a, 512, b, 512, c, 512, 100, 512)
    assert torch.equal(c, a + b)


# ---------------------------------------------------------------------------------------------
# inter-modal correlation (quirk included)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,M", [(1, 3), (2, 3), (3, 3), (4, 3), (5, 3), (8, 3),
                                 (1, 6), (2, 6), (5, 6), (6, 6), (8, 6), (3, 2), (4, 4), (7, 5)])
def test_inter_corr_fwd_bwd_vs_oracle(B, M):
    """M = 3 is the reference (the closed form is pinned to the reference's own inter_attn at B = 1..4 by
    tests/golden/inter_attn.npz); M = 6 is BASELINE.json configs[4] (every 3-band group its own modality); the
    batch-mixing view couples (modality, batch) differently for every (B, M) pair, B = M and gcd(B, M) > 1 included."""
    S_, C_ = 24, 64
    rng = np.random.default_rng(B * 10 + M)
    qkv = rng.standard_normal((M, B, S_, 3 * C_)).astype(np.float32)
    skip = rng.standard_normal((M, B, S_, C_)).astype(np.float32)
    g_tok = rng.standard_normal((B, (M + 1) * S_, C_)).astype(np.float32)
    q, k, v = qkv[..., :C_], qkv[..., C_:2 * C_], qkv[..., 2 * C_:]
    ref = O.inter_corr_fwd_np(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64),
                              skip.astype(np.float64))
    tq, ts = torch.from_numpy(qkv).to(dev()), torch.from_numpy(skip).to(dev())
    tokens = torch.full((B, (M + 1) * S_, C_), float("nan"), device=dev())
    ops.inter_corr_fwd(tq, ts, tokens, M, B, S_, C_)
    got = tokens.view(B, M + 1, S_, C_)[:, :M].permute(1, 0, 2, 3).cpu().numpy()
    assert np.isnan(tokens.view(B, M + 1, S_, C_)[:, M].cpu().numpy()).all()   # 4th group untouched
    assert rel_l2(got, ref) < 2e-6
    gX = g_tok.reshape(B, M + 1, S_, C_)[:, :M].transpose(1, 0, 2, 3).astype(np.float64)
    dq, dk, dv = O.inter_corr_bwd_np(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64), gX)
    dqkv = torch.full_like(tq, float("nan"))
    ops.inter_corr_bwd(tq, torch.from_numpy(g_tok).to(dev()), dqkv, M, B, S_, C_)
    d = dqkv.cpu().numpy()
    assert rel_l2(d[..., :C_], dq) < 5e-6
    assert rel_l2(d[..., C_:2 * C_], dk) < 5e-6
    assert rel_l2(d[..., 2 * C_:], dv) < 5e-6


@pytest.mark.parametrize("B", [1, 2, 3, 4])
def test_inter_corr_matches_reference_fixture(B):
    g = np.load(os.path.join(GOLDEN, "inter_attn.npz"))
    q, k, v, skip, ref = (g[f"b{B}/{n}"] for n in ("q", "k", "v", "skip", "out"))
    M, _, S_, C_ = q.shape
    qkv = np.concatenate([q, k, v], axis=-1)
    tokens = torch.zeros(B, (M + 1) * S_, C_, device=dev())
    ops.inter_corr_fwd(torch.from_numpy(qkv).to(dev()).contiguous(), torch.from_numpy(skip).to(dev()),
                       tokens, M, B, S_, C_)
    got = tokens.view(B, M + 1, S_, C_)[:, :M].permute(1, 0, 2, 3).cpu().numpy()
    assert rel_l2(got, ref) < 2e-6


# ---------------------------------------------------------------------------------------------
# Jaccard / confusion matrix: bit-exact
# ---------------------------------------------------------------------------------------------
def test_jaccard_bit_exact_vs_reference_fixture():
    from corrif_b200 import metrics
    g = np.load(os.path.join(GOLDEN, "jaccard.npz"))
    names = sorted({k.split("/")[0] for k in g.files if "/" in k})
    for n in names:
        y = torch.from_numpy(g[f"{n}/y"]).to(dev()).view(-1, 1)
        yp = torch.from_numpy(g[f"{n}/y_pred"]).to(dev()).view(-1, 1)
        out, sums = metrics.jaccard_all(y, yp)
        out = out.cpu().numpy()
        hard = n.startswith("hard") or n in ("all_ones", "tiny")
        refs = [g[f"{n}/Jaccard"][0], g[f"{n}/Jaccard2"][0], g[f"{n}/JaccardAndF1"][0]]
        if hard:
            assert [out[0], out[1], out[2]] == refs, (n, out, refs)       # bit-exact
        else:
            for a, b in zip(out, refs):
                assert abs(a - b) <= 2e-6 * abs(b) + 1e-12, (n, out, refs)
        assert metrics.Jaccard2(y, yp).shape == (1,)


def test_confusion_matrix_bit_exact_full_size():
    from corrif_b200 import metrics
    rng = np.random.default_rng(0)
    n = 64 * 256 * 256                                     # BASELINE config 4
    label = rng.integers(0, 10, n).astype(np.uint8)
    label[label == 7] = 3
    pred = np.where(rng.random(n) < 0.7, label, rng.integers(0, 10, n)).astype(np.uint8)
    cm = metrics.confusion_matrix(torch.from_numpy(label).to(dev()), torch.from_numpy(pred).to(dev()), 10)
    ref, tp, fp, fn = O.confusion_counts_np(label, pred, 10)
    assert np.array_equal(cm.cpu().numpy(), ref)
    assert int(cm.sum()) == n
    # per-class Jaccard2 through the float path equals the integer cells
    for c in (0, 3, 7):
        y = torch.from_numpy((label == c).astype(np.float32)).to(dev())
        yp = torch.from_numpy((pred == c).astype(np.float32)).to(dev())
        out, sums = metrics.jaccard_all(y, yp)
        s = sums.cpu().numpy()
        assert s[2] == tp[c] and s[0] - s[2] == fp[c] and s[1] - s[2] == fn[c] and s[3] == n
        assert out[1].item() == O.jaccard2_np((label == c), (pred == c))
    # ragged size
    cm2 = metrics.confusion_matrix(torch.from_numpy(label[:1003]).to(dev()), torch.from_numpy(pred[:1003]).to(dev()), 10)
    assert np.array_equal(cm2.cpu().numpy(), O.confusion_counts_np(label[:1003], pred[:1003], 10)[0])


def test_bce_and_adam():
    g = np.load(os.path.join(GOLDEN, "bce_loss.npz"))
    x, y = torch.from_numpy(g["probs"]).to(dev()), torch.from_numpy(g["masks"]).to(dev())
    n = x.numel()
    loss = torch.zeros(1, dtype=torch.float64, device=dev())
    dx = torch.empty_like(x)
    ops.bce_probs_fwd_bwd(x, y, n, 1.0 / n, loss, dx)
    assert abs(loss.item() / n - float(g["loss"])) < 1e-6
    assert np.abs(dx.cpu().numpy() - g["grad"]).max() < 1e-9
    p = _rand(10000, seed=1)
    gr = _rand(10000, seed=2)
    pr = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([pr], lr=1e-3)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in (1, 2, 3):
        pr.grad = gr.clone()
        opt.step()
        ops.adam_step(p, gr, m, v, p.numel(), 1e-3, 0.9, 0.999, 1e-8, 1.0, step)
    assert torch.allclose(p, pr.detach(), rtol=1e-5, atol=1e-7)


# ---------------------------------------------------------------------------------------------
# fused attention (tcgen05): forward and backward against torch fp64, with and without dropout
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,p", [(1, 128, 0.0), (2, 512, 0.0), (1, 2048, 0.0), (2, 256, 0.1), (1, 512, 0.1)])
def test_fused_attention_fwd_bwd(B, N, p):
    H, d, Cc = 8, 64, 512
    qkv = _rand(B * N, 3 * Cc, seed=N + 1)
    dO = _rand(B * N, Cc, seed=N + 2)
    O = torch.full((B * N, Cc), float("nan"), device=dev())
    lse = torch.empty(B * H, N, device=dev())
    delta = torch.empty(B * H, N, device=dev())
    bits = torch.zeros(B * H, N, N // 32, dtype=torch.int32, device=dev()) if p > 0 else None
    dqkv = torch.full((B * N, 3 * Cc), float("nan"), device=dev())
    ops.attention_fwd(qkv, O, lse, bits, B, N, H, d, 0.125, p, seed=31, site=8)
    ops.attention_bwd(qkv, O, dO, lse, bits, delta, dqkv, B, N, H, d, 0.125, p)
    torch.cuda.synchronize()
    x = qkv.double().requires_grad_(True)
    q, k, v = (x.view(B, N, 3, H, d).permute(2, 0, 3, 1, 4)[i] for i in range(3))
    P = torch.softmax(0.125 * q @ k.transpose(-1, -2), -1)
    if p > 0:
        m = torch.empty(B * H * N * N, device=dev())
        ops.dropout_mask(m, m.numel(), p, 31, 8)
        m = m.view(B, H, N, N)
        # the keep bits saved by the forward are the same Philox decisions
        w = bits.view(B, H, N, N // 32).long() & 0xFFFFFFFF
        unpacked = ((w.unsqueeze(-1) >> torch.arange(32, device=dev())) & 1).reshape(B, H, N, N)
        assert torch.equal(unpacked.float(), m)
        P = P * m.double() / (1 - p)
    ref = (P @ v).transpose(1, 2).reshape(B * N, Cc)
    ref.backward(dO.double())
    assert rel_l2(O.cpu().numpy(), ref.detach().cpu().numpy()) < 1.5e-3
    lse_ref = torch.logsumexp(0.125 * q @ k.transpose(-1, -2), -1) / math.log(2.0)
    assert rel_l2(lse.view(B, H, N).cpu().numpy(), lse_ref.detach().cpu().numpy()) < 5e-4
    g = x.grad.view(B * N, 3, Cc)
    got = dqkv.view(B * N, 3, Cc)
    for i, nm in enumerate("qkv"):
        e = rel_l2(got[:, i].cpu().numpy(), g[:, i].cpu().numpy())
        assert e < 3e-3, f"d{nm}: {e:.3e}"


@pytest.mark.parametrize("B,N,G", [(2, 256, 0), (6, 512, 2), (1, 2048, 0)])
def test_attention_keepbits_and_premasked_forward(B, N, G):
    """Stand-alone keep-bit pass == the bits the fused forward stores; the forward that READS them gives the
    same O and lse bit for bit (also with several modules per launch: group_batches / group_site_stride)."""
    H, d, Cc, p = 8, 64, 512, 0.1
    qkv = _rand(B * N, 3 * Cc, seed=N + 7)
    O1, O2 = torch.empty(B * N, Cc, device=dev()), torch.empty(B * N, Cc, device=dev())
    l1, l2 = torch.empty(B * H, N, device=dev()), torch.empty(B * H, N, device=dev())
    b1 = torch.zeros(B * H, N, N // 32, dtype=torch.int32, device=dev())
    b2 = torch.zeros_like(b1)
    sd = torch.tensor([3], dtype=torch.int64, device=dev())
    ops.attention_fwd(qkv, O1, l1, b1, B, N, H, d, 0.125, p, seed=40, seed_dev=sd, site=5, group_batches=G, group_site_stride=8)
    ops.attention_keepbits(b2, B, N, H, p, 40, 5, seed_dev=sd, group_batches=G, group_site_stride=8)
    assert torch.equal(b1, b2)
    ops.attention_fwd_premasked(qkv, O2, l2, b2, B, N, H, d, 0.125, p)
    assert torch.equal(O1, O2) and torch.equal(l1, l2)


def test_fused_loss_and_jaccard_tail():
    """F4_TRAIN.py:58-71 in one pass: BCEWithLogits over all elements (+ gradient) and Jaccard2 of
    channel 0, against torch and against the stand-alone Jaccard2 kernel (bit-exact on {0,1} masks and
    hard predictions, within fp32 rounding on probabilities)."""
    from corrif_b200 import metrics
    g = torch.Generator(device="cpu").manual_seed(5)
    B, CH, H = 3, 3, 224
    masks = (torch.rand(B, 1, 1, H, H, generator=g) < 0.3).float().repeat(1, CH, 1, 1, 1).to(dev())
    for hard in (False, True):
        probs = torch.rand(B, CH, 1, H, H, generator=g)
        if hard:
            probs = (probs < 0.4).float()
        probs = probs.to(dev()).requires_grad_(True)
        ref_loss = torch.nn.functional.binary_cross_entropy_with_logits(probs, masks)
        ref_loss.backward()
        loss, dx, jac, pixels = metrics.loss_and_jaccard(probs, masks)
        assert pixels == B * H * H
        assert abs(loss.item() - ref_loss.item()) < 2e-6 * abs(ref_loss.item())
        assert rel_l2(dx.cpu().numpy(), probs.grad.cpu().numpy()) < 1e-6
        load = B * H * H
        j_ref = metrics.Jaccard2(masks[:, 0].reshape(load, 1), probs.detach()[:, 0].reshape(load, 1))
        if hard:
            assert jac.item() == j_ref.item()
        else:
            assert abs(jac.item() - j_ref.item()) < 1e-6
    # empty mask: the inversion branch of F5_JACCARD2.py:12-14
    z = torch.zeros(2, 3, 1, 8, 8, device=dev())
    p = (torch.rand(2, 3, 1, 8, 8, generator=g) < 0.5).float().to(dev())
    _, _, jac, _ = metrics.loss_and_jaccard(p, z, want_grad=False)
    assert jac.item() == metrics.Jaccard2(z[:, 0].reshape(-1, 1), p[:, 0].reshape(-1, 1)).item()
