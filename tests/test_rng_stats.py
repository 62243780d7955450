"""The dropout counter RNG (csrc/common.cuh: dropout_key / dropout_bits) restated in numpy: statistical
sanity on the CPU, and bit-for-bit agreement of the masks the kernels export on the GPU."""
import numpy as np
import pytest

U = np.uint64
M32 = U(0xFFFFFFFF)


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> U(30))) * U(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U(27))) * U(0x94D049BB133111EB)
    return z ^ (z >> U(31))


def dropout_key(seed, site):
    with np.errstate(over="ignore"):
        return splitmix64(U(seed) ^ (U(site + 1) * U(0xD6E8FEB86659FD93)))


def dropout_bits(key, quad):
    key = U(key)
    c0 = quad & M32
    c1 = ((quad >> U(32)) ^ (key >> U(32))) & M32
    k = key & M32
    for _ in range(5):
        p = c0 * U(0xD256D193)
        c0 = ((p >> U(32)) ^ c1 ^ k) & M32
        c1 = p & M32
        k = (k + U(0x9E3779B9)) & M32
    return (c1 << U(32)) | c0


def keep_mask(seed, site, n, p):
    """Keep decisions of elements [0, n) (n % 4 == 0) for dropout probability p."""
    thresh = min(int(p * 65536.0 + 0.5), 65535)
    r = dropout_bits(dropout_key(seed, site), np.arange(n // 4, dtype=np.uint64))
    draws = np.stack([(r >> U(16 * i)) & U(0xFFFF) for i in range(4)], 1).reshape(-1)
    return draws >= U(thresh)


def test_rng_statistics():
    n = 1 << 23
    keep = keep_mask(1234, 3, n, 0.1).astype(np.float64)
    assert abs(keep.mean() - (1 - 6554 / 65536)) < 4 * np.sqrt(0.09 / n)
    x = keep - keep.mean()
    for lag in (1, 2, 3, 4, 8, 64, 512, 2048, 8192):
        c = float(np.mean(x[:-lag] * x[lag:]) / x.var())
        assert abs(c) < 5.0 / np.sqrt(n), (lag, c)
    rows = keep.reshape(-1, 2048).sum(1)                         # binomial row counts
    assert 0.93 < rows.var() / (2048 * 0.9 * 0.1) < 1.07
    other = keep_mask(1234, 4, n, 0.1).astype(np.float64)        # neighbouring site, same seed
    assert abs(np.corrcoef(keep, other)[0, 1]) < 5.0 / np.sqrt(n)
    nxt = keep_mask(1235, 3, n, 0.1).astype(np.float64)          # next step's seed
    assert abs(np.corrcoef(keep, nxt)[0, 1]) < 5.0 / np.sqrt(n)


@pytest.mark.gpu
def test_exported_masks_match_numpy_restatement():
    import torch
    from corrif_b200 import ops
    n, p = 1 << 16, 0.1
    for seed, site in ((99, 0), (7, 13)):
        m = torch.empty(n, device="cuda")
        ops.dropout_mask(m, n, p, seed, site)
        got = m.cpu().numpy() != 0
        assert np.array_equal(got, keep_mask(seed, site, n, p))
