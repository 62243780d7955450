"""Per-shape timing of the tcgen05 GEMM (CUDA events, 20 reps after warm-up, L2 flushed by a 256 MB
write between reps).  CORRIF_GEMM_V1=1 selects the non-persistent kernel for A/B comparison."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(64 << 20, device=dev)

SHAPES = [  # name, M, N, K, a_mn, b_mn, split
    ("linear mm qkv   ", 32768, 1536, 512, 0, 0, 1),
    ("linear mm proj  ", 32768, 512, 512, 0, 0, 1),
    ("linear intra qkv", 8192, 1536, 512, 0, 0, 1),
    ("linear intra fc ", 8192, 512, 512, 0, 0, 1),
    ("linear decode   ", 8192, 192, 2048, 0, 0, 1),
    ("linear encode   ", 8192, 512, 64, 0, 0, 1),
    ("dgrad mm qkv    ", 32768, 512, 1536, 0, 1, 1),
    ("dgrad mm fc     ", 32768, 512, 512, 0, 1, 1),
    ("dgrad intra qkv ", 8192, 512, 1536, 0, 1, 1),
    ("dgrad decode    ", 8192, 2048, 192, 0, 1, 1),
    ("wgrad mm qkv    ", 1536, 512, 32768, 1, 1, 0),
    ("wgrad mm fc     ", 512, 512, 32768, 1, 1, 0),
    ("wgrad intra qkv ", 1536, 512, 8192, 1, 1, 0),
    ("wgrad intra fc  ", 512, 512, 8192, 1, 1, 0),
    ("wgrad decode    ", 192, 2048, 8192, 1, 1, 0),
]


def split_for(M, N, K):
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    want = max(1, (2 * 148 + tiles - 1) // tiles)
    return max(1, min(want, max(1, (K // 32) // 4), 64))


for name, M, N, K, a_mn, b_mn, split in SHAPES:
    A = torch.randn(K, M, device=dev) if a_mn else torch.randn(M, K, device=dev)
    B = torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)
    D = torch.zeros(M, N, device=dev)
    sk = split_for(M, N, K) if split == 0 else 1
    kw = dict(M=M, N=N, K=K, lda=M if a_mn else K, ldb=N if b_mn else K, ldd=N, a_mn=bool(a_mn), b_mn=bool(b_mn),
              split_k=sk, epilogue=ops.EPI_ATOMIC_ADD if sk > 1 else ops.EPI_STORE)
    for _ in range(3):
        ops.gemm(A, B, D, **kw)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(A, B, D, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    print("%s M%6d N%5d K%6d split%3d  %8.1f us  %7.1f TFLOP/s" % (name, M, N, K, sk, ms * 1e3, 2.0 * M * N * K / ms / 1e9))
