"""Correctness sweep of the tcgen05 GEMM variants against an fp64 torch product (operands pre-rounded
to TF32 so that only the accumulation order differs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def tf32(x):
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


CASES = [  # M, N, K, a_mn, b_mn, split
    (256, 256, 32, 0, 0, 1), (256, 256, 64, 0, 0, 1), (512, 512, 512, 0, 0, 1), (8192, 512, 512, 0, 0, 1),
    (8192, 1536, 512, 0, 0, 1), (8192, 192, 2048, 0, 0, 1), (8192, 512, 1536, 0, 1, 1), (8192, 2048, 192, 0, 1, 1),
    (1536, 512, 8192, 1, 1, 6), (512, 512, 8192, 1, 1, 19), (192, 2048, 8192, 1, 1, 10), (384, 320, 96, 1, 0, 1),
    (200, 260, 100, 0, 0, 1), (32768, 1536, 512, 0, 0, 1),
]
bad = 0
for M, N, K, a_mn, b_mn, split in CASES:
    A = tf32(torch.randn(K, M, device=dev)) if a_mn else tf32(torch.randn(M, K, device=dev))
    B = tf32(torch.randn(K, N, device=dev)) if b_mn else tf32(torch.randn(N, K, device=dev))
    D = torch.zeros(M, N, device=dev)
    ops.gemm(A, B, D, M=M, N=N, K=K, lda=M if a_mn else K, ldb=N if b_mn else K, ldd=N, a_mn=bool(a_mn),
             b_mn=bool(b_mn), split_k=split, epilogue=ops.EPI_ATOMIC_ADD if split > 1 else ops.EPI_STORE)
    torch.cuda.synchronize()
    Am = A.t() if a_mn else A
    Bm = B.t() if b_mn else B
    ref = Am.double() @ Bm.double().t()
    err = float((D.double() - ref).norm() / ref.norm())
    ok = err < 1e-5
    bad += not ok
    print("M%6d N%5d K%6d a%d b%d split%3d  rel err %.2e %s" % (M, N, K, a_mn, b_mn, split, err,
                                                                 "ok" if ok else "FAIL"), flush=True)
sys.exit(1 if bad else 0)
