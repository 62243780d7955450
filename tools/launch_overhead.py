"""Fixed device-side cost of one GEMM launch: tiny problems launched back to back (the queue stays full,
so host launch latency is hidden) - what a kernel costs before it does any real work."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for name, M, N, K in (("pair 1 tile", 256, 256, 32), ("pair 74 tiles", 256 * 74, 256, 32), ("pair 74 tiles K=512", 256 * 74, 256, 512),
                      ("v2 single-CTA 1 tile", 128, 128, 32), ("v2 148 tiles", 128 * 148, 128, 32)):
    A, B, D = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.empty(M, N, device=dev)
    kw = dict(M=M, N=N, K=K, lda=K, ldb=K, ldd=N)
    for _ in range(20):
        ops.gemm(A, B, D, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 400
    e0.record()
    for _ in range(n):
        ops.gemm(A, B, D, **kw)
    e1.record()
    torch.cuda.synchronize()
    print("%-24s M%6d N%4d K%4d: %6.2f us per launch" % (name, M, N, K, e0.elapsed_time(e1) / n * 1e3))
x = torch.empty(1 << 20, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(400):
    ops.round_tf32(x, x, 1024)
e1.record()
torch.cuda.synchronize()
print("%-24s %6.2f us per launch" % ("trivial elementwise", e0.elapsed_time(e1) / 400 * 1e3))
