"""Stand-alone times of the keep-bit pass and of the forward that reads the bits, at the bench shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for B, N in ((16, 2048), (48, 512)):
    H, d, C = 8, 64, 512
    qkv = torch.randn(B * N, 3 * C, device=dev)
    O = torch.empty(B * N, C, device=dev)
    lse = torch.empty(B * H, N, device=dev)
    bits = torch.zeros(B * H, N, N // 32, dtype=torch.int32, device=dev)

    def t(fn, n=5):
        for _ in range(2):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3
    t_f = t(lambda: ops.attention_fwd(qkv, O, lse, bits, B, N, H, d, 0.125, 0.1, seed=1, site=8))
    t_k = t(lambda: ops.attention_keepbits(bits, B, N, H, 0.1, 1, 8))
    t_p = t(lambda: ops.attention_fwd_premasked(qkv, O, lse, bits, B, N, H, d, 0.125, 0.1))
    print("B%d N%d: fused forward %.1f us | keep-bit pass %.1f us + premasked forward %.1f us" % (B, N, t_f, t_k, t_p))
