"""Event log of one mid-grid CTA of the attention backward (CORRIF_ATTN_TIMING=1) at a bench shape.
Needs a library built with the log compiled in:
  CORRIF_NVCC_EXTRA=-DCORRIF_ATTN_EVLOG python corrifnet-*_b200/build.py --force"""
import os
import sys

os.environ["CORRIF_ATTN_TIMING"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
shapes = ((16, 2048),) if len(sys.argv) < 2 else ((int(sys.argv[1]), int(sys.argv[2])),)
for B, N in shapes:
    H, d, C = 8, 64, 512
    qkv = torch.randn(B * N, 3 * C, device=dev)
    dO = torch.randn(B * N, C, device=dev)
    O = torch.empty(B * N, C, device=dev)
    lse = torch.empty(B * H, N, device=dev)
    delta = torch.empty(B * H, N, device=dev)
    bits = torch.zeros(B * H, N, N // 32, dtype=torch.int32, device=dev)
    dqkv = torch.empty(B * N, 3 * C, device=dev)
    for it in range(2):
        ops.attention_fwd(qkv, O, lse, bits, B, N, H, d, 0.125, 0.1, seed=it, site=8)
        ops.attention_bwd(qkv, O, dO, lse, bits, delta, dqkv, B, N, H, d, 0.125, 0.1)
        torch.cuda.synchronize()
