"""One launch each of the tcgen05 convolution kernels at the decoder's 128^3 shapes (for `ncu --set full`):
forward 32 -> 8 (cat(24, 8)), data gradient 8 -> 32, weight gradient 32 -> 8."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from corrif_b200 import volume as V  # noqa: E402

dev = torch.device("cuda:0")
B, n = 8, 128
xs = [torch.randn(B, n, n, n, 24, device=dev), torch.randn(B, n, n, n, 8, device=dev)]
w = torch.randn(8, 32, 3, 3, 3, device=dev) * 0.05
out = torch.empty(B, n, n, n, 8, device=dev)
g = torch.randn(B, n, n, n, 8, device=dev)
dx = torch.empty(B, n, n, n, 32, device=dev)
dW = torch.zeros_like(w)
stats = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64)
for _ in range(2):
    V.conv3d_forward_auto(xs, w, None, 8, 3, V.PAD_REPLICATE, True, out, stats)
    V.conv3d_dgrad(g, w, 32, 3, V.PAD_REPLICATE, dx)
    V.conv3d_wgrad(xs, g, dW, 3, V.PAD_REPLICATE)
torch.cuda.synchronize()
