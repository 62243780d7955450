"""One forward and one data-gradient launch of the tcgen05 line convolution at a decoder shape (for ncu):
python tools/conv_tc_one.py [cin cout n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from corrif_b200 import volume as V  # noqa: E402

cin, cout, n = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (32, 8, 128)
dev = torch.device("cuda:0")
B = 8
x = torch.randn(B, n, n, n, cin, device=dev)
w = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
out = torch.empty(B, n, n, n, cout, device=dev)
g = torch.randn(B, n, n, n, cout, device=dev)
dx = torch.empty(B, n, n, n, cin, device=dev)
stats = torch.zeros(B, cout, 2, device=dev, dtype=torch.float64)
for _ in range(2):
    V.conv3d_forward_auto([x], w, None, cout, 3, V.PAD_REPLICATE, True, out, stats)
    V.conv3d_dgrad(g, w, cin, 3, V.PAD_REPLICATE, dx)
torch.cuda.synchronize()
