"""One weight-gradient launch of the tcgen05 kernel at a decoder shape (for ncu): python tools/wgrad_tc_one.py [cin cout n]"""
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
g = torch.randn(B, n, n, n, cout, device=dev)
dW = torch.zeros(cout, cin, 3, 3, 3, device=dev)
for _ in range(3):
    V.conv3d_wgrad([x], g, dW, 3, V.PAD_REPLICATE)
torch.cuda.synchronize()
