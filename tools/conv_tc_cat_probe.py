import os, sys, torch
sys.path.insert(0, '/root/repo')
from corrif_b200 import volume as V
dev = torch.device("cuda:0"); B, n = 8, 128
xs = [torch.randn(B, n, n, n, 24, device=dev), torch.randn(B, n, n, n, 8, device=dev)]
w = torch.randn(8, 32, 3, 3, 3, device=dev) * 0.05
out = torch.empty(B, n, n, n, 8, device=dev); g = torch.randn(B, n, n, n, 8, device=dev)
dx = torch.empty(B, n, n, n, 32, device=dev); stats = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64)
def timed(fn, k=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print("fwd cat(24,8)->8 %.3f ms   dgrad 8->32 %.3f ms" % (
    timed(lambda: V.conv3d_forward_auto(xs, w, None, 8, 3, V.PAD_REPLICATE, True, out, stats)),
    timed(lambda: V.conv3d_dgrad(g, w, 32, 3, V.PAD_REPLICATE, dx))))
