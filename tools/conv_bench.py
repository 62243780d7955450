"""Per-layer timing of the channels-last convolution kernels at the decoder's shapes (batch 8): forward, data gradient
(+ border pass) and weight gradient, CUDA events over 5 launches each.  Prints TFLOP/s and the achieved fraction of
HBM bandwidth on the ALGORITHMIC bytes (every source read once + the result written once)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from corrif_b200 import volume as V  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = [((32,), 8, 3, 128), ((16,), 8, 3, 128), ((64,), 16, 3, 64), ((32,), 16, 3, 64), ((128,), 32, 3, 32), ((64,), 32, 3, 32),
          ((320,), 64, 3, 16), ((8,), 8, 1, 128)]
if len(sys.argv) > 1:
    SHAPES = SHAPES[:int(sys.argv[1])]
B = 8


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("%-26s %9s %9s %9s   (ms; TFLOP/s; GB/s on algorithmic bytes)" % ("layer", "fwd", "dgrad", "wgrad"))
for chans, cout, k, n in SHAPES:
    cin = sum(chans)
    xs = [torch.randn(B, n, n, n, c, device=dev) for c in chans]
    w = torch.randn(cout, cin, k, k, k, device=dev) * 0.05
    bias = torch.zeros(cout, device=dev)
    out = torch.empty(B, n, n, n, cout, device=dev)
    g = torch.randn(B, n, n, n, cout, device=dev)
    dx = torch.empty(B, n, n, n, cin, device=dev)
    dW = torch.zeros_like(w)
    stats = torch.zeros(B, cout, 2, device=dev, dtype=torch.float64)
    pad = V.PAD_REPLICATE
    t_f = timed(lambda: V.conv3d_forward_auto(xs, w, bias, cout, k, pad, True, out, stats))
    t_d = timed(lambda: V.conv3d_dgrad(g, w, cin, k, pad, dx))
    t_w = timed(lambda: V.conv3d_wgrad(xs, g, dW, k, pad))
    flop = 2.0 * B * n ** 3 * (27 if k == 3 else 1) * cin * cout
    byts = 4.0 * B * n ** 3 * (cin + cout)
    print("%-26s %9.3f %9.3f %9.3f   fwd %6.1f TF/s %6.0f GB/s | dgrad %6.1f TF/s | wgrad %6.1f TF/s" % (
        "k%d %d^3 %d->%d" % (k, n, cin, cout), t_f, t_d, t_w, flop / t_f / 1e9, byts / t_f / 1e6, flop / t_d / 1e9,
        flop / t_w / 1e9))
    del xs, out, g, dx
    torch.cuda.empty_cache()
