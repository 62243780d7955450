"""Per-launch timing of one bench step (B=16, dropout 0.1): every library launch with its shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import fusion, module, ops  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
PD = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
torch.manual_seed(0)
blk = module.CorrIFusionBlock(dropout_rate=PD).to(dev)
named = dict(blk.named_parameters())
eng = fusion.FusionBlockEngine({n: named[n].detach() for n in fusion.param_names()}, dropout_p=PD)
x6 = [torch.randn(B, 64, 8, 8, 8, device=dev) for _ in range(3)]
fused, gout = torch.randn(B, 192, 8, 8, 8, device=dev), torch.randn(B, 192, 8, 8, 8, device=dev)
for i in range(3):
    eng.seed = i
    eng.forward(x6, fused)
    eng.backward(gout)
with ops.profile() as rec:
    eng.seed = 9
    eng.forward(x6, fused)
    eng.backward(gout)
tot = 0.0
agg = {}
for cls, det, ms, work in rec.details():
    tot += ms
    key = (cls, det)
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += ms; a[2] += work
print("total %.3f ms" % tot)
for (cls, det), (n, ms, work) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    rate = work / (ms * 1e-3) / 1e12 if cls.startswith(("gemm", "attn")) else work / (ms * 1e-3) / 1e9
    print("%-22s %-44s x%2d %8.1f us  %8.1f %s" % (cls, det, n, ms * 1e3, rate, "TF/s" if cls.startswith(("gemm", "attn")) else "GB/s"))
