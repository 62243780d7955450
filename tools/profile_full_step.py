"""Profiling driver for the whole train step: a few warm-up micro-batch steps of the drop-in MMVit4 (batch 8, 256^2
tiles, TrainStep + Adam), then exactly one step between cudaProfilerStart/Stop.  Use with
  ncu --profile-from-start off -k regex:corrif --metrics gpu__time_duration.sum,dram__bytes_read.sum,... python tools/profile_full_step.py
`--conv CIN COUT K N` runs one stand-alone convolution block (forward + backward) instead, for the `ncu --set full`
capture of the dominant volume kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200", "dropin"))
os.environ.setdefault("CORRIF_NO_GRAPHS", "1")          # individual launches, so that every kernel is listed
dev = torch.device("cuda:0")

if len(sys.argv) > 1 and sys.argv[1] == "--conv":
    from corrif_b200 import volume as V
    cin, cout, k, n = (int(a) for a in sys.argv[2:6])
    B = 8
    x = torch.randn(B, n, n, n, cin, device=dev, requires_grad=True)
    w = (torch.randn(cout, cin, k, k, k, device=dev) * 0.05).requires_grad_(True)
    b = torch.zeros(cout, device=dev, requires_grad=True)
    go = torch.randn(B, n, n, n, cout, device=dev)
    for it in range(2):
        if it == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        V.conv_block([x], w, b, k, V.PAD_REPLICATE).backward(go)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)

import mmvit4  # noqa: E402
from corrif_b200 import ops, train  # noqa: E402

B = 8
torch.manual_seed(0)
model = mmvit4.MMVit4(num_cls=1).to(dev).train()
step = train.TrainStep(model, torch.optim.Adam(model.parameters(), 1e-4), lim=224)
images = torch.randn(B, 3, 3, 256, 256, device=dev)
masks = (torch.rand(B, 1, 1, 224, 224, device=dev) < 0.3).float().repeat(1, 3, 1, 1, 1)
for _ in range(3):
    step((images, masks))
torch.cuda.synchronize()
torch.cuda.profiler.start()
step((images, masks))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one micro-batch step; library launches so far:", ops.launch_count())
