"""torch.profiler table of one whole-model train step (batch 8, 256x256): where the time outside the
fusion block goes (encoders / early fusion / decoder on stock PyTorch) - input for the "next" rows N1/N2."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200", "dropin"))
import mmvit4  # noqa: E402
from corrif_b200 import train  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = os.environ.get("CORRIF_CUDNN_BENCHMARK") == "1"
torch.manual_seed(0)
model = mmvit4.MMVit4(num_cls=1).to(dev).train()
optim = torch.optim.Adam(model.parameters(), 1e-4)
step = train.TrainStep(model, optim, lim=224)
images = torch.randn(B, 3, 3, 256, 256, device=dev)
masks = (torch.rand(B, 1, 1, 224, 224, device=dev) < 0.3).float().repeat(1, 3, 1, 1, 1)
for _ in range(6):
    step((images, masks))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step((images, masks))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=35, max_name_column_width=60))

# per-launch CUDA-event times of the library's own kernels in one step (volume ops with their shapes)
from corrif_b200 import ops  # noqa: E402
with ops.profile() as rec:
    step((images, masks))
rows = rec.details()
agg = {}
for cls, detail, ms, work in rows:
    key = (cls, detail.split(" bytes=")[0])
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += ms; a[2] += work
print("\n# libcorrif_b200 launches of one micro-batch step (CUDA events per launch; serialised, so the sum exceeds the step)")
print("%-22s %-34s %6s %10s %12s" % ("kernel", "shape", "n", "ms", "TFLOP/s|GB/s"))
tot = 0.0
for (cls, detail), (n, ms, work) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("CORRIF_PROFILE_ROWS", "60"))]:
    rate = work / (ms * 1e-3) / (1e12 if cls.startswith(("conv3d", "gemm", "attn")) else 1e9) if ms > 0 else 0
    print("%-22s %-34s %6d %10.3f %12.1f" % (cls, detail[:34], n, ms, rate))
    tot += ms
print("total of listed: %.2f ms" % tot)
