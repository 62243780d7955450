"""Kernel launch census of one micro-batch step of the drop-in model (stream launches): count, total and mean
GPU time per kernel name - the small-kernel overhead that CUDA graphs still pay (a few us each)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200", "dropin"))
import mmvit4  # noqa: E402
from corrif_b200 import train  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = mmvit4.MMVit4(num_cls=1).to(dev).train()
step = train.TrainStep(model, torch.optim.Adam(model.parameters(), 1e-4), lim=224)
images = torch.randn(8, 3, 3, 256, 256, device=dev)
masks = (torch.rand(8, 1, 1, 224, 224, device=dev) < 0.3).float().repeat(1, 3, 1, 1, 1)
for _ in range(4):
    step((images, masks))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step((images, masks))
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: -e.count)
print("total kernels: %d, total GPU time %.2f ms" % (sum(e.count for e in ev), sum(e.device_time_total for e in ev) / 1e3))
for e in ev[:28]:
    print("%5d  %9.1f us  %7.1f us/launch  %s" % (e.count, e.device_time_total, e.device_time_total / e.count, e.key[:120]))
# which ATen ops launch the copy kernels
ops_ = [e for e in prof.key_averages(group_by_input_shape=True) if e.key in ("aten::copy_", "aten::contiguous", "aten::clone", "aten::zeros", "aten::zero_", "aten::fill_")]
ops_.sort(key=lambda e: -e.count)
for e in ops_[:25]:
    print("%5d  %s  %s" % (e.count, e.key, str(e.input_shapes)[:140]))
