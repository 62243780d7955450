"""Does the stock-PyTorch encoder (ResNet-50 inflated to (1,k,k) 3-D convs, mmvit4.py:113-212) run faster on this B200
in channels_last_3d memory format (cuDNN's native layout: no nchwToNhwc / nhwcToNchw kernels around every conv)?
Times fwd+bwd of ONE encoder at batch 8, 256^2, for {contiguous, channels_last_3d} x {cudnn.benchmark off, on}."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "corrifnet-correlation-aware-interactive-fusion-multimodal-learning-for-multispectral-images_b200", "dropin"))
import mmvit4  # noqa: E402

dev = torch.device("cuda:0")
res = {}
for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    for fmt_name, fmt in (("contiguous", torch.contiguous_format), ("channels_last_3d", torch.channels_last_3d)):
        torch.manual_seed(0)
        enc = mmvit4.Encoder().to(dev).to(memory_format=fmt).train()
        x = torch.randn(8, 1, 3, 256, 256, device=dev).to(memory_format=fmt)

        def step():
            for p in enc.parameters():
                p.grad = None
            outs = enc(x)
            sum(o.square().mean() for o in outs).backward()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        res["%s/benchmark=%s" % (fmt_name, bench)] = e0.elapsed_time(e1) / 5
        del enc, x
        torch.cuda.empty_cache()
print(json.dumps(res, indent=1))
