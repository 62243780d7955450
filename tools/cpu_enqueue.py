"""How long does the HOST take to enqueue one fusion-block step (no sync) versus the GPU to run it?"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import fusion, module, ops  # noqa: E402

dev = torch.device("cuda:0")
B = 16
torch.manual_seed(0)
blk = module.CorrIFusionBlock(dropout_rate=0.1).to(dev)
named = dict(blk.named_parameters())
eng = fusion.FusionBlockEngine({n: named[n].detach() for n in fusion.param_names()}, dropout_p=0.1)
x6 = [torch.randn(B, 64, 8, 8, 8, device=dev) for _ in range(3)]
fused, gout = torch.randn(B, 192, 8, 8, 8, device=dev), torch.randn(B, 192, 8, 8, 8, device=dev)
flat, grads = eng.new_grad_buffers()
for i in range(3):
    eng.forward(x6, fused)
    eng.backward(gout, grads)
torch.cuda.synchronize()
n = 10
t0 = time.perf_counter()
for i in range(n):
    eng.forward(x6, fused)
    eng.backward(gout, grads)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.3f ms/step, until GPU done %.3f ms/step, launches/step %d" %
      ((t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3, ops.launch_count() // (n + 3)))
