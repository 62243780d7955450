"""Does a CUDA graph of the whole fusion-block step (fwd + bwd, 95 launches) beat stream launches?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import fusion, module  # noqa: E402

dev = torch.device("cuda:0")
B = 16
torch.manual_seed(0)
blk = module.CorrIFusionBlock(dropout_rate=0.1).to(dev)
named = dict(blk.named_parameters())
eng = fusion.FusionBlockEngine({n: named[n].detach() for n in fusion.param_names()}, dropout_p=0.1)
eng.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
x6 = [torch.randn(B, 64, 8, 8, 8, device=dev) for _ in range(3)]
fused, gout = torch.randn(B, 192, 8, 8, 8, device=dev), torch.randn(B, 192, 8, 8, 8, device=dev)
flat, grads = eng.new_grad_buffers()


def step():
    flat.zero_()
    eng.forward(x6, fused)
    eng.backward(gout, grads)
    eng.seed_dev += 1


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("stream launches: %.3f ms/step" % timeit(step))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
print("graph replay   : %.3f ms/step" % timeit(g.replay))
