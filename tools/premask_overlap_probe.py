"""How much does the backward slow down when next step's attention keep-bit passes (one block per SM) run
beside it on another stream?  (Feasibility probe for pre-generating the dropout decisions of step i+1 during
the backward of step i.)  Prints ms per fwd+bwd step without / with the concurrent passes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from corrif_b200 import fusion, module, ops  # noqa: E402

dev = torch.device("cuda:0")
B = 16
torch.manual_seed(0)
blk = module.CorrIFusionBlock(dropout_rate=0.1).to(dev)
named = dict(blk.named_parameters())
eng = fusion.FusionBlockEngine({n: named[n].detach() for n in fusion.param_names()}, dropout_p=0.1)
flat, grads = eng.new_grad_buffers()
x6 = [torch.randn(B, 64, 8, 8, 8, device=dev) for _ in range(3)]
fused, gout = torch.randn(B, 192, 8, 8, 8, device=dev), torch.randn(B, 192, 8, 8, 8, device=dev)
m1 = torch.zeros(3 * B * 8, 512, 16, dtype=torch.int32, device=dev)
m2 = torch.zeros(B * 8, 2048, 64, dtype=torch.int32, device=dev)
side = torch.cuda.Stream(dev)
sms = torch.cuda.get_device_properties(dev).multi_processor_count


def step(i, concurrent, blocks):
    eng.seed = i
    eng.forward(x6, fused)
    if concurrent:
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.attention_keepbits(m1, 3 * B, 512, 8, 0.1, i + 1, 0, group_batches=B, group_site_stride=8, max_blocks=blocks)
            ops.attention_keepbits(m2, B, 2048, 8, 0.1, i + 1, 24, max_blocks=blocks)
    flat.zero_()
    eng.backward(gout, grads)
    if concurrent:
        torch.cuda.current_stream().wait_stream(side)


for concurrent, blocks in ((False, 0), (True, sms), (True, 2 * sms), (True, 4 * sms), (False, 0)):
    for i in range(3):
        step(i, concurrent, blocks)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(10):
        step(10 + i, concurrent, blocks)
    b.record()
    torch.cuda.synchronize()
    print("concurrent keep-bit passes: %-5s blocks %4d   %.3f ms per step (stream launches, no graphs)" %
          (concurrent, blocks, a.elapsed_time(b) / 10))
