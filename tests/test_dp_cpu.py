"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2 (no GPU, no kernels):
bucketed gradient averaging from autograd hooks, exclusion of never-used parameters, micro-batch
accumulation, parameter broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(6, 8)
        self.b = nn.Linear(8, 5)
        self.unused = nn.Linear(3, 3)          # never touched in forward: like MMVit4's *_decode_conv
        self.c = nn.Linear(5, 3 * 4)

    def forward(self, x):
        y = self.c(torch.relu(self.b(torch.relu(self.a(x)))))
        return torch.sigmoid(y).view(x.shape[0], 3, 1, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _soft_jaccard(y, yp, eps=1e-8):          # CPU stand-in with the reference's formula (no GPU here)
    tp = (yp * y).sum(0)
    return (tp + eps) / (tp + ((1 - yp) * y).sum(0) + ((1 - y) * yp).sum(0) + eps)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from corrif_b200.train import TrainStep, broadcast_module
    torch.manual_seed(100 + rank)              # different init per rank -> broadcast must fix it
    model = Tiny()
    broadcast_module(model)
    w0 = [p.detach().clone() for p in model.parameters()]
    optim = torch.optim.SGD(model.parameters(), lr=0.1)
    step = TrainStep(model, optim, lim=2, jaccard_fn=_soft_jaccard, bucket_bytes=256)
    g = torch.Generator().manual_seed(7)
    data = [(torch.randn(4, 6, generator=g), (torch.rand(4, 3, 1, 2, 2, generator=g) < 0.4).float())
            for _ in range(world * 2 * 3)]
    outs = []
    for it in range(3):                        # step 0 builds the buckets, 1-2 use the overlapped path
        mbs = [data[it * world * 2 + rank * 2 + k] for k in range(2)]      # 2 micro-batches per rank
        outs.append(step(mbs))
    ret[rank] = {"w0": w0, "w": [p.detach().clone() for p in model.parameters()],
                 "nb": len(step.buckets.buckets), "skipped": len(step.buckets.skipped),
                 "loss": [o["loss"].item() for o in outs], "data": data}
    dist.destroy_process_group()


def test_dp_world2_matches_single_process_reference():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    for a, b in zip(r0["w0"], r1["w0"]):
        assert torch.equal(a, b)                                   # broadcast worked
    for a, b in zip(r0["w"], r1["w"]):
        assert torch.allclose(a, b, atol=1e-7)                     # ranks stay in lock-step
    assert r0["nb"] > 1 and r0["skipped"] == 2                     # several buckets; unused.{weight,bias} left out
    # single-process oracle: mean gradient over the 4 micro-batches of each step
    model = Tiny()
    with torch.no_grad():
        for p, w in zip(model.parameters(), r0["w0"]):
            p.copy_(w)
    optim = torch.optim.SGD(model.parameters(), lr=0.1)
    data = r0["data"]
    for it in range(3):
        optim.zero_grad()
        for k in range(4):
            x, m = data[it * 4 + k]
            (nn.functional.binary_cross_entropy_with_logits(model(x), m) / 4).backward()
        optim.step()
    for p, w in zip(model.parameters(), r0["w"]):
        assert torch.allclose(p.detach(), w, atol=1e-6), (p - w).abs().max()


def _ragged_worker(rank, world, port, ret):
    """5 micro-batches, groups of 2 on 2 ranks: the last optimizer step has ONE micro-batch, on rank 0 only."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from corrif_b200.train import TrainStep, broadcast_module, shard_micro_batches
    torch.manual_seed(5)
    model = Tiny()
    broadcast_module(model)
    w0 = [p.detach().clone() for p in model.parameters()]
    step = TrainStep(model, torch.optim.SGD(model.parameters(), lr=0.1), lim=2, jaccard_fn=_soft_jaccard, bucket_bytes=256)
    g = torch.Generator().manual_seed(9)
    data = [(torch.randn(4, 6, generator=g), (torch.rand(4, 3, 1, 2, 2, generator=g) < 0.4).float()) for _ in range(5)]
    for mine, total in shard_micro_batches(len(data), world, rank, world):
        step([data[j] for j in mine], total_micro_batches=total)
    ret[rank] = {"w0": w0, "w": [p.detach().clone() for p in model.parameters()], "data": data}
    dist.destroy_process_group()


def test_ragged_last_group_keeps_ranks_in_lock_step():
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_ragged_worker, args=(world, port, ret), nprocs=world, join=True)
    for a, b in zip(ret[0]["w"], ret[1]["w"]):
        assert torch.allclose(a, b, atol=1e-7)
    model = Tiny()
    with torch.no_grad():
        for p, w in zip(model.parameters(), ret[0]["w0"]):
            p.copy_(w)
    optim = torch.optim.SGD(model.parameters(), lr=0.1)
    data = ret[0]["data"]
    for group in ([0, 1], [2, 3], [4]):                 # single-process reference: mean gradient over each group
        optim.zero_grad()
        for j in group:
            (nn.functional.binary_cross_entropy_with_logits(model(data[j][0]), data[j][1]) / len(group)).backward()
        optim.step()
    for p, w in zip(model.parameters(), ret[0]["w"]):
        assert torch.allclose(p.detach(), w, atol=1e-6), (p - w).abs().max()
