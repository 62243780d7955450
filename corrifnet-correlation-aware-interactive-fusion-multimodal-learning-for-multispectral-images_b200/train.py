"""The data-parallel unit of the reference: the step body of ``train_model`` (F4_TRAIN.py:52-71),
without its two host synchronisations, plus batch-sharded data parallelism over NCCL.

Reference step:  zero_grad -> model(images) -> BCEWithLogitsLoss(outputs, masks) on the already
sigmoided outputs -> backward -> optim.step -> loss.item() [sync] -> Jaccard2(...) [sync inside the
``if y.sum(0)==0``].  Here loss and metric stay on the device (corrif_b200.metrics resolves the
empty-mask branch in the kernel) and are read once per epoch.

Data parallelism (new: the reference is single-device).  One process per GPU; a rank runs its own
micro-batches (the semantic micro-batch must not be split: train-mode BatchNorm and the inter_attn
batch-mixing view couple the samples of a batch) and gradients are averaged with
bucketed all-reduces launched from autograd hooks while the backward is still running.  Gradients
live in one flat buffer per bucket (``param.grad`` are views), so there is no copy-in/copy-out.
The parameters that never receive a gradient in MMVit4 (18 tensors: ``*_decode_conv``, ``seg_*``,
``fusion5``) are discovered on the first step and left out of the buckets.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import os

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .metrics import loss_and_jaccard


def _view_like(flat: torch.Tensor, off: int, p: torch.Tensor) -> torch.Tensor:
    """A view of ``flat[off : off + p.numel()]`` with p's shape AND strides.  The encoders' convolution weights live in
    channels_last_3d (cuDNN runs them on the volume layout); a plain ``view_as`` would hand autograd / cuDNN a
    contiguous-strided parameter and every (1,3,3) convolution would convert its weight on every call (measured: 177
    copy kernels per micro-batch step)."""
    seg = flat[off:off + p.numel()]
    if p.dim() == 5 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last_3d):
        return seg.as_strided(p.shape, p.stride())
    return seg.view_as(p) if p.is_contiguous() else seg.view(p.shape)


class GradBuckets:
    """Flat gradient storage + overlapped all-reduce.  Works with any backend (gloo on CPU for tests,
    NCCL over NVLink on the GPUs)."""

    ALIGN = 64      # elements

    def __init__(self, params: List[nn.Parameter], bucket_bytes: int = 64 << 20,
                 process_group=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes
        self.buckets: List[dict] = []
        self._built = False
        self._handles = []
        self._sync_now = True
        self._hooks = []

    # -- first step: plain autograd, then find out who actually got a gradient --------------------
    def build(self):
        live = [p for p in self.params if p.grad is not None]
        self.skipped = [p for p in self.params if p.grad is None]
        order = list(reversed(live))                       # backward produces the last layers first
        cur, cur_bytes = [], 0
        groups = []
        for p in order:
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= self.bucket_bytes:
                groups.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            groups.append(cur)
        for gi, grp in enumerate(groups):
            # every tensor starts on a 256-byte boundary (FlatAdam re-points the PARAMETERS into a buffer
            # of the same layout and the kernels need 16-byte aligned weights / biases); the gaps stay zero
            offs, n = [], 0
            for p in grp:
                offs.append(n)
                n += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            flat = torch.zeros(n, dtype=grp[0].dtype, device=grp[0].device)
            views = []
            for p, off in zip(grp, offs):
                view = _view_like(flat, off, p)
                view.copy_(p.grad)
                p.grad = view
                views.append(view)
            b = {"flat": flat, "params": grp, "offsets": offs, "views": views, "pending": len(grp), "index": gi,
                 "launched": False}
            self.buckets.append(b)
            for p in grp:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))
        self._built = True

    def _make_hook(self, bucket):
        """Runs after autograd has stored a parameter's gradient of the current micro-batch.  ``param.grad`` is None
        while a micro-batch is running (begin_micro_batch), so autograd ADOPTS the incoming gradient instead of
        launching one add kernel per parameter (~600 launches and ~7 ms of host time per micro-batch); when the
        bucket's last parameter has arrived its gradients are added into the flat buffer with ONE multi-tensor call,
        and on the step's last micro-batch the bucket's all-reduce starts right there, under the rest of the backward."""
        def hook(_param):
            bucket["pending"] -= 1
            if bucket["pending"] == 0:
                self._fold(bucket)
                if self._sync_now:
                    self._launch(bucket)
        return hook

    def _fold(self, bucket):
        got = [(v, p.grad) for v, p in zip(bucket["views"], bucket["params"]) if p.grad is not None and p.grad is not v]
        if got:
            torch._foreach_add_([v for v, _ in got], [g for _, g in got])
        for v, p in zip(bucket["views"], bucket["params"]):
            p.grad = v if self._sync_now else None          # after the last micro-batch .grad is the bucket view again

    def begin_micro_batch(self, last: bool):
        """Call before every micro-batch's forward: ``last`` = the step's final micro-batch (all-reduces start from
        its backward)."""
        self._sync_now = last
        if self._built:
            for b in self.buckets:
                b["pending"] = len(b["params"])
                for p in b["params"]:
                    p.grad = None

    def _launch(self, bucket):
        bucket["launched"] = True
        if self.world > 1:
            self._handles.append(dist.all_reduce(bucket["flat"], group=self.group, async_op=True))

    # -- per step ---------------------------------------------------------------------------------
    def zero(self):
        if self._built:
            for b in self.buckets:
                b["flat"].zero_()
                b["pending"] = len(b["params"])
                b["launched"] = False
            for p in self.skipped:
                p.grad = None
        else:
            for p in self.params:
                p.grad = None

    def set_sync(self, sync: bool):
        """Kept for callers of the first version: same as begin_micro_batch(last=sync)."""
        self.begin_micro_batch(sync)

    def finish(self, divisor: float = 1.0, total: Optional[float] = None):
        """Wait for the in-flight all-reduces and average.  ``divisor`` = local micro-batches of this step
        (gradients are then divided by world * divisor); ``total`` overrides that product with the number of
        micro-batches ALL ranks ran in this step (ragged last group of an epoch, where some ranks ran fewer or
        none).  Buckets whose hooks did not all fire on this rank (no local micro-batch, or a parameter without
        a gradient this step) are launched here, so every rank issues the same collectives in the same order.
        On the first step (buckets not built yet) do one blocking all-reduce per gradient, then build the
        buckets."""
        denom = float(total) if total is not None else self.world * divisor
        if not self._built:
            if any(p.grad is None for p in self.params) and all(p.grad is None for p in self.params):
                raise RuntimeError("GradBuckets: the first step needs at least one micro-batch on every rank")
            if self.world > 1:
                for p in self.params:
                    if p.grad is not None:
                        dist.all_reduce(p.grad, group=self.group)
            scale = 1.0 / denom
            if scale != 1.0:
                for p in self.params:
                    if p.grad is not None:
                        p.grad.mul_(scale)
            self.build()
            return
        for b in self.buckets:
            if not b["launched"]:
                if b["pending"] != 0:                      # a parameter got no gradient in the last micro-batch
                    self._sync_now = True
                    self._fold(b)
                self._launch(b)
        for h in self._handles:
            h.wait()
        self._handles.clear()
        scale = 1.0 / denom
        if scale != 1.0:
            for b in self.buckets:
                b["flat"].mul_(scale)

    def grad_bytes(self) -> int:
        return sum(b["flat"].numel() * b["flat"].element_size() for b in self.buckets)


class FlatAdam:
    """``torch.optim.Adam.step()`` of the reference (F2_MAIN.py:168-169, F4_TRAIN.py:62) as ONE kernel launch
    per gradient bucket.  The stock foreach implementation needs 13.9 ms per step for CorrIFNet's ~600
    parameter tensors (85 M elements: 2.4 GB of traffic, 0.4 ms at HBM speed); here parameters are re-pointed
    into flat buffers laid out like the gradient buckets, with flat exp_avg / exp_avg_sq beside them, and
    ``corrif_adam_step`` walks each bucket once.  The wrapped optimizer stays the owner of the hyper-parameters
    (LR schedulers keep working on its ``param_groups``) and its ``state`` is filled with views of the flat
    moments, so ``state_dict()`` is what stock Adam would hold.  Once FlatAdam has stepped, stepping the
    wrapped optimizer directly is not supported (its per-parameter ``step`` entries are one shared tensor)."""

    @staticmethod
    def eligible(optim) -> bool:
        if type(optim) is not torch.optim.Adam or len(optim.param_groups) != 1 or len(optim.state) != 0:
            return False
        g = optim.param_groups[0]
        return (g.get("weight_decay", 0) == 0 and not g.get("amsgrad", False) and not g.get("maximize", False)
                and not g.get("capturable", False) and not g.get("differentiable", False)
                and all(p.is_cuda and p.dtype == torch.float32 for p in g["params"]))

    def __init__(self, optim, buckets: "GradBuckets"):
        self.optim, self.buckets, self.t, self.slabs = optim, buckets, 0, None
        self.step_t = torch.zeros((), dtype=torch.float32)

    def _build(self):
        self.slabs = []
        for b in self.buckets.buckets:
            flat_p = torch.zeros_like(b["flat"])
            m, v = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
            for p, off in zip(b["params"], b["offsets"]):
                view = _view_like(flat_p, off, p)
                view.copy_(p.data)
                self.optim.state[p] = {"step": self.step_t, "exp_avg": _view_like(m, off, p),
                                       "exp_avg_sq": _view_like(v, off, p)}
                p.data = view
            self.slabs.append((flat_p, b["flat"], m, v))
        from . import module
        module._ENGINES.clear()              # engines are keyed by parameter addresses, which just moved

    def step(self):
        from . import ops
        if self.slabs is None:
            self._build()
        g = self.optim.param_groups[0]
        self.t += 1
        self.step_t.fill_(float(self.t))
        for flat_p, flat_g, m, v in self.slabs:
            ops.adam_step(flat_p, flat_g, m, v, flat_p.numel(), float(g["lr"]), float(g["betas"][0]),
                          float(g["betas"][1]), float(g["eps"]), 1.0, self.t)
        from . import volume
        volume.bump_weight_epoch()           # the weights changed behind autograd's version counters


def broadcast_module(model: nn.Module, src: int = 0, process_group=None, buffers_only: bool = False):
    """Make every rank hold rank ``src``'s parameters and buffers (as DDP does at construction);
    ``buffers_only`` re-synchronises just the buffers (BatchNorm running statistics are rank-local during
    training; evaluation and checkpoints use rank 0's)."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    tensors = list(model.buffers()) if buffers_only else list(model.parameters()) + list(model.buffers())
    for t in tensors:
        dist.broadcast(t.data, src=src, group=process_group)
    from . import volume
    volume.bump_weight_epoch()               # parameters were written through .data


def shard_micro_batches(n_batches: int, group: int, rank: int, world: int):
    """SURVEY.md section 8e partition.  The ``n_batches`` micro-batches of an epoch are consumed ``group`` at a
    time (one optimizer step per group, ``group`` a multiple of ``world``); inside a group rank r takes
    micro-batches r, r + world, ...  A micro-batch is never split (train-mode BatchNorm and the inter_attn
    batch-mixing view couple its samples).  Returns one entry per optimizer step:
    ``(indices of this rank, micro-batches of ALL ranks in the step)``; the last group may be ragged."""
    if group % world != 0:
        raise ValueError("micro-batches per step (%d) must be a multiple of the world size (%d)" % (group, world))
    steps = []
    for start in range(0, n_batches, group):
        members = list(range(start, min(start + group, n_batches)))
        steps.append((members[rank::world], len(members)))
    return steps


class TrainStep:
    """One optimisation step over ``len(micro_batches)`` local micro-batches (gradient accumulation)
    followed by the data-parallel gradient average and ``optim.step()``.

    Returns device tensors only: ``loss`` (mean over local micro-batches), ``jaccard_sum``
    (sum of Jaccard2 * batchLoad, F4_TRAIN.py:70-71) and ``pixels`` (sum of batchLoad)."""

    def __init__(self, model: nn.Module, optim: torch.optim.Optimizer, lim: int = 224,
                 jaccard_fn=None, process_group=None, bucket_bytes: int = 64 << 20, flat_adam: bool = True,
                 graphs: bool = False):
        self.model, self.optim, self.lim = model, optim, lim
        # graphs=True: after the first step (buckets and flat parameters in place) the model's forward and backward are
        # captured as two CUDA graphs (torch.cuda.make_graphed_callables) and replayed: ~2500 host-side launches per
        # micro-batch become two.  The step is host-bound without it.
        self.graphs, self._graph_shape, self._eager_forward = bool(graphs), None, None
        # The encoder trunks stay on cuDNN.  With static shapes (the graphs need them anyway) its
        # autotuner picks faster convolution kernels than the heuristics: -0.5 ms of a 54 ms micro-batch step, measured.
        # Tuning runs during the eager first step, before the capture.  CORRIF_CUDNN_BENCHMARK=0 leaves the flag alone.
        if self.graphs and os.environ.get("CORRIF_CUDNN_BENCHMARK", "1") != "0":
            torch.backends.cudnn.benchmark = True
        self.buckets = GradBuckets(list(model.parameters()), bucket_bytes, process_group)
        # a stock Adam with the reference's settings runs as one kernel per bucket (see FlatAdam)
        self.flat_adam = FlatAdam(optim, self.buckets) if flat_adam and FlatAdam.eligible(optim) else None
        # the default metric runs fused with the loss; a custom jaccard_fn keeps the two-step tail
        self.fused_tail = jaccard_fn is None
        if jaccard_fn is None:
            from .metrics import Jaccard2 as jaccard_fn
        self.jaccard_fn = jaccard_fn

    def _forward(self, images):
        if not self.graphs or not images.is_cuda:
            return self.model(images)
        if self._graph_shape is None and self.buckets._built and (self.flat_adam is None or self.flat_adam.slabs is not None):
            self._capture(images)
        if self._graph_shape == tuple(images.shape) and self.model.training:
            from . import ops
            ops._count(self._graph_launches)                 # kernels replayed by the two graphs of this micro-batch
            return self.model(images)                        # the graphed forward
        return (self._eager_forward or self.model)(images)   # another shape (ragged last batch) or eval mode

    def _capture(self, images):
        """Capture model forward + backward as CUDA graphs.  Needs static parameter addresses (hence after FlatAdam has
        re-pointed them) and nothing host-computed that changes per step: the fusion block's dropout seed moves to a
        device-resident counter (``model.device_seed``)."""
        import warnings
        eager = self.model.forward
        try:
            if hasattr(self.model, "device_seed"):
                self.model.device_seed = True
            torch.cuda.synchronize()
            from . import ops
            n0 = ops.launch_count()
            torch.cuda.make_graphed_callables(self.model, (images.detach().clone(),), allow_unused_input=True)
            # library kernels inside one forward + backward replay: make_graphed_callables ran 3 warm-up passes and
            # one capturing pass, each launching the same kernel sequence
            self._graph_launches = (ops.launch_count() - n0) // 4
            self._graph_shape, self._eager_forward = tuple(images.shape), eager
        except Exception as e:                               # stay on stream launches; never lose the step
            warnings.warn("TrainStep: CUDA-graph capture failed (%r); continuing with stream launches" % (e,))
            if hasattr(self.model, "device_seed"):
                self.model.device_seed = False
            self.model.forward = eager
            self.graphs, self._graph_shape = False, None

    def __call__(self, micro_batches, total_micro_batches: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """``micro_batches``: this rank's (images, masks) pairs for the step (a tuple = one micro-batch; may be
        empty on a rank that has no data in a ragged last group).  ``total_micro_batches``: how many micro-batches
        all ranks run in this step together (default world * len(micro_batches))."""
        if isinstance(micro_batches, tuple):
            micro_batches = [micro_batches]
        self.buckets.zero()
        n = len(micro_batches)
        if n == 0:
            self.buckets.finish(total=float(total_micro_batches))
            (self.flat_adam or self.optim).step()
            return {"loss": None, "jaccard_sum": None, "pixels": 0}
        loss_acc, jac_acc, pixels = None, None, 0
        for k, (images, masks) in enumerate(micro_batches):
            self.buckets.set_sync(k == n - 1)
            outputs = self._forward(images)
            if self.fused_tail:
                # loss (F4_TRAIN.py:58-60), its gradient and the Jaccard sums of :65-71 in one pass
                loss, dout, jac1, load = loss_and_jaccard(outputs, masks)
                outputs.backward(dout)                                       # :61
                jac = jac1 * load
            else:
                loss = F.binary_cross_entropy_with_logits(outputs, masks)    # F4_TRAIN.py:58-60
                loss.backward()                                              # :61
                with torch.no_grad():
                    load = masks.shape[0] * self.lim * self.lim              # :65
                    jac = self.jaccard_fn(masks[:, 0].reshape(load, 1), outputs.detach()[:, 0].reshape(load, 1)) * load
            with torch.no_grad():
                loss_acc = loss.detach() if loss_acc is None else loss_acc + loss.detach()
                jac_acc = jac if jac_acc is None else jac_acc + jac
                pixels += load
        self.buckets.finish(divisor=float(n), total=total_micro_batches)
        if self.flat_adam is not None:
            self.flat_adam.step()                                            # :62
        else:
            self.optim.step()                                                # :62
        return {"loss": loss_acc / n, "loss_sum": loss_acc, "micro_batches": n, "jaccard_sum": jac_acc, "pixels": pixels}
