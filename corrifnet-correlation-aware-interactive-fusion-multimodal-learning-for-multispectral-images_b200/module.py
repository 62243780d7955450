"""torch-facing layer: the fusion block as a registered custom op with autograd, and an nn.Module
whose parameters carry the reference's state_dict keys (mmvit4.py:398-426)."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import fusion
from .fusion import FusionBlockEngine, param_names

_ENGINES: Dict[Tuple, FusionBlockEngine] = {}
# forward / backward of the op are captured into CUDA graphs once the same input buffers have been seen
# three times (a training loop with static staging buffers); CORRIF_NO_GRAPHS=1 keeps stream launches
import os as _os
USE_GRAPHS = _os.environ.get("CORRIF_NO_GRAPHS") is None


def _engine_for(params: Sequence[Tensor], dropout_p: float, precision: str) -> FusionBlockEngine:
    """One engine (workspace + saved activations) per parameter set; keyed by storage addresses so
    in-place optimizer updates keep hitting the same engine."""
    key = (tuple(p.data_ptr() for p in params), float(dropout_p), precision)
    eng = _ENGINES.get(key)
    if eng is None:
        if len(_ENGINES) > 8:
            _ENGINES.clear()
        named = {n: p.detach() for n, p in zip(param_names(), params)}
        eng = FusionBlockEngine(named, dropout_p=dropout_p, precision=precision, use_graphs=USE_GRAPHS)
        eng._flat_grads = eng.new_grad_buffers()          # persistent: graph replay needs fixed addresses
        eng.generation = 0                                # forward counter: the workspace holds ONE forward
        _ENGINES[key] = eng
    return eng


class StaleForwardError(RuntimeError):
    """corrif::fusion_block_backward was asked for the backward of a forward whose saved activations are gone."""


def default_base_seed() -> int:
    """Base of the dropout seeds: follows ``torch.manual_seed`` and differs between data-parallel ranks (the
    masks are a pure function of (seed, site, element), so equal seeds would mean equal masks on every rank)."""
    rank = int(_os.environ.get("RANK", "0"))
    return (torch.initial_seed() ^ (0x9E3779B97F4A7C15 * (rank + 1))) & 0x3FFFFFFFFFFFFFFF


@torch.library.custom_op("corrif::fusion_block", mutates_args=(), device_types="cuda")
def fusion_block_op(x6_rgb: Tensor, x6_nir: Tensor, x6_swir: Tensor, fused_x6: Tensor,
                    params: Sequence[Tensor], dropout_p: float, seed: int, precision: str) -> Tensor:
    """x6_inter = CorrIFNet fusion block (mmvit4.py:456-529).  ``params`` in ``param_names()`` order."""
    eng = _engine_for(params, dropout_p, precision)
    eng.set_seed(seed)
    if seed < 0:
        eng.advance_device_seed()        # device-resident seed counter: a fresh mask set per call / per graph replay
    eng.generation += 1
    out = eng.forward([x6_rgb.contiguous(), x6_nir.contiguous(), x6_swir.contiguous()],
                      fused_x6.contiguous())
    # the output is saved by the decoder's autograd nodes and must survive the next forward: one 0.4 MB/sample copy
    return out.clone()


@fusion_block_op.register_fake
def _(x6_rgb, x6_nir, x6_swir, fused_x6, params, dropout_p, seed, precision):
    return torch.empty_like(fused_x6)


@torch.library.custom_op("corrif::fusion_block_backward", mutates_args=(), device_types="cuda")
def fusion_block_backward_op(gout: Tensor, params: Sequence[Tensor], dropout_p: float, seed: int,
                             precision: str, generation: int) -> List[Tensor]:
    """Backward of the corrif::fusion_block call number ``generation`` on the same parameter set.  Returns
    [d x6_rgb, d x6_nir, d x6_swir, d fused_x6, flat parameter gradients (param_names() order)].

    The saved activations (and keep-bits) live in the engine's single workspace, so only the MOST RECENT forward
    can be differentiated; anything else (two forwards then two backwards, activation checkpointing, a second
    model call before ``.backward()``, a forward with another batch size in between) raises StaleForwardError
    instead of silently returning gradients of the wrong forward.

    The five results are fresh tensors (autograd may adopt a returned tensor as ``.grad`` without copying, which
    must never alias the workspace the next backward overwrites): 4 x 0.1-0.4 MB/sample of input gradients and
    the 41 MB flat parameter-gradient buffer - 13 us of copy at HBM speed per backward."""
    eng = _engine_for(params, dropout_p, precision)
    if generation != eng.generation or gout.shape[0] != getattr(eng, "_B", gout.shape[0]):
        raise StaleForwardError(
            "corrif::fusion_block_backward: forward #%d is no longer in the workspace (the latest forward on this "
            "parameter set is #%d, batch %s); run backward before the next forward of the same block"
            % (generation, eng.generation, getattr(eng, "_B", "?")))
    eng.set_seed(seed)
    flat, views = eng._flat_grads
    flat.zero_()
    dx6, dfused, _ = eng.backward(gout.contiguous(), views)
    return [dx6[0].clone(), dx6[1].clone(), dx6[2].clone(), dfused.clone(), flat.clone()]


@fusion_block_backward_op.register_fake
def _(gout, params, dropout_p, seed, precision, generation):
    b = gout.shape[0]
    x = gout.new_empty(b, 64, 8, 8, 8)
    return [x, x.clone(), x.clone(), torch.empty_like(gout),
            gout.new_empty(sum(p.numel() for p in params))]


def _setup_ctx(ctx, inputs, output):
    _, _, _, _, params, dropout_p, seed, precision = inputs
    ctx.params = list(params)
    ctx.dropout_p, ctx.seed, ctx.precision = dropout_p, seed, precision
    # setup_context runs right after the forward op, on the same thread: the engine's counter is this forward's id
    ctx.generation = _engine_for(params, dropout_p, precision).generation


def _backward(ctx, gout):
    res = torch.ops.corrif.fusion_block_backward(gout, ctx.params, ctx.dropout_p, ctx.seed, ctx.precision,
                                                 ctx.generation)
    pgrads, off = [], 0
    for p in ctx.params:          # slice the flat buffer: one memset + one kernel set wrote all of it
        pgrads.append(res[4][off:off + p.numel()].view_as(p))
        off += p.numel()
    return res[0], res[1], res[2], res[3], pgrads, None, None, None


fusion_block_op.register_autograd(_backward, setup_context=_setup_ctx)


# --------------------------------------------------------------------------------------------------
def _attach(root: nn.Module, dotted: str, param: nn.Parameter):
    """Register ``param`` under a dotted state_dict key, creating plain container modules."""
    parts = dotted.split(".")
    mod = root
    for part in parts[:-1]:
        child = mod._modules.get(part)
        if child is None:
            child = nn.Module()
            mod.add_module(part, child)
        mod = child
    mod.register_parameter(parts[-1], param)


def fusion_param_shapes() -> Dict[str, Tuple[int, ...]]:
    C, E, S = fusion.C, fusion.ENC, fusion.S
    shapes: Dict[str, Tuple[int, ...]] = {}
    for n in param_names():
        if n.endswith("_pos"):
            shapes[n] = (1, S, C)
        elif n.startswith("fused6_encode_conv"):
            shapes[n] = (C, 3 * E, 1, 1, 1) if n.endswith("weight") else (C,)
        elif "_encode_conv" in n:
            shapes[n] = (C, E, 1, 1, 1) if n.endswith("weight") else (C,)
        elif n.startswith("qkv_"):
            shapes[n] = (3 * C, C, 1, 1, 1) if n.endswith("weight") else (3 * C,)
        elif n.startswith("multimodal_decode_conv"):
            shapes[n] = (3 * E, 4 * C, 1, 1, 1) if n.endswith("weight") else (3 * E,)
        elif n.endswith("qkv.weight"):
            shapes[n] = (3 * C, C)
        elif n.endswith(".weight") and "norm" not in n:
            shapes[n] = (C, C)
        else:
            shapes[n] = (C,)
    return shapes


class CorrIFusionBlock(nn.Module):
    """Stand-alone fusion block.  ``state_dict()`` keys and shapes equal the corresponding entries
    of the reference ``MMVit4`` (strict loading both ways for this subset).  forward(x6 list,
    fused_x6) -> x6_inter; ``self.training`` selects dropout p=0.1 as in the reference
    (mmvit4.py:361)."""

    def __init__(self, dropout_rate: float = 0.1, precision: str = "tf32"):
        super().__init__()
        self.dropout_rate = dropout_rate
        self.precision = precision
        self._step = 0
        self.base_seed = default_base_seed()
        for name, shape in fusion_param_shapes().items():
            p = nn.Parameter(torch.zeros(shape))
            if name.endswith("norm.weight"):
                nn.init.ones_(p)
            elif name.endswith(".weight"):
                nn.init.kaiming_normal_(p) if p.dim() == 5 else nn.init.kaiming_uniform_(p, a=5 ** 0.5)
            _attach(self, name, p)

    def ordered_params(self) -> List[nn.Parameter]:
        named = dict(self.named_parameters())
        return [named[n] for n in param_names()]

    def forward(self, x6: Sequence[Tensor], fused_x6: Tensor) -> Tensor:
        p = self.dropout_rate if self.training else 0.0
        self._step += 1
        return torch.ops.corrif.fusion_block(x6[0], x6[1], x6[2], fused_x6, self.ordered_params(), p,
                                             self.base_seed + self._step, self.precision)
