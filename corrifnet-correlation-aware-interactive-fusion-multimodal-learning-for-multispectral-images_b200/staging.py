"""Host <-> device staging for the fusion block: pinned host buffers, a dedicated copy stream and
double-buffered device inputs, so that the H2D copy of step i+1 and the D2H read of step i-1 overlap
the kernels of step i (the reference does ``.to(device)`` on the compute stream, F4_TRAIN.py:55-56,
and reads the loss back synchronously, :64)."""
from __future__ import annotations

from typing import List, Sequence

import torch


class PinnedPipeline:
    """``get(host_tensors)`` returns device copies of pinned host tensors whose transfer was enqueued on
    the copy stream (ideally one step earlier through ``prefetch``); ``put(dev_tensor, host_out)``
    reads a result back without stalling the compute stream.  Buffers are re-used round-robin over
    ``depth`` slots; events keep a slot from being overwritten while a step still reads it."""

    def __init__(self, device: torch.device, depth: int = 2):
        self.dev = device
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=device)
        self._slots: List[dict] = [dict(bufs=None, ready=None, done=None, busy=False) for _ in range(depth)]
        self._cur = None
        self._next_fill = 0
        self._next_take = 0
        self._pending = 0
        self._out_events: List[torch.cuda.Event] = []

    def prefetch(self, host_tensors: Sequence[torch.Tensor]) -> None:
        """Enqueue the H2D copies of one step's inputs on the copy stream."""
        if self._pending >= self.depth:
            raise RuntimeError("PinnedPipeline: all staging slots are in flight")
        slot = self._slots[self._next_fill]
        if slot["busy"]:
            raise RuntimeError("PinnedPipeline: the step that took this slot has not called release() yet")
        # a ragged last batch (or any change of shape / dtype) gets buffers of its own shape: copy_() into the
        # first batch's buffers would raise for a smaller batch and silently BROADCAST a batch of one
        if slot["bufs"] is None or len(slot["bufs"]) != len(host_tensors) or any(
                d.shape != h.shape or d.dtype != h.dtype for d, h in zip(slot["bufs"], host_tensors)):
            if slot["done"] is not None:
                slot["done"].synchronize()               # the old buffers may still be read by a running step
            slot["bufs"] = [torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in host_tensors]
        with torch.cuda.stream(self.copy_stream):
            if slot["done"] is not None:                 # the step that last used this slot has finished
                self.copy_stream.wait_event(slot["done"])
            for d, h in zip(slot["bufs"], host_tensors):
                d.copy_(h, non_blocking=True)
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(self.copy_stream)
        self._next_fill = (self._next_fill + 1) % self.depth
        self._pending += 1

    def get(self) -> List[torch.Tensor]:
        """Device tensors of the oldest prefetched step (the compute stream waits for their copy)."""
        if self._pending == 0:
            raise RuntimeError("PinnedPipeline.get() without a prefetch")
        slot = self._slots[self._next_take]
        torch.cuda.current_stream(self.dev).wait_event(slot["ready"])
        if self._cur is not None and self._cur["busy"]:
            self.release()                               # the previous step forgot: its kernels are all enqueued by now
        slot["busy"] = True
        self._cur = slot
        self._next_take = (self._next_take + 1) % self.depth
        self._pending -= 1
        return slot["bufs"]

    def release(self) -> None:
        """Call after the step's last kernel has been enqueued: its input slot may be refilled."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self._cur["done"] = ev
        self._cur["busy"] = False

    def put(self, dev_tensor: torch.Tensor, host_out: torch.Tensor) -> None:
        """Read a result back into pinned memory on the copy stream."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ev)
            host_out.copy_(dev_tensor, non_blocking=True)
        dev_tensor.record_stream(self.copy_stream)

    def synchronize(self) -> None:
        self.copy_stream.synchronize()
