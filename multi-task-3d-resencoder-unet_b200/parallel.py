"""Data-parallel training plumbing: bucketed gradient all-reduce overlapped with backward.

The reference trainer is single-process (train.py:131); this is the `torchrun` variant SURVEY 8(e)
describes.  One process per GPU, identical replicas, InstanceNorm has no cross-sample statistics, so
the only exchange is one mean all-reduce of the gradients per step:

  * unique parameters only (the state_dict aliases `conv`/`all_modules.0` and the per-decoder
    `encoder` copies are the same Parameter objects);
  * gradients live as views into a few flat fp32 buckets, filled in the order backward produces them
    (heads -> decoders -> deep encoder -> shallow encoder); a bucket is all-reduced on a side stream
    as soon as its last gradient has been accumulated, overlapping the remaining backward kernels;
  * parameters that never receive a gradient (the unused deep-supervision heads,
    builders/decoder.py:128-131) sit in a tail bucket that is skipped when nothing touched it.
"""
from __future__ import annotations

import contextlib
from typing import List

import torch
import torch.distributed as dist


class GradientBuckets:
    def __init__(self, model: torch.nn.Module, bucket_bytes: int = 64 << 20, process_group=None, average: bool = True):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        params = [p for p in model.parameters() if p.requires_grad]      # .parameters() de-duplicates
        self.params = params
        # backward visits parameters roughly in reverse registration order
        order = list(reversed(params))
        self.buckets: List[dict] = []
        cur, cur_bytes = [], 0
        for p in order:
            n = p.numel() * 4
            if cur and cur_bytes + n > bucket_bytes:
                self.buckets.append({"params": cur})
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += n
        if cur:
            self.buckets.append({"params": cur})
        dev = params[0].device
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._index = {}
        for bi, b in enumerate(self.buckets):
            total = sum(p.numel() for p in b["params"])
            b["flat"] = torch.zeros(total, dtype=torch.float32, device=dev)
            off = 0
            for p in b["params"]:
                p.grad = b["flat"][off:off + p.numel()].view_as(p)
                off += p.numel()
                self._index[id(p)] = bi
            b["pending"] = len(b["params"])
            b["touched"] = 0
            b["work"] = None
        self._sync = True
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.bytes_per_step = sum(b["flat"].numel() * 4 for b in self.buckets)

    # -- per step ----------------------------------------------------------------------------
    def zero_grad(self):
        """Keeps `.grad` as views into the flat buckets (use instead of optimizer.zero_grad())."""
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
            b["touched"] = 0
            b["work"] = None
        for p in self.params:
            bi = self._index[id(p)]
            if p.grad is None or p.grad.untyped_storage().data_ptr() != self.buckets[bi]["flat"].untyped_storage().data_ptr():
                raise RuntimeError("a gradient was re-allocated outside its bucket; call GradientBuckets.zero_grad(), "
                                   "not optimizer.zero_grad(set_to_none=True)")

    def _launch(self, b):
        if self.world == 1:
            return
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                if self.average:
                    b["flat"].div_(self.world)
                b["work"] = dist.all_reduce(b["flat"], group=self.group, async_op=True)
        else:
            if self.average:
                b["flat"].div_(self.world)
            b["work"] = dist.all_reduce(b["flat"], group=self.group, async_op=True)

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation: backward passes inside this context only add into the flat buckets; the
        all-reduce happens in the first backward outside it (like DistributedDataParallel.no_sync)."""
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev

    def _on_grad(self, p):
        if not self._sync:
            return
        b = self.buckets[self._index[id(p)]]
        b["pending"] -= 1
        b["touched"] += 1
        if b["pending"] == 0:
            self._launch(b)

    def finish(self):
        """Call after loss.backward(): flushes partially filled buckets (identical on every rank because
        the replicas are identical) and makes the reduced gradients visible to the current stream."""
        for b in self.buckets:
            if b["work"] is None and b["pending"] > 0 and b["touched"] > 0:
                self._launch(b)
        for b in self.buckets:
            if b["work"] is not None:
                b["work"].wait()
        if self.comm_stream is not None and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def remove(self):
        for h in self._hooks:
            h.remove()
