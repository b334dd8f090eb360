"""Data-parallel training plumbing: bucketed gradient all-reduce overlapped with backward.

The reference trainer is single-process (train.py:131); this is the `torchrun` variant SURVEY 8(e)
describes.  One process per GPU, identical replicas, InstanceNorm has no cross-sample statistics, so
the only exchange is one mean all-reduce of the gradients per step:

  * unique parameters only (the state_dict aliases `conv`/`all_modules.0` and the per-decoder
    `encoder` copies are the same Parameter objects);
  * gradients live in a few flat fp32 buckets (each gradient is moved into its slot as autograd produces it and the
    slot becomes the parameter's `.grad`); a bucket is all-reduced on a side stream as soon as its last gradient
    has arrived, overlapping the remaining backward kernels;
  * bucket membership follows the order in which backward actually produces the gradients: the first
    backward pass records the arrival order, the next `zero_grad()` re-cuts the buckets in that order
    (heads -> decoders -> deep encoder -> shallow encoder), so every bucket completes - and launches -
    while backward is still running;
  * parameters that received no gradient in that pass (the unused deep-supervision heads,
    builders/decoder.py:128-131: 8 tensors per decoder) are taken out of the buckets and keep
    `.grad = None`, exactly what the single-GPU path (`zero_grad(set_to_none=True)`) gives the optimiser,
    so AdamW's weight decay treats them identically for any world size.  Should one of them receive a
    gradient later, it is reduced on its own in `finish()` and the buckets are re-cut at the next step;
  * `comm_dtype=torch.bfloat16` halves the NVLink traffic (942 -> 471 MB per step for the 128^3 network):
    the bucket is divided by the world size, rounded to bf16, summed by NCCL and widened back (the
    compression DistributedDataParallel's bf16 hook applies); the default keeps fp32 on the wire.
  * `broadcast_parameters()` makes the replicas identical to rank 0's at start-up / after a resume
    (parameters and buffers), so differently seeded ranks cannot silently train different models.
"""
from __future__ import annotations

import contextlib
from typing import List, Optional

import torch
import torch.distributed as dist


def broadcast_parameters(model: torch.nn.Module, src: int = 0, process_group=None):
    """Rank `src`'s parameters and buffers to every rank (a no-op outside torch.distributed / world 1)."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return 0
    n = 0
    seen = set()
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            if id(t) in seen:
                continue
            seen.add(id(t))
            dist.broadcast(t.data, src, group=process_group)
            n += 1
    return n


def parameter_checksum_mismatch(model: torch.nn.Module, process_group=None) -> bool:
    """True when the replicas differ: compares a cheap per-rank checksum (sum and sum of squares over all parameters)
    across ranks with a min / max all-reduce."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return False
    ps = [p for p in model.parameters()]
    dev = ps[0].device
    cs = torch.zeros(2, dtype=torch.float64, device=dev)
    with torch.no_grad():
        for p in ps:
            v = p.detach().double()
            cs[0] += v.sum()
            cs[1] += (v * v).sum()
    lo, hi = cs.clone(), cs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=process_group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=process_group)
    return bool((lo != hi).any())


def never_used_parameters(model: torch.nn.Module):
    """Parameters of the drop-in network that no forward pass touches: the deep-supervision heads
    `task_decoders.<t>.seg_layers[:-1]` of decoders built with deep_supervision=False (builders/decoder.py:128-131
    always builds them so that checkpoints stay loadable)."""
    out = []
    for m in model.modules():
        heads = getattr(m, "seg_layers", None)
        if heads is not None and hasattr(m, "deep_supervision") and not m.deep_supervision:
            for h in list(heads)[:-1]:
                out.extend(h.parameters())
    return out


class GradientBuckets:
    def __init__(self, model: torch.nn.Module, bucket_bytes: int = 64 << 20, process_group=None, average: bool = True,
                 comm_dtype: Optional[torch.dtype] = None, unused=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        self.bucket_bytes = int(bucket_bytes)
        self.comm_dtype = comm_dtype
        self._nccl_avg = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        params = [p for p in model.parameters() if p.requires_grad]      # .parameters() de-duplicates
        self.params = params
        dev = params[0].device
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._sync = True
        self._arrival: List[torch.nn.Parameter] = []      # parameters in the order the last backward produced them
        self._arrived = set()
        self._stray: List[torch.nn.Parameter] = []        # got a gradient although they sit in no bucket
        self._recut_pending = False
        self.rebuilds = 0
        self.launched_in_backward = 0                     # buckets whose all-reduce started before finish() (last step)
        self.launched_in_finish = 0
        self._in_finish = False
        # backward visits parameters roughly in reverse registration order: the initial cut, without the parameters
        # known (or declared through `unused`) never to receive a gradient
        skip = {id(p) for p in (never_used_parameters(model) if unused is None else unused)}
        self._cut([p for p in reversed(params) if id(p) not in skip])
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]

    # -- bucket layout -------------------------------------------------------------------------
    def _cut(self, order):
        """(Re)build the flat buckets for `order`; parameters outside `order` get `.grad = None`."""
        dev = self.params[0].device
        groups, cur, cur_bytes = [], [], 0
        for p in order:
            n = p.numel() * 4
            if cur and cur_bytes + n > self.bucket_bytes:
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += n
        if cur:
            groups.append(cur)
        self.buckets: List[dict] = []
        self._index = {}
        self._slots = {}
        for bi, ps in enumerate(groups):
            # every slot starts on a 16-byte boundary (4 fp32): a 1- or 3-element head bias would otherwise leave all
            # later slots of its bucket misaligned for the 16-byte vector paths of the optimiser kernels
            offs, off = [], 0
            for p in ps:
                offs.append(off)
                off += (p.numel() + 3) & ~3
            b = {"params": ps, "flat": torch.zeros(off, dtype=torch.float32, device=dev), "pending": len(ps),
                 "touched": 0, "work": None, "wire": None}
            for p, o in zip(ps, offs):
                self._slots[id(p)] = b["flat"][o:o + p.numel()].view_as(p)
                p.grad = None
                self._index[id(p)] = bi
            self.buckets.append(b)
        for p in self.params:
            if id(p) not in self._index:
                p.grad = None
        self._filled = set()
        self.bytes_per_step = sum(b["flat"].numel() * (2 if self.comm_dtype == torch.bfloat16 else 4) for b in self.buckets)

    @property
    def skipped(self):
        """Parameters that sit in no bucket (no gradient in the recorded backward pass)."""
        return [p for p in self.params if id(p) not in self._index]

    # -- per step ----------------------------------------------------------------------------
    def zero_grad(self):
        """Use instead of optimizer.zero_grad().  Bucketed parameters get `.grad = None`: the first gradient autograd
        produces for a parameter is MOVED into its bucket slot by the post-accumulate hook (one read + one write) instead
        of being added into a zero-filled bucket (zero-fill + read-modify-write: twice the HBM traffic, 0.94 GB of
        gradients per step); later micro-batches of a gradient-accumulation window add into the slot in place."""
        if self._recut_pending and self._arrival:
            self._cut(list(self._arrival))
            self._recut_pending = False
            self.rebuilds += 1
        for b in self.buckets:
            b["pending"] = len(b["params"])
            b["touched"] = 0
            b["work"] = None
        for p in self.params:
            p.grad = None
        self._arrival, self._arrived, self._stray = [], set(), []
        self.launched_in_backward = self.launched_in_finish = 0
        self._filled = set()

    def _reduce(self, flat):
        """mean all-reduce of one flat fp32 tensor on the current stream; returns the async work handle.  NCCL averages
        inside the collective (ncclAvg: no separate division pass over the 0.94 GB of gradients); other backends (gloo in
        the CPU tests) divide first."""
        op = dist.ReduceOp.SUM
        if self.average:
            if self._nccl_avg:
                op = dist.ReduceOp.AVG
            else:
                flat.div_(self.world)
        if self.comm_dtype is not None and self.comm_dtype != flat.dtype:
            wire = flat.to(self.comm_dtype)
            work = dist.all_reduce(wire, op=op, group=self.group, async_op=True)
            return work, wire
        return dist.all_reduce(flat, op=op, group=self.group, async_op=True), None

    def _launch(self, b):
        if self._in_finish:
            self.launched_in_finish += 1
        else:
            self.launched_in_backward += 1
        if self.world == 1:
            b["work"] = True
            return
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                b["work"], b["wire"] = self._reduce(b["flat"])
        else:
            b["work"], b["wire"] = self._reduce(b["flat"])

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
        if id(p) not in self._arrived:
            self._arrived.add(id(p))
            self._arrival.append(p)
        bi = self._index.get(id(p))
        if bi is not None and id(p) not in self._filled:
            # first gradient of this window: move it into the bucket slot and make the slot the parameter's .grad
            slot = self._slots[id(p)]
            if p.grad.data_ptr() != slot.data_ptr():
                slot.copy_(p.grad)
                p.grad = slot
            self._filled.add(id(p))
        if not self._sync:
            return
        if bi is None:
            self._stray.append(p)
            return
        b = self.buckets[bi]
        b["pending"] -= 1
        b["touched"] += 1
        if b["pending"] == 0:
            self._launch(b)

    def finish(self):
        """Call after loss.backward(): flushes partially filled buckets (identical on every rank because
        the replicas are identical) and makes the reduced gradients visible to the current stream."""
        self._in_finish = True
        try:
            for b in self.buckets:
                if b["work"] is None and b["pending"] > 0 and b["touched"] > 0:
                    for p in b["params"]:          # slots nobody filled this step hold last step's values: send zeros
                        if id(p) not in self._filled:
                            self._slots[id(p)].zero_()
                    self._launch(b)
        finally:
            self._in_finish = False
        strays = []
        if self._stray and self.world > 1:
            for p in self._stray:
                if self.comm_stream is not None:
                    self.comm_stream.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(self.comm_stream):
                        strays.append((p, *self._reduce(p.grad)))
                else:
                    strays.append((p, *self._reduce(p.grad)))
        if self.world > 1:
            cm = torch.cuda.stream(self.comm_stream) if self.comm_stream is not None else contextlib.nullcontext()
            for b in self.buckets:
                if b["work"] is not None:
                    b["work"].wait()
                    if b["wire"] is not None:
                        with cm:
                            b["flat"].copy_(b["wire"])
                        b["wire"] = None
            for p, work, wire in strays:
                work.wait()
                if wire is not None:
                    with cm:
                        p.grad.copy_(wire)
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        # the recorded arrival order differs from the current cut (first step, or a parameter (dis)appeared): re-cut
        # at the next zero_grad()
        if self._sync and self._arrival:
            cur = [p for b in self.buckets for p in b["params"]]
            if len(cur) != len(self._arrival) or any(a is not c for a, c in zip(cur, self._arrival)):
                self._recut_pending = True

    def remove(self):
        for h in self._hooks:
            h.remove()
