"""`torchrun`-aware training step around the drop-in network: the data-parallel variant of the
reference's single-GPU `BaseTrainer.train` loop body (train.py:122-254).

What is kept from the reference loop (train.py line numbers):
  * optimiser choice AdamW | SGD(momentum 0.9, nesterov) with `initial_lr` / `weight_decay` (:69-84),
    CosineAnnealingLR(T_max = max_epoch, eta_min = 0) stepped once per epoch (:87-91);
  * per-task loss x task weight, summed (:204-218); division by the accumulation count (:222); clip_grad_norm_(3)
    then optimiser step every `grad_accumulate_n` micro-batches (:226-230);
  * checkpoint dictionary {'model', 'optimizer', 'scheduler', 'epoch'} (:249-254); `_orig_mod.` prefixes written
    by torch.compile'd reference checkpoints are accepted on load (inference.py:37-44).
What differs, on purpose:
  * one process per GPU under torchrun (RANK / LOCAL_RANK / WORLD_SIZE), gradients mean-all-reduced in flat
    buckets overlapped with backward (`parallel.GradientBuckets`); clipping runs after the reduce, so every rank
    clips by the same norm and the replicas stay identical;
  * bf16 kernels with fp32 master weights: no autocast context and no GradScaler (the reference's fp16 autocast
    needs one, train.py:94-96);
  * optionally the whole micro-step (forward + losses + backward + all-reduce + clip + optimiser) is captured once
    as a CUDA graph and replayed: the ~700 kernel launches of a step are otherwise host bound on the 4^3 / 8^3 layers;
  * the sampler is sharded by rank (`shard_indices`); rank 0 alone writes checkpoints.
The data pipeline (zarr datasets, augmentation), TensorBoard and the debug GIFs of the reference trainer are not
part of the hot path and stay with the caller.
"""
from __future__ import annotations

import contextlib
import os
from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import ops
from .losses import build_task_losses, task_losses
from . import _lib as L
from .optim import ClippedAdamW
from .parallel import GradientBuckets, broadcast_parameters

CLIP_NORM = 3.0          # train.py:227


def init_distributed(backend: Optional[str] = None):
    """(rank, local_rank, world_size) from the torchrun environment; creates the default process group when
    WORLD_SIZE > 1 (NCCL on GPUs, gloo otherwise) and binds the process to its GPU."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if cuda else "gloo")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def shard_indices(indices: Sequence[int], rank: int, world_size: int, drop_last: bool = True) -> List[int]:
    """This rank's share of a (pre-shuffled, identical on every rank) index list: a strided slice, trimmed so that
    every rank runs the same number of steps (an unequal count would dead-lock the gradient all-reduce)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    idx = list(indices)
    if drop_last:
        idx = idx[:len(idx) - len(idx) % world_size]
    elif len(idx) % world_size:
        idx = idx + idx[:world_size - len(idx) % world_size]       # wrap around, like DistributedSampler
    return idx[rank::world_size]


def strip_compile_prefix(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Checkpoints written by the reference carry torch.compile's `_orig_mod.` prefix (train.py:133,250)."""
    return {(k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k): v for k, v in state_dict.items()}


class DataParallelTrainer:
    """One optimisation step of the multi-task network, replicated over the ranks of a torchrun job.

    `mgr` is the reference's config manager (or any object with `tasks`, and optionally `optimizer`, `initial_lr`,
    `weight_decay`, `max_epoch`, `gradient_accumulation`); `model` is `NetworkFromConfig(mgr)` already on its GPU.
    """

    def __init__(self, model: torch.nn.Module, mgr, use_cuda_graph: bool = False, fused_losses: bool = True,
                 process_group=None, grad_comm_dtype: Optional[torch.dtype] = None, fused_optimizer: bool = True,
                 manage_packs: bool = False):
        self.model = model
        self.group = process_group
        self.mgr = mgr
        self.tasks = mgr.tasks
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.accumulate = max(1, int(getattr(mgr, "gradient_accumulation", 1) or 1))
        self.use_graph = bool(use_cuda_graph)
        if self.use_graph and self.accumulate != 1:
            raise NotImplementedError("whole-step CUDA graph capture with gradient accumulation > 1")
        # tasks that name a `loss_fn` follow the reference's table (train.py:45-66); without one the BASELINE pairing
        # applies (MaskedCosineLoss for a 3-channel "normals" task, BCEDiceLoss(0.5, 0.5) otherwise)
        if any("loss_fn" in info for info in self.tasks.values()):
            self.criteria = build_task_losses(self.tasks, fused=fused_losses)
        else:
            self.criteria = task_losses(self.tasks, fused=fused_losses)
        self.weights = {t: float(info.get("weight", 1.0)) for t, info in self.tasks.items()}
        lr = float(getattr(mgr, "initial_lr", 1e-3))
        wd = float(getattr(mgr, "weight_decay", 1e-4))
        # replicas must start identical whatever each rank's seed was: rank 0's parameters and buffers win
        if self.world > 1:
            broadcast_parameters(model, 0, process_group)
        dev = next(model.parameters()).device
        on_gpu = dev.type == "cuda"
        self._is_sgd = getattr(mgr, "optimizer", "AdamW") == "SGD"
        if self._is_sgd:
            # SGD bakes a Python-float lr into the captured kernels: the graph is re-captured whenever the schedule moves
            # the lr (end_epoch), see _lr_signature
            self.optimizer = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9, nesterov=True, weight_decay=wd)
            self._clip_in_step = False
        else:
            if self.use_graph and on_gpu:
                lr = torch.tensor(lr, dtype=torch.float32, device=dev)   # a tensor lr stays adjustable after capture
            if fused_optimizer and on_gpu:
                # clip_grad_norm_(3) + AdamW as two multi-tensor passes (optim.ClippedAdamW; same state_dict format)
                self.optimizer = ClippedAdamW(model.parameters(), lr=lr, weight_decay=wd, max_grad_norm=CLIP_NORM,
                                              manage_packs=bool(manage_packs))
            else:
                self.optimizer = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=wd, fused=on_gpu,
                                                   capturable=self.use_graph and on_gpu)
        self._clip_in_step = isinstance(self.optimizer, ClippedAdamW)
        self._lr_tensors = [g["lr"] for g in self.optimizer.param_groups if torch.is_tensor(g["lr"])]
        self.scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(self.optimizer, T_max=int(getattr(mgr, "max_epoch", 1000)),
                                                                    eta_min=0)
        self.buckets = (GradientBuckets(model, process_group=process_group, comm_dtype=grad_comm_dtype)
                        if self.world > 1 else None)
        self.params = list(model.parameters())
        self._micro = 0
        self._graph = None
        self.epoch = 0

    # -- one micro-batch -------------------------------------------------------------------------
    def _zero_grad(self):
        if self.buckets is not None:
            self.buckets.zero_grad()
        else:
            self.optimizer.zero_grad(set_to_none=True)

    def _loss(self, outputs, targets):
        total, per = 0.0, {}
        for t, gt in targets.items():
            per[t] = self.criteria[t](outputs[t], gt) * self.weights[t]
            total = total + per[t]
        return total, per

    def _micro_step(self, inputs, targets, do_update: bool):
        if self._micro == 0:
            self._zero_grad()
        outputs = self.model(inputs)
        total, per = self._loss(outputs, targets)
        # micro-batches before the update only accumulate into the flat buckets; the all-reduce rides the last backward
        hold = self.buckets.no_sync() if (self.buckets is not None and not do_update) else contextlib.nullcontext()
        with hold:
            (total / self.accumulate).backward()
        self._micro += 1
        if do_update:
            if self.buckets is not None:
                self.buckets.finish()
            if not self._clip_in_step:
                torch.nn.utils.clip_grad_norm_([p for p in self.params if p.grad is not None], CLIP_NORM)
            self.optimizer.step()
            self._micro = 0
        return total, per

    def train_step(self, inputs: torch.Tensor, targets: Dict[str, torch.Tensor], last_in_epoch: bool = False):
        """Forward, weighted multi-task loss, backward; every `gradient_accumulation`-th call (or when
        `last_in_epoch`) the gradients are all-reduced, clipped and applied.  Returns (total_loss, {task: loss})
        as device tensors (no host synchronisation).  Under data parallelism the accumulated gradient is reduced
        once, during the backward pass of the updating micro-batch."""
        self.model.train()
        if self.use_graph:
            return self._graphed_step(inputs, targets)
        do_update = (self._micro + 1) % self.accumulate == 0 or last_in_epoch
        return self._micro_step(inputs, targets, do_update)

    # -- whole-step CUDA graph ------------------------------------------------------------------------
    def _lr_signature(self):
        """Learning rates that are Python floats are constants of a captured graph; device-tensor rates are read by
        every replay.  The signature of the former decides whether a captured step is still valid."""
        return tuple(float(g["lr"]) for g in self.optimizer.param_groups if not torch.is_tensor(g["lr"]))

    def _graphed_step(self, inputs, targets):
        if self._graph is not None and self._graph_lr != self._lr_signature():
            self._graph = None                     # the schedule moved a baked-in lr (SGD): capture again
        if self._graph is None:
            self._capture(inputs, targets)
            self._graph_lr = self._lr_signature()
        self._static_in.copy_(inputs, non_blocking=True)
        for k, v in targets.items():
            self._static_tg[k].copy_(v, non_blocking=True)
        self._graph[0].replay()
        # a replay updates the parameters without running any Python (no Tensor._version bump, no optimiser hook):
        # packs cached by an earlier eager / eval forward must not survive it
        ops.invalidate_weight_packs()
        return self._graph[1]

    def _capture(self, inputs, targets):
        """Warm up (optimiser state allocation, kernel attributes, NCCL channels) and capture one whole step on static
        input buffers, then put parameters and optimiser state back to where they were: capturing costs no update."""
        self._static_in = inputs.detach().clone()
        self._static_tg = {k: v.detach().clone() for k, v in targets.items()}
        snap_p = [p.detach().clone() for p in self.params]
        snap_o = {id(p): {k: v.clone() for k, v in st.items() if torch.is_tensor(v)}
                  for p, st in self.optimizer.state.items()}
        ops.PACK_CACHE = False            # the weight (re)packing kernels must be part of every replay
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._micro_step(self._static_in, self._static_tg, True)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._micro_step(self._static_in, self._static_tg, True)
            self._graph = (g, out)
        finally:
            ops.PACK_CACHE = True
        with torch.no_grad():
            for p, s_ in zip(self.params, snap_p):
                p.copy_(s_)
            for p, st in self.optimizer.state.items():
                old = snap_o.get(id(p), {})
                for k, v in st.items():
                    if torch.is_tensor(v):
                        v.copy_(old[k]) if k in old else v.zero_()     # zero == freshly initialised AdamW / SGD state
        if self._clip_in_step:
            self.optimizer.refresh_packs()     # the captured step reads the optimiser-managed operand packs

    # -- epoch bookkeeping ------------------------------------------------------------------------
    def _restore_device_lr(self):
        """`optimizer.load_state_dict` (and a scheduler stepping a float) can replace the device-tensor lr of a
        capturable optimiser by a float or a CPU tensor, which a re-captured graph would bake in as a constant: put the
        loaded value back into the original device tensors."""
        if not self._lr_tensors:
            return
        for g, t in zip(self.optimizer.param_groups, self._lr_tensors):
            if g["lr"] is not t:
                t.fill_(float(g["lr"]))
                g["lr"] = t

    def check_device(self):
        """Raise if any kernel recorded a pipeline time-out since the last check (synchronises the current stream)."""
        if next(self.model.parameters()).is_cuda:
            L.device_error_check()

    def end_epoch(self):
        self.check_device()
        self.scheduler.step()
        self._restore_device_lr()
        self.epoch += 1

    def state_dict(self):
        """The reference's checkpoint dictionary; 'epoch' is the 0-based index of the last finished epoch
        (train.py:249-254 saves the loop variable, :164 resumes at checkpoint['epoch'] + 1)."""
        return {"model": self.model.state_dict(), "optimizer": self.optimizer.state_dict(),
                "scheduler": self.scheduler.state_dict(), "epoch": self.epoch - 1}

    def save_checkpoint(self, path: str, keep_newest: Optional[int] = None) -> bool:
        """Rank 0 writes the reference's checkpoint dictionary (train.py:249-254) atomically; other ranks only
        synchronise.  `keep_newest` = N prunes older `<model_name>_*.pth` siblings like train.py:256-265 (N = 10 there)."""
        wrote = False
        self.check_device()           # never write weights that a timed-out kernel may have corrupted
        if self.rank == 0:
            tmp = f"{path}.tmp"
            torch.save(self.state_dict(), tmp)
            os.replace(tmp, path)
            wrote = True
            if keep_newest:
                prune_checkpoints(path, keep_newest)
        if self.world > 1:
            dist.barrier()
        return wrote

    def load_checkpoint(self, path: str, strict: bool = True, weights_only: bool = False):
        ck = torch.load(path, map_location="cpu", weights_only=False)
        self.model.load_state_dict(strip_compile_prefix(ck["model"]), strict=strict)
        if not weights_only:
            if "optimizer" in ck:
                self.optimizer.load_state_dict(ck["optimizer"])
            if "scheduler" in ck:
                self.scheduler.load_state_dict(ck["scheduler"])
            self.epoch = int(ck.get("epoch", -1)) + 1          # the epoch to run next (train.py:164)
            self._restore_device_lr()
        if self.world > 1:
            broadcast_parameters(self.model, 0, self.group)    # every rank read the file; rank 0's copy is authoritative
        ops.invalidate_weight_packs()
        self._graph = None
        return ck


def prune_checkpoints(latest_path: str, keep: int = 10):
    """Keep the `keep` newest `<stem>_*.pth` files next to `latest_path` (train.py:256-265: sorted by mtime)."""
    import glob
    import re
    d, base = os.path.split(os.path.abspath(latest_path))
    m = re.match(r"(.*)_\d+\.pth$", base)
    if not m:
        return []
    files = sorted(glob.glob(os.path.join(d, glob.escape(m.group(1)) + "_*.pth")), key=os.path.getmtime)
    removed = []
    while len(files) > keep:
        oldest = files.pop(0)
        os.unlink(oldest)
        removed.append(oldest)
    return removed


def iterate_sharded(dataset_len: int, epoch: int, rank: int, world_size: int, seed: int = 0) -> Iterable[int]:
    """Indices of this rank for `epoch`: one permutation shared by all ranks (seeded by epoch), strided by rank."""
    g = torch.Generator().manual_seed(seed + epoch)
    perm = torch.randperm(dataset_len, generator=g).tolist()
    return shard_indices(perm, rank, world_size)
