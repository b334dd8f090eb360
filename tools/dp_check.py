#!/usr/bin/env python
"""2-GPU correctness check of the data-parallel trainer under NCCL (torchrun --nproc-per-node 2 tools/dp_check.py):
three optimiser steps of DataParallelTrainer (eager and CUDA-graph) on rank-specific batches; afterwards
  * the replicas are bit-identical across ranks (although every rank seeded its model differently),
  * the parameters equal (to bf16-noise tolerance) a single-process run on rank 0 that averages the two ranks' gradients by
    accumulating both batches (gradient_accumulation = 2 reproduces the mean of two equal-sized batches)."""
import contextlib, io, os, sys
from types import SimpleNamespace
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
T = rb.training
rank, local, world = T.init_distributed()
dev = torch.device("cuda", local)
tasks = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
patch = [32, 32, 32]


def build(seed):
    torch.manual_seed(seed)
    mgr = SimpleNamespace(tasks=tasks, train_patch_size=patch, train_batch_size=2, in_channels=1, vram_max=16.0,
                          autoconfigure=True, model_config={}, optimizer="SGD", initial_lr=0.02, weight_decay=0.0, max_epoch=10)
    with contextlib.redirect_stdout(io.StringIO()):
        return rb.NetworkFromConfig(mgr).to(dev), mgr


def batch(step, r):
    g = torch.Generator().manual_seed(100 * step + r)
    x = torch.rand(2, 1, *patch, generator=g)
    tg = {"sheet": (torch.rand(2, 1, *patch, generator=g) > 0.8).float(),
          "normals": torch.nn.functional.normalize(torch.randn(2, 3, *patch, generator=g), dim=1)}
    return x.to(dev), {k: v.to(dev) for k, v in tg.items()}


ok = True
for graph in (False, True):
    model, mgr = build(7 * rank)                       # different seeds: the trainer must broadcast rank 0's weights
    tr = T.DataParallelTrainer(model, mgr, use_cuda_graph=graph)
    losses = []
    for step in range(3):
        x, tg = batch(step, rank)
        total, _ = tr.train_step(x, tg)
        losses.append(float(total))
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    other = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(other, flat)
    same = all(torch.equal(other[0], o) for o in other)
    msg = f"[{'graph' if graph else 'eager'}] rank {rank}: losses {[round(v, 4) for v in losses]}, replicas identical: {same}"
    if rank == 0:
        ref_model, ref_mgr = build(0)
        ref_mgr.gradient_accumulation = 2
        # single process, no process group: both ranks' batches per update
        ref = T.DataParallelTrainer.__new__(T.DataParallelTrainer)
        T.DataParallelTrainer.__init__.__globals__["dist"]          # (same module)
        saved = dist.is_initialized
        try:
            dist.is_initialized = lambda: False
            ref.__init__(ref_model, ref_mgr, use_cuda_graph=False)
        finally:
            dist.is_initialized = saved
        for step in range(3):
            for r in range(world):
                x, tg = batch(step, r)
                ref.train_step(x, tg)
        a = torch.cat([p.detach().flatten() for p in ref_model.parameters()])
        rel = float((flat - a).norm() / a.norm())
        upd = float((flat - torch.cat([p.detach().flatten() for p in build(0)[0].parameters()])).norm() / a.norm())
        msg += f"; vs single-process mean-gradient run: rel-L2 of parameters {rel:.2e} (size of the 3 updates: {upd:.2e})"
        ok = ok and rel < 0.25 * upd
    ok = ok and same
    print(msg, flush=True)
    rb._lib.device_error_check()
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DP CHECK", "PASSED" if float(t) == 1.0 else "FAILED", flush=True)
dist.barrier()
os._exit(0 if float(t) == 1.0 else 1)
