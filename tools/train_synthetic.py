#!/usr/bin/env python
"""Synthetic-data training loop through the public trainer API (`training.DataParallelTrainer`), the torchrun variant
of the reference's `BaseTrainer.train` loop body (train.py:195-254).

    python tools/train_synthetic.py --patch 128 --batch 2 --steps 20
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_synthetic.py --steps 20 --graph

Every rank draws its own synthetic batches (sheet + normals targets); rank 0 prints the per-task losses
(one host sync per `--log-every` steps, not per step as the reference's `.item()` calls do) and writes a checkpoint in
the reference's format at the end.
"""
import argparse
import contextlib
import io
import os
import sys
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb   # noqa: E402

TASKS = {"sheet": {"channels": 1, "activation": "sigmoid", "weight": 1.0},
         "normals": {"channels": 3, "activation": "none", "weight": 1.0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patch", type=int, default=128)
    ap.add_argument("--batch", type=int, default=2, help="per GPU")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--graph", action="store_true", help="replay the whole step as one CUDA graph")
    ap.add_argument("--log-every", type=int, default=5)
    ap.add_argument("--stochastic-depth", type=float, default=0.0)
    ap.add_argument("--checkpoint", default=None)
    args = ap.parse_args()
    T = rb.training
    rank, local, world = T.init_distributed()
    dev = torch.device("cuda", local)
    P, B = args.patch, args.batch
    mgr = SimpleNamespace(tasks=TASKS, train_patch_size=[P] * 3, train_batch_size=B, in_channels=1, vram_max=16.0,
                          autoconfigure=True, model_config={"stochastic_depth_p": args.stochastic_depth},
                          optimizer="AdamW", initial_lr=1e-3, weight_decay=1e-4, max_epoch=100, model_name="Synthetic")
    torch.manual_seed(0)                      # identical replicas
    with contextlib.redirect_stdout(io.StringIO()):
        model = rb.NetworkFromConfig(mgr).to(dev)
    trainer = T.DataParallelTrainer(model, mgr, use_cuda_graph=args.graph)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    t0 = None
    for step in range(args.steps):
        x = torch.rand(B, 1, P, P, P, device=dev, generator=gen)
        tgt = {"sheet": (torch.rand(B, 1, P, P, P, device=dev, generator=gen) > 0.8).float(),
               "normals": torch.nn.functional.normalize(torch.randn(B, 3, P, P, P, device=dev, generator=gen), dim=1)}
        total, per = trainer.train_step(x, tgt)
        if step == 2:
            torch.cuda.synchronize()
            t0 = time.perf_counter()          # after warm-up / capture
        if rank == 0 and (step + 1) % args.log_every == 0:
            print(f"step {step + 1}: total {float(total):.4f} " + " ".join(f"{k} {float(v):.4f}" for k, v in per.items()),
                  flush=True)
    torch.cuda.synchronize()
    if rank == 0 and t0 is not None and args.steps > 3:
        dt = (time.perf_counter() - t0) / (args.steps - 3)
        print(f"{dt * 1e3:.1f} ms per step including batch synthesis, {world * B * P ** 3 / dt / 1e6:.1f} M voxels/s on {world} GPU(s)")
    trainer.end_epoch()
    if args.checkpoint:
        trainer.save_checkpoint(args.checkpoint)
    rb._lib.device_error_check()
    if world > 1:
        import torch.distributed as dist
        if args.graph:
            trainer._graph = None             # drop the captured collectives before tearing the communicator down
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
