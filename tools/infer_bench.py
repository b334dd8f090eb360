#!/usr/bin/env python
"""Sliding-window inference benchmark (BASELINE config 3 shape: 128^3 patches, 50 % overlap, sheet + normals,
z-slab sharded across ranks).  Synthetic uint8 volume served by a numpy array (zarr's slicing interface),
random-init weights.  Reports output voxels/s (volume voxels / sweep+merge+finalise time, device timed, max over
ranks) and patch voxels/s; `e2e` includes the host->device upload of the volume slab and the device->host read of
the finalised uint8/uint16 slabs.

    python tools/infer_bench.py --vol 384                       # 1 GPU
    torchrun --nproc-per-node 8 tools/infer_bench.py --vol 1024
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb   # noqa: E402

TASKS = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vol", type=int, default=384)
    ap.add_argument("--patch", type=int, default=128)
    ap.add_argument("--overlap", type=float, default=0.5)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--weight", default="gaussian", choices=["uniform", "gaussian"])
    ap.add_argument("--warmup-patches", type=int, default=4)
    ap.add_argument("--precise", action="store_true", help="split-precision (bf16x3) forward: fp32-accurate logits")
    ap.add_argument("--precise-impl", default=None, choices=[None, "mma", "auto"])
    args = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    inf = rb.inference
    P, V = args.patch, args.vol
    torch.manual_seed(0)
    mgr = SimpleNamespace(tasks=TASKS, train_patch_size=[P] * 3, train_batch_size=args.batch, in_channels=1, vram_max=16.0,
                          autoconfigure=True, model_config={})
    with contextlib.redirect_stdout(io.StringIO()):
        model = rb.NetworkFromConfig(mgr).to(dev).eval()
    rng = np.random.default_rng(0)
    volume = rng.integers(0, 256, size=(V, V, V), dtype=np.uint8)
    # the model applies the training-config activation in eval mode (build_network_from_config.py:322-323); the
    # inference config adds none on top (avoids the reference's double-sigmoid quirk)
    targets = {"sheet": {"channels": 1, "activation": "none"}, "normals": {"channels": 3, "activation": "none"}}
    sw = inf.SlidingWindowInferer(model, targets, (P,) * 3, overlap=args.overlap, batch_size=args.batch, weight=args.weight,
                                  rank=rank, world_size=world, device=dev, precise=args.precise)
    if args.precise_impl:
        rb.precise.IMPL = args.precise_impl
    positions, z_lo, z_hi, (zs, ys, xs) = sw.plan(volume.shape)
    # warm-up: a few patches through the network (weight packing, kernel attribute setup)
    with torch.no_grad():
        x = torch.rand(args.batch, 1, P, P, P, device=dev)
        for _ in range(max(1, args.warmup_patches // args.batch)):
            sw._forward(x)
    torch.cuda.synchronize()
    if world > 1:
        # establish the point-to-point connections the slab exchange uses (one-time NCCL setup, not part of a sweep)
        pairs, _ = inf.plan_slab_exchange(zs, P, V, world)
        one = torch.zeros(1, device=dev)
        for src, dst, _, _ in pairs:
            if rank == src:
                dist.send(one, dst)
            elif rank == dst:
                dist.recv(one, src)
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    blender = sw.sweep(volume)                      # includes the H2D upload of this rank's slab
    e1.record()
    own = inf.merge_slabs(blender, zs, rank, world) if world > 1 else (0, V)
    out = blender.finalize(*own) if own[1] > own[0] else {}
    e2.record()
    torch.cuda.synchronize()
    host = {t: v.cpu() for t, v in out.items()}     # D2H of the finalised slab
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms_sweep, ms_total = e0.elapsed_time(e1), e0.elapsed_time(e2)
    tt = torch.tensor([ms_sweep, ms_total, (t1 - t0) * 1e3], dtype=torch.float64, device=dev)
    npatch = torch.tensor([len(positions)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(npatch, op=dist.ReduceOp.SUM)
    rb._lib.device_error_check()
    if rank == 0:
        ms_sweep, ms_total, ms_wall = (float(v) for v in tt)
        n = int(npatch[0])
        d2h = sum(v.numel() * v.element_size() for v in host.values())
        line = {
            "metric": "infer output voxels/s", "value": V ** 3 / (ms_total * 1e-3), "unit": "voxels/s", "n_gpus": world,
            "patch_voxels_per_s": n * P ** 3 / (ms_total * 1e-3), "patches": n, "ms_total": ms_total, "ms_sweep": ms_sweep,
            "ms_per_patch": ms_sweep / max(1, n / world),
            "config": {"workload": f"sliding window {V}^3 uint8, {P}^3 patches, overlap {args.overlap}, {args.weight} blend, "
                                   f"batch {args.batch}, z-slab sharded over {world} GPU(s), sheet(1) + normals(3)"},
            "e2e": {"value": V ** 3 / (ms_wall * 1e-3), "unit": "voxels/s", "ms": ms_wall,
                    "h2d_bytes": getattr(sw, "h2d_bytes", 0), "d2h_bytes_rank0": d2h},
            "dtype": "bf16x3 (split precision, fp32-accurate)" if args.precise else "bf16", "data": "synthetic",
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
