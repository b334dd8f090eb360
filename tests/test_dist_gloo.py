"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): the bucketed gradient all-reduce used for
data-parallel training and the end-of-sweep z-slab exchange of the sliding-window inference."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(8, 16)
        self.b = torch.nn.Linear(16, 4)
        self.unused = torch.nn.Linear(16, 2)      # like the decoder's deep-supervision heads: never gets a gradient
        self.alias = self.a                        # aliased module: parameters must be bucketed once

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def _grad_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    rb = importlib.import_module("resenc_b200")
    par = importlib.import_module(rb._pkg.__name__ + ".parallel")
    _init(rank, world, port)
    torch.manual_seed(0)
    model = _Tiny()
    buckets = par.GradientBuckets(model, bucket_bytes=256)     # tiny buckets: several per step
    assert len(buckets.params) == 6 and len(buckets.buckets) >= 3
    res = []
    for step in range(2):
        torch.manual_seed(100 * step + rank)
        x = torch.randn(5, 8)
        buckets.zero_grad()
        model(x).square().sum().backward()
        buckets.finish()
        res.append(torch.cat([p.grad.flatten() for p in model.parameters()]).clone())
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_gradient_buckets_allreduce_mean(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    for step in range(2):
        ref = []
        for rank in range(2):
            torch.manual_seed(0)
            m = _Tiny()
            torch.manual_seed(100 * step + rank)
            x = torch.randn(5, 8)
            m(x).square().sum().backward()
            ref.append(torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten() for p in m.parameters()]))
        assert torch.allclose(got[step], (ref[0] + ref[1]) / 2, atol=1e-6)


def _slab_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    rb = importlib.import_module("resenc_b200")
    inf = rb.inference
    _init(rank, world, port)
    vol, patch = (40, 12, 10), (16, 8, 8)
    targets = {"sheet": {"channels": 1}, "normals": {"channels": 3}}
    zs, ys, xs = inf.axis_positions(vol, patch, 0.5)
    mine = inf.shard_z_starts(zs, world)[rank]
    z_lo, z_hi = mine[0], mine[-1] + patch[0]
    bl = inf.SlabBlender(targets, vol, patch, z_lo, z_hi, "cpu", "uniform")
    rng = np.random.default_rng(3)          # same stream on both ranks: every rank draws every patch
    for z in zs:
        for y in ys:
            for x in xs:
                p1 = torch.from_numpy(rng.random((1, *patch), dtype=np.float32))
                p3 = torch.from_numpy(rng.random((3, *patch), dtype=np.float32))
                if z in mine:          # host-side stand-in for rb_blend_accumulate (CUDA only)
                    a = z - z_lo
                    bl.sums["sheet"][:, a:a + 16, y:y + 8, x:x + 8] += p1
                    bl.sums["normals"][:, a:a + 16, y:y + 8, x:x + 8] += p3
                    bl.wsum[a:a + 16, y:y + 8, x:x + 8] += 1
    own = inf.merge_slabs(bl, zs, rank, world, add_fn=lambda d, s: d.add_(s))
    a, b = own[0] - z_lo, own[1] - z_lo
    torch.save({"own": own, "sheet": bl.sums["sheet"][:, a:b].clone(), "normals": bl.sums["normals"][:, a:b].clone(),
                "wsum": bl.wsum[a:b].clone()}, out + f".{rank}")
    dist.destroy_process_group()


def test_slab_exchange_matches_single_rank(tmp_path, built_lib):
    out = str(tmp_path / "s.pt")
    mp.spawn(_slab_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    parts = [torch.load(out + f".{r}") for r in range(2)]
    assert parts[0]["own"][0] == 0 and parts[0]["own"][1] == parts[1]["own"][0] and parts[1]["own"][1] == 40
    rb = importlib.import_module("resenc_b200")
    inf = rb.inference
    vol, patch = (40, 12, 10), (16, 8, 8)
    zs, ys, xs = inf.axis_positions(vol, patch, 0.5)
    rng = np.random.default_rng(3)
    s1 = np.zeros((1, *vol), np.float32); s3 = np.zeros((3, *vol), np.float32); w = np.zeros(vol, np.float32)
    for z in zs:
        for y in ys:
            for x in xs:
                p1 = rng.random((1, *patch), dtype=np.float32); p3 = rng.random((3, *patch), dtype=np.float32)
                s1[:, z:z + 16, y:y + 8, x:x + 8] += p1
                s3[:, z:z + 16, y:y + 8, x:x + 8] += p3
                w[z:z + 16, y:y + 8, x:x + 8] += 1
    got1 = torch.cat([p["sheet"] for p in parts], 1).numpy()
    got3 = torch.cat([p["normals"] for p in parts], 1).numpy()
    gotw = torch.cat([p["wsum"] for p in parts], 0).numpy()
    assert np.array_equal(gotw, w)
    assert np.allclose(got1, s1, atol=1e-5) and np.allclose(got3, s3, atol=1e-5)
