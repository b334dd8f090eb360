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


def _grad_worker(rank, world, port, out, wire=None):
    import sys
    sys.path.insert(0, ROOT)
    rb = importlib.import_module("resenc_b200")
    par = importlib.import_module(rb._pkg.__name__ + ".parallel")
    _init(rank, world, port)
    torch.manual_seed(0)
    model = _Tiny()
    buckets = par.GradientBuckets(model, bucket_bytes=256, comm_dtype=wire)     # tiny buckets: several per step
    assert len(buckets.params) == 6 and len(buckets.buckets) >= 3
    res, stats = [], []
    for step in range(3):
        torch.manual_seed(100 * step + rank)
        x = torch.randn(5, 8)
        buckets.zero_grad()
        model(x).square().sum().backward()
        buckets.finish()
        res.append(torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten()
                              for p in model.parameters()]).clone())
        stats.append({"in_backward": buckets.launched_in_backward, "in_finish": buckets.launched_in_finish,
                      "buckets": len(buckets.buckets), "skipped": len(buckets.skipped), "rebuilds": buckets.rebuilds,
                      "unused_grad_is_none": all(p.grad is None for p in model.unused.parameters())})
    if rank == 0:
        torch.save({"grads": res, "stats": stats}, out)
    dist.destroy_process_group()


def test_gradient_buckets_allreduce_mean(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    saved = torch.load(out)
    got, stats = saved["grads"], saved["stats"]
    # step 0 runs on the registration-order cut: the never-used head sits inside a bucket, which therefore only launches
    # in finish() (its unfilled slots sent as zeros).  Its arrival order re-cuts the buckets at step 1: from then on the
    # unused parameters sit in no bucket and EVERY bucket launches while backward is still running.  A parameter without a
    # gradient has .grad None from the first step on (as on the single-GPU path), so the optimiser treats it identically.
    assert stats[0]["in_finish"] >= 1 and stats[0]["unused_grad_is_none"]
    for st in stats[1:]:
        assert st["rebuilds"] == 1 and st["skipped"] == 2 and st["unused_grad_is_none"]
        assert st["in_finish"] == 0 and st["in_backward"] == st["buckets"] >= 2, st
    for step in range(3):
        ref = []
        for rank in range(2):
            torch.manual_seed(0)
            m = _Tiny()
            torch.manual_seed(100 * step + rank)
            x = torch.randn(5, 8)
            m(x).square().sum().backward()
            ref.append(torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten() for p in m.parameters()]))
        assert torch.allclose(got[step], (ref[0] + ref[1]) / 2, atol=1e-6)


def test_gradient_buckets_bf16_wire(tmp_path):
    """comm_dtype=bfloat16: the mean gradient travels as bf16 (half the bytes) and comes back widened; equal to the
    fp32 mean within bf16 rounding."""
    out = str(tmp_path / "g16.pt")
    mp.spawn(_grad_worker, args=(2, _free_port(), out, torch.bfloat16), nprocs=2, join=True)
    got = torch.load(out)["grads"]
    for step in range(3):
        ref = []
        for rank in range(2):
            torch.manual_seed(0)
            m = _Tiny()
            torch.manual_seed(100 * step + rank)
            x = torch.randn(5, 8)
            m(x).square().sum().backward()
            ref.append(torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten() for p in m.parameters()]))
        mean = (ref[0] + ref[1]) / 2
        assert torch.allclose(got[step], mean, rtol=2e-2, atol=1e-3 * float(mean.abs().max()))
        assert not torch.equal(got[step], mean)          # it really went through bf16


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


# ------------------------------------------------------------------------------------------
# DataParallelTrainer on CPU: replicas stay identical and equal the single-process
# step on the mean gradient; rank 0 alone writes the reference-format checkpoint
# ------------------------------------------------------------------------------------------
class _TinyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.body = torch.nn.Conv3d(1, 4, 3, padding=1)
        self.sheet = torch.nn.Conv3d(4, 1, 1)
        self.normals = torch.nn.Conv3d(4, 3, 1)

    def forward(self, x):
        h = torch.nn.functional.leaky_relu(self.body(x), 0.01)
        return {"sheet": self.sheet(h), "normals": self.normals(h)}


_TASKS = {"sheet": {"channels": 1, "activation": "sigmoid", "weight": 2.0}, "normals": {"channels": 3, "activation": "none"}}


def _tiny_batch(step, rank):
    g = torch.Generator().manual_seed(1000 * step + rank)
    x = torch.rand(2, 1, 6, 6, 6, generator=g)
    sheet = (torch.rand(2, 1, 6, 6, 6, generator=g) > 0.7).float()
    normals = torch.nn.functional.normalize(torch.randn(2, 3, 6, 6, 6, generator=g), dim=1)
    return x, {"sheet": sheet, "normals": normals}


def _trainer_worker(rank, world, port, out, accumulate=1):
    import sys
    from types import SimpleNamespace
    sys.path.insert(0, ROOT)
    rb = importlib.import_module("resenc_b200")
    T = rb.training
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    assert T.init_distributed("gloo") == (rank, rank, world)
    torch.manual_seed(7 * rank)      # every rank seeds differently: the trainer must broadcast rank 0's replica
    model = _TinyNet()
    mgr = SimpleNamespace(tasks=_TASKS, optimizer="AdamW", initial_lr=1e-2, weight_decay=1e-4, max_epoch=5,
                          gradient_accumulation=accumulate)
    tr = T.DataParallelTrainer(model, mgr, fused_losses=False)
    assert tr.world == world and tr.rank == rank
    losses = []
    for step in range(3 * accumulate):
        x, tgt = _tiny_batch(step, rank)
        total, per = tr.train_step(x, tgt)
        losses.append(float(total))
    tr.end_epoch()
    wrote = tr.save_checkpoint(out + ".ckpt")
    assert wrote == (rank == 0)
    torch.save({"params": [p.detach().clone() for p in model.parameters()], "losses": losses}, out + f".{rank}")
    dist.destroy_process_group()


def test_data_parallel_trainer_matches_mean_gradient_step(tmp_path):
    from types import SimpleNamespace
    out = str(tmp_path / "t")
    mp.spawn(_trainer_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    parts = [torch.load(out + f".{r}") for r in range(2)]
    for a, b in zip(parts[0]["params"], parts[1]["params"]):
        assert torch.equal(a, b)                       # replicas stay bit-identical
    rb = importlib.import_module("resenc_b200")
    L = rb.losses
    torch.manual_seed(0)
    model = _TinyNet()
    crit = L.task_losses(_TASKS, fused=False)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=1e-4)
    for step in range(3):
        grads = None
        for rank in range(2):
            x, tgt = _tiny_batch(step, rank)
            o = model(x)
            loss = 2.0 * crit["sheet"](o["sheet"], tgt["sheet"]) + crit["normals"](o["normals"], tgt["normals"])
            assert abs(float(loss) - parts[rank]["losses"][step]) < 1e-5
            g = torch.autograd.grad(loss, list(model.parameters()))
            grads = g if grads is None else [a + b for a, b in zip(grads, g)]
        for p, g in zip(model.parameters(), grads):
            p.grad = g / 2
        torch.nn.utils.clip_grad_norm_(model.parameters(), 3.0)
        opt.step()
    for a, b in zip(parts[0]["params"], model.parameters()):
        assert torch.allclose(a, b, atol=1e-6)
    ck = torch.load(out + ".ckpt", weights_only=False)
    assert set(ck) == {"model", "optimizer", "scheduler", "epoch"} and ck["epoch"] == 0     # train.py:249-254: 0-based
    assert set(ck["model"]) == set(model.state_dict())
    # a torch.compile'd reference checkpoint (keys prefixed with _orig_mod.) loads
    ck["model"] = {"_orig_mod." + k: v for k, v in ck["model"].items()}
    torch.save(ck, out + ".ckpt2")
    mgr = SimpleNamespace(tasks=_TASKS, optimizer="AdamW", initial_lr=1e-2, weight_decay=1e-4, max_epoch=5)
    tr = rb.training.DataParallelTrainer(_TinyNet(), mgr, fused_losses=False)
    tr.load_checkpoint(out + ".ckpt2")
    assert tr.epoch == 1                                 # resumes at checkpoint['epoch'] + 1 (train.py:164)
    # rolling window of checkpoints (train.py:256-265)
    import time
    for e in range(1, 6):
        tr.epoch = e
        tr.save_checkpoint(str(tmp_path / f"Model_{e}.pth"), keep_newest=3)
        time.sleep(0.02)
    assert sorted(os.listdir(tmp_path)).count("Model_1.pth") == 0
    assert [f for f in sorted(os.listdir(tmp_path)) if f.startswith("Model_")] == ["Model_3.pth", "Model_4.pth", "Model_5.pth"]
    for a, b in zip(tr.model.parameters(), parts[0]["params"]):
        assert torch.equal(a, b)


def test_shard_indices_cover_and_balance():
    rb = importlib.import_module("resenc_b200")
    T = rb.training
    idx = list(range(23))
    for world in (1, 2, 4, 8):
        shards = [T.shard_indices(idx, r, world) for r in range(world)]
        assert len({len(s) for s in shards}) == 1                     # equal step counts (no all-reduce dead-lock)
        flat = sorted(i for s in shards for i in s)
        assert flat == idx[:len(idx) - len(idx) % world]
        shards = [T.shard_indices(idx, r, world, drop_last=False) for r in range(world)]
        assert len({len(s) for s in shards}) == 1 and set(i for s in shards for i in s) == set(idx)
    a = list(T.iterate_sharded(10, 3, 0, 2)) + list(T.iterate_sharded(10, 3, 1, 2))
    assert sorted(a) == list(range(10))
    with pytest.raises(ValueError):
        T.shard_indices(idx, 2, 2)


def test_data_parallel_trainer_gradient_accumulation(tmp_path):
    """gradient_accumulation = 2 under data parallelism (train.py:222-230): micro-batches accumulate locally
    (GradientBuckets.no_sync), the all-reduce rides the updating backward; equals the single-process step on the
    mean gradient over ranks x micro-batches."""
    out = str(tmp_path / "a")
    mp.spawn(_trainer_worker, args=(2, _free_port(), out, 2), nprocs=2, join=True)
    parts = [torch.load(out + f".{r}") for r in range(2)]
    for a, b in zip(parts[0]["params"], parts[1]["params"]):
        assert torch.equal(a, b)
    rb = importlib.import_module("resenc_b200")
    crit = rb.losses.task_losses(_TASKS, fused=False)
    torch.manual_seed(0)
    model = _TinyNet()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=1e-4)
    for upd in range(3):
        grads = None
        for micro in range(2):
            for rank in range(2):
                x, tgt = _tiny_batch(2 * upd + micro, rank)
                o = model(x)
                loss = 2.0 * crit["sheet"](o["sheet"], tgt["sheet"]) + crit["normals"](o["normals"], tgt["normals"])
                assert abs(float(loss) - parts[rank]["losses"][2 * upd + micro]) < 1e-5
                g = torch.autograd.grad(loss, list(model.parameters()))
                grads = g if grads is None else [a + b for a, b in zip(grads, g)]
        for p, g in zip(model.parameters(), grads):
            p.grad = g / 4
        torch.nn.utils.clip_grad_norm_(model.parameters(), 3.0)
        opt.step()
    for a, b in zip(parts[0]["params"], model.parameters()):
        assert torch.allclose(a, b, atol=1e-6)
