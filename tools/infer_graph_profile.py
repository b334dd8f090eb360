#!/usr/bin/env python
"""Kernel table of ONE graph-replayed inference forward (4 patches of 128^3) + the per-batch extract / blend launches."""
import collections, contextlib, io, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
import bench
P, B = 128, int(os.environ.get("BATCH", 4))
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = rb.NetworkFromConfig(bench.make_mgr(P, 2)).cuda().eval()
targets = {"sheet": {"channels": 1, "activation": "none"}, "normals": {"channels": 3, "activation": "none"}}
sw = rb.inference.SlidingWindowInferer(model, targets, (P,) * 3, overlap=0.5, batch_size=B, weight="gaussian")
vol = np.random.default_rng(0).integers(0, 256, size=(256, 256, 384), dtype=np.uint8)
sw.sweep(vol)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    sw._graph[0].replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0, t1 = min(e.time_range.start for e in evs), max(e.time_range.end for e in evs)
print(f"forward of {B} patches: span {(t1 - t0) / 1e3:.3f} ms, {len(evs)} activities, sum {sum(e.device_time for e in evs) / 1e3:.3f} ms")
tab = collections.defaultdict(lambda: [0.0, 0])
for e in evs:
    k = e.name.replace("void ", "").replace("rb::", "").split("(")[0][:60]
    tab[k][0] += e.device_time; tab[k][1] += 1
for k, (us, n) in sorted(tab.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{us / 1e3:8.3f} ms {n:5d}  {k}")
