#!/usr/bin/env python
"""Kernel timeline of ONE CUDA-graph replay of the config-2 training step (what bench.py times): per-kernel table,
idle gaps between consecutive kernels, concurrency.  Usage: python tools/graph_profile.py [out.json]"""
import collections, contextlib, importlib, io, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
import bench
P, B = int(os.environ.get("PATCH", 128)), 2
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = rb.NetworkFromConfig(bench.make_mgr(P, B)).cuda().train()
crit = rb.losses.task_losses(bench.make_mgr(P, B).tasks)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True, capturable=True)
x, tgt = bench.synthetic_batch(B, P, "cpu", 0)
x = x.cuda(); tgt = {k: v.cuda() for k, v in tgt.items()}
params = list(model.parameters())
def step():
    out = model(x)
    loss = bench.gpu_losses(out, tgt, crit)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 3.0)
    opt.step()
    return loss
for _ in range(3):
    step()
rb.ops.PACK_CACHE = False
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
rb.ops.PACK_CACHE = True
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1) / 5:.3f} ms per step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
print(f"profiled replay: span {(t1 - t0) / 1e3:.3f} ms, {len(evs)} device activities, sum of durations {sum(e.device_time for e in evs) / 1e3:.3f} ms")
# coverage: time during which at least one kernel runs
cov, cur_s, cur_e = 0.0, None, None
for e in evs:
    s, en = e.time_range.start, e.time_range.end
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            cov += cur_e - cur_s
        cur_s, cur_e = s, en
    else:
        cur_e = max(cur_e, en)
cov += cur_e - cur_s
print(f"busy (>= 1 kernel running): {cov / 1e3:.3f} ms; idle gaps: {(t1 - t0 - cov) / 1e3:.3f} ms")
def short(n):
    n = n.replace("void ", "").replace("rb::", "")
    return n.split("(")[0][:60]
tab = collections.defaultdict(lambda: [0.0, 0])
for e in evs:
    tab[short(e.name)][0] += e.device_time; tab[short(e.name)][1] += 1
rows = sorted(tab.items(), key=lambda kv: -kv[1][0])
for k, (us, n) in rows[:45]:
    print(f"{us / 1e3:8.3f} ms {n:5d}  {k}")
# gap histogram
gaps = []
cur_e = None
for e in evs:
    if cur_e is not None and e.time_range.start > cur_e:
        gaps.append(e.time_range.start - cur_e)
    cur_e = max(cur_e or 0, e.time_range.end)
gaps.sort()
if gaps:
    print(f"gaps: n {len(gaps)}, median {gaps[len(gaps)//2]:.1f} us, mean {sum(gaps)/len(gaps):.1f} us, p90 {gaps[int(0.9*len(gaps))]:.1f} us, max {gaps[-1]:.1f} us")
if len(sys.argv) > 1:
    seq = [(short(e.name), round(e.time_range.start - t0, 1), round(e.device_time, 1)) for e in evs]
    json.dump(seq, open(sys.argv[1], "w"))
