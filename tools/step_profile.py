#!/usr/bin/env python
"""torch.profiler kernel table of two eager training steps (config 2) — where the non-conv time goes."""
import contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
import bench
import importlib
P, B = 128, 2
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = rb.NetworkFromConfig(bench.make_mgr(P, B)).cuda().train()
crit = importlib.import_module(rb._pkg.__name__ + ".losses").task_losses(bench.make_mgr(P, B).tasks)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
x, tgt = bench.synthetic_batch(B, P, "cpu", 0)
x = x.cuda(); tgt = {k: v.cuda() for k, v in tgt.items()}
def step():
    out = model(x)
    loss = bench.gpu_losses(out, tgt, crit)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(list(model.parameters()), 3.0)
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 2e3, e.count // 2) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device time per step: {tot:.2f} ms")
for k, ms, n in rows[:40]:
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}% {n:5d}  {k[:110]}")

# per-launch durations of the conv kernels in launch order (second profiled step), to spot the slow geometry classes
if os.environ.get("STEP_PROFILE_LAUNCHES"):
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and
           any(k in e.name for k in ("gather_conv", "slab_conv", "wgrad"))]
    evs.sort(key=lambda e: e.time_range.start)
    half = evs[len(evs) // 2:]
    out = []
    for e in half:
        short = "tc5t" if "tc5t" in e.name else "slab" if "slab" in e.name else "wgrad2" if "wgrad2" in e.name else \
            "wgrad" if "wgrad" in e.name else "tc5" if "tc5_gather" in e.name else "mma"
        out.append(f"{short}:{e.device_time:.0f}")
    print("launch order (us):", " ".join(out))
