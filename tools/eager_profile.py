#!/usr/bin/env python
"""Host-side profile (cProfile) of eager training steps at config 2: where the Python / ctypes time of the eager route
goes (the CUDA-graph route does not pay it)."""
import cProfile, contextlib, io, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
import bench
P, B = 128, 2
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = rb.NetworkFromConfig(bench.make_mgr(P, B)).cuda().train()
crit = rb.losses.task_losses(bench.make_mgr(P, B).tasks)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
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
for _ in range(3):
    step()
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"5 eager steps: host issue time {1e3 * (t1 - t0) / 5:.1f} ms/step, wall {1e3 * (t2 - t0) / 5:.1f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print("\n".join(l[:150] for l in s.getvalue().splitlines()[:50]))
