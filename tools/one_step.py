#!/usr/bin/env python
"""N eager training steps of BASELINE config 2 with nothing else attached (for `ncu -k regex:... --metrics ...`)."""
import contextlib, importlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
import bench
P, B = 128, 2
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = rb.NetworkFromConfig(bench.make_mgr(P, B)).cuda().train()
crit = rb.losses.task_losses(bench.make_mgr(P, B).tasks)
opt = rb.optim.ClippedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, max_grad_norm=3.0)   # clip + AdamW (train.py:227-228)
x, tgt = bench.synthetic_batch(B, P, "cpu", 0)
x = x.cuda(); tgt = {k: v.cuda() for k, v in tgt.items()}
for _ in range(steps):
    out = model(x)
    loss = bench.gpu_losses(out, tgt, crit)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
