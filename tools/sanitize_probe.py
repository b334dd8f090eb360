#!/usr/bin/env python
"""Smallest run that launches every hand-written kernel family once, for compute-sanitizer (one tool per gpurun call):
forward + losses + backward of the drop-in network at a [48, 32, 32] patch (stem, tcgen05 gather convs in both
orientations, the slab kernel, the tap-split deep layers + finish pass, both tcgen05 weight-gradient kernels, norm /
pool / head / loss / pack / unpack kernels) and a two-patch sliding-window blend (extract, accumulate, finalise)."""
import contextlib, io, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
from types import SimpleNamespace
tasks = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
patch = [48, 32, 32]
mgr = SimpleNamespace(tasks=tasks, train_patch_size=patch, train_batch_size=2, in_channels=1, vram_max=16.0, autoconfigure=True,
                      model_config={"squeeze_excitation": bool(int(os.environ.get("SE", "0")))})
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = rb.NetworkFromConfig(mgr).cuda().train()
crit = rb.losses.task_losses(tasks)
x = torch.rand(2, 1, *patch, device="cuda")
tgt = {"sheet": (torch.rand(2, 1, *patch, device="cuda") > 0.8).float(),
       "normals": torch.nn.functional.normalize(torch.randn(2, 3, *patch, device="cuda"), dim=1)}
l0 = rb._lib.launch_count()
out = model(x)
loss = sum(crit[t](out[t], tgt[t]) for t in tasks)
loss.backward()
torch.cuda.synchronize()
rb._lib.device_error_check()
model.eval()
targets = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
sw = rb.inference.SlidingWindowInferer(model, targets, patch, overlap=0.5, batch_size=2, weight="gaussian", use_cuda_graph=False)
vol = np.random.default_rng(0).integers(0, 256, size=(72, 32, 32), dtype=np.uint8)
res = sw.run(vol)
torch.cuda.synchronize()
rb._lib.device_error_check()
print(f"loss {float(loss):.5f}; {rb._lib.launch_count() - l0} launches; blend checksum {int(res['sheet'].to(torch.int64).sum())}")
