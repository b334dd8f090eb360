#!/usr/bin/env python
"""A short sliding-window sweep (256^3 volume -> 27 patches of 128^3, 2 per forward, Gaussian blend) + finalise, eager
launches: the command profiled by ncu for the blend / extract kernels (profiles/r2_ncu_infer_*.csv)."""
import contextlib, io, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
import bench
V = int(sys.argv[1]) if len(sys.argv) > 1 else 256
P = 128
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = rb.NetworkFromConfig(bench.make_mgr(P, 2)).cuda().eval()
targets = {"sheet": {"channels": 1, "activation": "none"}, "normals": {"channels": 3, "activation": "none"}}
sw = rb.inference.SlidingWindowInferer(model, targets, (P,) * 3, overlap=0.5, batch_size=2, weight="gaussian", use_cuda_graph=False)
vol = np.random.default_rng(0).integers(0, 256, size=(V, V, V), dtype=np.uint8)
bl = sw.sweep(vol, max_patches=int(os.environ.get("MAX_PATCHES", 8)))
out = bl.finalize()
torch.cuda.synchronize()
print({t: int(v.to(torch.int64).sum()) for t, v in out.items()})
