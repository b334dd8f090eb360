#!/usr/bin/env python
"""Per-parameter gradient comparison of the CUDA path against the oracle's fp32 autograd on one golden case
(diagnostic script, not collected by pytest; it lives under tests/ because it calls the oracle; prints rel-L2 of every
parameter gradient in registration order)."""
import importlib
import sys
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import resenc_b200 as rb                                   # noqa: E402
from helpers import (NET_CASES, case_mgr, golden_state, load_net_golden, quiet_build, rel_l2,  # noqa: E402
                     state_dict_from_params)
from oracle import resenc_oracle as O                      # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "sheet_normals_16"
impl = sys.argv[2] if len(sys.argv) > 2 else None
if impl:
    os.environ["RESENC_CONV_IMPL"] = impl
mgr, rd = case_mgr(case)
importlib.import_module(rb.builders.__name__ + ".resblocks").SE_REDUCE_DIMS = rd
model = quiet_build(rb.NetworkFromConfig, mgr)
model.load_state_dict(state_dict_from_params(model, golden_state(case)))
model = model.cuda().train()
gold = load_net_golden(case)
x = torch.from_numpy(gold["x"])
tg = {t: torch.from_numpy(gold["target::" + t]) for t in mgr.tasks}
loss_of = lambda t, p, y: O.masked_cosine_loss(p, y) if t == "normals" else O.bce_dice_loss(p, y)
params = {k: v.clone().requires_grad_(True) for k, v in golden_state(case).items()}
patch, cin, tasks, mc, _, batch = NET_CASES[case]
out = O.net_forward(params, O.autoconfig(patch), x, tasks, training=True, se=bool(mc.get("squeeze_excitation")), reduce_dims=rd)
sum(loss_of(t, out[t], tg[t]) for t in tasks).backward()
outp = model(x.cuda())
sum(loss_of(t, outp[t], tg[t].cuda()) for t in tasks).backward()
for n, p in model.named_parameters():
    g = params[n].grad
    if p.grad is None or g is None:
        print(f"{n:75s} none: ours {p.grad is None} oracle {g is None}")
        continue
    print(f"{n:75s} |g| {float(g.norm()):9.3e}  rel-L2 {rel_l2(p.grad, g):9.3e}")
