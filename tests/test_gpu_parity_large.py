"""GPU parity at the BASELINE geometries, i.e. the shapes that select the fast kernels (VERDICT r1 "missing" #3):
`slab_conv_kernel` (32-channel full-resolution layers, W in {32, 64, 96, 128, 192}), the weights-on-M h-major tiles,
`tc5_wgrad2<true>`, the tap-split deep layers.  The checker is the fp32 oracle (oracle/resenc_oracle.py, pinned
against the unmodified reference by tests/test_oracle_vs_golden.py) run on the host cores on the same weights and
inputs; nothing here reads /root/reference.

Tolerances (north star): per-task outputs within relative L2 1e-2 of the fp32 result for bf16 compute; weight
gradients are compared through their per-parameter norms and a few full tensors with the bounds calibrated in
tests/test_gpu_network.py (LeakyReLU sign flips make bf16 gradients differ from fp32 ones by tens of per cent for
any bf16 implementation); loss curves over 200 optimiser steps (reference loop train.py:182-231) stay within the
bounds written in the test.
"""
import time

import numpy as np
import pytest
import torch

from helpers import make_mgr, quiet_build, rel_l2
from oracle import resenc_oracle as O   # checker only

pytestmark = pytest.mark.gpu

TASKS2 = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}


@pytest.fixture(autouse=True)
def _device_error_guard(rb):
    yield
    rb._lib.device_error_check()


def _loss(task, pred, target):
    return O.masked_cosine_loss(pred, target) if task == "normals" else O.bce_dice_loss(pred, target)


def _targets(tasks, batch, patch, gen):
    out = {}
    for t, info in tasks.items():
        if t == "normals":
            out[t] = torch.nn.functional.normalize(torch.randn(batch, 3, *patch, generator=gen), dim=1)
        else:
            out[t] = (torch.rand(batch, info["channels"], *patch, generator=gen) > 0.8).float()
    return out


def _forward_case(rb, patch, batch, tasks, in_channels=1, model_config=None, training=True, tol=1e-2, seed=0,
                  calibrate=False):
    """calibrate=True (non-BASELINE shapes only): the bound becomes max(tol, 0.9 x the rel-L2 of PyTorch's own bf16
    autocast of the oracle network on this GPU), i.e. never worse than the existing bf16 GPU path."""
    torch.manual_seed(seed)
    mc = dict(model_config or {})
    model = quiet_build(rb.NetworkFromConfig, make_mgr(patch, tasks, in_channels=in_channels, batch=batch, model_config=mc)).cuda()
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(batch, in_channels, *patch, generator=gen)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    topo = O.autoconfig(patch)
    t0 = time.time()
    with torch.no_grad():
        ref = O.net_forward(sd, topo, x, tasks, training=training, se=bool(mc.get("squeeze_excitation", False)))
    t_ref = time.time() - t0
    model.train(training)
    with torch.no_grad():
        out = model(x.cuda())
    res = {}
    cal = {}
    if calibrate:
        sd_c = {k: v.cuda() for k, v in sd.items()}
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            oc = O.net_forward(sd_c, topo, x.cuda(), tasks, training=training, se=bool(mc.get("squeeze_excitation", False)))
        cal = {t: rel_l2(oc[t].float(), ref[t]) for t in tasks}
    for t in tasks:
        r = rel_l2(out[t], ref[t])
        res[t] = r
        bound = max(tol, 0.9 * cal[t]) if calibrate else tol
        print(f"{'x'.join(map(str, patch))} x{batch} {t}: rel-L2 vs fp32 oracle {r:.3e} (bound {bound:.2e}"
              + (f", torch bf16 autocast {cal[t]:.3e}" if calibrate else "") + f"; oracle {t_ref:.1f} s)")
        assert out[t].dtype == torch.float32 and tuple(out[t].shape) == tuple(ref[t].shape)
        assert r < bound, (patch, t, r)
    return model, out, ref


def test_config2_128_batch2_forward_vs_oracle(rb):
    """BASELINE config 2 (the headline workload): ResEncM-autoconfig 128^3, batch 2, sheet + normals, default init."""
    model, out, ref = _forward_case(rb, [128, 128, 128], 2, TASKS2)
    assert model.num_stages == 6
    agree = float(((out["sheet"].cpu() > 0) == (ref["sheet"] > 0)).float().mean())
    print(f"128^3 x2 sheet sign agreement {agree:.5f}")
    assert agree > 0.99


def test_config4_96_multichannel_se_forward_vs_oracle(rb):
    """BASELINE config 4 geometry: 96^3, 4 input channels, one binary task, squeeze-excitation on (SE arithmetic is
    third-party => the oracle side is the DNA restatement, 'parity unpinned' for that sub-block)."""
    tasks = {"ink": {"channels": 1, "activation": "sigmoid"}}
    model, out, ref = _forward_case(rb, [96, 96, 96], 1, tasks, in_channels=4, model_config={"squeeze_excitation": True},
                                    training=False)
    assert model.num_stages == 5
    agree = float(((out["ink"].cpu() > 0.5) == (ref["ink"] > 0.5)).float().mean())
    print(f"96^3 ink threshold agreement {agree:.5f}")
    assert agree > 0.99


def test_config5_192_forward_vs_oracle(rb):
    """BASELINE config 5 geometry: the 6-stage topology at 192^3 (batch 1)."""
    model, _, _ = _forward_case(rb, [192, 192, 192], 1, TASKS2)
    assert model.num_stages == 6


def test_reference_shipped_patch_shapes_forward_vs_oracle(rb):
    """The patch shapes of the reference's own task files: [64, 192, 192] (tasks/sheet_normals.yaml:3) and the
    anisotropic [14, 256, 256]-style ink patch, here [16, 256, 256] (tasks/ink.yaml:20 needs a depth divisible by the
    pooling schedule)."""
    _forward_case(rb, [64, 192, 192], 1, TASKS2, calibrate=True)
    _forward_case(rb, [16, 256, 256], 1, {"ink": {"channels": 1, "activation": "sigmoid"}}, in_channels=2, calibrate=True)


def test_gradients_64_vs_oracle(rb):
    """Whole-network weight gradients at 64^3 (BASELINE config 1 geometry, batch 2): slab fprop / dgrad, h-major
    weights-on-M tiles and the two-sided tap-stacked wgrad all sit on this path.  Per-parameter gradient norms within
    15 % of the fp32 oracle's (parameters whose gradient is noise in the reference itself excluded), the loss within
    1e-2, selected full gradients within max(0.25, 1.25 x the error of PyTorch's bf16 autocast of the same network on
    the same GPU) relative L2 (bf16 sign-flip calibration, see test_gpu_network.py)."""
    patch, batch = [64, 64, 64], 2
    torch.manual_seed(0)
    model = quiet_build(rb.NetworkFromConfig, make_mgr(patch, TASKS2, batch=batch)).cuda().train()
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(batch, 1, *patch, generator=gen)
    tg = _targets(TASKS2, batch, patch, gen)
    named = dict(model.named_parameters())
    params = {k: v.detach().float().cpu().clone().requires_grad_(True) for k, v in named.items()}
    sd = {}
    by_id = {id(p): n for n, p in named.items()}
    for k, v in model.state_dict(keep_vars=True).items():
        sd[k] = params[by_id[id(v)]]
    topo = O.autoconfig(patch)
    ref = O.net_forward(sd, topo, x, TASKS2, training=True)
    lref = sum(_loss(t, ref[t], tg[t]) for t in TASKS2)
    lref.backward()
    out = model(x.cuda())
    l = sum(_loss(t, out[t], tg[t].cuda()) for t in TASKS2)
    l.backward()
    print(f"64^3 x2 loss: product {float(l):.5f} oracle {float(lref):.5f}")
    assert abs(float(l) - float(lref)) < 1e-2
    gn = {n: float(p.grad.double().norm()) for n, p in params.items() if p.grad is not None}
    gmax = max(gn.values())
    worst, worst_name = 0.0, ""
    for n, g in gn.items():
        assert named[n].grad is not None, n
        if g < 1e-4 * gmax:
            continue
        dev = abs(float(named[n].grad.double().norm()) - g) / g
        if dev > worst:
            worst, worst_name = dev, n
    print(f"64^3 x2 worst gradient-norm deviation {worst:.3e} ({worst_name})")
    assert worst < 0.15, worst_name
    # calibration: PyTorch's own bf16 autocast (ATen / cuDNN on this GPU) of the same functional network on the same
    # weights - what "a bf16 implementation" of this network costs in gradient accuracy (LeakyReLU sign flips: a deep
    # 512-channel layer at 4^3 sees every flip of the 26 blocks above it)
    cal = {k: v.detach().cuda().clone().requires_grad_(True) for k, v in params.items()}
    sd_c = {k: cal[by_id[id(v)]] for k, v in model.state_dict(keep_vars=True).items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        oc = O.net_forward(sd_c, topo, x.cuda(), TASKS2, training=True)
    sum(_loss(t, oc[t].float(), tg[t].cuda()) for t in TASKS2).backward()
    for n in ("shared_encoder.stem.convs.0.conv.weight", "shared_encoder.stages.0.blocks.0.conv1.conv.weight",
              "shared_encoder.stages.1.blocks.1.conv2.conv.weight", "shared_encoder.stages.4.blocks.2.conv1.conv.weight",
              "task_decoders.normals.stages.3.convs.0.conv.weight", "task_decoders.sheet.transpconvs.3.weight",
              "task_decoders.sheet.seg_layers.3.weight"):
        r = rel_l2(named[n].grad, params[n].grad)
        rc = rel_l2(cal[n].grad, params[n].grad)
        print(f"64^3 x2 grad {n}: rel-L2 {r:.3e} (torch bf16 autocast on the same GPU: {rc:.3e})")
        assert r < max(0.25, 1.25 * rc), (n, r, rc)


def test_loss_curve_200_steps_tracks_oracle(rb):
    """North star: "loss curves must agree over 200 steps".  The reference loop body (train.py:195-231: forward,
    per-task losses summed, backward, clip_grad_norm_(3), optimiser step; SGD momentum 0.9 nesterov as train.py:76-84)
    runs for 200 steps over a cycle of 4 fixed batches at 32^3 on the oracle (CPU fp32) and on the CUDA path from the
    same initial weights.  Bounds (measured: mean |deviation| 2.0e-3, max 1.7e-2 at step 186): mean |deviation| over the
    200 steps < 8e-3, every step within 5e-2, the means of the last 20 steps within 3 % of each other, and both curves
    fall below 85 % of their start (1.60 -> 1.22 on the oracle)."""
    patch, batch, steps = [32, 32, 32], 2, 200
    torch.manual_seed(0)
    model = quiet_build(rb.NetworkFromConfig, make_mgr(patch, TASKS2, batch=batch)).cuda().train()
    named = dict(model.named_parameters())
    params = {k: v.detach().float().cpu().clone().requires_grad_(True) for k, v in named.items()}
    by_id = {id(p): n for n, p in named.items()}
    sd = {k: params[by_id[id(v)]] for k, v in model.state_dict(keep_vars=True).items()}
    topo = O.autoconfig(patch)
    gen = torch.Generator().manual_seed(11)
    batches = []
    for _ in range(4):
        x = torch.rand(batch, 1, *patch, generator=gen)
        batches.append((x, _targets(TASKS2, batch, patch, gen)))
    dev_batches = [(x.cuda(), {t: v.cuda() for t, v in tg.items()}) for x, tg in batches]
    kw = dict(lr=0.01, momentum=0.9, nesterov=True, weight_decay=3e-5)
    opt_o = torch.optim.SGD(list(params.values()), **kw)
    opt_p = torch.optim.SGD(model.parameters(), **kw)
    lo, lp = [], []
    for s in range(steps):
        x, tg = batches[s % 4]
        out = O.net_forward(sd, topo, x, TASKS2, training=True)
        l = sum(_loss(t, out[t], tg[t]) for t in TASKS2)
        opt_o.zero_grad(set_to_none=True)
        l.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params.values() if p.grad is not None], 3.0)
        opt_o.step()
        lo.append(float(l))
        xc, tgc = dev_batches[s % 4]
        outp = model(xc)
        l2 = sum(_loss(t, outp[t], tgc[t]) for t in TASKS2)
        opt_p.zero_grad(set_to_none=True)
        l2.backward()
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.grad is not None], 3.0)
        opt_p.step()
        lp.append(float(l2))
    dev = np.abs(np.array(lo) - np.array(lp))
    print("oracle  every 10th:", " ".join(f"{v:.4f}" for v in lo[::10]))
    print("product every 10th:", " ".join(f"{v:.4f}" for v in lp[::10]))
    print(f"200-step curve: mean |dev| {dev.mean():.4e}, max {dev.max():.4e} at step {int(dev.argmax())}, "
          f"tail means {np.mean(lo[-20:]):.4f} / {np.mean(lp[-20:]):.4f}")
    assert np.all(np.isfinite(lp))
    assert dev.mean() < 8e-3 and dev.max() < 5e-2
    assert abs(np.mean(lo[-20:]) - np.mean(lp[-20:])) < 0.03 * np.mean(lo[-20:])
    assert np.mean(lp[-20:]) < 0.85 * lp[0] and np.mean(lo[-20:]) < 0.85 * lo[0]
