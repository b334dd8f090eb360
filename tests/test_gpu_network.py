"""GPU parity of the whole drop-in network against fixtures generated from the UNMODIFIED reference
(tests/golden/net_*.npz, made by oracle/make_golden.py) and against the oracle on fresh inputs.

Tolerances (north star): per-task outputs within relative L2 1e-2 of the reference fp32 output for
bf16 compute.  The 16^3 fixture networks normalise over planes of only 64 voxels, where bf16 operand
rounding alone costs more than that: PyTorch's own bf16 autocast of the UNMODIFIED reference measures
1.2e-2 .. 1.8e-2 on them (stored in the fixtures as `autocast_bf16_rel::<task>` by oracle/make_golden.py).
There the bound is max(1e-2, 0.8 x that figure), i.e. strictly better than autocast; at 64^3 (BASELINE
config 1) the plain 1e-2 bound is asserted.  Losses within 1e-2 absolute; weight-gradients no worse than PyTorch's bf16 autocast of the reference
(calibration stored in the fixtures); threshold agreement reported and asserted >= 99 % at these
tiny random-init sizes (SURVEY 0.10 shows PyTorch's own bf16 autocast reaches 99.65 %).
"""
import importlib

import numpy as np
import pytest
import torch

from helpers import (NET_CASES, NET_EXTRA, case_mgr, golden_eval, golden_state, load_net_golden, make_mgr, quiet_build, rel_l2,
                     state_dict_from_params)
from oracle import resenc_oracle as O   # checker only

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _device_error_guard(rb):
    yield
    rb._lib.device_error_check()


def _set_se_dims(rb, rd):
    mod = importlib.import_module(rb.builders.__name__ + ".resblocks")
    mod.SE_REDUCE_DIMS = rd


def _build(rb, case):
    mgr, rd = case_mgr(case)
    _set_se_dims(rb, rd)
    model = quiet_build(rb.NetworkFromConfig, mgr)
    params = golden_state(case)
    missing, unexpected = model.load_state_dict(state_dict_from_params(model, params), strict=True)
    assert not missing and not unexpected
    gold = load_net_golden(case)
    for k in gold.files:          # replay the reference run's stochastic-depth draws (training mode only)
        if k.startswith("drop::"):
            model.get_submodule(k[6:]).drop_path.forced_factor = torch.from_numpy(gold[k])
    return model.cuda(), mgr


def _loss(task, pred, target):
    return O.masked_cosine_loss(pred, target) if task == "normals" else O.bce_dice_loss(pred, target)


@pytest.mark.parametrize("case", list(NET_CASES))
def test_network_matches_reference_golden(rb, case):
    try:
        model, mgr = _build(rb, case)
        gold = load_net_golden(case)
        x = torch.from_numpy(gold["x"]).cuda()
        model.train()
        out = model(x)
        assert list(out.keys()) == list(mgr.tasks.keys())
        total = 0.0
        for t in mgr.tasks:
            ref = torch.from_numpy(gold["train::" + t])
            assert out[t].dtype == torch.float32 and tuple(out[t].shape) == tuple(ref.shape)
            r = rel_l2(out[t], ref)
            tol = max(1e-2, 0.8 * float(gold["autocast_bf16_rel::" + t]))
            print(f"{case}/{t}: train rel-L2 {r:.3e} (bound {tol:.3e}, torch bf16 autocast of the reference "
                  f"{float(gold['autocast_bf16_rel::' + t]):.3e})")
            assert r < tol, (case, t, r)
            l = _loss(t, out[t], torch.from_numpy(gold["target::" + t]).cuda())
            assert abs(float(l) - float(gold["loss::" + t])) < 1e-2
            total = total + l
        total.backward()
        names = [str(n) for n in gold["param_names"]]
        named = dict(model.named_parameters())
        has = np.array([named[n].grad is not None for n in names])
        assert np.array_equal(has, gold["has_grad"]), [n for n, a, b in zip(names, has, gold["has_grad"]) if a != b]
        worst, worst_name, worst_se = 0.0, "", 0.0
        gmax = float(np.max(gold["grad_norms"]))
        for n, gn in zip(names, gold["grad_norms"]):
            if named[n].grad is None or gn < 1e-4 * gmax:
                # cancelled conv biases (reference: rounding noise ~1e-9, ours: exact 0) and SE fc1 weights behind
                # a global pool that is identically 0: both are noise in the reference itself
                continue
            mine = float(named[n].grad.double().norm())
            if "squeeze_excitation" in n:
                # gate parameters see the bf16 noise of a whole stage through one pooled scalar per channel: PyTorch's
                # own bf16 autocast of the reference deviates by 0.25-0.30 on them in the SE fixtures, and the figure
                # moves by +-0.05 between runs of the CUDA path (atomics order), so they get their own bound
                worst_se = max(worst_se, abs(mine - gn) / gn)
                continue
            if abs(mine - gn) / gn > worst:
                worst, worst_name = abs(mine - gn) / gn, n
        # gradients of a LeakyReLU network are discontinuous in the activations: every sign flip caused by bf16
        # forward rounding changes a local derivative 100x, so bf16 gradients differ from fp32 ones by tens of per
        # cent in rel-L2 whatever the kernel (PyTorch's bf16 autocast of the reference is stored as calibration)
        ac_dev = float(gold["autocast_bf16_gradnorm_dev"])
        print(f"{case}: worst grad-norm deviation {worst:.3e} ({worst_name}); torch bf16 autocast: {ac_dev:.3e}")
        # floor 0.15: the figure moves by +-0.05 between runs (atomics order), most on the half-dropped batch of the
        # stochastic-depth fixture
        assert worst < max(0.15, 2.0 * ac_dev), worst_name
        print(f"{case}: worst SE-gate grad-norm deviation {worst_se:.3e}")
        assert worst_se < max(0.35, 2.0 * ac_dev)
        for k in gold.files:
            if k.startswith("grad::"):
                r = rel_l2(named[k[6:]].grad, gold[k])
                ac = float(gold["autocast_bf16_gradrel::" + k[6:]]) if "autocast_bf16_gradrel::" + k[6:] in gold.files else 0.0
                if float(np.linalg.norm(gold[k])) < 1e-4 * gmax:
                    continue     # noise in the reference itself (see above)
                print(f"{case}: {k} rel-L2 {r:.3e} (torch bf16 autocast {ac:.3e})")
                assert r < max(5e-2, 1.5 * ac), (k, r)
        model.eval()
        with torch.no_grad():
            ev = model(x)
        for t, info in mgr.tasks.items():
            ref = golden_eval(gold, t, info["activation"])
            r = rel_l2(ev[t], ref)
            print(f"{case}/{t}: eval rel-L2 {r:.3e}")
            assert r < max(1e-2, 0.8 * float(gold["autocast_bf16_rel::" + t]))
            if info["activation"] == "sigmoid":
                agree = float(((ev[t].cpu() > 0.5) == (ref > 0.5)).float().mean())
                print(f"{case}/{t}: threshold agreement {agree:.5f}")
                assert agree >= 0.99
            if info["activation"] == "softmax":
                agree = float((ev[t].cpu().argmax(1) == ref.argmax(1)).float().mean())
                print(f"{case}/{t}: argmax agreement {agree:.5f}")
                assert agree >= 0.99
    finally:
        _set_se_dims(rb, "all")


def test_network_64_vs_oracle_default_init(rb):
    """BASELINE config 1 (64^3, batch 1, sheet + normals) with PyTorch default init under seed 0:
    forward, loss and a step of SGD agree with the oracle on the same weights."""
    tasks = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
    torch.manual_seed(0)
    model = quiet_build(rb.NetworkFromConfig, make_mgr([64, 64, 64], tasks)).cuda()
    x = torch.rand(1, 1, 64, 64, 64)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    topo = O.autoconfig([64, 64, 64])
    with torch.no_grad():
        ref = O.net_forward(sd, topo, x, tasks, training=True)
    model.train()
    out = model(x.cuda())
    for t in tasks:
        r = rel_l2(out[t], ref[t])
        print(f"64^3 {t}: rel-L2 {r:.3e}")
        assert r < 1e-2
    agree = float(((out["sheet"].cpu() > 0) == (ref["sheet"] > 0)).float().mean())
    print(f"64^3 sheet sign agreement {agree:.5f}")
    assert agree > 0.99


def test_training_loss_decreases_and_tracks_oracle(rb):
    """30 SGD steps on a fixed batch (16^3 net): the loss curve of the CUDA path follows the oracle's
    fp32 curve (same init, same data): identical to 5e-3 over the first six steps, 2e-2 over the first ten, within 0.1
    later (mean deviation < 3e-2), same end point.  The bounds carry the run-to-run spread of the CUDA path itself
    (fp32 atomics order -> occasional bf16 rounding flips, amplified by momentum SGD on a single batch)."""
    case = "sheet_normals_16"
    model, mgr = _build(rb, case)
    gold = load_net_golden(case)
    x = torch.from_numpy(gold["x"])
    tg = {t: torch.from_numpy(gold["target::" + t]) for t in mgr.tasks}
    # oracle side: functional forward over leaf tensors with autograd
    params = {k: v.clone().requires_grad_(True) for k, v in golden_state(case).items()}
    topo = O.autoconfig(NET_CASES[case][0])
    opt_o = torch.optim.SGD(list(params.values()), lr=0.05, momentum=0.9)
    opt_p = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    xc = x.cuda()
    tgc = {t: v.cuda() for t, v in tg.items()}
    model.train()
    lo, lp = [], []
    for step in range(30):
        out = O.net_forward(params, topo, x, mgr.tasks, training=True)
        l = sum(_loss(t, out[t], tg[t]) for t in mgr.tasks)
        opt_o.zero_grad()
        l.backward()
        opt_o.step()
        lo.append(float(l))
        outp = model(xc)
        l2 = sum(_loss(t, outp[t], tgc[t]) for t in mgr.tasks)
        opt_p.zero_grad(set_to_none=True)
        l2.backward()
        opt_p.step()
        lp.append(float(l2))
    print("oracle :", " ".join(f"{v:.4f}" for v in lo))
    print("product:", " ".join(f"{v:.4f}" for v in lp))
    assert lp[-1] < 0.3 * lp[0]
    # SGD with momentum on one batch is mildly chaotic: the first ten steps must coincide, later ones stay close
    dev = [abs(a - b) for a, b in zip(lo, lp)]
    assert max(dev[:6]) < 5e-3
    assert max(dev[:10]) < 2e-2
    assert max(dev) < 0.1 and sum(dev) / len(dev) < 3e-2
    assert abs(lo[-1] - lp[-1]) < 0.15 * lo[-1]


def test_state_dict_roundtrip_and_compile_wrapper(rb):
    model, mgr = _build(rb, "sheet_normals_16")
    x = torch.rand(2, 1, 16, 16, 16, device="cuda")
    model.eval()
    with torch.no_grad():
        a = model(x)
    sd = {("_orig_mod." + k): v.clone() for k, v in model.state_dict().items()}   # checkpoint of a compiled model
    m2 = quiet_build(rb.NetworkFromConfig, mgr).cuda()
    m2.load_state_dict({k[len("_orig_mod."):]: v for k, v in sd.items()})
    m2.eval()
    with torch.no_grad():
        b = m2(x)
    # run-to-run differences: fp32 statistics / split-K partial sums are combined with atomics, whose order
    # moves the last bits and occasionally a bf16 rounding
    for t in a:
        assert rel_l2(a[t], b[t]) < 5e-3
    cm = torch.compile(m2)
    with torch.no_grad(), torch.amp.autocast("cuda"):
        c = cm(x)
    for t in a:
        assert c[t].dtype == torch.float32 and rel_l2(a[t], c[t]) < 5e-3


def test_eval_activation_semantics(rb):
    model, mgr = _build(rb, "aniso_8x32x32")
    x = torch.rand(1, 1, 8, 32, 32, device="cuda")
    model.train()
    with torch.no_grad():
        raw = model(x)["sheet"]
    model.eval()
    with torch.no_grad():
        act = model(x)["sheet"]
    # two forward passes differ by atomics-order rounding (a few bf16 ulps on a few activations)
    assert rel_l2(act, torch.softmax(raw, 1)) < 5e-3
    assert torch.allclose(act.sum(1), torch.ones_like(act.sum(1)), atol=1e-5)


def test_inference_tail_fusion_is_transparent(rb):
    """no_grad forward: the decoders' last norm + act + head run as one pass (ops.conv_norm_act_head, no stored
    activation); switching the fusion off (the two-pass path the training forward uses) gives the same logits up to
    the run-to-run atomics noise of two forward passes (5e-3, as in test_eval_activation_semantics)."""
    model, mgr = _build(rb, "aniso_8x32x32")
    x = torch.rand(2, 1, 8, 32, 32, device="cuda")
    model.eval()
    with torch.no_grad():
        model(x)                                        # weight packs are built (and cached) by the first forward
    n0 = rb._lib.launch_count()
    with torch.no_grad():
        fused = {t: v.clone() for t, v in model(x).items()}
    n_fused = rb._lib.launch_count() - n0
    rb.ops.FUSE_HEAD = False
    try:
        n0 = rb._lib.launch_count()
        with torch.no_grad():
            plain = model(x)
        n_plain = rb._lib.launch_count() - n0
    finally:
        rb.ops.FUSE_HEAD = True
    assert n_fused == n_plain - len(mgr.tasks)          # one launch less per task decoder
    for t in mgr.tasks:
        assert rel_l2(fused[t], plain[t]) < 5e-3, t


def test_training_step_is_cuda_graph_capturable(rb):
    """bench.py replays the whole step (fwd + loss + bwd + clip + AdamW) as one CUDA graph: nothing on the path may
    synchronise, copy from pageable host memory or allocate index tensors on the host (weight packs included, which are
    rebuilt inside the graph when ops.PACK_CACHE is off).  Strided convs (merged data gradient) and the 32-channel
    full-resolution layers are part of this 32^3 network."""
    tasks = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
    torch.manual_seed(0)
    model = quiet_build(rb.NetworkFromConfig, make_mgr([32, 32, 32], tasks)).cuda().train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, capturable=True, fused=True)
    x = torch.rand(2, 1, 32, 32, 32, device="cuda")
    tg = {"sheet": (torch.rand(2, 1, 32, 32, 32, device="cuda") > 0.5).float(),
          "normals": torch.nn.functional.normalize(torch.randn(2, 3, 32, 32, 32, device="cuda"), dim=1)}
    params = [p for p in model.parameters()]

    def step():
        out = model(x)
        loss = sum(_loss(t, out[t], tg[t]) for t in tasks)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 3.0)
        opt.step()
        return loss

    rb.ops.PACK_CACHE = False
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            eager = float(step())
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            gl = step()
        losses = []
        for _ in range(3):
            g.replay()
            torch.cuda.synchronize()
            losses.append(float(gl))
    finally:
        rb.ops.PACK_CACHE = True
    print("eager", eager, "graph replays", losses)
    assert all(np.isfinite(losses)) and losses[0] <= eager + 1e-2 and losses[-1] < losses[0]


# ------------------------------------------------------------------------------------------
# split-precision ("bf16x3") inference tier: north star "within relative L2 1e-4 with fp32 accumulation",
# ">= 99.9 % argmax agreement"
# ------------------------------------------------------------------------------------------
PRECISE_TOL = 1e-4


@pytest.mark.parametrize("impl", ["mma", "auto"])
@pytest.mark.parametrize("case", [c for c in NET_CASES if NET_CASES[c][4] == "all"
                                  and not NET_EXTRA.get(c, {}).get("residual_decoder", False)])
def test_precise_tier_matches_reference_golden(rb, case, impl):
    """Eval-mode outputs of the drop-in under ops.precise_inference() against the UNMODIFIED reference's fp32
    outputs (tests/golden): rel-L2 < 1e-4 per task, threshold / argmax agreement >= 99.9 %."""
    model, mgr = _build(rb, case)
    gold = load_net_golden(case)
    x = torch.from_numpy(gold["x"]).cuda()
    model.eval()
    with rb.ops.precise_inference(impl=impl):
        ev = model(x)
    for t, info in mgr.tasks.items():
        ref = golden_eval(gold, t, info["activation"])
        assert ev[t].dtype == torch.float32 and tuple(ev[t].shape) == tuple(ref.shape)
        r = rel_l2(ev[t], ref)
        print(f"precise[{impl}] {case}/{t}: eval rel-L2 {r:.3e}")
        assert r < PRECISE_TOL, (case, t, r)
        if info["activation"] == "sigmoid":
            agree = float(((ev[t].cpu() > 0.5) == (ref > 0.5)).float().mean())
            assert agree >= 0.999, agree
        if info["activation"] == "softmax":
            agree = float((ev[t].cpu().argmax(1) == ref.argmax(1)).float().mean())
            assert agree >= 0.999, agree
    # the tier is a context: outside it the same module runs the bf16 path again
    with torch.no_grad():
        ev2 = model(x)
    t0 = next(iter(mgr.tasks))
    assert rel_l2(ev2[t0], golden_eval(gold, t0, mgr.tasks[t0]["activation"])) > PRECISE_TOL


def test_precise_tier_rejects_what_it_does_not_cover(rb):
    try:
        model, mgr = _build(rb, "ink_se23_16")      # SE squeeze over (2, 3): bf16 tier only
        x = torch.from_numpy(load_net_golden("ink_se23_16")["x"]).cuda()
        model.eval()
        with pytest.raises(NotImplementedError):
            with rb.ops.precise_inference():
                model(x)
    finally:
        _set_se_dims(rb, "all")


@pytest.mark.parametrize("impl", ["mma", "auto"])
def test_precise_tier_64_vs_oracle(rb, impl):
    """BASELINE config 1 geometry (64^3, sheet + normals, PyTorch default init): the split-precision forward agrees
    with the fp32 oracle to < 1e-4 rel-L2 and on >= 99.9 % of the thresholded voxels."""
    tasks = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
    torch.manual_seed(0)
    model = quiet_build(rb.NetworkFromConfig, make_mgr([64, 64, 64], tasks)).cuda().eval()
    x = torch.rand(1, 1, 64, 64, 64)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    topo = O.autoconfig([64, 64, 64])
    with torch.no_grad():
        ref = O.net_forward(sd, topo, x, tasks, training=False)
    with rb.ops.precise_inference(impl=impl):
        out = model(x.cuda())
    for t in tasks:
        r = rel_l2(out[t], ref[t])
        print(f"precise[{impl}] 64^3 {t}: rel-L2 {r:.3e}")
        assert r < PRECISE_TOL
    agree = float(((out["sheet"].cpu() > 0.5) == (ref["sheet"] > 0.5)).float().mean())
    print(f"precise 64^3 sheet threshold agreement {agree:.6f}")
    assert agree >= 0.999


@pytest.mark.parametrize("optimizer", ["SGD", "AdamW"])
def test_trainer_step_eager_and_graph(rb, optimizer):
    """training.DataParallelTrainer on one GPU: the CUDA-graph step equals the eager step, and
    capturing it costs no optimiser update (the first returned loss is the loss at the initial weights).  SGD curves
    coincide; AdamW's sign-like first updates amplify the atomics-order noise of near-zero gradients, so only its
    first steps are compared tightly."""
    from types import SimpleNamespace
    case = "sheet_normals_16"
    gold = load_net_golden(case)
    x = torch.from_numpy(gold["x"]).cuda()
    tgt = {t: torch.from_numpy(gold["target::" + t]).cuda() for t in ("sheet", "normals")}
    curves = {}
    for mode in ("eager", "graph"):
        model, mgr = _build(rb, case)
        tm = SimpleNamespace(tasks=mgr.tasks, optimizer=optimizer, initial_lr=0.05 if optimizer == "SGD" else 1e-3,
                             weight_decay=0.0, max_epoch=10)
        tr = rb.training.DataParallelTrainer(model, tm, use_cuda_graph=(mode == "graph"))
        curve = []
        for step in range(6):
            total, per = tr.train_step(x, tgt)
            curve.append(float(total.detach()))
            assert set(per) == {"sheet", "normals"}
        curves[mode] = curve
        rb._lib.device_error_check()
    print(f"trainer curves ({optimizer}):", {k: [round(v, 4) for v in c] for k, c in curves.items()})
    assert abs(curves["eager"][0] - float(gold["loss_total"])) < 1e-2
    assert abs(curves["graph"][0] - curves["eager"][0]) < 2e-3          # no hidden updates during capture
    assert min(curves["eager"][1:]) < curves["eager"][0]                # it learns
    # momentum SGD / Adam on one batch amplify the atomics-order noise of the CUDA path: first steps coincide, later
    # ones stay close (same bounds as test_training_loss_decreases_and_tracks_oracle)
    for a, b in zip(curves["eager"][:2], curves["graph"][:2]):
        assert abs(a - b) < 1e-2
    for a, b in zip(curves["eager"], curves["graph"]):
        assert abs(a - b) < 8e-2
    assert curves["graph"][-1] < 0.9 * curves["graph"][0] and curves["eager"][-1] < 0.9 * curves["eager"][0]
