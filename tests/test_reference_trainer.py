"""The reference's own trainer on top of the drop-in ("train.py/BaseTrainer run unchanged"; VERDICT r1
"missing" #6) and the torch.library custom-op layer (north star; VERDICT row N1)."""
import collections
import os

import pytest
import torch

import reference_stubs as RS
from helpers import make_mgr, quiet_build, rel_l2

needs_reference = pytest.mark.skipif(RS.reference_root() is None, reason="no reference tree (oracle/_ref not built)")

TASKS2 = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}


@pytest.fixture
def stubs(tmp_path):
    yield tmp_path
    RS.uninstall()


@needs_reference
def test_reference_trainer_builds_the_drop_in(rb, stubs):
    """BaseTrainer._build_model / _build_loss / _get_optimizer / _get_scheduler (train.py:29-91) through
    `install_as_builders()`: the model class is the drop-in, its state_dict has the reference's keys and shapes."""
    import json
    from helpers import GOLDEN
    cfg = RS.make_config(stubs)
    train, trainer = RS.load_reference_trainer(cfg)
    import builders.build_network_from_config as bn
    assert bn.NetworkFromConfig is rb.NetworkFromConfig and train.NetworkFromConfig is rb.NetworkFromConfig
    model = quiet_build(trainer._build_model)
    assert type(model) is rb.NetworkFromConfig
    with open(os.path.join(GOLDEN, "keys_sheet_normals_16.json")) as f:
        ref_keys = json.load(f)["state_dict"]
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert got == ref_keys                               # the UNMODIFIED reference's key census for this config
    losses = trainer._build_loss()
    assert type(losses["sheet"]).__name__ == "BCEDiceLoss" and type(losses["normals"]).__name__ == "MaskedCosineLoss"
    assert type(losses["sheet"]).__module__ == "training.losses.losses"      # the reference's class, not ours
    opt = trainer._get_optimizer(model)
    assert isinstance(opt, torch.optim.AdamW) and len(opt.param_groups[0]["params"]) == len(list(model.parameters()))
    assert trainer._get_scheduler(opt).T_max == cfg.max_epoch


def test_torch_compile_sees_custom_ops():
    """Dynamo traces NetworkFromConfig.forward without a graph break (strict export fails on one) and the fused units
    show up as opaque `resenc_b200::*` operator nodes (fake kernels: no GPU needed)."""
    import resenc_b200 as rb
    model = quiet_build(rb.NetworkFromConfig, make_mgr([32, 32, 32], TASKS2, batch=2))
    model.train()
    ep = torch.export.export(model, (torch.rand(2, 1, 32, 32, 32),), strict=True)
    cnt = collections.Counter()

    def walk(gm):
        for n in gm.graph.nodes:
            if n.op == "call_function":
                cnt[str(n.target)] += 1
            if n.op == "get_attr" and isinstance(getattr(gm, n.target, None), torch.fx.GraphModule):
                walk(getattr(gm, n.target))
    walk(ep.graph_module)
    # 32^3 autoconfig: 4 stages, blocks [1, 3, 4, 6]: stem + 28 block convs + 3 projection skips + 2 x 3 decoder convs
    assert cnt["resenc_b200.conv_norm_act.default"] == 38
    assert cnt["resenc_b200.conv_transpose3d.default"] == 6 and cnt["resenc_b200.avg_pool3d.default"] == 3
    assert cnt["resenc_b200.head_conv1x1.default"] == 2
    for name in rb.ops.custom_ops.OPERATORS:
        assert hasattr(torch.ops.resenc_b200, name)
    # fake kernels propagate the channels-last layout and the fp32 logits
    outs = ep.graph_module.graph.output_node().args[0]
    metas = [o.meta["val"] for o in outs if hasattr(o, "meta") and "val" in o.meta]
    assert all(m.dtype == torch.float32 for m in metas[-2:])


def test_custom_op_backward_graph_under_fake_tensors(monkeypatch):
    """Forward + backward of three network variants through the custom ops with FakeTensors (CPU, no kernels run): every
    parameter that the eager path gives a gradient gets one of its own shape."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    import resenc_b200 as rb
    monkeypatch.setattr(rb.ops, "FORCE_CUSTOM_OPS", True)
    manual = dict(features_per_stage=[32, 64, 128], num_stages=3, n_blocks_per_stage=[1, 2, 2], kernel_sizes=[[3, 3, 3]] * 3,
                  n_conv_per_stage_decoder=[2, 1], strides=[[1, 1, 1], [2, 2, 2], [2, 2, 2]], basic_encoder_block="BasicBlockD",
                  bottleneck_block="BasicBlockD", basic_decoder_block="ResidualBlock")
    for mc, auto, unused in (({}, True, 8),
                             ({"squeeze_excitation": True, "stochastic_depth_p": 0.2, "conv_bias": True,
                               "norm_op_kwargs": {"affine": True, "eps": 1e-5}}, True, 8), (manual, False, 4)):
        with FakeTensorMode():
            model = quiet_build(rb.NetworkFromConfig, make_mgr([32, 32, 32], TASKS2, batch=2, model_config=mc, autoconfigure=auto))
            model.train()
            out = model(torch.rand(2, 1, 32, 32, 32))
            sum(v.mean() for v in out.values()).backward()
            params = list(model.parameters())
            assert sum(p.grad is None for p in params) == unused          # the never-used deep-supervision heads
            assert all(p.grad.shape == p.shape for p in params if p.grad is not None)


@pytest.mark.gpu
def test_custom_ops_opcheck_and_match_eager_functions(rb):
    """torch.library.opcheck (schema, fake kernel, autograd registration, AOT dispatch) on every forward operator, and the
    custom-op route gives the eager autograd.Functions' results (same kernels; two runs differ by atomics order)."""
    torch.manual_seed(0)
    dev = "cuda"
    x = rb.ops.as_cl(torch.randn(2, 32, 8, 8, 8, device=dev))
    w = (torch.randn(32, 32, 3, 3, 3, device=dev) * 0.05).requires_grad_(True)
    res = rb.ops.as_cl(torch.randn(2, 32, 8, 8, 8, device=dev))
    ops = torch.ops.resenc_b200
    args = (x.detach().requires_grad_(True), w, None, res.detach().requires_grad_(True), None, None, None, None, None, None, None,
            [1, 1, 1], 1e-5, True, 0.01, "all", False, "")
    # test_autograd_registration / fake / schema; gradcheck-style numerical tests are meaningless for bf16 kernels
    torch.library.opcheck(ops.conv_norm_act.default, args, test_utils=("test_schema", "test_autograd_registration", "test_faketensor"))
    wt = (torch.randn(32, 16, 2, 2, 2, device=dev) * 0.1).requires_grad_(True)
    torch.library.opcheck(ops.conv_transpose3d.default, (x, wt, [2, 2, 2], ""),
                          test_utils=("test_schema", "test_autograd_registration", "test_faketensor"))
    torch.library.opcheck(ops.avg_pool3d.default, (x, [2, 2, 2]), test_utils=("test_schema", "test_autograd_registration", "test_faketensor"))
    hw = torch.randn(3, 32, 1, 1, 1, device=dev, requires_grad=True)
    hb = torch.randn(3, device=dev, requires_grad=True)
    torch.library.opcheck(ops.head_conv1x1.default, (x, hw, hb, 0), test_utils=("test_schema", "test_autograd_registration", "test_faketensor"))
    # whole network: eager Functions vs forced custom ops
    model = quiet_build(rb.NetworkFromConfig, make_mgr([16, 16, 16], TASKS2, batch=2, model_config={"squeeze_excitation": True})).cuda().train()
    xin = torch.rand(2, 1, 16, 16, 16, device=dev)

    def run():
        for p in model.parameters():
            p.grad = None
        out = model(xin)
        (out["sheet"].mean() + out["normals"].square().mean()).backward()
        return {k: v.detach().clone() for k, v in out.items()}, {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    o1, g1 = run()
    o1b, g1b = run()                       # run-to-run spread of one route (fp32 atomics order in statistics / split-K)
    rb.ops.FORCE_CUSTOM_OPS = True
    try:
        o2, g2 = run()
    finally:
        rb.ops.FORCE_CUSTOM_OPS = False
    for t in o1:
        assert rel_l2(o2[t], o1[t]) < 5e-3, t
    assert set(g1) == set(g2)
    names = [n for n in g1 if float(g1[n].norm()) > 1e-6]
    noise = max(rel_l2(g1b[n], g1[n]) for n in names)
    worst = max(rel_l2(g2[n], g1[n]) for n in names)
    print(f"custom-op route vs autograd.Function route: worst gradient rel-L2 {worst:.3e} (two runs of one route: {noise:.3e})")
    assert worst < max(0.05, 3.0 * noise)
    rb._lib.device_error_check()


@pytest.mark.gpu
def test_torch_compile_forward_backward_matches_eager(rb):
    """torch.compile(model) as the reference wraps it (train.py:133), under fp16 autocast with a GradScaler
    (train.py:94-96,203,224-230): outputs and the first optimiser steps match the eager drop-in."""
    torch.manual_seed(0)
    mgr = make_mgr([16, 16, 16], TASKS2, batch=2)
    eager = quiet_build(rb.NetworkFromConfig, mgr).cuda().train()
    comp_base = quiet_build(rb.NetworkFromConfig, mgr).cuda().train()
    comp_base.load_state_dict(eager.state_dict())
    compiled = torch.compile(comp_base)
    x = torch.rand(2, 1, 16, 16, 16, device="cuda")
    tgt = (torch.rand(2, 1, 16, 16, 16, device="cuda") > 0.8).float()
    scaler = torch.amp.GradScaler("cuda")
    oe = torch.optim.SGD(eager.parameters(), lr=0.05)
    oc = torch.optim.SGD(comp_base.parameters(), lr=0.05)
    le, lc = [], []
    for step in range(3):
        out_e = eager(x)
        l = torch.nn.functional.binary_cross_entropy_with_logits(out_e["sheet"], tgt) + out_e["normals"].square().mean()
        oe.zero_grad(set_to_none=True)
        l.backward()
        oe.step()
        le.append(float(l))
        with torch.amp.autocast("cuda"):
            out_c = compiled(x)
            l2 = torch.nn.functional.binary_cross_entropy_with_logits(out_c["sheet"], tgt) + out_c["normals"].square().mean()
        assert out_c["sheet"].dtype == torch.float32
        oc.zero_grad(set_to_none=True)
        scaler.scale(l2).backward()
        scaler.step(oc)
        scaler.update()
        lc.append(float(l2))
        if step == 0:
            for t in out_e:
                assert rel_l2(out_c[t], out_e[t]) < 5e-3, t
    print("eager   :", le, "\ncompiled:", lc)
    assert all(abs(a - b) < 2e-2 for a, b in zip(le, lc)) and lc[-1] < lc[0]
    rb._lib.device_error_check()


@pytest.mark.gpu
@needs_reference
def test_reference_trainer_runs_unchanged_on_the_drop_in(rb, stubs):
    """BaseTrainer.train() of the reference's train.py, unmodified: torch.compile(model) (:133), fp16 autocast (:203),
    GradScaler (:224-229), clip_grad_norm_ (:227), AdamW, CosineAnnealingLR, checkpoint (:249-254), validation in eval
    mode (:268-320) - one epoch of three steps on synthetic 16^3 batches."""
    cfg = RS.make_config(stubs, max_steps_per_epoch=3)
    train, trainer = RS.load_reference_trainer(cfg)
    cwd = os.getcwd()
    os.chdir(stubs)                       # train.py writes `<model_name>_final.pth` into the working directory (:338)
    try:
        l0 = rb._lib.launch_count()
        trainer.train()
        launched = rb._lib.launch_count() - l0
    finally:
        os.chdir(cwd)
    assert launched > 300, "the reference loop must have run on the sm_100a kernels"
    ck = torch.load(cfg.ckpt_out_base / "Model_1.pth", weights_only=False)
    assert set(ck) == {"model", "optimizer", "scheduler", "epoch"} and ck["epoch"] == 0
    keys = list(ck["model"].keys())
    assert all(k.startswith("_orig_mod.") for k in keys)          # torch.compile'd module, as in the reference
    final = torch.load(stubs / "Model_final.pth", weights_only=False)
    m = quiet_build(rb.NetworkFromConfig, cfg)
    m.load_state_dict(rb.training.strip_compile_prefix(final))
    assert all(torch.isfinite(p).all() for p in m.parameters())
    rb._lib.device_error_check()
