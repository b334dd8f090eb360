"""Pins the oracle (oracle/resenc_oracle.py) against fixtures generated from the UNMODIFIED
reference code (oracle/make_golden.py).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from helpers import (GOLDEN, NET_CASES, golden_eval, golden_state, load_keys, load_net_golden, oracle_kwargs, oracle_topology,
                     rel_l2)
from oracle import resenc_oracle as O


def sha16(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


with open(os.path.join(GOLDEN, "host_goldens.json")) as f:
    HOST = json.load(f)


@pytest.mark.parametrize("rec", HOST["positions"], ids=lambda r: "x".join(map(str, r["vol"])))
def test_oracle_positions(rec):
    pos = O.all_positions(rec["vol"], rec["patch"], rec["overlap"])
    table = np.array(pos, dtype=np.int64)
    assert len(pos) == rec["count"]
    assert sha16(table) == rec["sha1"]


@pytest.mark.parametrize("rec", HOST["gaussian"], ids=lambda r: "x".join(map(str, r["tile"])))
def test_oracle_gaussian(rec):
    g = O.gaussian_map(rec["tile"])
    assert sha16(g) == rec["sha1"]
    assert float(g.max()) == rec["max"] and float(g.min()) == rec["min"]


@pytest.mark.parametrize("rec", HOST["topology"], ids=lambda r: "x".join(map(str, r["patch"])))
def test_oracle_topology(rec):
    npool, strides, kernels = O.pool_and_conv_props(rec["patch"])
    assert list(npool) == rec["num_pool"]
    assert [list(s) for s in strides] == rec["strides"]
    assert [list(k) for k in kernels] == rec["kernels"]
    assert O.blocks_per_stage(len(strides)) == rec["blocks"]


@pytest.mark.parametrize("case", list(NET_CASES))
def test_oracle_network_forward_and_loss(case):
    patch, cin, tasks, mc, rd, batch = NET_CASES[case]
    gold = load_net_golden(case)
    sd_unique = golden_state(case)
    # the oracle walks reference key names; build the aliased dict the reference's state_dict has
    sd = dict(sd_unique)
    topo = oracle_topology(case)
    kw = oracle_kwargs(case)
    x = torch.from_numpy(gold["x"])
    se = bool(mc.get("squeeze_excitation", False))
    drop = {k[6:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("drop::")} or None
    assert (drop is not None) == bool(mc.get("stochastic_depth_p", 0.0))
    with torch.no_grad():
        out_t = O.net_forward(sd, topo, x, tasks, training=True, se=se, reduce_dims=rd, drop=drop, **kw)
        out_e = O.net_forward(sd, topo, x, tasks, training=False, se=se, reduce_dims=rd, **kw)
    total = 0.0
    for t in tasks:
        assert rel_l2(out_t[t], gold["train::" + t]) < 2e-5, (case, t)
        assert rel_l2(out_e[t], golden_eval(gold, t, tasks[t]["activation"])) < 2e-5, (case, t)
        tgt = torch.from_numpy(gold["target::" + t])
        l = O.masked_cosine_loss(out_t[t], tgt) if t == "normals" else O.bce_dice_loss(out_t[t], tgt)
        assert abs(float(l) - float(gold["loss::" + t])) < 2e-5
        total += float(l)
    assert abs(total - float(gold["loss_total"])) < 5e-5


def test_oracle_param_census():
    for case in NET_CASES:
        keys = load_keys(case)
        assert len(keys["parameters"]) > 0 and len(keys["state_dict"]) >= len(keys["parameters"])


def test_oracle_blend_roundtrip():
    """uniform blend of a constant field returns the constant (count normalisation, :207-210)."""
    targets = {"sheet": {"channels": 1}, "normals": {"channels": 3}}
    vol, patch = (24, 20, 28), (16, 16, 16)
    pos = O.all_positions(vol, patch, 0.5)
    rng = np.random.default_rng(0)
    preds = {"sheet": np.full((len(pos), 1, *patch), 0.5, np.float32),
             "normals": np.tile(np.array([0.0, 0.6, 0.8], np.float32).reshape(1, 3, 1, 1, 1), (len(pos), 1, *patch))}
    sums, counts = O.blend_reference(preds, pos, vol, targets)
    assert counts["sheet"].min() >= 1
    out = O.finalize_reference(sums, counts, targets)
    assert out["sheet"].dtype == np.uint8 and (out["sheet"] == 127).all()
    assert out["normals"].dtype == np.uint16
    exp = ((np.array([0.0, 0.6, 0.8], np.float32) + 1) / 2 * 65535)
    got = out["normals"][:, 3, 3, 3].astype(np.float64)
    assert np.all(np.abs(got - exp) <= 1.0)


def test_whole_net_sanity_value_of_the_survey():
    """SURVEY 8(c) "whole-net sanity": torch.manual_seed(0); build the 64^3 two-task network; x = torch.rand(1,1,64^3);
    sheet target (rand > 0.8), normals target normalize(randn); BCEDice + MaskedCosine = 1.66644 on the reference
    (torch 2.11 CPU).  Reproducing it needs (a) the drop-in's constructors to consume the RNG exactly like the
    reference's (same module creation order, PyTorch default init) and (b) the oracle's forward and losses."""
    import contextlib
    import io
    from types import SimpleNamespace

    import resenc_b200 as rb
    tasks = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
    mgr = SimpleNamespace(tasks=tasks, train_patch_size=[64] * 3, train_batch_size=1, in_channels=1, vram_max=16.0,
                          autoconfigure=True, model_config={})
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = rb.NetworkFromConfig(mgr)
    x = torch.rand(1, 1, 64, 64, 64)
    ts = (torch.rand(1, 1, 64, 64, 64) > 0.8).float()
    tn = torch.nn.functional.normalize(torch.randn(1, 3, 64, 64, 64), dim=1)
    assert abs(sum(p.numel() for p in model.parameters()) / 1e6 - 118.09) < 0.01          # SURVEY topology row
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    topo = O.autoconfig([64] * 3)
    with torch.no_grad():
        out = O.net_forward(sd, topo, x, tasks, training=True)
    loss = O.bce_dice_loss(out["sheet"], ts) + O.masked_cosine_loss(out["normals"], tn)
    assert abs(float(loss) - 1.66644) < 1e-4
    assert abs(float(out["sheet"].mean()) - 0.22254) < 1e-4 and abs(float(out["sheet"].std()) - 0.31666) < 1e-4
    assert abs(float(out["normals"].mean()) + 0.28349) < 1e-4 and abs(float(out["normals"].std()) - 0.38100) < 1e-4
