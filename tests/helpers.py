"""Shared test helpers: seeded weights / inputs (same generators as oracle/make_golden.py), metrics."""
import contextlib
import io
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle.make_golden import (NET_CASES, NET_EXTRA, oracle_kwargs, oracle_topology, seeded_inputs,  # noqa: E402,F401
                                seeded_state)                                                       # (test infrastructure)


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten().cpu()
    b = torch.as_tensor(b).double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def make_mgr(patch, tasks, in_channels=1, batch=1, model_config=None, autoconfigure=True):
    return SimpleNamespace(tasks=tasks, train_patch_size=list(patch), train_batch_size=batch,
                           in_channels=in_channels, vram_max=16.0, autoconfigure=autoconfigure,
                           model_config=dict(model_config or {}))


def case_mgr(case):
    patch, cin, tasks, mc, rd, batch = NET_CASES[case]
    auto = NET_EXTRA.get(case, {}).get("autoconfigure", True)
    return make_mgr(patch, tasks, in_channels=cin, batch=batch, model_config=mc, autoconfigure=auto), rd


def golden_eval(gold, task, activation):
    """Reference eval-mode output of `task`: stored, or (fixtures without `eval::`, i.e. no stochastic depth) the head
    activation of build_network_from_config.py:322-323 applied to the stored training-mode logits."""
    if "eval::" + task in gold.files:
        return torch.from_numpy(gold["eval::" + task])
    logits = torch.from_numpy(gold["train::" + task])
    act = str(activation).lower()
    return torch.sigmoid(logits) if act == "sigmoid" else torch.softmax(logits, 1) if act == "softmax" else logits


def load_keys(case):
    with open(os.path.join(GOLDEN, f"keys_{case}.json")) as f:
        return json.load(f)


def load_net_golden(case):
    return np.load(os.path.join(GOLDEN, f"net_{case}.npz"))


def quiet_build(ctor, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return ctor(*a, **k)


def golden_state(case):
    keys = load_keys(case)
    return seeded_state([(n, tuple(s)) for n, s in keys["parameters"]], seed=7)


def state_dict_from_params(model, params):
    """Full (aliased) state_dict from unique named parameters."""
    sd = {}
    named = dict(model.named_parameters())
    by_id = {id(p): n for n, p in named.items()}
    for k, v in model.state_dict(keep_vars=True).items():
        sd[k] = params[by_id[id(v)]]
    return sd
