"""Learning-rate schedule under whole-step CUDA graphs: the captured step must follow
CosineAnnealingLR (train.py:87-91) after end_epoch() and after a resume from a checkpoint, for AdamW (device-tensor
lr read by every replay) and SGD (float lr baked into the capture -> re-captured when the schedule moves it)."""
from types import SimpleNamespace

import pytest
import torch

from helpers import case_mgr, golden_state, load_net_golden, quiet_build, state_dict_from_params

pytestmark = pytest.mark.gpu


def _trainer(rb, optimizer, lr, graph=True):
    case = "sheet_normals_16"
    mgr, _ = case_mgr(case)
    model = quiet_build(rb.NetworkFromConfig, mgr)
    model.load_state_dict(state_dict_from_params(model, golden_state(case)))
    model = model.cuda()
    tm = SimpleNamespace(tasks=mgr.tasks, optimizer=optimizer, initial_lr=lr, weight_decay=0.0, max_epoch=2)
    return rb.training.DataParallelTrainer(model, tm, use_cuda_graph=graph), model


def _batch():
    gold = load_net_golden("sheet_normals_16")
    x = torch.from_numpy(gold["x"]).cuda()
    return x, {t: torch.from_numpy(gold["target::" + t]).cuda() for t in ("sheet", "normals")}


def _update_norm(tr, model, x, tgt):
    probe = dict(model.named_parameters())["shared_encoder.stages.1.blocks.0.conv1.conv.weight"]
    before = probe.detach().clone()
    tr.train_step(x, tgt)
    torch.cuda.synchronize()
    return float((probe.detach() - before).norm())


@pytest.mark.parametrize("optimizer,lr", [("AdamW", 1e-3), ("SGD", 1e-2)])
def test_graph_step_follows_the_schedule(rb, optimizer, lr):
    tr, model = _trainer(rb, optimizer, lr)
    x, tgt = _batch()
    for _ in range(3):
        tr.train_step(x, tgt)
    a = _update_norm(tr, model, x, tgt)
    tr.end_epoch()                                  # cosine, T_max 2: lr -> lr / 2
    cur = float(tr.optimizer.param_groups[0]["lr"])
    assert abs(cur - lr / 2) < 1e-6 * max(1.0, lr)
    tr.train_step(x, tgt)
    b = _update_norm(tr, model, x, tgt)
    print(f"{optimizer}: update norm before end_epoch {a:.4e}, after {b:.4e} (lr halved)")
    assert 0.25 * a < b < 0.8 * a, (a, b)
    rb._lib.device_error_check()


def test_graph_step_follows_the_schedule_after_resume(rb, tmp_path):
    tr, model = _trainer(rb, "AdamW", 1e-3)
    x, tgt = _batch()
    for _ in range(3):
        tr.train_step(x, tgt)
    a = _update_norm(tr, model, x, tgt)
    tr.end_epoch()
    path = str(tmp_path / "ck.pth")
    tr.save_checkpoint(path)
    tr2, model2 = _trainer(rb, "AdamW", 1e-3)
    tr2.load_checkpoint(path)
    g = tr2.optimizer.param_groups[0]
    assert torch.is_tensor(g["lr"]) and g["lr"].is_cuda, "a resumed capturable optimiser must keep its device lr tensor"
    assert abs(float(g["lr"]) - 5e-4) < 1e-9 and tr2.epoch == 1
    tr2.train_step(x, tgt)
    b = _update_norm(tr2, model2, x, tgt)
    print(f"resume: update norm at lr 1e-3 {a:.4e}, after resume at 5e-4 {b:.4e}")
    assert 0.25 * a < b < 0.8 * a, (a, b)
    # and a checkpoint written with a float lr (the reference's, train.py:249-254) resumes the same way
    ck = torch.load(path, weights_only=False)
    for pg in ck["optimizer"]["param_groups"]:
        pg["lr"] = float(pg["lr"])
    torch.save(ck, path)
    tr3, _ = _trainer(rb, "AdamW", 1e-3)
    tr3.load_checkpoint(path)
    g3 = tr3.optimizer.param_groups[0]
    assert torch.is_tensor(g3["lr"]) and g3["lr"].is_cuda and abs(float(g3["lr"]) - 5e-4) < 1e-9


def _managed_packs_match_weights(rb, model):
    """Every optimiser-managed operand buffer holds exactly the pack of the CURRENT parameter value."""
    n = 0
    for name, p in model.named_parameters():
        ent = getattr(p, "_rb_opt_packs", None)
        if ent is None or ent["ptr"] != p.data_ptr():
            continue
        f, d = rb.ops._pack_kernel(p, True, True)
        assert torch.equal(ent["f"], f) and torch.equal(ent["d"], d), name
        n += 1
    return n


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_with_optimizer_managed_packs(rb, graph):
    """ClippedAdamW(manage_packs=True) inside the trainer: the (captured) step reads operand packs that only the
    optimiser kernel and `refresh_packs()` (after the capture's parameter restore) write.  The invariant that matters
    is deterministic: after every step - eager or replayed, with a schedule step and an eager eval forward in between -
    each managed buffer equals the pack of the current weight, bit for bit.  (Loss curves of two runs of this 16^3
    network differ by ~1e-2 after a few Adam steps from the statistics atomics alone, so they only get a loose bound.)"""
    x, tgt = _batch()
    curves = {}
    for managed in (False, True):
        case = "sheet_normals_16"
        mgr, _ = case_mgr(case)
        model = quiet_build(rb.NetworkFromConfig, mgr)
        model.load_state_dict(state_dict_from_params(model, golden_state(case)))
        model = model.cuda()
        tm = SimpleNamespace(tasks=mgr.tasks, optimizer="AdamW", initial_lr=1e-3, weight_decay=1e-4, max_epoch=4)
        tr = rb.training.DataParallelTrainer(model, tm, use_cuda_graph=graph, manage_packs=managed)
        assert tr.optimizer.manage_packs == managed
        cur = []
        for step in range(6):
            total, _ = tr.train_step(x, tgt)
            cur.append(float(total))
            torch.cuda.synchronize()
            n = _managed_packs_match_weights(rb, model)
            assert (n > 0) == managed, (managed, step, n)        # the fused update really ran for the conv weights
            if step == 2:
                tr.end_epoch()
                model.eval()
                with torch.no_grad():
                    model(x)                       # an eager forward between replays must not disturb the managed packs
                _managed_packs_match_weights(rb, model)
        curves[managed] = cur
    dev = max(abs(a - b) for a, b in zip(curves[True], curves[False]))
    print(f"graph={graph}: managed packs vs pack kernel, max loss deviation {dev:.2e}; {curves[True]}")
    assert dev < 5e-2
    assert curves[True][-1] < curves[True][0]
    rb._lib.device_error_check()
