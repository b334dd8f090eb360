"""optim.ClippedAdamW: clip_grad_norm_(3) + torch.optim.AdamW.step() (reference train.py:79-83,227-228) as two
multi-tensor passes of the library.  CPU part: hyper-parameter validation, state_dict interchange with torch.optim.AdamW
and the float64 restatement used as the checker; GPU part: the kernels against PyTorch's own clip + fused AdamW."""
import math
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import optim_oracle  # noqa: E402  (checker only)


def _mk(shapes, seed, device):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g).to(device) for s in shapes]


def test_hyperparameter_validation_and_state_dict_keys(rb):
    O = rb.optim
    p = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(NotImplementedError):
        O.ClippedAdamW([p], amsgrad=True)
    with pytest.raises(ValueError):
        O.ClippedAdamW([p], betas=(1.0, 0.9))
    with pytest.raises(ValueError):
        O.ClippedAdamW([p], max_grad_norm=0.0)
    opt = O.ClippedAdamW([p], lr=2e-3, weight_decay=1e-4, max_grad_norm=3.0)
    ref = torch.optim.AdamW([p], lr=2e-3, weight_decay=1e-4)
    mine, theirs = opt.state_dict()["param_groups"][0], ref.state_dict()["param_groups"][0]
    assert set(theirs) <= set(mine)                                      # every key torch.optim.AdamW writes is there
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad", "maximize"):
        assert mine[k] == theirs[k], k
    # a CPU parameter is refused loudly (there is no fallback path)
    p.grad = torch.ones(4)
    with pytest.raises(NotImplementedError):
        opt.step()
    # a torch.optim.AdamW checkpoint loads (and the other way round)
    ref.step()
    opt.load_state_dict(ref.state_dict())
    assert set(opt.state[p]) == {"step", "exp_avg", "exp_avg_sq"}
    ref.load_state_dict(opt.state_dict())


def test_reference_step_is_clip_then_adamw(rb):
    """The float64 restatement the GPU test checks against == clip_grad_norm_ + torch.optim.AdamW on the CPU."""
    O = rb.optim
    shapes = [(7, 5), (33,), (2, 3, 3, 3, 3)]
    ps = [torch.nn.Parameter(t.double()) for t in _mk(shapes, 1, "cpu")]
    ref = torch.optim.AdamW(ps, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.05)
    mine_p = [p.detach().clone() for p in ps]
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    for step in range(1, 5):
        gs = [t.double() * (4.0 if step % 2 else 0.05) for t in _mk(shapes, 10 + step, "cpu")]     # clipped / not clipped
        for p, g in zip(ps, gs):
            p.grad = g.clone()
        total_ref = torch.nn.utils.clip_grad_norm_(ps, 3.0)
        ref.step()
        out, total = optim_oracle.reference_step(mine_p, gs, m, v, step, 3e-3, (0.9, 0.99), 1e-8, 0.05, 3.0)
        mine_p, m, v = [o[0] for o in out], [o[1] for o in out], [o[2] for o in out]
        assert abs(total - float(total_ref)) < 1e-9 * total
        for a, b in zip(mine_p, ps):
            assert torch.allclose(a, b.detach(), rtol=1e-12, atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("max_norm", [3.0, None])
def test_clipped_adamw_matches_torch(rb, max_norm):
    """Five steps on tensors that cover the vector path, its scalar tail, a misaligned view (gradient bucket slot), a
    tensor larger than one launch chunk's grid and > 48 tensors (two launches); gradients alternate between clipped
    and unclipped.  Bound: 2e-6 relative to PyTorch's clip_grad_norm_ + AdamW(fused) (both fp32; PyTorch evaluates
    the scalar coefficients in double), 1e-6 against the float64 restatement."""
    O = rb.optim
    dev = "cuda"
    shapes = [(512, 64, 3, 3, 3), (1 << 20,), (33, 7), (5,), (1,), (64, 32, 1, 1, 1)] + [(17 + i,) for i in range(50)]
    init = _mk(shapes, 2, dev)
    big = torch.zeros(4 * 1000 + 3, device=dev)
    ps = [torch.nn.Parameter(t.clone()) for t in init]
    qs = [torch.nn.Parameter(t.clone()) for t in init]
    opt = O.ClippedAdamW(ps, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=max_norm)
    ref = torch.optim.AdamW(qs, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, fused=True)
    p64 = [t.double() for t in init]
    m64 = [torch.zeros_like(t) for t in p64]
    v64 = [torch.zeros_like(t) for t in p64]
    for step in range(1, 6):
        gs = [t * (3.0 if step % 2 else 1e-3) for t in _mk(shapes, 20 + step, dev)]
        for i, (p, q, g) in enumerate(zip(ps, qs, gs)):
            if i == 2:      # a gradient that lives at a 4-byte-aligned offset of a flat buffer (DDP bucket slot)
                slot = big[1:1 + g.numel()].view_as(g)
                slot.copy_(g)
                p.grad = slot
            else:
                p.grad = g.clone()
            q.grad = g.clone()
        before = [p.grad.clone() for p in ps]
        opt.step()
        if max_norm is not None:
            total_ref = torch.nn.utils.clip_grad_norm_(qs, max_norm)
            assert abs(float(opt.last_grad_norm()) - float(total_ref)) < 1e-5 * float(total_ref)
        ref.step()
        for p, b in zip(ps, before):
            assert torch.equal(p.grad, b)                    # gradients are read, never rescaled in place
        out, _ = optim_oracle.reference_step(p64, [g.double() for g in gs], m64, v64, step, 2e-3, (0.9, 0.999), 1e-8, 1e-2, max_norm)
        p64, m64, v64 = [o[0] for o in out], [o[1] for o in out], [o[2] for o in out]
        for i, (p, q, e) in enumerate(zip(ps, qs, p64)):
            scale = float(e.abs().max()) + 1e-12
            assert float((p.detach().double() - e).abs().max()) < 1e-6 * scale + 1e-7, (step, i)
            assert float((p.detach() - q.detach()).abs().max()) < 2e-6 * scale + 2e-7, (step, i)
            for key in ("exp_avg", "exp_avg_sq"):      # absolute bound: the lerp cancels for entries near zero
                a, b = opt.state[p][key], ref.state[q][key]
                assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max()) + 1e-12, (step, i, key)
        assert float(opt.state[ps[0]]["step"]) == step
    rb._lib.device_error_check()


@pytest.mark.gpu
def test_clipped_adamw_under_cuda_graph_follows_a_tensor_lr(rb):
    """A captured step reads lr and the step count from device memory: replays keep counting, and filling the lr
    tensor changes the size of the next update."""
    O = rb.optim
    p = torch.nn.Parameter(torch.ones(1024, device="cuda"))
    lr = torch.tensor(1e-2, device="cuda")
    opt = O.ClippedAdamW([p], lr=lr, weight_decay=0.0, max_grad_norm=None)
    p.grad = torch.ones_like(p)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        opt.step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        opt.step()
    a = p.detach().clone()
    g.replay()
    b = p.detach().clone()
    lr.fill_(1e-3)
    g.replay()
    c = p.detach().clone()
    torch.cuda.synchronize()
    assert float(opt.state[p]["step"]) == 3.0                 # one eager step + two replays (capturing runs nothing)
    d1, d2 = float((a - b).abs().mean()), float((b - c).abs().mean())
    assert abs(d1 - 1e-2) < 1e-4 and abs(d2 - 1e-3) < 1e-5     # constant gradient: every Adam update is lr


@pytest.mark.gpu
def test_fused_update_writes_next_steps_operand_packs(rb):
    """rb_adamw_clip_pack_step == rb_adamw_clip_step followed by rb_pack_conv_weights, bit for bit (weights, moments and
    both bf16 operand layouts), including a Cout that is not a multiple of the 32-row tile and a 1x3x3 kernel."""
    O, ops = rb.optim, rb.ops
    for shape in ((64, 32, 3, 3, 3), (40, 64, 3, 3, 3), (32, 32, 1, 3, 3), (16, 96, 1, 1, 1)):
        torch.manual_seed(sum(shape))
        w0 = torch.randn(shape, device="cuda") * 0.1
        a, b = torch.nn.Parameter(w0.clone()), torch.nn.Parameter(w0.clone())
        ops.pack_conv_fprop(a)                                  # the network consumes `a` as a packed operand
        assert getattr(a, "_rb_wants_fd", False)
        fused = O.ClippedAdamW([a], lr=1e-2, weight_decay=1e-2, max_grad_norm=1.0, manage_packs=True)
        plain = O.ClippedAdamW([b], lr=1e-2, weight_decay=1e-2, max_grad_norm=1.0)
        for step in range(3):
            g = torch.randn(shape, device="cuda") * (2.0 if step != 1 else 1e-3)
            a.grad, b.grad = g.clone(), g.clone()
            plain.step()
            fused.step()                                         # last: any optimiser step opens a new pack epoch
            assert torch.equal(a.detach(), b.detach()), (shape, step)
            for key in ("exp_avg", "exp_avg_sq"):
                assert torch.equal(fused.state[a][key], plain.state[b][key]), (shape, step, key)
            ent = ops.opt_packs(a)
            assert ent is not None                               # fresh for the epoch the step opened
            f_ref, d_ref = ops._pack_kernel(b, True, True)
            assert torch.equal(ent["f"], f_ref) and torch.equal(ent["d"], d_ref), (shape, step)
            assert ops.pack_conv_fprop(a) is ent["f"] and ops.pack_conv_dgrad_full(a) is ent["d"]
            if step == 1:                                        # somebody else's optimiser step: stale, pack kernel again
                other = torch.nn.Parameter(torch.zeros(8, device="cuda"))
                other.grad = torch.ones_like(other)
                torch.optim.SGD([other], lr=0.1).step()
                assert ops.opt_packs(a) is None
                assert torch.equal(ops.pack_conv_fprop(a), f_ref)
        # anything that can change the parameter behind the optimiser's back makes the entry stale ...
        with torch.no_grad():
            a.mul_(0.5)
        assert ops.opt_packs(a) is None
        assert torch.equal(ops.pack_conv_fprop(a), ops._pack_kernel(a, True, False)[0])
        # ... and refresh_packs() re-packs into the SAME buffers (what a captured graph keeps reading)
        f_buf = a._rb_opt_packs["f"]
        fused.refresh_packs()
        ent = ops.opt_packs(a)
        assert ent is not None and ent["f"] is f_buf and torch.equal(ent["f"], ops._pack_kernel(a, True, False)[0])
        ops.invalidate_weight_packs()
        assert ops.opt_packs(a) is None
    rb._lib.device_error_check()


@pytest.mark.gpu
def test_training_with_managed_packs_tracks_the_pack_kernel_path(rb):
    """Ten eager steps of a small network: ClippedAdamW(manage_packs=True) (no pack kernel launches after the first
    step) against manage_packs=False.  Same kernels otherwise; the bit-level equivalence is the test above, here the
    launch count must drop and the curves must stay within the run-to-run noise of the statistics atomics."""
    from helpers import make_mgr, quiet_build
    losses = {}
    launches = {}
    for managed in (False, True):
        torch.manual_seed(0)
        mgr = make_mgr((32, 32, 32), {"sheet": {"channels": 1, "activation": "none"}, "normals": {"channels": 3, "activation": "none"}})
        model = quiet_build(rb.NetworkFromConfig, mgr).cuda().train()
        opt = rb.optim.ClippedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, max_grad_norm=3.0, manage_packs=managed)
        g = torch.Generator().manual_seed(1)
        x = torch.rand(2, 1, 32, 32, 32, generator=g).cuda()
        ts = (torch.rand(2, 1, 32, 32, 32, generator=g) > 0.5).float().cuda()
        tn = torch.nn.functional.normalize(torch.randn(2, 3, 32, 32, 32, generator=g), dim=1).cuda()
        crit = rb.losses.task_losses(mgr.tasks)
        cur = []
        for step in range(10):
            if step == 9:
                n0 = rb._lib.launch_count()
            out = model(x)
            loss = crit["sheet"](out["sheet"], ts) + crit["normals"](out["normals"], tn)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            cur.append(float(loss))
        launches[managed] = rb._lib.launch_count() - n0
        losses[managed] = cur
        n_managed = 0
        for name, p in model.named_parameters():
            ent = rb.ops.opt_packs(p)
            if ent is not None:                                   # fresh after the last step, and exactly pack(weight)
                f, d = rb.ops._pack_kernel(p, True, True)
                assert torch.equal(ent["f"], f) and torch.equal(ent["d"], d), name
                n_managed += 1
        assert (n_managed > 0) == managed
    assert launches[True] <= launches[False]                     # one fused update per weight replaces one pack launch
    dev = max(abs(a - b) for a, b in zip(losses[True], losses[False]))
    print("managed-pack loss curve deviation", dev, losses[True][-1], losses[False][-1])
    assert dev < 5e-2            # two runs of one configuration already differ by ~1e-2 (atomics order + Adam's sign-like steps)
    assert losses[True][-1] < losses[True][0]
