"""The "existing GPU path" comparison SURVEY 8(d) asks for: the same network (BASELINE config 2: 128^3, batch 2,
sheet + normals, 6 stages) run by PyTorch eager under bf16 autocast with channels_last_3d tensors - i.e. ATen / cuDNN
kernels driven by the oracle's functional restatement of the reference - against the drop-in, forward + losses +
backward, same weights and inputs, CUDA events after warm-up.  Prints both times (recorded in profiles/README.md) and
asserts the hand-written path is the faster one and that the two agree."""
import pytest
import torch

from helpers import make_mgr, quiet_build, rel_l2
from oracle import resenc_oracle as O   # baseline / checker only

pytestmark = pytest.mark.gpu


def _time(fn, warm=2, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def test_faster_than_pytorch_cudnn_autocast(rb):
    P, B = 128, 2
    tasks = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
    torch.manual_seed(0)
    model = quiet_build(rb.NetworkFromConfig, make_mgr([P, P, P], tasks, batch=B)).cuda().train()
    x = torch.rand(B, 1, P, P, P, device="cuda")
    tg = {"sheet": (torch.rand(B, 1, P, P, P, device="cuda") > 0.8).float(),
          "normals": torch.nn.functional.normalize(torch.randn(B, 3, P, P, P, device="cuda"), dim=1)}
    crit = rb.losses.task_losses(tasks)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    topo = O.autoconfig([P, P, P])
    xcl = x.contiguous(memory_format=torch.channels_last_3d)
    keep = {}

    def ours():
        out = model(x)
        loss = sum(crit[t](out[t], tg[t]) for t in tasks)
        model.zero_grad(set_to_none=True)
        loss.backward()
        keep["ours"] = {t: out[t].detach() for t in tasks}

    def eager():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = O.net_forward(params, topo, xcl, tasks, training=True)
        loss = sum(crit[t](out[t].float(), tg[t]) for t in tasks)
        for p in params.values():
            p.grad = None
        loss.backward()
        keep["eager"] = {t: out[t].detach().float() for t in tasks}

    t_ours = _time(ours)
    t_eager = _time(eager)
    # the drop-in's eager launches are host bound at the deep 4^3 / 8^3 layers: also replay it as one CUDA graph, the way
    # bench.py runs the step
    rb.ops.PACK_CACHE = False
    try:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ours()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            ours()
        t_graph = _time(g.replay)
    finally:
        rb.ops.PACK_CACHE = True
    vox = B * P ** 3
    print(f"\nfwd+loss+bwd at {P}^3 x{B}: drop-in {t_ours:.1f} ms eager launches / {t_graph:.1f} ms as a CUDA graph "
          f"({vox / t_graph / 1e3:.1f} M voxels/s), PyTorch eager bf16 autocast channels_last_3d (ATen/cuDNN) {t_eager:.1f} ms "
          f"({vox / t_eager / 1e3:.1f} M voxels/s), ratio {t_eager / t_graph:.2f}x")
    for t in tasks:
        print(f"  {t}: rel-L2 drop-in vs autocast eager {rel_l2(keep['ours'][t], keep['eager'][t]):.3e}")
        assert rel_l2(keep["ours"][t], keep["eager"][t]) < 3e-2
    assert t_ours < t_eager
