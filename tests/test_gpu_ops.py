"""GPU parity tests of the individual operators (through ops -> C ABI -> sm_100a kernels) against
plain PyTorch fp32 references on the same bf16-rounded operands.

Tolerances (stated here, relative L2 unless noted):
  bf16 outputs (one rounding of an fp32 accumulator)      4e-3
  fp32 outputs of bf16 x bf16 products (wgrad, heads)      2e-4
Implementation under test is selected by RESENC_CONV_IMPL (auto | mma | tc5), default auto.
"""
import itertools
import os

import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu
TOL_BF16 = 4e-3
TOL_F32 = 2e-4


def q(t):
    """bf16-round, keep fp32 (what the kernels see)."""
    return t.to(torch.bfloat16).float()


@pytest.fixture(autouse=True)
def _device_error_guard(rb):
    yield
    rb._lib.device_error_check()


CONV_CASES = [
    # n, cin, cout, dims, kernel, stride
    (2, 32, 32, (16, 16, 16), (3, 3, 3), (1, 1, 1)),
    (2, 32, 64, (16, 16, 16), (3, 3, 3), (2, 2, 2)),
    (1, 64, 128, (8, 8, 8), (3, 3, 3), (1, 1, 1)),
    (2, 64, 32, (8, 12, 20), (1, 1, 1), (1, 1, 1)),
    (1, 32, 64, (8, 16, 16), (1, 3, 3), (1, 2, 2)),
    (1, 8, 16, (6, 10, 14), (3, 3, 3), (1, 1, 1)),
    (2, 128, 128, (4, 4, 4), (3, 3, 3), (1, 1, 1)),
    (1, 32, 32, (10, 12, 18), (3, 3, 3), (2, 2, 2)),
    (2, 512, 512, (4, 4, 4), (3, 3, 3), (1, 1, 1)),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: f"n{c[0]}_{c[1]}to{c[2]}_{'x'.join(map(str, c[3]))}_k{c[4][1]}s{c[5][1]}")
def test_conv3d_fwd_bwd(rb, case):
    n, cin, cout, dims, k, s = case
    torch.manual_seed(0)
    x = q(torch.randn(n, cin, *dims, device="cuda"))
    w = (torch.randn(cout, cin, *k, device="cuda") / (cin * k[0] * k[1] * k[2]) ** 0.5).requires_grad_(True)
    pad = tuple((kk - 1) // 2 for kk in k)
    xr = x.clone().requires_grad_(True)
    ref = F.conv3d(xr, q(w.detach()), None, s, pad)
    xp = x.clone().requires_grad_(True)
    y = rb.ops.conv3d(xp, w, s)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16
    assert rel_l2(y.float(), ref) < TOL_BF16
    g = q(torch.randn_like(ref))
    ref.backward(g)
    gw_ref = torch.autograd.grad(F.conv3d(x, w, None, s, pad), w, g)[0]
    y.backward(g.to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d))
    assert rel_l2(xp.grad.float(), xr.grad) < TOL_BF16
    assert rel_l2(w.grad, gw_ref) < TOL_F32


def test_conv3d_two_sources_equals_cat(rb):
    torch.manual_seed(1)
    a = q(torch.randn(2, 32, 8, 8, 8, device="cuda"))
    b = q(torch.randn(2, 32, 8, 8, 8, device="cuda"))
    w = (torch.randn(32, 64, 3, 3, 3, device="cuda") / 40).requires_grad_(True)
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv3d(torch.cat((ar, br), 1), q(w.detach()), None, 1, 1)
    ap, bp = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = rb.ops.conv3d(ap, w, 1, x_cat=bp)
    assert rel_l2(y.float(), ref) < TOL_BF16
    g = q(torch.randn_like(ref))
    ref.backward(g)
    gw_ref = torch.autograd.grad(F.conv3d(torch.cat((a, b), 1), w, None, 1, 1), w, g)[0]
    y.backward(g.to(torch.bfloat16))
    assert rel_l2(ap.grad.float(), ar.grad) < TOL_BF16
    assert rel_l2(bp.grad.float(), br.grad) < TOL_BF16
    assert rel_l2(w.grad, gw_ref) < TOL_F32


@pytest.mark.parametrize("stride", [(2, 2, 2), (1, 2, 2)])
@pytest.mark.parametrize("cin,cout", [(64, 32), (128, 64), (16, 8)])
def test_conv_transpose3d(rb, stride, cin, cout):
    torch.manual_seed(2)
    x = q(torch.randn(2, cin, 4, 6, 8, device="cuda"))
    w = (torch.randn(cin, cout, *stride, device="cuda") / cin ** 0.5).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.conv_transpose3d(xr, q(w.detach()), None, stride)
    xp = x.clone().requires_grad_(True)
    y = rb.ops.conv_transpose3d(xp, w, stride)
    assert y.shape == ref.shape
    assert rel_l2(y.float(), ref) < TOL_BF16
    g = q(torch.randn_like(ref))
    ref.backward(g)
    gw_ref = torch.autograd.grad(F.conv_transpose3d(x, w, None, stride), w, g)[0]
    y.backward(g.to(torch.bfloat16))
    assert rel_l2(xp.grad.float(), xr.grad) < TOL_BF16
    assert rel_l2(w.grad, gw_ref) < TOL_F32


@pytest.mark.parametrize("affine", [False, True])
@pytest.mark.parametrize("with_res,act", [(False, True), (True, True), (False, False)])
@pytest.mark.parametrize("shape", [(2, 32, 8, 8, 8), (1, 64, 4, 6, 10), (2, 512, 4, 4, 4), (1, 8, 16, 16, 16)])
def test_instance_norm_act(rb, affine, with_res, act, shape):
    torch.manual_seed(3)
    n, c = shape[:2]
    y = q(torch.randn(shape, device="cuda") * 2 + 0.5)
    res = q(torch.randn(shape, device="cuda")) if with_res else None
    gamma = (1 + 0.2 * torch.randn(c, device="cuda")).requires_grad_(True) if affine else None
    beta = (0.1 * torch.randn(c, device="cuda")).requires_grad_(True) if affine else None

    def reference(yy, rr, ga, be):
        o = F.instance_norm(yy, None, None, ga, be, True, 0.0, 1e-5)
        if rr is not None:
            o = o + rr
        return F.leaky_relu(o, 0.01) if act else o

    yr = y.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if with_res else None
    gr = gamma.detach().clone().requires_grad_(True) if affine else None
    br = beta.detach().clone().requires_grad_(True) if affine else None
    ref = reference(yr, rr, gr, br)
    yp = y.clone().requires_grad_(True)
    rp = res.clone().requires_grad_(True) if with_res else None
    z = rb.ops.instance_norm_act(yp, rp, gamma, beta, 1e-5, act, 0.01)
    assert rel_l2(z.float(), ref) < TOL_BF16
    g = q(torch.randn_like(ref))
    ref.backward(g)
    z.backward(g.to(torch.bfloat16))
    # the kernel's lrelu mask comes from the bf16-rounded output; elements within rounding of 0 may flip
    assert rel_l2(yp.grad.float(), yr.grad) < 8e-3
    if with_res:
        assert rel_l2(rp.grad.float(), rr.grad) < 8e-3
    if affine:
        assert rel_l2(gamma.grad, gr.grad) < 5e-3
        assert rel_l2(beta.grad, br.grad) < 5e-3


@pytest.mark.parametrize("reduce_dims", ["all", (2, 3)])
@pytest.mark.parametrize("affine", [False, True])
def test_instance_norm_se_act(rb, reduce_dims, affine):
    torch.manual_seed(4)
    n, c, rd = 2, 32, 8
    shape = (n, c, 6, 8, 10)
    y = q(torch.randn(shape, device="cuda") * 1.5 + 0.3)
    res = q(torch.randn(shape, device="cuda"))
    params = [torch.randn(rd, c, 1, 1, 1, device="cuda") * 0.3, torch.randn(rd, device="cuda") * 0.1,
              torch.randn(c, rd, 1, 1, 1, device="cuda") * 0.3, torch.randn(c, device="cuda") * 0.1]
    gamma = (1 + 0.2 * torch.randn(c, device="cuda")) if affine else None
    beta = (0.3 * torch.randn(c, device="cuda")) if affine else None

    def reference(yy, rr, ga, be, w1, b1, w2, b2):
        o = F.instance_norm(yy, None, None, ga, be, True, 0.0, 1e-5)
        dims = (2, 3, 4) if reduce_dims == "all" else reduce_dims
        s = o.mean(dims, keepdim=True)
        s = F.conv3d(F.relu(F.conv3d(s, w1, b1)), w2, b2)
        return F.leaky_relu(o * torch.sigmoid(s) + rr, 0.01)

    leaf = lambda t: None if t is None else t.detach().clone().requires_grad_(True)
    ref_in = [leaf(t) for t in [y, res, gamma, beta] + params]
    ref = reference(*ref_in)
    prod_in = [leaf(t) for t in [y, res, gamma, beta] + params]
    z = rb.ops.instance_norm_se_act(prod_in[0], prod_in[1], prod_in[2], prod_in[3], *prod_in[4:], eps=1e-5, act=True,
                                    slope=0.01, reduce_dims=reduce_dims)
    assert rel_l2(z.float(), ref) < TOL_BF16
    g = q(torch.randn_like(ref))
    ref.backward(g)
    z.backward(g.to(torch.bfloat16))
    names = ["y", "res", "gamma", "beta", "fc1.w", "fc1.b", "fc2.w", "fc2.b"]
    for nm, a, b in zip(names, prod_in, ref_in):
        if a is None:
            continue
        assert a.grad is not None, nm
        if float(b.grad.norm()) < 1e-6 * max(1.0, float(b.norm())):
            # global pooling of a non-affine InstanceNorm output is exactly 0: the reference's fc1.weight gradient
            # is rounding noise, ours is exactly 0
            assert float(a.grad.float().norm()) < 1e-5, nm
            continue
        assert rel_l2(a.grad.float(), b.grad) < 1e-2, nm


@pytest.mark.parametrize("with_se", [False, True])
@pytest.mark.parametrize("affine", [False, True])
def test_instance_norm_drop_path(rb, with_se, affine):
    """Stochastic depth (DropPath, reference resblocks.py:109-112) folded into the block tail: per-sample factor
    0 or 1/keep between the norm and the SE gate; dropped samples pass the residual only and get zero gradient."""
    torch.manual_seed(5)
    n, c, rd = 4, 32, 8
    shape = (n, c, 6, 8, 8)
    y = q(torch.randn(shape, device="cuda") * 1.5 + 0.3)
    res = q(torch.randn(shape, device="cuda"))
    drop = torch.tensor([1.25, 0.0, 1.25, 0.0], device="cuda")
    params = [torch.randn(rd, c, 1, 1, 1, device="cuda") * 0.3, torch.randn(rd, device="cuda") * 0.1,
              torch.randn(c, rd, 1, 1, 1, device="cuda") * 0.3, torch.randn(c, device="cuda") * 0.1] if with_se else []
    gamma = (1 + 0.2 * torch.randn(c, device="cuda")) if affine else None
    beta = (0.3 * torch.randn(c, device="cuda")) if affine else None

    def reference(yy, rr, ga, be, *se):
        o = F.instance_norm(yy, None, None, ga, be, True, 0.0, 1e-5) * drop.view(-1, 1, 1, 1, 1)
        if se:
            s = o.mean((2, 3, 4), keepdim=True)
            o = o * torch.sigmoid(F.conv3d(F.relu(F.conv3d(s, se[0], se[1])), se[2], se[3]))
        return F.leaky_relu(o + rr, 0.01)

    leaf = lambda t: None if t is None else t.detach().clone().requires_grad_(True)
    ref_in = [leaf(t) for t in [y, res, gamma, beta] + params]
    ref = reference(*ref_in)
    prod_in = [leaf(t) for t in [y, res, gamma, beta] + params]
    if with_se:
        z = rb.ops.instance_norm_se_act(*prod_in, eps=1e-5, act=True, slope=0.01, reduce_dims="all", drop=drop)
    else:
        z = rb.ops.instance_norm_act(*prod_in, eps=1e-5, act=True, slope=0.01, drop=drop)
    assert rel_l2(z.float(), ref) < TOL_BF16
    # dropped samples: lrelu(res) exactly
    assert torch.equal(z[1].float(), F.leaky_relu(res[1].float(), 0.01).to(torch.bfloat16).float())
    g = q(torch.randn_like(ref))
    ref.backward(g)
    z.backward(g.to(torch.bfloat16))
    assert float(prod_in[0].grad[1].float().abs().max()) == 0.0 and float(prod_in[0].grad[3].float().abs().max()) == 0.0
    names = ["y", "res", "gamma", "beta", "fc1.w", "fc1.b", "fc2.w", "fc2.b"]
    for nm, a, b in zip(names, prod_in, ref_in):
        if a is None:
            continue
        assert a.grad is not None, nm
        if float(b.grad.norm()) < 1e-6 * max(1.0, float(b.norm())):
            assert float(a.grad.float().norm()) < 1e-5, nm
            continue
        assert rel_l2(a.grad.float(), b.grad) < 1e-2, nm


@pytest.mark.parametrize("stride", [(2, 2, 2), (1, 2, 2)])
def test_avg_pool(rb, stride):
    torch.manual_seed(5)
    x = q(torch.randn(2, 32, 4, 8, 12, device="cuda"))
    xr = x.clone().requires_grad_(True)
    ref = F.avg_pool3d(xr, stride, stride)
    xp = x.clone().requires_grad_(True)
    out = rb.ops.avg_pool3d(xp, stride)
    assert rel_l2(out.float(), ref) < TOL_BF16
    g = q(torch.randn_like(ref))
    ref.backward(g)
    out.backward(g.to(torch.bfloat16))
    assert rel_l2(xp.grad.float(), xr.grad) < TOL_BF16


@pytest.mark.parametrize("k,act", [(1, None), (3, None), (1, "sigmoid"), (3, "softmax"), (2, "softmax")])
def test_head(rb, k, act):
    torch.manual_seed(6)
    x = q(torch.randn(2, 32, 6, 8, 10, device="cuda"))
    w = (torch.randn(k, 32, 1, 1, 1, device="cuda") * 0.2).requires_grad_(True)
    b = (torch.randn(k, device="cuda") * 0.1).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    ref = F.conv3d(xr, wr, br)
    if act == "sigmoid":
        ref = torch.sigmoid(ref)
    elif act == "softmax":
        ref = torch.softmax(ref, 1)
    xp = x.clone().requires_grad_(True)
    out = rb.ops.head_conv1x1(xp, w, b, act)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == ref.shape
    assert rel_l2(out, ref) < 1e-5
    if act is None:
        g = torch.randn_like(ref)
        ref.backward(g)
        out.backward(g)
        assert rel_l2(xp.grad.float(), xr.grad) < TOL_BF16
        assert rel_l2(w.grad, wr.grad) < TOL_F32
        assert rel_l2(b.grad, br.grad) < TOL_F32


@pytest.mark.parametrize("c", [8, 24, 64, 256])
@pytest.mark.parametrize("k", [1, 3, 8])
def test_head_channel_counts(rb, c, k):
    """C / 8 a power of two takes the (voxel, channel-group) lane kernel, C = 24 the per-voxel one; S = 5*7*9 is not a
    multiple of the voxels a warp instruction covers (ragged tail)."""
    torch.manual_seed(60 + c + k)
    x = q(torch.randn(2, c, 5, 7, 9, device="cuda"))
    w = torch.randn(k, c, 1, 1, 1, device="cuda") * 0.2
    b = torch.randn(k, device="cuda") * 0.1
    for act, fn in ((None, lambda t: t), ("sigmoid", torch.sigmoid), ("softmax", lambda t: torch.softmax(t, 1))):
        out = rb.ops.head_conv1x1(x, w, b, act)
        assert rel_l2(out, fn(F.conv3d(x, w, b))) < 1e-5


@pytest.mark.parametrize("case", [
    # n, cin, cat, cout, dims, k, with_res, head activation
    (2, 32, 0, 32, (8, 8, 8), 1, False, None),
    (1, 32, 32, 32, (5, 6, 10), 3, False, None),
    (2, 16, 0, 64, (4, 6, 6), 3, True, "softmax"),
    (1, 32, 0, 256, (4, 4, 4), 8, False, "sigmoid"),
    (1, 8, 0, 8, (3, 5, 7), 2, True, None),
])
def test_fused_norm_act_head_tail(rb, case):
    """Inference tail of a decoder: conv -> IN -> LeakyReLU -> 1x1x1 head with the activation never stored
    (ops.conv_norm_act_head / rb_norm_act_head_fwd) against (a) the two-pass path of the same library - same bf16
    rounding of the activation, only the summation order of the 1x1 dot product differs: 1e-5 - and (b) torch fp32."""
    n, cin, cat, cout, dims, k, with_res, hact = case
    torch.manual_seed(61)
    x = q(torch.randn(n, cin, *dims, device="cuda"))
    xc = q(torch.randn(n, cat, *dims, device="cuda")) if cat else None
    w = torch.randn(cout, cin + cat, 3, 3, 3, device="cuda") / (27 * (cin + cat)) ** 0.5
    gamma = 1 + 0.2 * torch.randn(cout, device="cuda")
    beta = 0.1 * torch.randn(cout, device="cuda")
    res = q(torch.randn(n, cout, *dims, device="cuda")) if with_res else None
    hw = torch.randn(k, cout, 1, 1, 1, device="cuda") * 0.2
    hb = torch.randn(k, device="cuda") * 0.1
    head = (hw, hb, hact)
    with torch.no_grad():
        assert rb.ops.can_fuse_head(w, head)
        fused = rb.ops.conv_norm_act_head(x, w, 1, xc, res, gamma, beta, 1e-5, True, 0.01, False, head)
        z = rb.ops.conv_norm_act(x, w, 1, x_cat=xc, res=res, gamma=gamma, beta=beta, act=True)
        two_pass = rb.ops.head_conv1x1(z, hw, hb, hact)
        xin = torch.cat((x, xc), 1) if cat else x
        o = F.instance_norm(F.conv3d(xin, q(w), None, 1, 1), None, None, gamma, beta, True, 0.0, 1e-5)
        ref = F.conv3d(F.leaky_relu(o + res if with_res else o, 0.01), hw, hb)
        ref = torch.sigmoid(ref) if hact == "sigmoid" else torch.softmax(ref, 1) if hact == "softmax" else ref
    assert fused.shape == two_pass.shape == ref.shape and fused.dtype == torch.float32 and fused.is_contiguous()
    assert rel_l2(fused, two_pass) < 1e-5
    assert rel_l2(fused, ref) < TOL_BF16
    with torch.enable_grad():
        assert not rb.ops.can_fuse_head(w, head)     # training keeps the stored activation for backward


@pytest.mark.parametrize("cin", [1, 2, 4])
def test_stem_conv(rb, cin):
    torch.manual_seed(7)
    x = q(torch.rand(2, cin, 8, 10, 12, device="cuda"))
    w = (torch.randn(32, cin, 3, 3, 3, device="cuda") / (27 * cin) ** 0.5).requires_grad_(True)
    ref = F.conv3d(x, q(w.detach()), None, 1, 1)
    y = rb.ops.stem_conv3d(x, w)
    assert rel_l2(y.float(), ref) < TOL_BF16
    g = q(torch.randn_like(ref))
    gw_ref = torch.autograd.grad(F.conv3d(x, w, None, 1, 1), w, g)[0]
    y.backward(g.to(torch.bfloat16))
    assert rel_l2(w.grad, gw_ref) < TOL_F32


def test_layout_roundtrip(rb):
    x = q(torch.randn(2, 16, 3, 5, 7, device="cuda"))
    cl = rb.ops.as_cl(x)
    assert rb.ops.is_cl(cl) and torch.equal(cl.float(), x)
    back = rb.ops.cl_to_ncdhw_f32(cl)
    assert back.is_contiguous() and torch.equal(back, x)


def test_conv_linearity_at_full_resolution(rb):
    """Size-independent property at the headline layer shape (32 -> 32 @ 128^3): conv(2x) == 2 conv(x)
    exactly in bf16 (scaling by a power of two commutes with every rounding)."""
    torch.manual_seed(8)
    x = rb.ops.as_cl(torch.randn(1, 32, 128, 128, 128, device="cuda"))
    w = torch.randn(32, 32, 3, 3, 3, device="cuda") / 30
    a = rb.ops.conv3d(x, w, 1)
    b = rb.ops.conv3d((x.float() * 2).to(torch.bfloat16), w, 1)
    assert torch.equal(b.float(), a.float() * 2)
    # translation equivariance away from the border: shifting the input by one voxel along W
    xs = torch.roll(x, 1, dims=4)
    c = rb.ops.conv3d(xs, w, 1)
    assert torch.equal(c[..., 2:-2].float(), torch.roll(a, 1, dims=4)[..., 2:-2].float())


@pytest.mark.parametrize("case", [(2, 32, 32, (8, 8, 8), 1, False), (1, 32, 64, (8, 12, 16), 2, False),
                                  (2, 64, 64, (4, 4, 4), 1, True), (1, 16, 16, (6, 6, 6), 1, True)])
def test_fused_conv_norm_act_unit(rb, case):
    """The unit the network is built from: act(IN(conv(x)) * gamma + beta + res), fp32 pre-norm inside."""
    n, cin, cout, dims, s, with_res = case
    torch.manual_seed(9)
    x = q(torch.randn(n, cin, *dims, device="cuda"))
    w = (torch.randn(cout, cin, 3, 3, 3, device="cuda") / (27 * cin) ** 0.5).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(cout, device="cuda")).requires_grad_(True)
    beta = (0.1 * torch.randn(cout, device="cuda")).requires_grad_(True)
    od = tuple((d - 1) // s + 1 for d in dims)
    res = q(torch.randn(n, cout, *od, device="cuda")) if with_res else None

    xr = x.clone().requires_grad_(True)
    wr, gr, br = (t.detach().clone().requires_grad_(True) for t in (w, gamma, beta))
    rr = res.clone().requires_grad_(True) if with_res else None
    o = F.instance_norm(F.conv3d(xr, q(wr.detach()) + (wr - wr.detach()), None, s, 1), None, None, gr, br, True, 0.0, 1e-5)
    ref = F.leaky_relu(o + rr if with_res else o, 0.01)
    xp = x.clone().requires_grad_(True)
    rp = res.clone().requires_grad_(True) if with_res else None
    z = rb.ops.conv_norm_act(xp, w, s, res=rp, gamma=gamma, beta=beta, act=True)
    assert rb.ops.is_cl(z) and rel_l2(z.float(), ref) < TOL_BF16
    g = q(torch.randn_like(ref))
    ref.backward(g)
    z.backward(g.to(torch.bfloat16))
    assert rel_l2(xp.grad.float(), xr.grad) < 1e-2
    assert rel_l2(w.grad, wr.grad) < 1e-2
    assert rel_l2(gamma.grad, gr.grad) < 1e-2 and rel_l2(beta.grad, br.grad) < 1e-2
    if with_res:
        assert rel_l2(rp.grad.float(), rr.grad) < 8e-3


@pytest.mark.parametrize("co,ci,k", [(32, 32, (3, 3, 3)), (64, 128, (3, 3, 3)), (40, 24, (1, 3, 3)), (512, 1024, (3, 3, 3)),
                                     (64, 32, (1, 1, 1)), (72, 64, (3, 3, 3)), (33, 32, (2, 2, 2)), (256, 512, (2, 2, 2))])
def test_weight_pack_unpack_kernels(rb, co, ci, k):
    """Tiled pack / unpack kernels == the torch permutes they replace (bit-exact)."""
    ops = rb.ops
    torch.manual_seed(11)
    w = torch.randn(co, ci, *k, device="cuda")
    T = k[0] * k[1] * k[2]
    f, d = ops._pack_kernel(w, True, True)
    assert torch.equal(f, w.permute(2, 3, 4, 0, 1).reshape(T, co, ci).to(torch.bfloat16))
    assert torch.equal(d, w.flip(2, 3, 4).permute(2, 3, 4, 1, 0).reshape(T, ci, co).to(torch.bfloat16))
    dw = torch.randn(T, co, ci, device="cuda")
    g = ops.unpack_wgrad(dw, co, ci, k)
    assert torch.equal(g, dw.view(*k, co, ci).permute(3, 4, 0, 1, 2))


def test_fused_losses_match_compositions(rb):
    """csrc/loss.cuh (one reduction pass + one gradient pass) against the PyTorch compositions of
    training/losses/losses.py in the package and against the oracle: loss values to 1e-6, gradients to rel-L2 1e-5.
    Includes masked-out normals, saturated logits, several channels and a grad_out != 1."""
    from oracle import resenc_oracle as O   # checker only
    L = rb.losses
    torch.manual_seed(5)
    for shape in [(2, 1, 16, 24, 40), (1, 3, 8, 8, 8), (2, 2, 5, 7, 9)]:
        z = (torch.randn(*shape, device="cuda") * 4).requires_grad_(True)
        t = (torch.rand(*shape, device="cuda") > 0.8).float()
        a = L.FusedBCEDiceLoss(0.5, 0.5)(z, t)
        (a * 1.7).backward()
        z2 = z.detach().clone().requires_grad_(True)
        b = L.BCEDiceLoss(0.5, 0.5)(z2, t)
        (b * 1.7).backward()
        assert abs(float(a) - float(b)) < 2e-6 and abs(float(a) - float(O.bce_dice_loss(z.detach().cpu(), t.cpu()))) < 2e-6
        assert rel_l2(z.grad, z2.grad) < 1e-5
    for shape in [(2, 3, 16, 24, 40), (1, 3, 7, 9, 11)]:
        p = torch.randn(*shape, device="cuda", requires_grad=True)
        tn = torch.nn.functional.normalize(torch.randn(*shape, device="cuda"), dim=1)
        tn[:, :, :3] = 0
        a = L.FusedMaskedCosineLoss()(p, tn)
        (a * 0.6).backward()
        p2 = p.detach().clone().requires_grad_(True)
        b = L.MaskedCosineLoss()(p2, tn)
        (b * 0.6).backward()
        assert abs(float(a) - float(b)) < 2e-6 and abs(float(a) - float(O.masked_cosine_loss(p.detach().cpu(), tn.cpu()))) < 2e-6
        assert rel_l2(p.grad, p2.grad) < 1e-5
    # CPU tensors keep the composition (the fused kernels are CUDA only)
    zc = torch.randn(1, 1, 4, 4, 4)
    assert torch.isfinite(L.FusedBCEDiceLoss(0.5, 0.5)(zc, (zc > 0).float()))


def test_norm_backward_sign_from_prenorm(rb):
    """Backward of conv + InstanceNorm + LeakyReLU that recomputes lrelu'(z) from the fp32 pre-norm tensor and
    the forward's folded scale / shift instead of reading the stored activation: same gradients."""
    ops = rb.ops
    torch.manual_seed(9)
    x = q(torch.randn(2, 32, 12, 16, 16, device="cuda"))
    w = (torch.randn(32, 32, 3, 3, 3, device="cuda") / 30).requires_grad_(True)
    g = q(torch.randn(2, 32, 12, 16, 16, device="cuda")).to(torch.bfloat16)
    outs = []
    for flag in (False, True):
        ops.SIGN_FROM_PRENORM = flag
        try:
            xp = x.clone().requires_grad_(True)
            w.grad = None
            z = ops.conv_norm_act(xp, w, (1, 1, 1), None, None, None, None, 1e-5, True, 0.01)
            z.backward(g)
            outs.append((z.detach().float(), xp.grad.float(), w.grad.clone()))
        finally:
            ops.SIGN_FROM_PRENORM = True
    assert rel_l2(outs[1][0], outs[0][0]) < 1e-3
    assert rel_l2(outs[1][1], outs[0][1]) < 5e-3 and rel_l2(outs[1][2], outs[0][2]) < 5e-3


def test_split_precision_layout_kernels(rb):
    """rb_split_apply / rb_avgpool_split / rb_stem_im2col_split (include/resenc_b200.h): the [hi | lo | hi] rows
    reproduce their fp32 source to 2^-16 relative and the fused fp32 arithmetic matches torch."""
    P = rb.precise
    torch.manual_seed(11)
    n, c, d, h, w = 2, 16, 4, 6, 8
    y = torch.randn(n, c, d, h, w, device="cuda") * 3 + 0.5
    ycl = y.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
    resv = torch.randn(n, c, d, h, w, device="cuda")
    res = P._split_apply(resv.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3), None, None, None, False, 0.0)
    assert tuple(res.shape) == (n, 3 * c, d, h, w) and rb.ops.is_cl(res)
    assert torch.equal(res[:, :c], res[:, 2 * c:])
    assert torch.equal(res[:, :c].float(), resv.to(torch.bfloat16).float())
    assert rel_l2(P.join(res), resv) < 2 ** -16
    scale = torch.rand(n, c, device="cuda") + 0.5
    shift = torch.randn(n, c, device="cuda")
    z = P._split_apply(ycl, res, scale, shift, True, 0.01)
    ref = F.leaky_relu(y * scale.view(n, c, 1, 1, 1) + shift.view(n, c, 1, 1, 1) + P.join(res), 0.01)
    assert rel_l2(P.join(z), ref) < 1e-5
    pooled = P.avg_pool3d(z, (2, 2, 2))
    assert rel_l2(P.join(pooled), F.avg_pool3d(P.join(z), 2, 2)) < 1e-5
    pooled = P.avg_pool3d(z, (1, 2, 2))
    assert rel_l2(P.join(pooled), F.avg_pool3d(P.join(z), (1, 2, 2), (1, 2, 2))) < 1e-5
    # whole unit against torch fp32: conv + InstanceNorm + residual + LeakyReLU, strided and concatenated variants
    wgt = torch.randn(24, c, 3, 3, 3, device="cuda") * 0.1
    for stride in (1, 2):
        out = P.conv_norm_act(z, wgt, stride=stride)
        ref = F.leaky_relu(F.instance_norm(F.conv3d(P.join(z), wgt, None, stride, 1)), 0.01)
        r = rel_l2(P.join(out), ref)
        print(f"split conv_norm_act stride {stride}: rel-L2 {r:.3e}")
        assert r < 3e-5
    w2 = torch.randn(16, 2 * c, 3, 3, 3, device="cuda") * 0.1
    out = P.conv_norm_act(z, w2, x_cat=res, res=res, act=True)
    ref = F.leaky_relu(F.instance_norm(F.conv3d(torch.cat((P.join(z), P.join(res)), 1), w2, None, 1, 1)) + P.join(res), 0.01)
    assert rel_l2(P.join(out), ref) < 3e-5
    wt = torch.randn(c, 8, 2, 2, 2, device="cuda") * 0.2
    bt = torch.randn(8, device="cuda")
    up = P.conv_transpose3d(z, wt, (2, 2, 2), bias=bt)
    assert rel_l2(P.join(up), F.conv_transpose3d(P.join(z), wt, bt, 2)) < 3e-5
    hw, hb = torch.randn(3, 8, 1, 1, 1, device="cuda"), torch.randn(3, device="cuda")
    assert rel_l2(P.head_conv1x1(up, hw, hb, "softmax"), torch.softmax(F.conv3d(P.join(up), hw, hb), 1)) < 3e-5
    # stem on the raw fp32 input (2 channels)
    xin = torch.rand(1, 2, 6, 8, 8, device="cuda")
    ws = torch.randn(8, 2, 3, 3, 3, device="cuda") * 0.3
    out = P.conv_norm_act(xin, ws, stem=True)
    ref = F.leaky_relu(F.instance_norm(F.conv3d(xin, ws, None, 1, 1)), 0.01)
    assert rel_l2(P.join(out), ref) < 3e-5


@pytest.mark.parametrize("case", [
    # cout, cin, kernel, stride, input dims
    (64, 32, (3, 3, 3), (2, 2, 2), (16, 16, 16)),
    (128, 64, (3, 3, 3), (2, 2, 2), (8, 8, 8)),
    (64, 32, (3, 3, 3), (1, 2, 2), (8, 16, 16)),
    (32, 32, (1, 3, 3), (1, 2, 2), (8, 16, 16)),
    (512, 512, (3, 3, 3), (2, 2, 2), (8, 8, 8)),
])
def test_merged_dgrad_pack_kernel_equals_the_torch_construction(rb, case):
    """rb_pack_conv_dgrad_merged (one launch) against the torch-op construction it replaces (flip, zero fill, one
    strided slice copy + cast per parity class): bit-identical operand, zero blocks included."""
    ops = rb.ops
    co, ci, k, s, dims = case
    torch.manual_seed(co + ci)
    w = torch.randn(co, ci, *k, device="cuda")
    pad = tuple((kk - 1) // 2 for kk in k)
    od = ops._conv_out_dims(dims, k, s)
    axes = ops._merged_dgrad_plan(k, s, pad, dims, od)
    assert axes is not None
    ref = ops.pack_conv_dgrad_merged(w, axes, s, force_torch=True)
    assert ops.DMERGE_KERNEL
    new = ops._pack_dgrad_merged_kernel(w, axes, s)
    assert new.shape == ref.shape and new.dtype == ref.dtype and new.is_contiguous()
    assert torch.equal(new, ref)
    assert int((ref == 0).sum()) > 0          # the zero blocks are part of the comparison
