"""The tcgen05 / TMEM / TMA gather-conv kernel against (a) PyTorch fp32 on bf16-rounded operands and
(b) the shape-generic mma.sync kernel, on every geometry class the network uses: 3x3x3 stride 1 / 2,
1x1x1, two K-segments (virtual concat), parity-class dgrad with strided stores, pixel-shuffle transposed
conv and its k2s2 data gradient, ragged grids (TMA out-of-bounds fill), batch-folded tiles (tn > 1),
N tiles of 32..256 and the fused InstanceNorm statistics epilogue."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu
TOL = 4e-3


def q(t):
    return t.to(torch.bfloat16).float()


@pytest.fixture(autouse=True)
def _device_error_guard(rb):
    yield
    rb._lib.device_error_check()


CASES = [
    (2, 32, 32, (16, 16, 16), (3, 3, 3), (1, 1, 1)),
    (1, 64, 64, (16, 16, 16), (3, 3, 3), (1, 1, 1)),
    (2, 32, 64, (16, 16, 16), (3, 3, 3), (2, 2, 2)),
    (1, 128, 128, (8, 8, 8), (3, 3, 3), (1, 1, 1)),
    (2, 64, 32, (8, 12, 20), (1, 1, 1), (1, 1, 1)),
    (1, 32, 64, (8, 16, 16), (1, 3, 3), (1, 2, 2)),
    (1, 16, 32, (6, 10, 14), (3, 3, 3), (1, 1, 1)),
    (2, 256, 256, (4, 4, 4), (3, 3, 3), (1, 1, 1)),
    (1, 32, 32, (10, 12, 18), (3, 3, 3), (2, 2, 2)),
    (2, 256, 512, (8, 8, 8), (3, 3, 3), (2, 2, 2)),
    (1, 48, 96, (8, 8, 24), (3, 3, 3), (1, 1, 1)),
    (1, 32, 32, (64, 64, 64), (3, 3, 3), (1, 1, 1)),
    # >= 64 tiles of 256 voxels with <= 128 output channels: the weights-on-M orientation (conv_tc5t.cuh)
    (1, 64, 64, (32, 32, 32), (3, 3, 3), (1, 1, 1)),
    (1, 32, 64, (64, 64, 64), (3, 3, 3), (2, 2, 2)),
    (2, 128, 128, (16, 32, 32), (3, 3, 3), (1, 1, 1)),
    (1, 32, 32, (40, 36, 44), (3, 3, 3), (1, 1, 1)),
    (1, 64, 32, (24, 40, 72), (1, 1, 1), (1, 1, 1)),
    (1, 32, 64, (16, 64, 64), (1, 3, 3), (1, 2, 2)),
    # 64 tiles (under half a wave of SMs) with 27 taps of 256 channels: tap split over two CTAs per tile
    (2, 256, 256, (16, 16, 16), (3, 3, 3), (1, 1, 1)),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[0]}_{c[1]}to{c[2]}_{'x'.join(map(str, c[3]))}_k{c[4][1]}s{c[5][1]}")
def test_tc5_conv_fwd_bwd(rb, case):
    n, cin, cout, dims, k, s = case
    torch.manual_seed(0)
    x = q(torch.randn(n, cin, *dims, device="cuda"))
    w = (torch.randn(cout, cin, *k, device="cuda") / (cin * k[0] * k[1] * k[2]) ** 0.5).requires_grad_(True)
    pad = tuple((kk - 1) // 2 for kk in k)
    xr = x.clone().requires_grad_(True)
    ref = F.conv3d(xr, q(w.detach()), None, s, pad)
    g = q(torch.randn_like(ref))
    ref.backward(g)
    outs = {}
    bwd_ok = cin % 32 == 0          # the data gradient is a conv with Nout = Cin; tc5 needs Nout % 32 == 0
    for impl in ("tc5", "mma"):
        xp = x.clone().requires_grad_(True)
        y = rb.ops.conv3d(xp, w, s, impl=impl)
        if bwd_ok:
            y.backward(g.to(torch.bfloat16))
        outs[impl] = (y.detach().float(), xp.grad.float() if bwd_ok else xr.grad)
        rb._lib.device_error_check()
    e_f, e_b = rel_l2(outs["tc5"][0], ref), rel_l2(outs["tc5"][1], xr.grad)
    print(f"tc5 fwd {e_f:.2e} bwd {e_b:.2e}; mma fwd {rel_l2(outs['mma'][0], ref):.2e}; "
          f"tc5-vs-mma {rel_l2(outs['tc5'][0], outs['mma'][0]):.2e}")
    assert e_f < TOL and e_b < TOL
    assert rel_l2(outs["tc5"][0], outs["mma"][0]) < 3e-3


@pytest.mark.parametrize("dims,c,co", [((8, 8, 8), 64, 64), ((32, 32, 32), 32, 32), ((16, 48, 40), 64, 64),
                                       ((16, 16, 16), 256, 256)])
def test_tc5_two_sources(rb, dims, c, co):
    torch.manual_seed(1)
    a = q(torch.randn(2, c, *dims, device="cuda"))
    b = q(torch.randn(2, c, *dims, device="cuda"))
    w = torch.randn(co, 2 * c, 3, 3, 3, device="cuda") / (54 * c) ** 0.5
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv3d(torch.cat((ar, br), 1), q(w), None, 1, 1)
    g = q(torch.randn_like(ref))
    ref.backward(g)
    ap, bp = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = rb.ops.conv3d(ap, w, 1, x_cat=bp, impl="tc5")
    y.backward(g.to(torch.bfloat16))
    assert rel_l2(y.float(), ref) < TOL
    assert rel_l2(ap.grad.float(), ar.grad) < TOL and rel_l2(bp.grad.float(), br.grad) < TOL


@pytest.mark.parametrize("stride", [(2, 2, 2), (1, 2, 2)])
@pytest.mark.parametrize("cin,cout", [(64, 32), (512, 256)])
def test_tc5_conv_transpose(rb, stride, cin, cout):
    torch.manual_seed(2)
    x = q(torch.randn(2, cin, 4, 6, 8, device="cuda"))
    w = torch.randn(cin, cout, *stride, device="cuda") / cin ** 0.5
    xr = x.clone().requires_grad_(True)
    ref = F.conv_transpose3d(xr, q(w), None, stride)
    g = q(torch.randn_like(ref))
    ref.backward(g)
    xp = x.clone().requires_grad_(True)
    y = rb.ops.conv_transpose3d(xp, w, stride, impl="tc5")
    y.backward(g.to(torch.bfloat16))
    assert rel_l2(y.float(), ref) < TOL
    assert rel_l2(xp.grad.float(), xr.grad) < TOL


def test_tc5_fused_statistics(rb):
    """Sum / sum-of-squares epilogue == reductions of the fp32 accumulators (compared with the stored
    bf16 output, so within bf16 rounding)."""
    torch.manual_seed(3)
    ops = rb.ops
    for n, c, dims in [(2, 32, (16, 16, 16)), (2, 128, (4, 4, 4)), (1, 64, (10, 12, 14)), (2, 32, (32, 32, 32)),
                       (1, 64, (20, 36, 40))]:
        x = ops.as_cl(torch.randn(n, c, *dims, device="cuda"))
        w = torch.randn(c, c, 3, 3, 3, device="cuda") / (27 * c) ** 0.5
        y = ops.new_cl(n, c, *dims, "cuda")
        ssum = torch.zeros(n, c, device="cuda")
        ssq = torch.zeros(n, c, device="cuda")
        ops._launch_gather(x, None, ops.pack_conv_fprop(w), y, None, in_dims=dims, taps=(3, 3, 3), off=(-1, -1, -1),
                           istr=(1, 1, 1), out_grid=dims, nout=c, impl="tc5", stats=(ssum, ssq))
        yf = y.float()
        assert rel_l2(ssum, yf.sum((2, 3, 4))) < 5e-3
        assert rel_l2(ssq, (yf * yf).sum((2, 3, 4))) < 5e-3
        # fp32 destination: same accumulators, no rounding; statistics then agree to fp32 summation error
        y32 = ops.new_cl_f32(n, c, *dims, "cuda")
        st = ops._launch_gather(x, None, ops.pack_conv_fprop(w), y32, None, in_dims=dims, taps=(3, 3, 3),
                                off=(-1, -1, -1), istr=(1, 1, 1), out_grid=dims, nout=c, impl="tc5", want_stats=True)
        assert st is not None
        assert rel_l2(y32, yf) < 3e-3 and torch.equal(y32.to(torch.bfloat16), y)
        assert rel_l2(st[0], y32.double().sum((2, 3, 4))) < 2e-5
        assert rel_l2(st[1], (y32.double() ** 2).sum((2, 3, 4))) < 2e-5


SLAB_CASES = [(1, (64, 64, 64)), (1, (24, 32, 128)), (2, (40, 32, 32)), (2, (10, 64, 64)),
              (1, (40, 48, 96)), (2, (18, 96, 96))]      # W = 96: tiles of two rows, N = 192 accumulator columns


@pytest.mark.parametrize("n,dims", SLAB_CASES, ids=lambda v: str(v).replace(" ", ""))
def test_slab_conv(rb, n, dims):
    """conv_slab.cuh (z-marching CTAs, kw taps stacked on M and shifted in the epilogue) against the shape-generic
    mma.sync kernel and PyTorch: bf16 and fp32 destinations, fused statistics, the two-source accumulate pass
    (virtual concat) and the two-destination data gradient.  Ragged z chunks included (D = 40, 10, 24)."""
    ops, L = rb.ops, rb._lib
    torch.manual_seed(11)
    a = q(torch.randn(n, 32, *dims, device="cuda"))
    b = q(torch.randn(n, 32, *dims, device="cuda"))
    w1 = torch.randn(32, 32, 3, 3, 3, device="cuda") / (27 * 32) ** 0.5
    w2 = torch.randn(32, 64, 3, 3, 3, device="cuda") / (27 * 64) ** 0.5
    kw = dict(in_dims=dims, taps=(3, 3, 3), off=(-1, -1, -1), istr=(1, 1, 1), out_grid=dims)
    acl, bcl = ops.as_cl(a), ops.as_cl(b)

    # the library really picks the slab kernel for these shapes
    y16 = ops.new_cl(n, 32, *dims, "cuda")
    d = ops._make_desc(acl, None, y16, None, nout=32, mode=0, ostr=(1, 1, 1), ooff=(0, 0, 0), full=None, ps=None, psC=0,
                       impl=None, **kw)
    assert L.load().rb_conv_gather_plan(ctypes.byref(d)) == L.IMPL_TCGEN05_SLAB

    # one source, bf16 destination
    ref1 = F.conv3d(a, q(w1), None, 1, 1)
    ops._launch_gather(acl, None, ops.pack_conv_fprop(w1), y16, None, nout=32, **kw)
    ym = ops.new_cl(n, 32, *dims, "cuda")
    ops._launch_gather(acl, None, ops.pack_conv_fprop(w1), ym, None, nout=32, impl="mma", **kw)
    assert rel_l2(y16.float(), ref1) < TOL and rel_l2(y16.float(), ym.float()) < 3e-3

    # one source, fp32 destination + statistics
    y32 = ops.new_cl_f32(n, 32, *dims, "cuda")
    st = ops._launch_gather(acl, None, ops.pack_conv_fprop(w1), y32, None, nout=32, want_stats=True, **kw)
    assert st is not None
    assert rel_l2(y32, ref1) < 1e-4
    assert rel_l2(st[0], y32.double().sum((2, 3, 4))) < 2e-5
    assert rel_l2(st[1], (y32.double() ** 2).sum((2, 3, 4))) < 2e-5

    # two sources (virtual concat): second launch accumulates into the fp32 destination, statistics of the sum
    ref2 = F.conv3d(torch.cat((a, b), 1), q(w2), None, 1, 1)
    z32 = ops.new_cl_f32(n, 32, *dims, "cuda")
    st2 = ops._launch_gather(acl, bcl, ops.pack_conv_fprop(w2), z32, None, nout=32, want_stats=True, **kw)
    assert st2 is not None
    assert rel_l2(z32, ref2) < 1e-4
    assert rel_l2(st2[0], z32.double().sum((2, 3, 4))) < 2e-5
    assert rel_l2(st2[1], (z32.double() ** 2).sum((2, 3, 4))) < 2e-5

    # fp16 destinations (the pre-norm layout of the fused unit): one source, then the accumulate pass of the second source
    y16h = torch.empty((n, *dims, 32), dtype=torch.float16, device="cuda").permute(0, 4, 1, 2, 3)
    sth = ops._launch_gather(acl, None, ops.pack_conv_fprop(w1), y16h, None, nout=32, want_stats=True, **kw)
    assert sth is not None and rel_l2(y16h.float(), ref1) < 6e-4
    assert rel_l2(sth[0], y32.double().sum((2, 3, 4))) < 2e-5          # statistics come from the fp32 accumulators
    z16h = torch.empty((n, *dims, 32), dtype=torch.float16, device="cuda").permute(0, 4, 1, 2, 3)
    sth2 = ops._launch_gather(acl, bcl, ops.pack_conv_fprop(w2), z16h, None, nout=32, want_stats=True, **kw)
    assert sth2 is not None and rel_l2(z16h.float(), ref2) < 1e-3

    # data gradient of the two-source conv: one 32-channel source, two 32-channel destinations
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    g = q(torch.randn_like(ref2))
    F.conv3d(torch.cat((ar, br), 1), q(w2), None, 1, 1).backward(g)
    ap, bp = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ops.conv3d(ap, w2, 1, x_cat=bp).backward(g.to(torch.bfloat16))
    assert rel_l2(ap.grad.float(), ar.grad) < TOL and rel_l2(bp.grad.float(), br.grad) < TOL


def test_tc5_support_query(rb, built_lib):
    d = rb._lib.ConvDesc()
    for f, v in dict(nsrc=1, srcC0=8, NB=1, ID=4, IH=4, IW=4, tapD=3, tapH=3, tapW=3, offD=-1, offH=-1, offW=-1, istrD=1,
                     istrH=1, istrW=1, OD=4, OH=4, OW=4, Nout=32, ostrD=1, ostrH=1, ostrW=1, FD=4, FH=4, FW=4, outC0=32,
                     psD=1, psH=1, psW=1).items():
        setattr(d, f, v)
    assert built_lib.rb_conv_gather_tc5_supported(ctypes.byref(d)) == 0    # C % 16 != 0 -> mma.sync path
    d.srcC0 = 32
    assert built_lib.rb_conv_gather_tc5_supported(ctypes.byref(d)) == 1


WGRAD_CASES = [
    # n, cin (per source), cout, dims, kernel, stride, two_sources
    (2, 32, 32, (16, 16, 16), (3, 3, 3), (1, 1, 1), False),
    (1, 64, 64, (16, 16, 16), (3, 3, 3), (1, 1, 1), False),
    (2, 32, 64, (16, 16, 16), (3, 3, 3), (2, 2, 2), False),
    (1, 128, 256, (8, 8, 8), (3, 3, 3), (2, 2, 2), False),
    (2, 32, 32, (8, 12, 20), (3, 3, 3), (1, 1, 1), True),
    (1, 64, 64, (10, 6, 12), (3, 3, 3), (1, 1, 1), True),
    (2, 512, 512, (4, 4, 4), (3, 3, 3), (1, 1, 1), False),
    (2, 64, 32, (8, 12, 20), (1, 1, 1), (1, 1, 1), False),
    (1, 32, 64, (8, 16, 16), (1, 3, 3), (1, 2, 2), False),
    (1, 48, 96, (6, 10, 14), (3, 3, 3), (1, 1, 1), False),
    # stride 1, channels % 32 == 0, >= 4096 voxels: two-sided tap stacking (wgrad2_tc5.cuh)
    (1, 32, 32, (32, 32, 32), (3, 3, 3), (1, 1, 1), False),
    (1, 32, 32, (16, 32, 32), (3, 3, 3), (1, 1, 1), True),
    (1, 64, 128, (16, 16, 16), (3, 3, 3), (1, 1, 1), False),
    (1, 128, 64, (12, 20, 24), (3, 3, 3), (1, 1, 1), False),
    (2, 64, 64, (10, 18, 22), (3, 3, 3), (1, 1, 1), True),
    (1, 32, 64, (8, 32, 32), (1, 3, 3), (1, 1, 1), False),
    (1, 96, 32, (16, 16, 16), (3, 3, 3), (1, 1, 1), False),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: f"n{c[0]}_{c[1]}{'x2' if c[6] else ''}to{c[2]}_{'x'.join(map(str, c[3]))}_k{c[4][1]}s{c[5][1]}")
def test_tc5_wgrad(rb, case):
    """tcgen05 weight gradient (MN-major operands straight from the channels-last tensors) == torch fp32 on
    the same bf16 operands, and == the mma.sync kernel."""
    n, cin, cout, dims, k, s, two = case
    torch.manual_seed(5)
    ops = rb.ops
    pad = tuple((kk - 1) // 2 for kk in k)
    x0 = q(torch.randn(n, cin, *dims, device="cuda"))
    x1 = q(torch.randn(n, cin, *dims, device="cuda")) if two else None
    xin = torch.cat((x0, x1), 1) if two else x0
    w = torch.randn(cout, xin.shape[1], *k, device="cuda", requires_grad=True)
    y = F.conv3d(xin, w, None, s, pad)
    g = q(torch.randn_like(y))
    gw_ref = torch.autograd.grad(y, w, g)[0]
    od = tuple(y.shape[2:])
    got = {}
    qct = xin.shape[1]
    tc5_ok = cout % 16 == 0 and cin % 16 == 0 and qct % 32 == 0
    for impl in ("tc5" if tc5_ok else "auto", "mma"):
        dw = ops._launch_wgrad(ops.as_cl(g), ops.as_cl(x0), ops.as_cl(x1) if two else None, grid=od, qdims=dims, taps=k,
                               off=tuple(-p for p in pad), istr=s, impl=impl)
        rb._lib.device_error_check()
        got[impl] = dw.view(*k, cout, xin.shape[1]).permute(3, 4, 0, 1, 2)
    e5, em = rel_l2(got["tc5" if tc5_ok else "auto"], gw_ref), rel_l2(got["mma"], gw_ref)
    print(f"wgrad tc5 {e5:.2e} mma {em:.2e}")
    assert e5 < 2e-4 and em < 2e-4


@pytest.mark.parametrize("cin,cout,stride", [(64, 32, (2, 2, 2)), (512, 256, (2, 2, 2)), (128, 64, (1, 2, 2))])
def test_tc5_wgrad_conv_transpose(rb, cin, cout, stride):
    torch.manual_seed(6)
    x = q(torch.randn(2, cin, 4, 6, 8, device="cuda"))
    w = (torch.randn(cin, cout, *stride, device="cuda") / cin ** 0.5).requires_grad_(True)
    gw_ref_in = w.detach().clone().requires_grad_(True)
    ref = F.conv_transpose3d(x, gw_ref_in, None, stride)
    g = q(torch.randn_like(ref))
    gw_ref = torch.autograd.grad(ref, gw_ref_in, g)[0]
    y = rb.ops.conv_transpose3d(x.clone().requires_grad_(True), w, stride, impl="tc5")
    y.backward(g.to(torch.bfloat16))
    assert rel_l2(w.grad, gw_ref) < 2e-4


# ------------------------------------------------------------------------------------------
# guard-band test: compute-sanitizer is closed on the B200 pool (profiles/r2_sanitize_racecheck_closed.log), so the
# out-of-bounds check is done by hand - every kernel family writes into the middle of a larger allocation whose borders
# hold a sentinel that must survive, and its statistics / workspace buffers get the same treatment
# ------------------------------------------------------------------------------------------
GUARD_CASES = [  # (n, cin, cout, dims, k, stride, note)
    (2, 32, 32, (40, 32, 32), 3, 1, "slab W=32"), (1, 32, 32, (24, 48, 96), 3, 1, "slab W=96 (N=192)"),
    (1, 64, 64, (32, 32, 32), 3, 1, "tc5t h-major"), (2, 32, 32, (16, 16, 16), 3, 1, "tc5"),
    (2, 512, 512, (4, 4, 4), 3, 1, "tap split 4^3"), (2, 256, 256, (16, 16, 16), 3, 1, "tap split 16^3"),
    (1, 32, 64, (16, 32, 32), 3, 2, "strided"), (1, 32, 32, (12, 20, 44), 3, 1, "ragged tiles"),
]


@pytest.mark.parametrize("case", GUARD_CASES, ids=lambda c: c[6].replace(" ", "_"))
def test_conv_kernels_stay_inside_their_buffers(rb, case):
    n, cin, cout, dims, k, s, _ = case
    ops = rb.ops
    torch.manual_seed(3)
    x = ops.as_cl(q(torch.randn(n, cin, *dims, device="cuda")))
    w = torch.randn(cout, cin, k, k, k, device="cuda") / (k ** 3 * cin) ** 0.5
    od = ops._conv_out_dims(dims, (k,) * 3, (s,) * 3)
    nel = n * od[0] * od[1] * od[2] * cout
    G = 4096                                      # guard elements either side (16-byte aligned for both dtypes)
    SENT = 12345.0
    for dt, tol in ((torch.bfloat16, TOL), (torch.float32, TOL), (torch.float16, TOL)):
        big = torch.full((nel + 2 * G,), SENT, dtype=dt, device="cuda")
        sent = big[0].clone()
        y = big[G:G + nel].view(n, *od, cout).permute(0, 4, 1, 2, 3)
        sbig = torch.full((2 * n * cout + 2 * 64,), SENT, dtype=torch.float32, device="cuda")
        st = sbig[64:64 + 2 * n * cout].view(2, n, cout)
        st.zero_()
        pad = (k - 1) // 2
        ops._launch_gather(x, None, ops.pack_conv_fprop(w), y, None, in_dims=dims, taps=(k,) * 3, off=(-pad,) * 3, istr=(s,) * 3,
                           out_grid=od, nout=cout, stats=(st[0], st[1]))
        torch.cuda.synchronize()
        rb._lib.device_error_check()
        assert bool((big[:G] == sent).all()) and bool((big[G + nel:] == sent).all()), "conv wrote outside its destination"
        assert bool((sbig[:64] == SENT).all()) and bool((sbig[64 + 2 * n * cout:] == SENT).all()), "statistics overran"
        ref = F.conv3d(x.float(), q(w), None, s, pad)
        r = rel_l2(y.float(), ref)
        assert r < (tol if dt == torch.bfloat16 else 6e-4 if dt == torch.float16 else 1e-4), (dt, r)
        assert rel_l2(st[0], ref.double().sum((2, 3, 4))) < 5e-3
    # data gradient and weight gradient through the same descriptors (allocated by the library's host side)
    xr = x.detach().clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    g = q(torch.randn(n, cout, *od, device="cuda"))
    ops.conv3d(xr, wr, s).backward(g.to(torch.bfloat16))
    rb._lib.device_error_check()
    xt = x.float().detach().clone().requires_grad_(True)
    wt = q(w).clone().requires_grad_(True)
    F.conv3d(xt, wt, None, s, pad).backward(g)
    assert rel_l2(xr.grad.float(), xt.grad) < TOL and rel_l2(wr.grad, wt.grad) < 2e-3
