"""Functional layer over the C ABI: autograd.Functions whose forward and backward launch the
hand-written sm_100a kernels.  PyTorch is used for storage, autograd bookkeeping and O(N*C)-sized
glue only; no torch operator touches a full-resolution activation on the hot path.

Tensor convention: activations travel as *logically* NCDHW tensors (the shapes the reference's
modules see) whose memory is dense channels-last NDHWC bf16 -- i.e. ``buf.permute(0, 4, 1, 2, 3)``
of a contiguous ``[N, D, H, W, C]`` buffer.  Heads return plain NCDHW fp32 like the reference.

Reference arithmetic replaced (path:line under the reference tree):
  conv3d / conv_transpose3d     builders/simple_conv_blocks.py:43-51, builders/decoder.py:110-113,147
  instance_norm + leaky_relu    builders/simple_conv_blocks.py:58-64, build_network_from_config.py:172,208-210
  residual add + SE gate        builders/resblocks.py:106-114
  avg_pool3d                    builders/resblocks.py:92-95
  1x1x1 head (+ activation)     builders/decoder.py:131,151-152, build_network_from_config.py:322-323
"""
from __future__ import annotations

import contextlib
import ctypes as C
import itertools

import torch

from . import _lib as L

BF16 = torch.bfloat16
LRELU_SLOPE_DEFAULT = 0.01


class _KernelTimer:
    """Optional CUDA-event spans around kernel launches on the launching stream (bench.py's roofline
    numbers: per-category device time and algorithmic FLOPs).  Disabled unless bench.py enables it."""

    def __init__(self):
        self.on = False
        self.records = []

    def enable(self, on=True):
        self.on = bool(on)
        if on:
            self.records = []

    @contextlib.contextmanager
    def span(self, kind, flops=0.0):
        if not self.on:
            yield
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        self.records.append((kind, flops, a, b))

    def summary(self):
        if not self.records:
            return {}
        torch.cuda.synchronize()
        out = {}
        for kind, flops, a, b in self.records:
            d = out.setdefault(kind, {"ms": 0.0, "flops": 0.0, "launches": 0})
            d["ms"] += a.elapsed_time(b)
            d["flops"] += flops
            d["launches"] += 1
        return out


KERNEL_TIMER = _KernelTimer()


# ------------------------------------------------------------------------------------------
# layout helpers
# ------------------------------------------------------------------------------------------
def new_cl(n, c, d, h, w, device):
    """Uninitialised activation: logical [n, c, d, h, w], memory NDHWC bf16."""
    return torch.empty((n, d, h, w, c), dtype=BF16, device=device).permute(0, 4, 1, 2, 3)


def zeros_cl(n, c, d, h, w, device):
    return torch.zeros((n, d, h, w, c), dtype=BF16, device=device).permute(0, 4, 1, 2, 3)


def is_cl(x: torch.Tensor) -> bool:
    return x.dim() == 5 and x.dtype == BF16 and x.permute(0, 2, 3, 4, 1).is_contiguous()


def as_cl(x: torch.Tensor) -> torch.Tensor:
    """Return `x` as a dense NDHWC bf16 activation (no copy when it already is one)."""
    if is_cl(x):
        return x
    L.require_cuda(x, "as_cl")
    if x.dim() != 5:
        raise ValueError(f"expected a 5-D [N, C, D, H, W] tensor, got shape {tuple(x.shape)}")
    n, c, d, h, w = x.shape
    if c % 8 != 0:
        raise ValueError(f"channels-last bf16 activations need C % 8 == 0, got C={c}")
    if x.dtype == torch.float32 and x.is_contiguous():
        out = new_cl(n, c, d, h, w, x.device)
        L.check(L.load().rb_ncdhw_to_cl(x.data_ptr(), out.data_ptr(), n, c, d * h * w, L.stream_ptr()), "rb_ncdhw_to_cl")
        return out
    # dtype / stride normalisation of foreign tensors (fp16 under autocast, sliced views): glue
    return x.to(BF16).contiguous(memory_format=torch.channels_last_3d)


def cl_to_ncdhw_f32(x: torch.Tensor) -> torch.Tensor:
    x = as_cl(x)
    n, c, d, h, w = x.shape
    out = torch.empty((n, c, d, h, w), dtype=torch.float32, device=x.device)
    L.check(L.load().rb_cl_to_ncdhw(x.data_ptr(), out.data_ptr(), n, c, d * h * w, L.stream_ptr()), "rb_cl_to_ncdhw")
    return out


def _triple(v):
    if isinstance(v, (tuple, list)):
        if len(v) != 3:
            raise NotImplementedError(f"only 3-D operators are implemented on the B200 path, got {v}")
        return tuple(int(i) for i in v)
    return (int(v),) * 3


# ------------------------------------------------------------------------------------------
# raw launches
# ------------------------------------------------------------------------------------------
def _launch_gather(src0, src1, wpk, out0, out1, *, in_dims, taps, off, istr, out_grid, nout, mode=0,
                   ostr=(1, 1, 1), ooff=(0, 0, 0), full=None, ps=None, psC=0, impl=None, stats=None):
    """One rb_conv_gather call.  src*/out* are NDHWC bf16 activations (logical NCDHW views)."""
    lib = L.load()
    d = L.ConvDesc()
    d.nsrc = 2 if src1 is not None else 1
    d.srcC0 = src0.shape[1]
    d.srcC1 = src1.shape[1] if src1 is not None else 0
    d.NB = src0.shape[0]
    d.ID, d.IH, d.IW = in_dims
    d.tapD, d.tapH, d.tapW = taps
    d.offD, d.offH, d.offW = off
    d.istrD, d.istrH, d.istrW = istr
    d.OD, d.OH, d.OW = out_grid
    d.Nout = nout
    d.mode = mode
    d.ostrD, d.ostrH, d.ostrW = ostr
    d.ooffD, d.ooffH, d.ooffW = ooff
    d.FD, d.FH, d.FW = full if full is not None else out_grid
    d.outC0 = out0.shape[1]
    d.outC1 = out1.shape[1] if out1 is not None else 0
    if ps is not None:
        d.psD, d.psH, d.psW = ps
        d.psC = psC
    else:
        d.psD = d.psH = d.psW = 1
        d.psC = 0
    d.impl = L.default_impl() if impl is None else L.impl_code(impl)
    d.splitK = 0
    ws_bytes = lib.rb_conv_gather_workspace(C.byref(d))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=src0.device) if ws_bytes else None
    ssum, ssq = (None, None) if stats is None else stats
    flops = 2.0 * d.NB * d.OD * d.OH * d.OW * nout * (d.srcC0 + d.srcC1) * taps[0] * taps[1] * taps[2]
    with KERNEL_TIMER.span("conv", flops):
        rc = lib.rb_conv_gather(C.byref(d), src0.data_ptr(), L.ptr(src1), wpk.data_ptr(), out0.data_ptr(), L.ptr(out1),
                                L.ptr(ssum), L.ptr(ssq), L.ptr(ws), ws_bytes, L.stream_ptr())
    L.check(rc, "rb_conv_gather")


def _launch_wgrad(P, Q0, Q1, *, grid, qdims, taps, off, istr):
    lib = L.load()
    d = L.WgradDesc()
    d.PC = P.shape[1]
    d.nq = 2 if Q1 is not None else 1
    d.QC0 = Q0.shape[1]
    d.QC1 = Q1.shape[1] if Q1 is not None else 0
    d.NB = P.shape[0]
    d.GD, d.GH, d.GW = grid
    d.QD, d.QH, d.QW = qdims
    d.tapD, d.tapH, d.tapW = taps
    d.offD, d.offH, d.offW = off
    d.istrD, d.istrH, d.istrW = istr
    d.splits = 0
    ntaps = taps[0] * taps[1] * taps[2]
    dw = torch.zeros((ntaps, d.PC, d.QC0 + d.QC1), dtype=torch.float32, device=P.device)
    flops = 2.0 * d.NB * d.GD * d.GH * d.GW * d.PC * (d.QC0 + d.QC1) * ntaps
    with KERNEL_TIMER.span("wgrad", flops):
        rc = lib.rb_wgrad_gather(C.byref(d), P.data_ptr(), Q0.data_ptr(), L.ptr(Q1), dw.data_ptr(), L.stream_ptr())
    L.check(rc, "rb_wgrad_gather")
    return dw


# ------------------------------------------------------------------------------------------
# weight packing (fp32 canonical parameter -> bf16 [taps][rows][cols]); small tensors, torch glue.
# The fprop pack is cached on the Parameter and reused until the optimiser touches it, so the
# sliding-window sweep packs each weight once.
# ------------------------------------------------------------------------------------------
def _cached_pack(weight, kind, fn):
    key = (weight._version, weight.data_ptr(), str(weight.device))
    cache = getattr(weight, "_rb_pack", None)
    if cache is None or cache.get("key") != key:
        cache = {"key": key}
        try:
            weight._rb_pack = cache
        except Exception:  # pragma: no cover - exotic tensor subclasses
            return fn()
    if kind not in cache:
        with torch.no_grad():
            cache[kind] = fn()
    return cache[kind]


def pack_conv_fprop(weight):
    """[Cout, Cin, kd, kh, kw] -> [taps][Cout][Cin] bf16."""
    co, ci, kd, kh, kw = weight.shape
    return _cached_pack(weight, "f", lambda: weight.detach().permute(2, 3, 4, 0, 1).reshape(kd * kh * kw, co, ci)
                        .to(BF16).contiguous())


def _axis_classes(K, s, pad, I):
    """Data-gradient decomposition along one axis: input coordinate i = s*j + r (class r) receives
    from output o = j + off + t through kernel index klist[t]  (i = o*s + k - pad)."""
    out = []
    for r in range(s):
        cnt = len(range(r, I, s))
        if cnt == 0:
            continue
        ks = [k for k in range(K - 1, -1, -1) if (r + pad - k) % s == 0]
        off = (r + pad - ks[0]) // s if ks else 0
        out.append((r, cnt, ks, off))
    return out


def pack_conv_dgrad_class(weight, kds, khs, kws):
    """taps selected by kernel-index lists per axis -> [taps][Cin][Cout] bf16."""
    w = weight.detach()
    dev = w.device
    w = w.index_select(2, torch.tensor(kds, device=dev)).index_select(3, torch.tensor(khs, device=dev)) \
         .index_select(4, torch.tensor(kws, device=dev))
    co, ci = w.shape[:2]
    return w.permute(2, 3, 4, 1, 0).reshape(len(kds) * len(khs) * len(kws), ci, co).to(BF16).contiguous()


# ------------------------------------------------------------------------------------------
# Conv3d (k in {1,3} per axis, stride in {1,2} per axis, pad (k-1)//2, one or two concatenated inputs)
# ------------------------------------------------------------------------------------------
def _conv_out_dims(in_dims, k, s):
    return tuple((i + 2 * ((kk - 1) // 2) - kk) // ss + 1 for i, kk, ss in zip(in_dims, k, s))


class _Conv3dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, stride, impl, x0, x1):
        x0 = as_cl(x0)
        x1 = as_cl(x1) if x1 is not None else None
        L.require_cuda(x0, "conv3d")
        co, ci, kd, kh, kw = weight.shape
        c_in = x0.shape[1] + (x1.shape[1] if x1 is not None else 0)
        if c_in != ci:
            raise ValueError(f"conv3d: weight expects {ci} input channels, got {c_in}")
        k = (kd, kh, kw)
        if any(kk not in (1, 3) for kk in k):
            raise NotImplementedError(f"conv3d: kernel sizes 1 and 3 are implemented, got {k}")
        n = x0.shape[0]
        in_dims = tuple(x0.shape[2:])
        od = _conv_out_dims(in_dims, k, stride)
        y = new_cl(n, co, *od, x0.device)
        _launch_gather(x0, x1, pack_conv_fprop(weight), y, None, in_dims=in_dims, taps=k,
                       off=tuple(-((kk - 1) // 2) for kk in k), istr=stride, out_grid=od, nout=co, impl=impl)
        ctx.save_for_backward(weight, x0, x1)
        ctx.stride, ctx.impl = stride, impl
        return y

    @staticmethod
    def backward(ctx, dy):
        weight, x0, x1 = ctx.saved_tensors
        dy = as_cl(dy)
        stride, impl = ctx.stride, ctx.impl
        co, ci, kd, kh, kw = weight.shape
        k = (kd, kh, kw)
        pad = tuple((kk - 1) // 2 for kk in k)
        n = x0.shape[0]
        in_dims = tuple(x0.shape[2:])
        od = tuple(dy.shape[2:])
        gw = gx0 = gx1 = None
        if ctx.needs_input_grad[0]:
            dw = _launch_wgrad(dy, x0, x1, grid=od, qdims=in_dims, taps=k, off=tuple(-p for p in pad), istr=stride)
            gw = dw.view(kd, kh, kw, co, ci).permute(3, 4, 0, 1, 2).contiguous()
        need0 = ctx.needs_input_grad[3]
        need1 = x1 is not None and ctx.needs_input_grad[4]
        if need0 or need1:
            c0 = x0.shape[1]
            classes = [_axis_classes(k[a], stride[a], pad[a], in_dims[a]) for a in range(3)]
            empty = any(len(cls[2]) == 0 for axis in classes for cls in axis)
            mk = zeros_cl if empty else new_cl
            gx0 = mk(n, c0, *in_dims, x0.device)
            gx1 = mk(n, x1.shape[1], *in_dims, x0.device) if x1 is not None else None
            for cd, ch, cw in itertools.product(*classes):
                if not (cd[2] and ch[2] and cw[2]):
                    continue
                wpk = pack_conv_dgrad_class(weight, cd[2], ch[2], cw[2])
                _launch_gather(dy, None, wpk, gx0, gx1, in_dims=od, taps=(len(cd[2]), len(ch[2]), len(cw[2])),
                               off=(cd[3], ch[3], cw[3]), istr=(1, 1, 1), out_grid=(cd[1], ch[1], cw[1]), nout=ci,
                               ostr=stride, ooff=(cd[0], ch[0], cw[0]), full=in_dims, impl=impl)
            if not need0:
                gx0 = None
            if not need1:
                gx1 = None
        return gw, None, None, gx0, gx1


def conv3d(x, weight, stride=1, x_cat=None, impl=None):
    """Pre-norm convolution output (bf16, channels-last).  `x_cat` is concatenated after `x`
    along channels without materialising the concatenation."""
    return _Conv3dFn.apply(weight, _triple(stride), impl, x, x_cat)


class _ZeroGradParamFn(torch.autograd.Function):
    """Ties a parameter whose effect cancels exactly (a conv bias in front of InstanceNorm) into the
    graph so that it receives the exact gradient, zero, as it (up to rounding) does in the reference."""

    @staticmethod
    def forward(ctx, y, p):
        ctx.pshape, ctx.pdev = tuple(p.shape), p.device
        return y.view_as(y)

    @staticmethod
    def backward(ctx, dy):
        return dy, torch.zeros(ctx.pshape, dtype=torch.float32, device=ctx.pdev)


def attach_cancelled_bias(y, bias):
    return y if bias is None else _ZeroGradParamFn.apply(y, bias)


# ------------------------------------------------------------------------------------------
# ConvTranspose3d with kernel == stride (non-overlapping "pixel shuffle" upsampling)
# ------------------------------------------------------------------------------------------
class _ConvT3dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, stride, impl, x):
        x = as_cl(x)
        L.require_cuda(x, "conv_transpose3d")
        ci, co, sd, sh, sw = weight.shape
        if (sd, sh, sw) != stride:
            raise NotImplementedError("conv_transpose3d: only kernel_size == stride is implemented")
        if x.shape[1] != ci:
            raise ValueError(f"conv_transpose3d: weight expects {ci} input channels, got {x.shape[1]}")
        n = x.shape[0]
        in_dims = tuple(x.shape[2:])
        full = tuple(i * s for i, s in zip(in_dims, stride))
        npar = sd * sh * sw
        y = new_cl(n, co, *full, x.device)
        wpk = _cached_pack(weight, "t", lambda: weight.detach().permute(2, 3, 4, 1, 0).reshape(1, npar * co, ci)
                           .to(BF16).contiguous())
        _launch_gather(x, None, wpk, y, None, in_dims=in_dims, taps=(1, 1, 1), off=(0, 0, 0), istr=(1, 1, 1),
                       out_grid=in_dims, nout=npar * co, mode=1, ostr=stride, full=full, ps=stride, psC=co, impl=impl)
        ctx.save_for_backward(weight, x)
        ctx.stride, ctx.impl = stride, impl
        return y

    @staticmethod
    def backward(ctx, dy):
        weight, x = ctx.saved_tensors
        dy = as_cl(dy)
        stride, impl = ctx.stride, ctx.impl
        ci, co, sd, sh, sw = weight.shape
        n = x.shape[0]
        in_dims = tuple(x.shape[2:])
        full = tuple(dy.shape[2:])
        npar = sd * sh * sw
        gw = gx = None
        if ctx.needs_input_grad[0]:
            dw = _launch_wgrad(x, dy, None, grid=in_dims, qdims=full, taps=stride, off=(0, 0, 0), istr=stride)
            gw = dw.view(sd, sh, sw, ci, co).permute(3, 4, 0, 1, 2).contiguous()
        if ctx.needs_input_grad[3]:
            gx = new_cl(n, ci, *in_dims, x.device)
            wpk = weight.detach().permute(2, 3, 4, 0, 1).reshape(npar, ci, co).to(BF16).contiguous()
            _launch_gather(dy, None, wpk, gx, None, in_dims=full, taps=stride, off=(0, 0, 0), istr=stride,
                           out_grid=in_dims, nout=ci, impl=impl)
        return gw, None, None, gx


def conv_transpose3d(x, weight, stride, impl=None):
    return _ConvT3dFn.apply(weight, _triple(stride), impl, x)


# ------------------------------------------------------------------------------------------
# InstanceNorm (+affine) + residual + LeakyReLU, closed-form backward (no gate)
# ------------------------------------------------------------------------------------------
def _plane_reduce(kind, y, dz, z, per_w, slope):
    n, c, d, h, w = y.shape
    g = w if per_w else 1
    out = torch.empty((n, g, c, 2), dtype=torch.float64, device=y.device)
    with KERNEL_TIMER.span("norm_reduce"):
        rc = L.load().rb_plane_reduce(kind, y.data_ptr(), L.ptr(dz), L.ptr(z), out.data_ptr(), n, d * h * w, c, w,
                                      1 if per_w else 0, float(slope), L.stream_ptr())
    L.check(rc, "rb_plane_reduce")
    return out


class _NormActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, res, gamma, beta, eps, act, slope):
        y = as_cl(y)
        L.require_cuda(y, "instance_norm")
        res = as_cl(res) if res is not None else None
        if res is not None and res.shape != y.shape:
            raise ValueError(f"residual shape {tuple(res.shape)} != {tuple(y.shape)}")
        n, c, d, h, w = y.shape
        S = d * h * w
        lib = L.load()
        st = L.stream_ptr()
        sums = _plane_reduce(0, y, None, None, False, slope)
        small = torch.empty((4, n, c), dtype=torch.float32, device=y.device)  # mean, rstd, scale, shift
        L.check(lib.rb_in_finalize_fwd(sums.data_ptr(), L.ptr(gamma), L.ptr(beta), small[0].data_ptr(), small[1].data_ptr(),
                                       small[2].data_ptr(), small[3].data_ptr(), n, c, float(S), float(eps), st),
                "rb_in_finalize_fwd")
        z = new_cl(n, c, d, h, w, y.device)
        with KERNEL_TIMER.span("norm_apply"):
            rc = lib.rb_norm_act_fwd(y.data_ptr(), L.ptr(res), z.data_ptr(), small[2].data_ptr(), small[3].data_ptr(),
                                     n, S, c, w, 0, 1 if act else 0, float(slope), st)
        L.check(rc, "rb_norm_act_fwd")
        ctx.save_for_backward(y, z if act else None, small, gamma)
        ctx.act, ctx.slope, ctx.has_res, ctx.has_beta = act, slope, res is not None, beta is not None
        return z

    @staticmethod
    def backward(ctx, dz):
        y, z, small, gamma = ctx.saved_tensors
        dz = as_cl(dz)
        n, c, d, h, w = y.shape
        S = d * h * w
        lib = L.load()
        st = L.stream_ptr()
        red = _plane_reduce(1, y, dz, z, False, ctx.slope)
        ks = torch.empty((3, n, c), dtype=torch.float32, device=y.device)
        dgamma = torch.zeros(c, dtype=torch.float32, device=y.device) if gamma is not None else None
        dbeta = torch.zeros(c, dtype=torch.float32, device=y.device) if ctx.has_beta else None
        L.check(lib.rb_in_finalize_bwd(red.data_ptr(), small[0].data_ptr(), small[1].data_ptr(), L.ptr(gamma),
                                       ks[0].data_ptr(), ks[1].data_ptr(), ks[2].data_ptr(), L.ptr(dgamma), L.ptr(dbeta),
                                       n, c, float(S), st), "rb_in_finalize_bwd")
        dy = new_cl(n, c, d, h, w, y.device)
        dres = new_cl(n, c, d, h, w, y.device) if (ctx.has_res and ctx.needs_input_grad[1]) else None
        with KERNEL_TIMER.span("norm_apply"):
            rc = lib.rb_norm_act_bwd(dz.data_ptr(), L.ptr(z), y.data_ptr(), dy.data_ptr(), L.ptr(dres), ks[0].data_ptr(),
                                     ks[1].data_ptr(), ks[2].data_ptr(), n, S, c, w, 0, 1 if ctx.act else 0,
                                     float(ctx.slope), st)
        L.check(rc, "rb_norm_act_bwd")
        return dy, dres, dgamma, dbeta, None, None, None


def instance_norm_act(y, res=None, gamma=None, beta=None, eps=1e-5, act=True, slope=LRELU_SLOPE_DEFAULT):
    """z = [LeakyReLU]( InstanceNorm(y) [* gamma + beta] [+ res] )."""
    return _NormActFn.apply(y, res, gamma, beta, float(eps), bool(act), float(slope))


# ------------------------------------------------------------------------------------------
# InstanceNorm + squeeze-excitation gate + residual + LeakyReLU.
# The O(N*W*C) part (statistics -> scale/shift, the SE bottleneck MLP and its backward) is a small
# differentiable torch graph; the two full-tensor passes per direction are the same kernels as above.
# ------------------------------------------------------------------------------------------
def _gate_small_graph(S1, S2, Pw, count, plane, gamma, beta, eps, w1, b1, w2, b2, per_w):
    mean = S1 / count
    var = (S2 / count - mean * mean).clamp_min(0.0)
    rstd = torch.rsqrt(var + eps)
    a0 = rstd if gamma is None else rstd * gamma.double()
    b0 = -mean * a0 if beta is None else beta.double() - mean * a0
    if per_w:                                     # squeeze over (D, H): one value per (n, w, c)
        sq = a0.unsqueeze(1) * (Pw / plane) + b0.unsqueeze(1)
    else:                                         # global average pool of the normalised tensor
        sq = (a0 * mean + b0).unsqueeze(1)
    hid = torch.relu(torch.nn.functional.linear(sq.float(), w1.flatten(1), b1))
    gate = torch.sigmoid(torch.nn.functional.linear(hid, w2.flatten(1), b2)).double()
    A = (a0.unsqueeze(1) * gate).float()
    B = (b0.unsqueeze(1) * gate).float()
    return A.contiguous(), B.contiguous()          # [N, G, C]


class _NormGateActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, res, gamma, beta, w1, b1, w2, b2, eps, act, slope, per_w):
        y = as_cl(y)
        L.require_cuda(y, "instance_norm_se")
        res = as_cl(res) if res is not None else None
        n, c, d, h, w = y.shape
        S = d * h * w
        lib = L.load()
        st = L.stream_ptr()
        sums = _plane_reduce(0, y, None, None, False, slope)[:, 0]          # [N, C, 2]
        pw = _plane_reduce(0, y, None, None, True, slope)[..., 0] if per_w else None   # [N, W, C]
        params = [gamma, beta, w1, b1, w2, b2]
        with torch.enable_grad():
            leaves = [sums[..., 0].detach().requires_grad_(True), sums[..., 1].detach().requires_grad_(True),
                      pw.detach().requires_grad_(True) if per_w else None]
            pl = [p.detach().requires_grad_(True) if p is not None else None for p in params]
            A, B = _gate_small_graph(leaves[0], leaves[1], leaves[2], float(S), float(d * h), pl[0], pl[1], eps,
                                     pl[2], pl[3], pl[4], pl[5], per_w)
        z = new_cl(n, c, d, h, w, y.device)
        L.check(lib.rb_norm_act_fwd(y.data_ptr(), L.ptr(res), z.data_ptr(), A.data_ptr(), B.data_ptr(), n, S, c, w,
                                    1 if per_w else 0, 1 if act else 0, float(slope), st), "rb_norm_act_fwd")
        ctx.save_for_backward(y, z if act else None)
        ctx.graph = (leaves, pl, A, B)
        ctx.cfg = (act, slope, per_w, res is not None)
        return z

    @staticmethod
    def backward(ctx, dz):
        y, z = ctx.saved_tensors
        dz = as_cl(dz)
        act, slope, per_w, has_res = ctx.cfg
        leaves, pl, A, B = ctx.graph
        n, c, d, h, w = y.shape
        S = d * h * w
        lib = L.load()
        st = L.stream_ptr()
        red = _plane_reduce(1, y, dz, z, per_w, slope)          # [N, G, C, 2] = (sum g, sum g*y)
        dB = red[..., 0].float()
        dA = red[..., 1].float()
        inputs = [t for t in leaves + pl if t is not None]
        grads = torch.autograd.grad([A, B], inputs, [dA, dB], allow_unused=True)
        it = iter(grads)
        gl = [next(it) if t is not None else None for t in leaves]
        gp = [next(it) if t is not None else None for t in pl]
        g = A.shape[1]
        zero = torch.zeros((n, c), dtype=torch.float64, device=y.device)
        dS1 = gl[0] if gl[0] is not None else zero
        dS2 = gl[1] if gl[1] is not None else zero
        k1 = A
        k2 = (2.0 * dS2).float().unsqueeze(1).expand(n, g, c).contiguous()
        k3 = dS1.unsqueeze(1).expand(n, g, c)
        if per_w and gl[2] is not None:
            k3 = k3 + gl[2]
        k3 = k3.float().contiguous()
        dy = new_cl(n, c, d, h, w, y.device)
        dres = new_cl(n, c, d, h, w, y.device) if (has_res and ctx.needs_input_grad[1]) else None
        L.check(lib.rb_norm_act_bwd(dz.data_ptr(), L.ptr(z), y.data_ptr(), dy.data_ptr(), L.ptr(dres), k1.data_ptr(),
                                    k2.data_ptr(), k3.data_ptr(), n, S, c, w, 1 if per_w else 0, 1 if act else 0,
                                    float(slope), st), "rb_norm_act_bwd")
        gp = [None if v is None else v for v in gp]
        return (dy, dres, gp[0], gp[1], gp[2], gp[3], gp[4], gp[5], None, None, None, None)


def instance_norm_se_act(y, res, gamma, beta, fc1_w, fc1_b, fc2_w, fc2_b, eps=1e-5, act=True,
                         slope=LRELU_SLOPE_DEFAULT, reduce_dims="all"):
    """z = [LeakyReLU]( SE(InstanceNorm(y)) + res ), SE(o) = o * sigmoid(fc2(relu(fc1(mean_dims(o)))))."""
    if reduce_dims in ("all", (2, 3, 4), [2, 3, 4]):
        per_w = False
    elif tuple(reduce_dims) == (2, 3):
        per_w = True
    else:
        raise NotImplementedError(f"SE squeeze over dims {reduce_dims} is not implemented (use 'all' or (2, 3))")
    return _NormGateActFn.apply(y, res, gamma, beta, fc1_w, fc1_b, fc2_w, fc2_b, float(eps), bool(act), float(slope), per_w)


# ------------------------------------------------------------------------------------------
# AvgPool3d(kernel == stride)
# ------------------------------------------------------------------------------------------
class _AvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, stride):
        x = as_cl(x)
        L.require_cuda(x, "avg_pool3d")
        n, c, d, h, w = x.shape
        sd, sh, sw = stride
        out = new_cl(n, c, d // sd, h // sh, w // sw, x.device)
        L.check(L.load().rb_avgpool_fwd(x.data_ptr(), out.data_ptr(), n, d, h, w, c, sd, sh, sw, L.stream_ptr()),
                "rb_avgpool_fwd")
        ctx.shape, ctx.stride = (n, c, d, h, w), stride
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = as_cl(dout)
        n, c, d, h, w = ctx.shape
        sd, sh, sw = ctx.stride
        din = new_cl(n, c, d, h, w, dout.device)
        L.check(L.load().rb_avgpool_bwd(dout.data_ptr(), din.data_ptr(), n, d, h, w, c, sd, sh, sw, L.stream_ptr()),
                "rb_avgpool_bwd")
        return din, None


def avg_pool3d(x, stride):
    return _AvgPoolFn.apply(x, _triple(stride))


# ------------------------------------------------------------------------------------------
# Task head: 1x1x1 conv with bias -> NCDHW fp32 logits (+ eval activation)
# ------------------------------------------------------------------------------------------
_ACT = {None: 0, "none": 0, "sigmoid": 1, "softmax": 2}


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        x = as_cl(x)
        L.require_cuda(x, "head")
        n, c, d, h, w = x.shape
        k = weight.shape[0]
        w2 = weight.detach().reshape(k, c).float().contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        out = torch.empty((n, k, d, h, w), dtype=torch.float32, device=x.device)
        L.check(L.load().rb_head_fwd(x.data_ptr(), w2.data_ptr(), L.ptr(b), out.data_ptr(), n, d * h * w, c, k, act,
                                     L.stream_ptr()), "rb_head_fwd")
        ctx.save_for_backward(x, w2)
        ctx.act, ctx.has_bias, ctx.wshape = act, bias is not None, tuple(weight.shape)
        return out

    @staticmethod
    def backward(ctx, dl):
        if ctx.act != 0:
            raise NotImplementedError("backward through the fused eval-mode activation is not implemented "
                                      "(the reference applies it only when not self.training)")
        x, w2 = ctx.saved_tensors
        n, c, d, h, w = x.shape
        k = w2.shape[0]
        dl = dl.float().contiguous()
        dx = new_cl(n, c, d, h, w, x.device)
        dw = torch.zeros((k, c), dtype=torch.float32, device=x.device)
        db = torch.zeros(k, dtype=torch.float32, device=x.device)
        L.check(L.load().rb_head_bwd(x.data_ptr(), w2.data_ptr(), dl.data_ptr(), dx.data_ptr(), dw.data_ptr(),
                                     db.data_ptr(), n, d * h * w, c, k, L.stream_ptr()), "rb_head_bwd")
        return dx, dw.view(ctx.wshape), (db if ctx.has_bias else None), None


def head_conv1x1(x, weight, bias, activation=None):
    if weight.shape[0] > 8:
        raise NotImplementedError("task heads with more than 8 output channels are not implemented")
    act = _ACT[activation if activation is None else str(activation).lower()]
    return _HeadFn.apply(x, weight, bias, act)


# ------------------------------------------------------------------------------------------
# Stem convolution on the raw NCDHW fp32 input (Cin not a multiple of 8): im2col + 1-tap GEMM
# ------------------------------------------------------------------------------------------
class _StemConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, impl):
        L.require_cuda(x, "stem conv")
        x = x.float().contiguous()
        n, ci, d, h, w = x.shape
        co, ci_w, kd, kh, kw = weight.shape
        if ci != ci_w:
            raise ValueError(f"stem conv: weight expects {ci_w} input channels, got {ci}")
        K = kd * kh * kw * ci
        kp = (K + 15) // 16 * 16
        col = new_cl(n, kp, d, h, w, x.device)
        L.check(L.load().rb_stem_im2col(x.data_ptr(), col.data_ptr(), n, ci, d, h, w, kd, kh, kw, kp, L.stream_ptr()),
                "rb_stem_im2col")

        def pack():
            wp = torch.zeros((1, co, kp), dtype=BF16, device=x.device)
            wp[0, :, :K] = weight.detach().permute(0, 2, 3, 4, 1).reshape(co, K).to(BF16)
            return wp
        y = new_cl(n, co, d, h, w, x.device)
        _launch_gather(col, None, _cached_pack(weight, "s", pack), y, None, in_dims=(d, h, w), taps=(1, 1, 1),
                       off=(0, 0, 0), istr=(1, 1, 1), out_grid=(d, h, w), nout=co, impl=impl)
        ctx.save_for_backward(col)
        ctx.wshape = tuple(weight.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        (col,) = ctx.saved_tensors
        dy = as_cl(dy)
        co, ci, kd, kh, kw = ctx.wshape
        K = kd * kh * kw * ci
        dims = tuple(dy.shape[2:])
        gw = None
        if ctx.needs_input_grad[1]:
            dw = _launch_wgrad(dy, col, None, grid=dims, qdims=dims, taps=(1, 1, 1), off=(0, 0, 0), istr=(1, 1, 1))
            gw = dw[0, :, :K].reshape(co, kd, kh, kw, ci).permute(0, 4, 1, 2, 3).contiguous()
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("gradient w.r.t. the network input is not implemented")
        return None, gw, None


def stem_conv3d(x, weight, impl=None):
    """Stride-1 'same' convolution of the raw network input (any Cin)."""
    if x.dtype == BF16 and is_cl(x) and x.shape[1] % 8 == 0:
        return conv3d(x, weight, 1, impl=impl)
    if weight.shape[1] * weight.shape[2] * weight.shape[3] * weight.shape[4] > 1024:
        raise NotImplementedError("stem im2col path supports up to 1024 (taps x input channels)")
    return _StemConvFn.apply(x, weight, impl)
