"""Split-precision ("bf16x3") inference tier: the north star's "within relative L2 1e-4 with fp32 accumulation"
bound and the >= 99.9 % argmax agreement, on the same tensor-core kernels as the bf16 path.

Every stored activation v travels as hi = bf16(v), lo = bf16(v - hi) in ONE channels-last tensor whose voxel row is
[hi(C) | lo(C) | hi(C)]; every weight as [hi_w | hi_w | lo_w] along its input-channel axis.  A convolution over the
3C-channel row with fp32 accumulation then evaluates hi*hi_w + lo*hi_w + hi*lo_w (16 + 16 mantissa bits, the dropped
lo*lo_w term is ~2^-18 relative) in a single `rb_conv_gather` launch; InstanceNorm statistics are double sums over the
fp32 result, and normalise / SE gate / residual / LeakyReLU run in fp32 inside `rb_split_apply`, which writes the next
split row.  `ops.precise_inference()` switches the four operators the network is built from to the functions below;
the modules, their parameters and the sliding-window driver are unchanged.  Inference only (no autograd).

Reference arithmetic replaced: the same call sites as ops.py (builders/simple_conv_blocks.py:43-64,
builders/resblocks.py:92-114, builders/decoder.py:110-113,131,147-152), evaluated to fp32 accuracy.
"""
from __future__ import annotations

import os

import torch

from . import _lib as L
from . import ops

BF16 = torch.bfloat16
# which gather-conv implementation runs the 3C-wide contractions ("auto" = the tcgen05 kernels where the widened
# shapes qualify, "mma" = the shape-generic mma.sync kernel)
IMPL = os.environ.get("RESENC_PRECISE_IMPL", "auto")


def _split(w: torch.Tensor):
    hi = w.to(BF16)
    return hi, (w - hi.float()).to(BF16)


def _pack3(w: torch.Tensor) -> torch.Tensor:
    """fp32 [..., K] -> bf16 [..., 3K] = [hi | hi | lo] (pairs with activation rows [hi | lo | hi])."""
    hi, lo = _split(w.float())
    return torch.cat((hi, hi, lo), -1)


def _logical_channels(x: torch.Tensor, what: str) -> int:
    if not ops.is_cl(x) or x.shape[1] % 24 != 0:
        raise ValueError(f"{what}: expected a split-precision activation (channels-last bf16, 3*C channels, C % 8 == 0), "
                         f"got shape {tuple(x.shape)} dtype {x.dtype}")
    return x.shape[1] // 3


def new_split(n, c, d, h, w, device):
    return ops.new_cl(n, 3 * c, d, h, w, device)


def join(x: torch.Tensor) -> torch.Tensor:
    """fp32 NCDHW value of a split activation (tests / debugging; a full-resolution torch expression)."""
    c = _logical_channels(x, "join")
    return (x[:, :c].float() + x[:, c:2 * c].float()).contiguous()


def _split_apply(y, res, scale, shift, act, slope):
    n, c, d, h, w = y.shape
    z = new_split(n, c, d, h, w, y.device)
    rc = L.load().rb_split_apply(y.data_ptr(), L.ptr(res), z.data_ptr(), L.ptr(scale), L.ptr(shift), n, d * h * w, c,
                                 1 if act else 0, float(slope), L.stream_ptr())
    L.check(rc, "rb_split_apply")
    return z


def _conv_pack(weight, c0, c1):
    co, ci, kd, kh, kw = weight.shape

    def pack():
        w = weight.detach().float().permute(2, 3, 4, 0, 1).reshape(kd * kh * kw, co, ci)
        parts = [_pack3(w[..., :c0])]
        if c1:
            parts.append(_pack3(w[..., c0:]))
        return torch.cat(parts, -1).contiguous()
    return ops._cached_pack(weight, f"x3:{c0}:{c1}", pack)


def conv_norm_act(x, weight, stride=1, x_cat=None, res=None, gamma=None, beta=None, eps=1e-5, act=True,
                  slope=ops.LRELU_SLOPE_DEFAULT, se=None, se_reduce_dims="all", stem=False, impl=None, drop=None):
    """Split-precision twin of `ops.conv_norm_act` (same arguments; activations are split rows)."""
    if drop is not None:
        raise NotImplementedError("the split-precision tier is inference only (stochastic depth is a training op)")
    if se is not None and ops._se_per_w(se_reduce_dims):
        raise NotImplementedError("split-precision tier: SE squeeze over (2, 3) is not implemented (use 'all')")
    impl = IMPL if impl is None else impl
    stride = ops._triple(stride)
    lib = L.load()
    co, ci, kd, kh, kw = weight.shape
    k = (kd, kh, kw)
    if stem:
        L.require_cuda(x, "stem conv")
        xf = x.detach().float().contiguous()
        n, cin, d, h, w = xf.shape
        if cin != ci:
            raise ValueError(f"stem conv: weight expects {ci} input channels, got {cin}")
        K = kd * kh * kw * ci
        kp = (K + 15) // 16 * 16
        src0 = ops.new_cl(n, 3 * kp, d, h, w, xf.device)
        L.check(lib.rb_stem_im2col_split(xf.data_ptr(), src0.data_ptr(), n, ci, d, h, w, kd, kh, kw, kp, L.stream_ptr()),
                "rb_stem_im2col_split")
        src1 = None

        def pack():
            wp = torch.zeros((1, co, kp), dtype=torch.float32, device=weight.device)
            wp[0, :, :K] = weight.detach().float().permute(0, 2, 3, 4, 1).reshape(co, K)
            return _pack3(wp).contiguous()
        wpk = ops._cached_pack(weight, "x3s", pack)
        in_dims, od, taps, off, istr = (d, h, w), (d, h, w), (1, 1, 1), (0, 0, 0), (1, 1, 1)
    else:
        src0 = ops.as_cl(x)
        L.require_cuda(src0, "conv3d")
        c0 = _logical_channels(src0, "conv3d")
        src1 = ops.as_cl(x_cat) if x_cat is not None else None
        c1 = _logical_channels(src1, "conv3d") if src1 is not None else 0
        if c0 + c1 != ci:
            raise ValueError(f"conv3d: weight expects {ci} input channels, got {c0 + c1}")
        if any(kk not in (1, 3) for kk in k):
            raise NotImplementedError(f"conv3d: kernel sizes 1 and 3 are implemented, got {k}")
        if src1 is not None and tuple(src1.shape[2:]) != tuple(src0.shape[2:]):
            raise ValueError("conv3d: concatenated inputs must share their spatial shape")
        n = src0.shape[0]
        in_dims = tuple(src0.shape[2:])
        od = ops._conv_out_dims(in_dims, k, stride)
        wpk = _conv_pack(weight, c0, c1)
        taps, off, istr = k, tuple(-((kk - 1) // 2) for kk in k), stride
    y = ops.new_cl_f32(n, co, *od, src0.device)
    ops._launch_gather(src0, src1, wpk, y, None, in_dims=in_dims, taps=taps, off=off, istr=istr, out_grid=od, nout=co,
                       impl=impl)
    S = od[0] * od[1] * od[2]
    sums = ops._plane_reduce(0, y, None, None, False, slope)           # [N, 1, C, 2] double
    if res is not None:
        res = ops.as_cl(res)
        if _logical_channels(res, "residual") != co or tuple(res.shape[2:]) != tuple(od) or res.shape[0] != n:
            raise ValueError(f"residual shape {tuple(res.shape)} does not match the split output [{n}, 3*{co}, {od}]")
    if se is None:
        small = torch.empty((4, n, co), dtype=torch.float32, device=y.device)
        L.check(lib.rb_in_finalize_fwd(sums.data_ptr(), None, None, L.ptr(gamma), L.ptr(beta), small[0].data_ptr(),
                                       small[1].data_ptr(), small[2].data_ptr(), small[3].data_ptr(), n, co, float(S),
                                       float(eps), L.stream_ptr()), "rb_in_finalize_fwd")
        scale, shift = small[2], small[3]
    else:
        with torch.no_grad():
            A, B = ops._gate_small_graph(sums[:, 0, :, 0], sums[:, 0, :, 1], None, float(S), float(od[0] * od[1]), gamma, beta,
                                         eps, se[0], se[1], se[2], se[3], False)
        scale, shift = A[:, 0].contiguous(), B[:, 0].contiguous()
    return _split_apply(y, res, scale, shift, act, slope)


def conv_transpose3d(x, weight, stride, impl=None, bias=None):
    """Split-precision twin of `ops.conv_transpose3d` (kernel == stride); the bias, if any, rides the split pass."""
    impl = IMPL if impl is None else impl
    stride = ops._triple(stride)
    x = ops.as_cl(x)
    L.require_cuda(x, "conv_transpose3d")
    ci, co, sd, sh, sw = weight.shape
    if (sd, sh, sw) != stride:
        raise NotImplementedError("conv_transpose3d: only kernel_size == stride is implemented")
    if _logical_channels(x, "conv_transpose3d") != ci:
        raise ValueError(f"conv_transpose3d: weight expects {ci} input channels, got {x.shape[1] // 3}")
    n = x.shape[0]
    in_dims = tuple(x.shape[2:])
    full = tuple(i * s for i, s in zip(in_dims, stride))
    npar = sd * sh * sw
    wpk = ops._cached_pack(weight, "x3t", lambda: _pack3(weight.detach().float().permute(2, 3, 4, 1, 0)
                                                         .reshape(1, npar * co, ci)).contiguous())
    y = ops.new_cl_f32(n, co, *full, x.device)
    ops._launch_gather(x, None, wpk, y, None, in_dims=in_dims, taps=(1, 1, 1), off=(0, 0, 0), istr=(1, 1, 1),
                       out_grid=in_dims, nout=npar * co, mode=1, ostr=stride, full=full, ps=stride, psC=co, impl=impl)
    scale = shift = None
    if bias is not None:
        scale = torch.ones((n, co), dtype=torch.float32, device=x.device)
        shift = bias.detach().float().reshape(1, co).expand(n, co).contiguous()
    return _split_apply(y, None, scale, shift, False, 0.0)


def avg_pool3d(x, stride):
    stride = ops._triple(stride)
    x = ops.as_cl(x)
    L.require_cuda(x, "avg_pool3d")
    c = _logical_channels(x, "avg_pool3d")
    n, _, d, h, w = x.shape
    sd, sh, sw = stride
    out = new_split(n, c, d // sd, h // sh, w // sw, x.device)
    L.check(L.load().rb_avgpool_split(x.data_ptr(), out.data_ptr(), n, d, h, w, c, sd, sh, sw, L.stream_ptr()),
            "rb_avgpool_split")
    return out


def head_conv1x1(x, weight, bias, activation=None):
    """1x1x1 head on a split activation: the head kernel multiplies in fp32, so the weight row [w | w | 0] over the
    3C-wide voxel row gives (hi + lo) . w exactly as the bf16 path gives x . w."""
    if weight.shape[0] > 8:
        raise NotImplementedError("task heads with more than 8 output channels are not implemented")
    act = ops._ACT[activation if activation is None else str(activation).lower()]
    x = ops.as_cl(x)
    L.require_cuda(x, "head")
    c = _logical_channels(x, "head")
    n, _, d, h, w = x.shape
    k = weight.shape[0]
    w2 = weight.detach().reshape(k, c).float()
    w3 = torch.cat((w2, w2, torch.zeros_like(w2)), 1).contiguous()
    b = bias.detach().float().contiguous() if bias is not None else None
    out = torch.empty((n, k, d, h, w), dtype=torch.float32, device=x.device)
    L.check(L.load().rb_head_fwd(x.data_ptr(), w3.data_ptr(), L.ptr(b), out.data_ptr(), n, d * h * w, 3 * c, k, act,
                                 L.stream_ptr()), "rb_head_fwd")
    return out
