"""`torch.library` custom-op layer over the C ABI (north star: "a thin C-ABI and torch custom-op layer").

Every fused unit of the network is registered as an operator in the `resenc_b200` namespace with
  * a CUDA implementation (the same primitives the eager autograd.Functions of ops.py launch),
  * a fake (meta) implementation: output shapes / dtypes / the channels-last strides, no data, so that
    `torch.compile(model)` (reference train.py:133, inference.py:37) traces the model and sees these units as
    opaque graph nodes instead of a disabled frame,
  * an autograd rule whose backward is itself a registered operator (`*_bwd`), so AOTAutograd can build the
    backward graph without looking inside,
  * an autocast rule: the operators are dtype-agnostic (bf16 channels-last activations in, fp32 logits out) and take
    their inputs as they come - under `torch.autocast("cuda")` (train.py:203, inference.py:118) nothing is cast.

Operators (schema names):
  resenc_b200::conv_norm_act       act( [SE]( [drop *] IN( conv3d(cat(x, x_cat)) ) ) + res )   (+ ::conv_norm_act_bwd)
  resenc_b200::conv_transpose3d    ConvTranspose3d(kernel == stride), pixel-shuffle store        (+ ::conv_transpose3d_bwd)
  resenc_b200::avg_pool3d          AvgPool3d(stride, stride) of the ResNet-D skip                 (+ ::avg_pool3d_bwd)
  resenc_b200::head_conv1x1        1x1x1 conv + bias (+ eval activation) -> NCDHW fp32            (+ ::head_conv1x1_bwd)

Reference arithmetic replaced: see ops.py.  Eager calls use the autograd.Functions directly (no dispatcher overhead);
`ops.conv_norm_act` & co. route here when `torch.compiler.is_compiling()`.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L
from . import ops

NS = "resenc_b200"


def _empty(dev, dtype=torch.float32):
    return torch.empty(0, dtype=dtype, device=dev)


def _opt(t: Tensor) -> Optional[Tensor]:
    """Operators return zero-element tensors where a gradient does not exist."""
    return None if t is None or t.numel() == 0 else t


def _dims(gate_dims: str):
    return "all" if gate_dims == "all" else tuple(int(v) for v in gate_dims.split(","))


def _dims_str(d) -> str:
    return "all" if d in ("all", (2, 3, 4), [2, 3, 4]) else ",".join(str(int(v)) for v in d)


def _impl(impl: str):
    return None if impl == "" else impl


# ------------------------------------------------------------------------------------------
# conv + InstanceNorm (+ SE gate, + stochastic depth) + residual + LeakyReLU
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::conv_norm_act", mutates_args=(), device_types="cuda")
def _conv_norm_act(x0: Tensor, weight: Tensor, x1: Optional[Tensor], res: Optional[Tensor], gamma: Optional[Tensor],
                   beta: Optional[Tensor], drop: Optional[Tensor], se_w1: Optional[Tensor], se_b1: Optional[Tensor],
                   se_w2: Optional[Tensor], se_b2: Optional[Tensor], stride: List[int], eps: float, act: bool,
                   slope: float, gate_dims: str, stem: bool, impl: str) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Returns (z, y, small, s12, pw): z = the activation; y = the pre-norm conv output (ops.PRENORM_DTYPE); small = [4, N, C]
    (mean, rstd, scale, shift) on the plain path; s12 = double [2, N, C] plane sums and pw = double [N, W, C] on the
    gated path (zero-element tensors where not applicable).  All but z exist for the backward operator."""
    stride_t = tuple(int(s) for s in stride)
    gate = (se_w1, se_b1, se_w2, se_b2) if se_w1 is not None else None
    if stem:
        src0, src1 = ops._stem_im2col(x0, tuple(weight.shape[2:])), None
        y, stats = ops._stem_forward(weight, _impl(impl), src0, out_f32=True, want_stats=True)
    else:
        src0 = ops.as_cl(x0)
        src1 = ops.as_cl(x1) if x1 is not None else None
        L.require_cuda(src0, "conv3d")
        y, stats = ops._conv_forward(weight, stride_t, _impl(impl), src0, src1, out_f32=True, want_stats=True)
    r = ops.as_cl(res) if res is not None else None
    if r is not None and r.shape != y.shape:
        raise ValueError(f"residual shape {tuple(r.shape)} != {tuple(y.shape)}")
    z, st = ops._norm_forward(y, r, gamma, beta, eps, act, slope, stats, gate, _dims(gate_dims), drop)
    dev = y.device
    if st.gate is None:
        return z, y, st.small, _empty(dev, torch.float64), _empty(dev, torch.float64)
    s1, s2, pw = st.sums
    return z, y, _empty(dev), torch.stack((s1, s2)), (pw.contiguous() if pw is not None else _empty(dev, torch.float64))


@_conv_norm_act.register_fake
def _(x0, weight, x1, res, gamma, beta, drop, se_w1, se_b1, se_w2, se_b2, stride, eps, act, slope, gate_dims, stem, impl):
    co = weight.shape[0]
    k = tuple(weight.shape[2:])
    n = x0.shape[0]
    od = tuple(x0.shape[2:]) if stem else ops._conv_out_dims(tuple(x0.shape[2:]), k, tuple(stride))
    z = x0.new_empty((n, *od, co), dtype=torch.bfloat16).permute(0, 4, 1, 2, 3)
    y = x0.new_empty((n, *od, co), dtype=ops.PRENORM_DTYPE).permute(0, 4, 1, 2, 3)
    gated = se_w1 is not None or drop is not None
    if not gated:
        return z, y, x0.new_empty((4, n, co), dtype=torch.float32), x0.new_empty(0, dtype=torch.float64), \
            x0.new_empty(0, dtype=torch.float64)
    per_w = se_w1 is not None and gate_dims != "all"
    pw = x0.new_empty((n, od[2], co), dtype=torch.float64) if per_w else x0.new_empty(0, dtype=torch.float64)
    return z, y, x0.new_empty(0, dtype=torch.float32), x0.new_empty((2, n, co), dtype=torch.float64), pw


@torch.library.custom_op(f"{NS}::conv_norm_act_bwd", mutates_args=(), device_types="cuda")
def _conv_norm_act_bwd(dz: Tensor, z: Tensor, y: Tensor, small: Tensor, s12: Tensor, pw: Tensor, x0: Tensor, weight: Tensor,
                       x1: Optional[Tensor], has_res: bool, gamma: Optional[Tensor], beta: Optional[Tensor],
                       drop: Optional[Tensor], se_w1: Optional[Tensor], se_b1: Optional[Tensor], se_w2: Optional[Tensor],
                       se_b2: Optional[Tensor], stride: List[int], eps: float, act: bool, slope: float, gate_dims: str,
                       stem: bool, impl: str, need: List[bool]) -> List[Tensor]:
    """need = [weight, x0, x1, res].  Returns [gw, gx0, gx1, dres, dgamma, dbeta, g_se_w1, g_se_b1, g_se_w2, g_se_b2]
    with zero-element tensors for gradients that do not exist."""
    stride_t = tuple(int(s) for s in stride)
    gate = (se_w1, se_b1, se_w2, se_b2) if se_w1 is not None else None
    n, c, d, h, w = y.shape
    st = ops._NormState()
    st.gamma, st.has_beta, st.act, st.slope, st.eps, st.has_res = gamma, beta is not None, act, slope, eps, has_res
    st.n_gate = 0 if gate is None else 4
    st.sums = None
    if small.numel():
        st.gate, st.per_w, st.small = None, False, small
    else:
        st.small = None
        st.per_w = ops._se_per_w(_dims(gate_dims)) if gate is not None else False
        ops._rebuild_gate_state(st, s12[0], s12[1], pw if pw.numel() else None, d * h * w, d * h, gamma, beta, gate, drop)
    dy, dres, dgamma, dbeta, ggate = ops._norm_backward(st, y, z if act else None, ops.as_cl(dz), has_res and need[3])
    if stem:
        col = ops._stem_im2col(x0, tuple(weight.shape[2:]))          # recomputed (0.1 ms) instead of carried as an op output
        gw = ops._stem_backward(tuple(weight.shape), col, dy) if need[0] else None
        gx0 = gx1 = None
    else:
        src0 = ops.as_cl(x0)
        src1 = ops.as_cl(x1) if x1 is not None else None
        gw, gx0, gx1 = ops._conv_backward(weight, stride_t, _impl(impl), src0, src1, dy, need[0], need[1], need[2])
    dev = y.device
    out = [gw, gx0, gx1, dres, dgamma, dbeta, *(ggate if ggate is not None else (None,) * 4)]
    return [t if t is not None else _empty(dev) for t in out]


@_conv_norm_act_bwd.register_fake
def _(dz, z, y, small, s12, pw, x0, weight, x1, has_res, gamma, beta, drop, se_w1, se_b1, se_w2, se_b2, stride, eps, act,
      slope, gate_dims, stem, impl, need):
    def like(t, cond=True):
        return torch.empty_like(t) if (t is not None and cond) else dz.new_empty(0, dtype=torch.float32)
    return [like(weight, need[0]), like(x0, need[1] and not stem), like(x1, need[2] and not stem), like(dz, has_res and need[3]),
            like(gamma), like(beta), like(se_w1), like(se_b1), like(se_w2), like(se_b2)]


def _cna_setup(ctx, inputs, output):
    (x0, weight, x1, res, gamma, beta, drop, se_w1, se_b1, se_w2, se_b2, stride, eps, act, slope, gate_dims, stem, impl) = inputs
    z, y, small, s12, pw = output
    ctx.save_for_backward(z, y, small, s12, pw, x0, weight, x1, gamma, beta, drop, se_w1, se_b1, se_w2, se_b2)
    ctx.has_res = res is not None
    ctx.cfg = (list(stride), eps, act, slope, gate_dims, stem, impl)
    ctx.mark_non_differentiable(y, small, s12, pw)
    ctx.set_materialize_grads(False)


def _cna_backward(ctx, gz, gy, gsmall, gs12, gpw):
    z, y, small, s12, pw, x0, weight, x1, gamma, beta, drop, se_w1, se_b1, se_w2, se_b2 = ctx.saved_tensors
    stride, eps, act, slope, gate_dims, stem, impl = ctx.cfg
    ni = ctx.needs_input_grad
    need = [bool(ni[1]), bool(ni[0]) and not stem, bool(ni[2]) and x1 is not None, bool(ni[3]) and ctx.has_res]
    g = torch.ops.resenc_b200.conv_norm_act_bwd(gz, z, y, small, s12, pw, x0, weight, x1, ctx.has_res, gamma, beta, drop,
                                                se_w1, se_b1, se_w2, se_b2, stride, eps, act, slope, gate_dims, stem, impl, need)
    gw, gx0, gx1, dres, dgamma, dbeta, g1, g2, g3, g4 = (_opt(t) for t in g)
    #       x0   weight x1  res   gamma   beta  drop  se_w1 se_b1 se_w2 se_b2  (non-tensor arguments)
    return (gx0, gw, gx1, dres, dgamma, dbeta, None, g1, g2, g3, g4, None, None, None, None, None, None, None)


torch.library.register_autograd(f"{NS}::conv_norm_act", _cna_backward, setup_context=_cna_setup)


def conv_norm_act(x, weight, stride, x_cat, res, gamma, beta, eps, act, slope, se, se_reduce_dims, stem, impl, drop):
    se = se or (None, None, None, None)
    out = torch.ops.resenc_b200.conv_norm_act(x, weight, x_cat, res, gamma, beta, drop, se[0], se[1], se[2], se[3],
                                              list(stride), eps, act, slope, _dims_str(se_reduce_dims), stem, impl or "")
    return out[0]


# ------------------------------------------------------------------------------------------
# ConvTranspose3d (kernel == stride)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::conv_transpose3d", mutates_args=(), device_types="cuda")
def _conv_transpose3d(x: Tensor, weight: Tensor, stride: List[int], impl: str) -> Tensor:
    with torch.no_grad():
        return ops._ConvT3dFn.apply(weight, tuple(int(s) for s in stride), _impl(impl), x)


@_conv_transpose3d.register_fake
def _(x, weight, stride, impl):
    n = x.shape[0]
    full = [int(i) * int(s) for i, s in zip(x.shape[2:], stride)]
    return x.new_empty((n, *full, weight.shape[1]), dtype=torch.bfloat16).permute(0, 4, 1, 2, 3)


@torch.library.custom_op(f"{NS}::conv_transpose3d_bwd", mutates_args=(), device_types="cuda")
def _conv_transpose3d_bwd(dy: Tensor, x: Tensor, weight: Tensor, stride: List[int], impl: str, need: List[bool]) -> List[Tensor]:
    """need = [weight, x]; returns [gw, gx]."""
    class _Ctx:
        pass
    ctx = _Ctx()
    ctx.saved_tensors = (weight, ops.as_cl(x))
    ctx.stride, ctx.impl = tuple(int(s) for s in stride), _impl(impl)
    ctx.needs_input_grad = (need[0], False, False, need[1])
    gw, _, _, gx = ops._ConvT3dFn.backward(ctx, dy)
    return [t if t is not None else _empty(dy.device) for t in (gw, gx)]


@_conv_transpose3d_bwd.register_fake
def _(dy, x, weight, stride, impl, need):
    e = dy.new_empty(0, dtype=torch.float32)
    gx = x.new_empty((x.shape[0], *x.shape[2:], x.shape[1]), dtype=torch.bfloat16).permute(0, 4, 1, 2, 3) if need[1] else e
    return [torch.empty_like(weight) if need[0] else e, gx]


def _ct_setup(ctx, inputs, output):
    x, weight, stride, impl = inputs
    ctx.save_for_backward(x, weight)
    ctx.cfg = (list(stride), impl)


def _ct_backward(ctx, gy):
    x, weight = ctx.saved_tensors
    stride, impl = ctx.cfg
    need = [bool(ctx.needs_input_grad[1]), bool(ctx.needs_input_grad[0])]
    gw, gx = (_opt(t) for t in torch.ops.resenc_b200.conv_transpose3d_bwd(gy, x, weight, stride, impl, need))
    return gx, gw, None, None


torch.library.register_autograd(f"{NS}::conv_transpose3d", _ct_backward, setup_context=_ct_setup)


def conv_transpose3d(x, weight, stride, impl):
    return torch.ops.resenc_b200.conv_transpose3d(x, weight, list(stride), impl or "")


# ------------------------------------------------------------------------------------------
# AvgPool3d(kernel == stride)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::avg_pool3d", mutates_args=(), device_types="cuda")
def _avg_pool3d(x: Tensor, stride: List[int]) -> Tensor:
    with torch.no_grad():
        return ops._AvgPoolFn.apply(x, tuple(int(s) for s in stride))


@_avg_pool3d.register_fake
def _(x, stride):
    n, c, d, h, w = x.shape
    return x.new_empty((n, d // stride[0], h // stride[1], w // stride[2], c), dtype=torch.bfloat16).permute(0, 4, 1, 2, 3)


@torch.library.custom_op(f"{NS}::avg_pool3d_bwd", mutates_args=(), device_types="cuda")
def _avg_pool3d_bwd(dout: Tensor, shape: List[int], stride: List[int]) -> Tensor:
    class _Ctx:
        pass
    ctx = _Ctx()
    ctx.shape, ctx.stride = tuple(int(v) for v in shape), tuple(int(s) for s in stride)
    return ops._AvgPoolFn.backward(ctx, dout)[0]


@_avg_pool3d_bwd.register_fake
def _(dout, shape, stride):
    n, c, d, h, w = shape
    return dout.new_empty((n, d, h, w, c), dtype=torch.bfloat16).permute(0, 4, 1, 2, 3)


def _ap_setup(ctx, inputs, output):
    x, stride = inputs
    ctx.shape, ctx.stride = [int(v) for v in x.shape], list(stride)


def _ap_backward(ctx, g):
    return torch.ops.resenc_b200.avg_pool3d_bwd(g, ctx.shape, ctx.stride), None


torch.library.register_autograd(f"{NS}::avg_pool3d", _ap_backward, setup_context=_ap_setup)


def avg_pool3d(x, stride):
    return torch.ops.resenc_b200.avg_pool3d(x, list(stride))


# ------------------------------------------------------------------------------------------
# task head
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::head_conv1x1", mutates_args=(), device_types="cuda")
def _head(x: Tensor, weight: Tensor, bias: Optional[Tensor], act: int) -> Tensor:
    with torch.no_grad():
        return ops._HeadFn.apply(x, weight, bias, act)


@_head.register_fake
def _(x, weight, bias, act):
    n, c, d, h, w = x.shape
    return x.new_empty((n, weight.shape[0], d, h, w), dtype=torch.float32)


@torch.library.custom_op(f"{NS}::head_conv1x1_bwd", mutates_args=(), device_types="cuda")
def _head_bwd(dl: Tensor, x: Tensor, weight: Tensor, has_bias: bool) -> List[Tensor]:
    """Returns [dx, dw, db] (db zero-element without a bias)."""
    class _Ctx:
        pass
    ctx = _Ctx()
    xc = ops.as_cl(x)
    k, c = weight.shape[0], xc.shape[1]
    ctx.saved_tensors = (xc, weight.detach().reshape(k, c).float().contiguous())
    ctx.act, ctx.has_bias, ctx.wshape = 0, has_bias, tuple(weight.shape)
    dx, dw, db, _ = ops._HeadFn.backward(ctx, dl)
    return [dx, dw, db if db is not None else _empty(dl.device)]


@_head_bwd.register_fake
def _(dl, x, weight, has_bias):
    n, c, d, h, w = x.shape
    dx = x.new_empty((n, d, h, w, c), dtype=torch.bfloat16).permute(0, 4, 1, 2, 3)
    return [dx, torch.empty_like(weight, dtype=torch.float32),
            dl.new_empty(weight.shape[0] if has_bias else 0, dtype=torch.float32)]


def _hd_setup(ctx, inputs, output):
    x, weight, bias, act = inputs
    if act != 0:
        ctx.fused_act = True
        return
    ctx.fused_act = False
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None


def _hd_backward(ctx, g):
    if ctx.fused_act:
        raise NotImplementedError("backward through the fused eval-mode activation is not implemented "
                                  "(the reference applies it only when not self.training)")
    x, weight = ctx.saved_tensors
    dx, dw, db = torch.ops.resenc_b200.head_conv1x1_bwd(g, x, weight, ctx.has_bias)
    return dx, dw, _opt(db), None


torch.library.register_autograd(f"{NS}::head_conv1x1", _hd_backward, setup_context=_hd_setup)


def head_conv1x1(x, weight, bias, act):
    return torch.ops.resenc_b200.head_conv1x1(x, weight, bias, int(act))


# ------------------------------------------------------------------------------------------
# autocast: run as called, whatever autocast state the caller is in (nothing is cast to the autocast dtype)
# ------------------------------------------------------------------------------------------
def _passthrough(name):
    op = getattr(torch.ops.resenc_b200, name).default

    def impl(*args, **kwargs):
        with torch._C._ExcludeDispatchKeyGuard(torch._C.DispatchKeySet(torch._C.DispatchKey.AutocastCUDA)):
            return op(*args, **kwargs)
    return impl


_AUTOCAST_LIB = torch.library.Library(NS, "IMPL")
for _name in ("conv_norm_act", "conv_norm_act_bwd", "conv_transpose3d", "conv_transpose3d_bwd", "avg_pool3d", "avg_pool3d_bwd",
              "head_conv1x1", "head_conv1x1_bwd"):
    _AUTOCAST_LIB.impl(_name, _passthrough(_name), "AutocastCUDA")

OPERATORS = ("conv_norm_act", "conv_norm_act_bwd", "conv_transpose3d", "conv_transpose3d_bwd", "avg_pool3d", "avg_pool3d_bwd",
             "head_conv1x1", "head_conv1x1_bwd")
