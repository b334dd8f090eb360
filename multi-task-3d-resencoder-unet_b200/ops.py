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
import os

import torch
from torch.optim.optimizer import register_optimizer_step_post_hook as _register_step_hook

from . import _lib as L

BF16 = torch.bfloat16
LRELU_SLOPE_DEFAULT = 0.01
# The fused units are registered twice: as torch.autograd.Functions (eager: no dispatcher overhead on ~700 launches per
# step) and as `torch.library` custom ops (custom_ops.py: fake + autograd + autocast rules) which Dynamo sees as opaque
# graph nodes under torch.compile (reference train.py:133, inference.py:37).  RESENC_FORCE_CUSTOM_OPS=1 sends eager
# calls through the custom ops as well (tests).
FORCE_CUSTOM_OPS = os.environ.get("RESENC_FORCE_CUSTOM_OPS") is not None


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
    def span(self, kind, flops=0.0, nbytes=0.0, moved=0.0):
        """flops / nbytes: ALGORITHMIC work of the launch; moved: bytes this implementation touches."""
        if not self.on:
            yield
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        self.records.append((kind, flops, a, b, nbytes, moved))

    def summary(self):
        if not self.records:
            return {}
        torch.cuda.synchronize()
        out = {}
        for kind, flops, a, b, nbytes, moved in self.records:
            d = out.setdefault(kind, {"ms": 0.0, "flops": 0.0, "launches": 0, "bytes": 0.0, "bytes_moved": 0.0})
            d["ms"] += a.elapsed_time(b)
            d["flops"] += flops
            d["bytes"] += nbytes
            d["bytes_moved"] += moved
            d["launches"] += 1
        return out


KERNEL_TIMER = _KernelTimer()


# ------------------------------------------------------------------------------------------
# precision tier switch: bf16 operands / fp32 accumulation (default, training and inference) or the
# split-precision "bf16x3" inference tier of precise.py (fp32-accurate results on the same kernels)
# ------------------------------------------------------------------------------------------
class _PreciseState:
    on = False
    impl = None


@contextlib.contextmanager
def precise_inference(impl=None):
    """Inside this context the network's operators (conv_norm_act, conv_transpose3d, avg_pool3d, head_conv1x1) run
    the split-precision tier: outputs within ~1e-5 relative L2 of an fp32 evaluation instead of ~1e-2.  Inference
    only: autograd is disabled, stochastic depth must be off (eval mode).  `impl` ("mma" | "auto") selects the
    gather-conv kernel family for the 3C-wide contractions (default: precise.IMPL)."""
    prev = (_PreciseState.on, _PreciseState.impl)
    _PreciseState.on, _PreciseState.impl = True, impl
    try:
        with torch.no_grad():
            yield
    finally:
        _PreciseState.on, _PreciseState.impl = prev


def precise_active() -> bool:
    return _PreciseState.on


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
    if torch.compiler.is_compiling():
        # traced by Dynamo (module-level glue such as the residual decoder's concatenation): torch ops only
        return x.to(BF16).contiguous(memory_format=torch.channels_last_3d)
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
def new_cl_f32(n, c, d, h, w, device):
    """Pre-norm conv output: logical [n, c, d, h, w], memory NDHWC fp32 (keeps the accumulator precision)."""
    return torch.empty((n, d, h, w, c), dtype=torch.float32, device=device).permute(0, 4, 1, 2, 3)


def is_cl_f32(x: torch.Tensor) -> bool:
    return x.dim() == 5 and x.dtype == torch.float32 and x.permute(0, 2, 3, 4, 1).is_contiguous()


# Element type of the pre-norm conv output inside the fused conv + norm + act unit.  The InstanceNorm statistics always
# come from the fp32 accumulators (conv epilogue / finish pass); only the tensor the normalise pass and the backward pass
# re-read is stored.  fp16 (default): 11-bit significand - 8x finer than the bf16 rounding every stored activation gets
# anyway, so the outputs move by ~1 % of the bf16 tier's own error - at half the bytes of fp32 on five passes per layer
# (conv store, normalise read, two backward reads; saturating stores: |y| is clamped to 65504 instead of overflowing).
# RESENC_PRENORM=f32 restores the round-1 layout.
PRENORM_DTYPE = {"f32": torch.float32, "f16": torch.float16}[os.environ.get("RESENC_PRENORM", "f16")]
_YMODE = {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2}


def new_prenorm(n, c, d, h, w, device):
    """Pre-norm conv output: logical [n, c, d, h, w], memory NDHWC in PRENORM_DTYPE."""
    return torch.empty((n, d, h, w, c), dtype=PRENORM_DTYPE, device=device).permute(0, 4, 1, 2, 3)


def is_prenorm(x: torch.Tensor) -> bool:
    return x.dim() == 5 and x.dtype in (torch.float32, torch.float16) and x.permute(0, 2, 3, 4, 1).is_contiguous()


def as_prenorm(y: torch.Tensor) -> torch.Tensor:
    """Pre-norm tensors are accepted as NDHWC fp32 (kept) or anything `as_cl` can turn into NDHWC bf16."""
    return y if is_prenorm(y) and y.shape[1] % 8 == 0 and y.is_cuda else as_cl(y)


def _make_desc(src0, src1, out0, out1, *, in_dims, taps, off, istr, out_grid, nout, mode, ostr, ooff, full, ps, psC, impl):
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
    d.outF32 = _YMODE[out0.dtype]
    return d


_DESC_CACHE = {}


def _launch_gather(src0, src1, wpk, out0, out1, *, in_dims, taps, off, istr, out_grid, nout, mode=0,
                   ostr=(1, 1, 1), ooff=(0, 0, 0), full=None, ps=None, psC=0, impl=None, stats=None, want_stats=False,
                   algo_flops=None):
    """One rb_conv_gather call.  src* are NDHWC bf16 activations, out* NDHWC bf16 or fp32 (logical NCDHW views).
    stats=(sum, sumsq) fp32 [NB, Nout] are filled by the tcgen05 epilogue; with want_stats=True they are
    allocated here when (and only when) the library will run the tcgen05 kernel, and returned (else None)."""
    lib = L.load()
    # descriptor, kernel choice and workspace size depend on shapes only: built once per distinct launch geometry (the
    # ~40 ctypes field writes + two library queries cost ~20 us per launch in the eager path otherwise)
    key = (tuple(src0.shape), None if src1 is None else tuple(src1.shape), tuple(out0.shape), out0.dtype,
           None if out1 is None else tuple(out1.shape), tuple(in_dims), tuple(taps), tuple(off), tuple(istr), tuple(out_grid), nout,
           mode, tuple(ostr), tuple(ooff), None if full is None else tuple(full), None if ps is None else tuple(ps), psC, impl,
           L.default_impl())
    hit = _DESC_CACHE.get(key)
    if hit is None:
        d = _make_desc(src0, src1, out0, out1, in_dims=in_dims, taps=taps, off=off, istr=istr, out_grid=out_grid, nout=nout,
                       mode=mode, ostr=ostr, ooff=ooff, full=full, ps=ps, psC=psC, impl=impl)
        plan = lib.rb_conv_gather_plan(C.byref(d))
        ws_bytes = lib.rb_conv_gather_workspace(C.byref(d)) if plan >= 0 else 0
        if len(_DESC_CACHE) > 4096:
            _DESC_CACHE.clear()
        hit = _DESC_CACHE[key] = (d, plan, ws_bytes)
    d, plan, ws_bytes = hit
    if want_stats and stats is None:
        if plan < 0:
            L.check(plan, "rb_conv_gather_plan")
        if plan in (L.IMPL_TCGEN05, L.IMPL_TCGEN05_SLAB, L.IMPL_TCGEN05_SPLITK):
            st = torch.zeros((2, d.NB, nout), dtype=torch.float32, device=src0.device)
            stats = (st[0], st[1])
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=src0.device) if ws_bytes else None
    ssum, ssq = (None, None) if stats is None else stats
    # `algo_flops`: callers whose operand carries structural zeros (the merged strided data gradient) pass the
    # algorithmic count instead of the padded GEMM size
    flops = algo_flops if algo_flops is not None else \
        2.0 * d.NB * d.OD * d.OH * d.OW * nout * (d.srcC0 + d.srcC1) * taps[0] * taps[1] * taps[2]
    with KERNEL_TIMER.span("conv", flops):
        rc = lib.rb_conv_gather(C.byref(d), src0.data_ptr(), L.ptr(src1), wpk.data_ptr(), out0.data_ptr(), L.ptr(out1),
                                L.ptr(ssum), L.ptr(ssq), L.ptr(ws), ws_bytes, L.stream_ptr())
    L.check(rc, "rb_conv_gather")
    return stats


def _launch_wgrad(P, Q0, Q1, *, grid, qdims, taps, off, istr, impl=None):
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
    d.impl = L.default_impl() if impl is None else L.impl_code(impl)
    ntaps = taps[0] * taps[1] * taps[2]
    dw = torch.empty((ntaps, d.PC, d.QC0 + d.QC1), dtype=torch.float32, device=P.device)   # initialised by the library
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
PACK_CACHE = True     # set False while capturing a CUDA graph so the (re)packing kernels are part of every replay
# strided data gradients as one pixel-shuffle launch (False: one launch per parity class, the cross-check path)
MERGED_STRIDED_DGRAD = os.environ.get("RESENC_NO_MERGED_DGRAD") is None
# InstanceNorm+LeakyReLU backward without re-reading the stored activation: layers without residual / gate recompute
# lrelu'(z) from the fp32 pre-norm tensor and the forward's folded scale / shift (2 of 8-10 bytes per element less in
# both backward passes; 32.5 -> 32.1 ms per step).  RESENC_NO_SIGN_FROM_PRENORM=1 restores the read of z.
SIGN_FROM_PRENORM = os.environ.get("RESENC_NO_SIGN_FROM_PRENORM") is None
# weight gradient of the deep (small-grid) layers on a side stream, concurrent with the data gradient
CONCURRENT_WGRAD = os.environ.get("RESENC_NO_CONCURRENT_WGRAD") is None
# same fork for the transposed convolutions: measured without gain (29.9 against 29.7 ms), opt-in
CONCURRENT_CONVT_WGRAD = CONCURRENT_WGRAD and os.environ.get("RESENC_CONCURRENT_CONVT_WGRAD") is not None
CONCURRENT_WGRAD_MAX_VOXELS = int(os.environ.get("RESENC_CONCURRENT_WGRAD_MAX_VOXELS", 1 << 40))


# Fused optimisers (torch.optim.AdamW(fused=True), `torch._fused_adamw_`) update parameters in place WITHOUT bumping
# `Tensor._version`, so the version alone cannot key the pack cache: every optimiser step anywhere in the process
# advances this epoch (global post-step hook) and invalidates all cached packs.
_PACK_EPOCH = [0]


def _on_optimizer_step(optimizer, args, kwargs):
    _PACK_EPOCH[0] += 1


_register_step_hook(_on_optimizer_step)


def invalidate_weight_packs():
    """Call after modifying parameters through an API that neither bumps `Tensor._version` nor is an optimiser step."""
    _PACK_EPOCH[0] += 1


def _cached_pack(weight, kind, fn):
    if not PACK_CACHE:
        with torch.no_grad():
            return fn()
    key = (weight._version, _PACK_EPOCH[0], weight.data_ptr(), str(weight.device))
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


def _pack_kernel(weight, want_f, want_d):
    """rb_pack_conv_weights: canonical fp32 [Cout, Cin, kd, kh, kw] -> bf16 [taps][Cout][Cin] (fprop operand) and /
    or [taps, flipped][Cin][Cout] (stride-1 data-gradient operand), one coalesced pass."""
    co, ci, kd, kh, kw = weight.shape
    T = kd * kh * kw
    w = weight.detach()
    if not w.is_contiguous():
        w = w.contiguous()
    f = torch.empty((T, co, ci), dtype=BF16, device=w.device) if want_f else None
    d = torch.empty((T, ci, co), dtype=BF16, device=w.device) if want_d else None
    L.check(L.load().rb_pack_conv_weights(w.data_ptr(), L.ptr(f), L.ptr(d), co, ci, T, L.stream_ptr()),
            "rb_pack_conv_weights")
    return f, d


# PACK_CACHE off (a CUDA graph of a training step is being captured or replayed eagerly): packs are not kept across
# optimiser steps, but within one step the fprop and the stride-1 data-gradient operand of a weight still come from ONE
# launch (one read of the fp32 parameter instead of two): forward packs both when the weight will need a gradient, the
# backward pass picks the second one up.  Entries die with the optimiser epoch.
_STEP_PACKS = {"epoch": -1, "packs": {}}


def _step_packs():
    if _STEP_PACKS["epoch"] != _PACK_EPOCH[0]:
        _STEP_PACKS["epoch"] = _PACK_EPOCH[0]
        _STEP_PACKS["packs"] = {}
    return _STEP_PACKS["packs"]


# Optimiser-managed packs (optim.ClippedAdamW(manage_packs=True)): the update kernel of a conv weight writes the bf16
# operands of the NEXT step itself (rb_adamw_clip_pack_step), into buffers that live as long as the parameter
# (`weight._rb_opt_packs`), so neither an eager step nor a captured one launches the pack kernel for that weight.  The
# entry is trusted only while nothing else can have touched the parameter: same storage, same `Tensor._version`, and
# the optimiser epoch it was written for (any other optimiser step, `invalidate_weight_packs()`, an in-place edit or a
# `load_state_dict` makes it stale and the pack kernel runs again).  CUDA-graph replays run no Python: whoever modifies
# parameters between replays behind the optimiser's back (the trainer restoring its snapshot after a capture) calls
# `ClippedAdamW.refresh_packs()`.
def opt_packs(weight):
    ent = getattr(weight, "_rb_opt_packs", None)
    if ent is None:
        return None
    if ent["epoch"] != _PACK_EPOCH[0] or ent["version"] != weight._version or ent["ptr"] != weight.data_ptr():
        return None
    return ent


def opt_packs_store(weight, f, d, epoch_after_step):
    weight._rb_opt_packs = {"f": f, "d": d, "epoch": epoch_after_step, "version": weight._version, "ptr": weight.data_ptr()}


def pack_conv_fprop(weight):
    """[Cout, Cin, kd, kh, kw] -> [taps][Cout][Cin] bf16."""
    if not weight.is_cuda or weight.dtype != torch.float32:
        co, ci, kd, kh, kw = weight.shape
        return _cached_pack(weight, "f", lambda: weight.detach().permute(2, 3, 4, 0, 1).reshape(kd * kh * kw, co, ci)
                            .to(BF16).contiguous())
    weight._rb_wants_fd = True           # tells the optimiser which parameters are consumed as packed conv operands
    ent = opt_packs(weight)
    if ent is not None:
        return ent["f"]
    if not PACK_CACHE:
        want_d = weight.requires_grad      # (grad mode is off inside autograd.Function.forward: not a usable signal)
        if not want_d:
            return _pack_kernel(weight, True, False)[0]
        f, d = _pack_kernel(weight, True, True)
        _step_packs()[id(weight)] = (weight, d)
        return f
    both = _cached_pack(weight, "fd", lambda: _pack_kernel(weight, True, True))
    return both[0]


def pack_conv_dgrad_full(weight):
    """Stride-1 data-gradient operand: all taps, flipped, [taps][Cin][Cout] bf16."""
    ent = opt_packs(weight) if weight.is_cuda else None
    if ent is not None:
        return ent["d"]
    if not PACK_CACHE:
        hit = _step_packs().pop(id(weight), None)
        if hit is not None and hit[0] is weight:
            return hit[1]
        return _pack_kernel(weight, False, True)[1]
    return _cached_pack(weight, "fd", lambda: _pack_kernel(weight, True, True))[1]


def unpack_wgrad(dw, a, b, kshape):
    """[taps][A][B] fp32 (kernel result) -> canonical [A, B, kd, kh, kw] fp32."""
    T = kshape[0] * kshape[1] * kshape[2]
    out = torch.empty((a, b, *kshape), dtype=torch.float32, device=dw.device)
    L.check(L.load().rb_unpack_wgrad(dw.data_ptr(), out.data_ptr(), a, b, T, L.stream_ptr()), "rb_unpack_wgrad")
    return out


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


def _flipped(weight):
    """w[..., ::-1, ::-1, ::-1]: every data-gradient class is a strided slice of this one tensor."""
    return _cached_pack(weight, "flip", lambda: weight.detach().flip(2, 3, 4))


def pack_conv_dgrad_class(weight, kds, khs, kws, strides=None):
    """taps selected by kernel-index lists per axis -> [taps][Cin][Cout] bf16.  The lists produced by
    `_axis_classes` are descending arithmetic progressions, i.e. strided slices of the flipped kernel."""
    wf = _flipped(weight)
    K = weight.shape[2:]

    def sl(ks, k):
        if len(ks) == 1:
            a = k - 1 - ks[0]
            return slice(a, a + 1)
        step = ks[0] - ks[1]
        return slice(k - 1 - ks[0], None, step)
    w = wf[:, :, sl(kds, K[0]), sl(khs, K[1]), sl(kws, K[2])]
    assert w.shape[2:] == (len(kds), len(khs), len(kws))
    co, ci = w.shape[:2]
    return w.permute(2, 3, 4, 1, 0).reshape(len(kds) * len(khs) * len(kws), ci, co).to(BF16).contiguous()


def _merged_dgrad_plan(k, stride, pad, in_dims, od):
    """Strided data gradient as ONE pixel-shuffle gather conv over dy (instead of one launch per parity class):
    every input voxel i = s*j + r of an axis receives from outputs j + off .. through a class-specific tap subset, so
    all prod(s) classes of a dy neighbourhood are the N columns [(rd, rh, rw), ci] of one GEMM whose weight blocks
    are zero where a (class, window tap) pair has no kernel index.  Returns per-axis (ntap, off, [(r, u0, ks)]) or
    None when the geometry does not tile exactly."""
    if all(s == 1 for s in stride) or any(s > 2 for s in stride):
        return None
    axes = []
    for a in range(3):
        cls = _axis_classes(k[a], stride[a], pad[a], in_dims[a])
        if len(cls) != stride[a] or any(c[1] != od[a] for c in cls):
            return None
        live = [c for c in cls if c[2]]
        if not live:
            return None
        lo = min(c[3] for c in live)
        hi = max(c[3] + len(c[2]) - 1 for c in live)
        axes.append((hi - lo + 1, lo, [(c[0], c[3] - lo, c[2]) for c in cls]))
    return axes


DMERGE_KERNEL = os.environ.get("RESENC_DMERGE_KERNEL", "1") != "0"


def _pack_dgrad_merged_kernel(weight, axes, stride):
    """rb_pack_conv_dgrad_merged: the tensor `pack_conv_dgrad_merged` describes in ONE launch (the torch-op form below
    costs a flip, a zero fill and a slice copy + cast per parity class: ~17 launches per strided conv and step)."""
    import ctypes as C
    co, ci, k0, k1, k2 = weight.shape
    nt = [a[0] for a in axes]
    kidx = (C.c_byte * 24)(*([-1] * 24))
    for a in range(3):
        for r, u0, ks in axes[a][2]:
            for j, k in enumerate(ks):
                kidx[(a * 2 + r) * 4 + u0 + j] = k
    out = torch.empty((nt[0] * nt[1] * nt[2], stride[0] * stride[1] * stride[2] * ci, co), dtype=BF16, device=weight.device)
    w = weight.detach()
    L.check(L.load().rb_pack_conv_dgrad_merged(w.data_ptr(), out.data_ptr(), co, ci, k0, k1, k2, (C.c_int * 3)(*nt),
                                                (C.c_int * 3)(*stride), kidx, L.stream_ptr()), "rb_pack_conv_dgrad_merged")
    return out


def pack_conv_dgrad_merged(weight, axes, stride, force_torch=False):
    """[window taps][(rd, rh, rw, ci)][co] bf16 for `_merged_dgrad_plan` (zero blocks included)."""
    def pack():
        co, ci = weight.shape[:2]
        if (DMERGE_KERNEL and not force_torch and weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous()
                and co % 2 == 0 and all(a[0] <= 4 for a in axes) and all(1 <= s <= 2 for s in stride)):
            return _pack_dgrad_merged_kernel(weight, axes, stride)
        nt = [a[0] for a in axes]
        wp = torch.zeros((nt[0], nt[1], nt[2], stride[0], stride[1], stride[2], ci, co), dtype=BF16, device=weight.device)
        wf = _flipped(weight)
        K = weight.shape[2:]

        def sl(ks, kk):   # descending kernel-index progression -> slice of the flipped kernel (no index tensors:
            a = kk - 1 - ks[0]   # nothing here may copy from the host, this runs inside CUDA graph capture)
            return slice(a, a + 1) if len(ks) == 1 else slice(a, a + (len(ks) - 1) * (ks[0] - ks[1]) + 1, ks[0] - ks[1])
        for rd, ud, kd in axes[0][2]:
            for rh, uh, kh in axes[1][2]:
                for rw, uw, kw in axes[2][2]:
                    if not (kd and kh and kw):
                        continue
                    blk = wf[:, :, sl(kd, K[0]), sl(kh, K[1]), sl(kw, K[2])]   # [co, ci, |kd|, |kh|, |kw|]
                    wp[ud:ud + len(kd), uh:uh + len(kh), uw:uw + len(kw), rd, rh, rw] = blk.permute(2, 3, 4, 1, 0).to(BF16)
        return wp.reshape(nt[0] * nt[1] * nt[2], stride[0] * stride[1] * stride[2] * ci, co)
    if force_torch:
        with torch.no_grad():
            return pack()
    return _cached_pack(weight, "dmerge", pack)



# ------------------------------------------------------------------------------------------
# primitives (no autograd): convolution forward / backward
# ------------------------------------------------------------------------------------------
def _conv_out_dims(in_dims, k, s):
    return tuple((i + 2 * ((kk - 1) // 2) - kk) // ss + 1 for i, kk, ss in zip(in_dims, k, s))


def _check_conv_args(weight, x0, x1):
    co, ci, kd, kh, kw = weight.shape
    c_in = x0.shape[1] + (x1.shape[1] if x1 is not None else 0)
    if c_in != ci:
        raise ValueError(f"conv3d: weight expects {ci} input channels, got {c_in}")
    if any(kk not in (1, 3) for kk in (kd, kh, kw)):
        raise NotImplementedError(f"conv3d: kernel sizes 1 and 3 are implemented, got {(kd, kh, kw)}")
    if x1 is not None and tuple(x1.shape[2:]) != tuple(x0.shape[2:]):
        raise ValueError("conv3d: concatenated inputs must share their spatial shape")


def _conv_forward(weight, stride, impl, x0, x1, out_f32=False, want_stats=False):
    _check_conv_args(weight, x0, x1)
    co, ci, kd, kh, kw = weight.shape
    k = (kd, kh, kw)
    n = x0.shape[0]
    in_dims = tuple(x0.shape[2:])
    od = _conv_out_dims(in_dims, k, stride)
    y = (new_prenorm if out_f32 else new_cl)(n, co, *od, x0.device)
    stats = _launch_gather(x0, x1, pack_conv_fprop(weight), y, None, in_dims=in_dims, taps=k,
                           off=tuple(-((kk - 1) // 2) for kk in k), istr=stride, out_grid=od, nout=co, impl=impl,
                           want_stats=want_stats)
    return y, stats


def _conv_backward(weight, stride, impl, x0, x1, dy, need_w, need0, need1):
    co, ci, kd, kh, kw = weight.shape
    k = (kd, kh, kw)
    pad = tuple((kk - 1) // 2 for kk in k)
    n = x0.shape[0]
    in_dims = tuple(x0.shape[2:])
    od = tuple(dy.shape[2:])
    gw = gx0 = gx1 = None
    need1 = need1 and x1 is not None
    if need_w:
        # the two gradient kernels only share their input dy: the weight gradient runs on a side stream next to the
        # data gradient (fork / join around this function, also inside CUDA graph capture).  The deep layers (16^3 /
        # 8^3 / 4^3) fill less than half of the SMs with either kernel (30.4 -> 29.7 ms per step); on the large layers
        # the tail of one persistent kernel overlaps the head of the other (-> 29.2 ms)
        # (not while bench.py brackets every launch with CUDA events: a span must time one kernel running alone)
        fork = (CONCURRENT_WGRAD and not KERNEL_TIMER.on and (need0 or need1) and dy.is_cuda
                and n * od[0] * od[1] * od[2] <= CONCURRENT_WGRAD_MAX_VOXELS)
        if fork:
            cur = torch.cuda.current_stream(dy.device)
            side = _side_stream(dy.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                dw = _launch_wgrad(dy, x0, x1, grid=od, qdims=in_dims, taps=k, off=tuple(-p for p in pad), istr=stride, impl=impl)
                gw = unpack_wgrad(dw, co, ci, k)
            try:
                return (gw, *_conv_backward_data(weight, stride, impl, x0, x1, dy, need0, need1))
            finally:
                cur.wait_stream(side)     # joined before anything that was allocated here can be freed or reused
        dw = _launch_wgrad(dy, x0, x1, grid=od, qdims=in_dims, taps=k, off=tuple(-p for p in pad), istr=stride, impl=impl)
        gw = unpack_wgrad(dw, co, ci, k)
    return (gw, *_conv_backward_data(weight, stride, impl, x0, x1, dy, need0, need1))


_SIDE_STREAMS = {}


def _side_stream(device):
    """The side stream paired with the CURRENT stream of `device` (decoders running on their own streams fork to
    their own side streams, so the forks add no dependencies between them)."""
    dev = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return st


def _conv_backward_data(weight, stride, impl, x0, x1, dy, need0, need1):
    """Data gradient(s) of `_conv_forward` w.r.t. its one or two sources; returns (gx0, gx1)."""
    co, ci, kd, kh, kw = weight.shape
    k = (kd, kh, kw)
    pad = tuple((kk - 1) // 2 for kk in k)
    n = x0.shape[0]
    in_dims = tuple(x0.shape[2:])
    od = tuple(dy.shape[2:])
    gx0 = gx1 = None
    if need0 or need1:
        c0 = x0.shape[1]
        merged = _merged_dgrad_plan(k, stride, pad, in_dims, od) if MERGED_STRIDED_DGRAD else None
        if merged is not None and ci % 32 == 0 and co % 16 == 0 and c0 % 8 == 0:
            gx0 = new_cl(n, c0, *in_dims, x0.device)
            gx1 = new_cl(n, x1.shape[1], *in_dims, x0.device) if x1 is not None else None
            npar = stride[0] * stride[1] * stride[2]
            _launch_gather(dy, None, pack_conv_dgrad_merged(weight, merged, stride), gx0, gx1, in_dims=od,
                           taps=tuple(a[0] for a in merged), off=tuple(a[1] for a in merged), istr=(1, 1, 1), out_grid=od,
                           nout=npar * ci, mode=1, ostr=stride, full=in_dims, ps=stride, psC=ci, impl=impl,
                           algo_flops=2.0 * n * od[0] * od[1] * od[2] * co * ci * k[0] * k[1] * k[2])
            return (gx0 if need0 else None), (gx1 if need1 else None)
        classes = [_axis_classes(k[a], stride[a], pad[a], in_dims[a]) for a in range(3)]
        empty = any(len(cls[2]) == 0 for axis in classes for cls in axis)
        mk = zeros_cl if empty else new_cl
        gx0 = mk(n, c0, *in_dims, x0.device)
        gx1 = mk(n, x1.shape[1], *in_dims, x0.device) if x1 is not None else None
        for cd, ch, cw in itertools.product(*classes):
            if not (cd[2] and ch[2] and cw[2]):
                continue
            full = all(len(c[2]) == kk and stride[a] == 1 for a, (c, kk) in enumerate(zip((cd, ch, cw), k)))
            wpk = pack_conv_dgrad_full(weight) if (full and weight.is_cuda and weight.dtype == torch.float32) \
                else pack_conv_dgrad_class(weight, cd[2], ch[2], cw[2])
            _launch_gather(dy, None, wpk, gx0, gx1, in_dims=od, taps=(len(cd[2]), len(ch[2]), len(cw[2])),
                           off=(cd[3], ch[3], cw[3]), istr=(1, 1, 1), out_grid=(cd[1], ch[1], cw[1]), nout=ci,
                           ostr=stride, ooff=(cd[0], ch[0], cw[0]), full=in_dims, impl=impl)
        if not need0:
            gx0 = None
        if not need1:
            gx1 = None
    return gx0, gx1


# stem: im2col of the raw NCDHW fp32 input, then a 1-tap GEMM over K = taps * Cin (padded to 16)
def _stem_im2col(x, kshape):
    L.require_cuda(x, "stem conv")
    if x.requires_grad:
        raise NotImplementedError("gradient w.r.t. the network input is not implemented")
    x = x.detach().float().contiguous()
    n, ci, d, h, w = x.shape
    kd, kh, kw = kshape
    K = kd * kh * kw * ci
    if K > 1024:
        raise NotImplementedError("stem im2col path supports up to 1024 (taps x input channels)")
    kp = (K + 15) // 16 * 16
    col = new_cl(n, kp, d, h, w, x.device)
    with KERNEL_TIMER.span("stem_im2col"):
        rc = L.load().rb_stem_im2col(x.data_ptr(), col.data_ptr(), n, ci, d, h, w, kd, kh, kw, kp, L.stream_ptr())
    L.check(rc, "rb_stem_im2col")
    return col


def _stem_pack(weight, kp):
    co, ci, kd, kh, kw = weight.shape
    K = kd * kh * kw * ci

    def pack():
        wp = torch.zeros((1, co, kp), dtype=BF16, device=weight.device)
        wp[0, :, :K] = weight.detach().permute(0, 2, 3, 4, 1).reshape(co, K).to(BF16)
        return wp
    return _cached_pack(weight, "s", pack)


def _stem_forward(weight, impl, col, out_f32=False, want_stats=False):
    co = weight.shape[0]
    n, kp, d, h, w = col.shape
    y = (new_prenorm if out_f32 else new_cl)(n, co, d, h, w, col.device)
    stats = _launch_gather(col, None, _stem_pack(weight, kp), y, None, in_dims=(d, h, w), taps=(1, 1, 1), off=(0, 0, 0),
                           istr=(1, 1, 1), out_grid=(d, h, w), nout=co, impl=impl, want_stats=want_stats)
    return y, stats


def _stem_backward(wshape, col, dy):
    co, ci, kd, kh, kw = wshape
    K = kd * kh * kw * ci
    dims = tuple(dy.shape[2:])
    dw = _launch_wgrad(dy, col, None, grid=dims, qdims=dims, taps=(1, 1, 1), off=(0, 0, 0), istr=(1, 1, 1))
    return dw[0, :, :K].reshape(co, kd, kh, kw, ci).permute(0, 4, 1, 2, 3).contiguous()


# ------------------------------------------------------------------------------------------
# primitives (no autograd): InstanceNorm (+affine, +SE gate) + residual + LeakyReLU
# ------------------------------------------------------------------------------------------
def _plane_reduce(kind, y, dz, z, per_w, slope, sign=None):
    """`sign` = (scale, shift) [N, C] fp32 of the forward pass: the activation's sign is recomputed from y instead of
    being read from the stored z (layers without residual / gate)."""
    n, c, d, h, w = y.shape
    g = w if per_w else 1
    out = torch.empty((n, g, c, 2), dtype=torch.float64, device=y.device)
    el = float(n * c * d * h * w)
    yb = 4 if y.dtype == torch.float32 else 2
    # algorithmic: forward statistics ride the conv epilogue (0 B); backward reduce reads dz + y in bf16 (4 B / element)
    with KERNEL_TIMER.span("norm_reduce", nbytes=el * (4 if kind == 1 else 2),
                           moved=el * (yb + (2 if kind == 1 else 0) + (2 if z is not None else 0))):
        rc = L.load().rb_plane_reduce(kind, y.data_ptr(), _YMODE[y.dtype], L.ptr(dz), L.ptr(z),
                                      L.ptr(sign[0]) if sign else None, L.ptr(sign[1]) if sign else None,
                                      out.data_ptr(), n, d * h * w, c, w, 1 if per_w else 0, float(slope), L.stream_ptr())
    L.check(rc, "rb_plane_reduce")
    return out


def _apply_fwd(y, res, A, B, per_w, act, slope):
    n, c, d, h, w = y.shape
    z = new_cl(n, c, d, h, w, y.device)
    el = float(n * c * d * h * w)
    yb = 4 if y.dtype == torch.float32 else 2
    with KERNEL_TIMER.span("norm_apply", nbytes=el * (4 + (2 if res is not None else 0)),
                           moved=el * (yb + 2 + (2 if res is not None else 0))):
        rc = L.load().rb_norm_act_fwd(y.data_ptr(), _YMODE[y.dtype], L.ptr(res), z.data_ptr(),
                                      A.data_ptr(), B.data_ptr(), n, d * h * w, c, w, 1 if per_w else 0, 1 if act else 0,
                                      float(slope), L.stream_ptr())
    L.check(rc, "rb_norm_act_fwd")
    return z


def _apply_bwd(dz, z, y, k1, k2, k3, per_w, act, slope, want_dres, sign=None):
    n, c, d, h, w = y.shape
    dy = new_cl(n, c, d, h, w, y.device)
    dres = new_cl(n, c, d, h, w, y.device) if want_dres else None
    el = float(n * c * d * h * w)
    yb = 4 if y.dtype == torch.float32 else 2
    # algorithmic: read dz + y (bf16), write dy (+ dres): 6 (+2) B / element; the activation sign comes with y
    with KERNEL_TIMER.span("norm_apply", nbytes=el * (6 + (2 if want_dres else 0)),
                           moved=el * (2 + yb + 2 + (2 if z is not None else 0) + (2 if want_dres else 0))):
        rc = L.load().rb_norm_act_bwd(dz.data_ptr(), L.ptr(z), L.ptr(sign[0]) if sign else None,
                                      L.ptr(sign[1]) if sign else None, y.data_ptr(), _YMODE[y.dtype],
                                      dy.data_ptr(), L.ptr(dres), k1.data_ptr(), k2.data_ptr(), k3.data_ptr(), n,
                                      d * h * w, c, w, 1 if per_w else 0, 1 if act else 0, float(slope), L.stream_ptr())
    L.check(rc, "rb_norm_act_bwd")
    return dy, dres


class _NormState:
    """What the backward of a norm (+gate) + act needs besides y, z."""
    __slots__ = ("small", "gamma", "has_beta", "act", "slope", "eps", "has_res", "gate", "per_w", "n_gate", "sums")


def _norm_forward(y, res, gamma, beta, eps, act, slope, stats=None, gate=None, reduce_dims="all", drop=None):
    """z = act( [gate *] [drop *] (IN(y) [* gamma + beta]) + res ).  `stats` = fp32 (sum, sumsq) from the conv epilogue.
    gate = (fc1_w, fc1_b, fc2_w, fc2_b) or None; drop = per-sample stochastic-depth factor [N] (0 or 1/keep) or None."""
    n, c, d, h, w = y.shape
    S = d * h * w
    lib = L.load()
    st = _NormState()
    st.gamma, st.has_beta, st.act, st.slope, st.eps, st.has_res = gamma, beta is not None, act, slope, eps, res is not None
    st.n_gate = 0 if gate is None else 4
    st.sums = None
    if gate is None and drop is None:
        st.gate, st.per_w = None, False
        sums = None if stats is not None else _plane_reduce(0, y, None, None, False, slope)
        small = torch.empty((4, n, c), dtype=torch.float32, device=y.device)     # mean, rstd, scale, shift
        L.check(lib.rb_in_finalize_fwd(L.ptr(sums), L.ptr(stats[0]) if stats else None, L.ptr(stats[1]) if stats else None,
                                       L.ptr(gamma), L.ptr(beta), small[0].data_ptr(), small[1].data_ptr(),
                                       small[2].data_ptr(), small[3].data_ptr(), n, c, float(S), float(eps),
                                       L.stream_ptr()), "rb_in_finalize_fwd")
        st.small = small
        z = _apply_fwd(y, res, small[2], small[3], False, act, slope)
        return z, st
    # gated (squeeze-excitation and / or stochastic-depth) path: the O(N*W*C) part is a small differentiable torch graph
    per_w = _se_per_w(reduce_dims) if gate is not None else False
    st.per_w = per_w
    if gate is None:
        gate = (None, None, None, None)
    if stats is not None:
        s1, s2 = stats[0].double(), stats[1].double()
    else:
        sums = _plane_reduce(0, y, None, None, False, slope)[:, 0]
        s1, s2 = sums[..., 0], sums[..., 1]
    pw = _plane_reduce(0, y, None, None, True, slope)[..., 0] if per_w else None      # [N, W, C]
    params = [gamma, beta, *gate]
    with torch.enable_grad():
        leaves = [s1.detach().clone().requires_grad_(True), s2.detach().clone().requires_grad_(True),
                  pw.detach().requires_grad_(True) if per_w else None]
        pl = [p.detach().requires_grad_(True) if p is not None else None for p in params]
        A, B = _gate_small_graph(leaves[0], leaves[1], leaves[2], float(S), float(d * h), pl[0], pl[1], eps,
                                 pl[2], pl[3], pl[4], pl[5], per_w, drop)
    st.gate = (leaves, pl, A, B)
    st.small = None
    st.sums = (s1.detach(), s2.detach(), pw.detach() if per_w else None)     # enough to rebuild the small graph
    z = _apply_fwd(y, res, A, B, per_w, act, slope)
    return z, st


@contextlib.contextmanager
def _autograd_dispatch_enabled():
    """Inside a `torch.library` operator implementation the dispatcher runs below autograd (the autograd keys sit in the
    thread-local exclude set), so `torch.enable_grad()` alone records nothing.  The gated path differentiates its
    O(N*W*C) torch graph with autograd inside the backward operator: lift the exclusion for that small graph."""
    K = torch._C.DispatchKey
    exc = torch._C._dispatch_tls_local_exclude_set()
    if not (exc.has(K.AutogradFunctionality) or exc.has(K.AutogradOther)):
        yield
        return
    new_exc = exc
    for k in (K.AutogradFunctionality, K.AutogradOther, K.AutogradNestedTensor, K.ADInplaceOrView):
        new_exc = new_exc.remove(k)
    with torch._C._ForceDispatchKeyGuard(torch._C._dispatch_tls_local_include_set(), new_exc):
        yield


def _rebuild_gate_state(st, s1, s2, pw, S, plane, gamma, beta, gate, drop):
    """The gated path's O(N*W*C) torch graph from its leaves (custom-op backward: the forward op cannot hand a
    Python graph to the backward op, only tensors)."""
    params = [gamma, beta, *(gate if gate is not None else (None, None, None, None))]
    with _autograd_dispatch_enabled(), torch.enable_grad():
        leaves = [s1.detach().clone().requires_grad_(True), s2.detach().clone().requires_grad_(True),
                  pw.detach().clone().requires_grad_(True) if pw is not None else None]
        pl = [p.detach().requires_grad_(True) if p is not None else None for p in params]
        A, B = _gate_small_graph(leaves[0], leaves[1], leaves[2], float(S), float(plane), pl[0], pl[1], st.eps,
                                 pl[2], pl[3], pl[4], pl[5], st.per_w, drop)
    st.gate = (leaves, pl, A, B)
    return st


def _norm_backward(st, y, z, dz, want_dres):
    """Returns dy, dres, dgamma, dbeta, gate parameter grads (4-tuple or None)."""
    n, c, d, h, w = y.shape
    S = d * h * w
    lib = L.load()
    if st.gate is None:
        # no residual, no gate: z = lrelu(fmaf(y, scale, shift)), so lrelu'(z) follows from y and the stored z is not read
        sign = (st.small[2], st.small[3]) if (st.act and not st.has_res and SIGN_FROM_PRENORM) else None
        zs = None if (sign is not None or not st.act) else z
        red = _plane_reduce(1, y, dz, zs, False, st.slope, sign)
        ks = torch.empty((3, n, c), dtype=torch.float32, device=y.device)
        dgamma = torch.zeros(c, dtype=torch.float32, device=y.device) if st.gamma is not None else None
        dbeta = torch.zeros(c, dtype=torch.float32, device=y.device) if st.has_beta else None
        L.check(lib.rb_in_finalize_bwd(red.data_ptr(), st.small[0].data_ptr(), st.small[1].data_ptr(), L.ptr(st.gamma),
                                       ks[0].data_ptr(), ks[1].data_ptr(), ks[2].data_ptr(), L.ptr(dgamma), L.ptr(dbeta),
                                       n, c, float(S), L.stream_ptr()), "rb_in_finalize_bwd")
        dy, dres = _apply_bwd(dz, zs, y, ks[0], ks[1], ks[2], False, st.act, st.slope, want_dres, sign)
        return dy, dres, dgamma, dbeta, None
    leaves, pl, A, B = st.gate
    per_w = st.per_w
    red = _plane_reduce(1, y, dz, z, per_w, st.slope)          # [N, G, C, 2] = (sum g, sum g*y)
    dB = red[..., 0].float()
    dA = red[..., 1].float()
    inputs = [t for t in leaves + pl if t is not None]
    with _autograd_dispatch_enabled():
        grads = torch.autograd.grad([A, B], inputs, [dA, dB], allow_unused=True)
    it = iter(grads)
    gl = [next(it) if t is not None else None for t in leaves]
    gp = [next(it) if t is not None else None for t in pl]
    g = A.shape[1]
    zero = torch.zeros((n, c), dtype=torch.float64, device=y.device)
    dS1 = gl[0] if gl[0] is not None else zero
    dS2 = gl[1] if gl[1] is not None else zero
    k2 = (2.0 * dS2).float().unsqueeze(1).expand(n, g, c).contiguous()
    k3 = dS1.unsqueeze(1).expand(n, g, c)
    if per_w and gl[2] is not None:
        k3 = k3 + gl[2]
    k3 = k3.float().contiguous()
    dy, dres = _apply_bwd(dz, z, y, A, k2, k3, per_w, st.act, st.slope, want_dres)
    return dy, dres, gp[0], gp[1], (tuple(gp[2:6]) if st.n_gate else None)


def _se_per_w(reduce_dims):
    if reduce_dims in ("all", (2, 3, 4), [2, 3, 4]):
        return False
    if tuple(reduce_dims) == (2, 3):
        return True
    raise NotImplementedError(f"SE squeeze over dims {reduce_dims} is not implemented (use 'all' or (2, 3))")


def _gate_small_graph(S1, S2, Pw, count, plane, gamma, beta, eps, w1, b1, w2, b2, per_w, drop=None):
    mean = S1 / count
    var = (S2 / count - mean * mean).clamp_min(0.0)
    rstd = torch.rsqrt(var + eps)
    a0 = rstd if gamma is None else rstd * gamma.double()
    b0 = -mean * a0 if beta is None else beta.double() - mean * a0
    if drop is not None:                          # DropPath sits between the norm and the SE gate (resblocks.py:109-112)
        f = drop.double().reshape(-1, 1)
        a0, b0 = a0 * f, b0 * f
    if w1 is None:                                # stochastic depth without squeeze-excitation
        return a0.unsqueeze(1).float().contiguous(), b0.unsqueeze(1).float().contiguous()
    if per_w:                                     # squeeze over (D, H): one value per (n, w, c)
        sq = a0.unsqueeze(1) * (Pw / plane) + b0.unsqueeze(1)
    else:                                         # global average pool of the normalised tensor
        sq = (a0 * mean + b0).unsqueeze(1)
    hid = torch.relu(torch.nn.functional.linear(sq.float(), w1.flatten(1), b1))
    gate = torch.sigmoid(torch.nn.functional.linear(hid, w2.flatten(1), b2)).double()
    A = (a0.unsqueeze(1) * gate).float()
    B = (b0.unsqueeze(1) * gate).float()
    return A.contiguous(), B.contiguous()          # [N, G, C]


# ------------------------------------------------------------------------------------------
# autograd: stand-alone operators (public functional API, tests)
# ------------------------------------------------------------------------------------------
class _Conv3dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, stride, impl, x0, x1):
        x0 = as_cl(x0)
        x1 = as_cl(x1) if x1 is not None else None
        L.require_cuda(x0, "conv3d")
        y, _ = _conv_forward(weight, stride, impl, x0, x1)
        ctx.save_for_backward(weight, x0, x1)
        ctx.stride, ctx.impl = stride, impl
        return y

    @staticmethod
    def backward(ctx, dy):
        weight, x0, x1 = ctx.saved_tensors
        gw, gx0, gx1 = _conv_backward(weight, ctx.stride, ctx.impl, x0, x1, as_cl(dy), ctx.needs_input_grad[0],
                                      ctx.needs_input_grad[3], ctx.needs_input_grad[4])
        return gw, None, None, gx0, gx1


def conv3d(x, weight, stride=1, x_cat=None, impl=None):
    """Pre-norm convolution output (bf16, channels-last).  `x_cat` is concatenated after `x`
    along channels without materialising the concatenation."""
    return _Conv3dFn.apply(weight, _triple(stride), impl, x, x_cat)


class _NormActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, res, gamma, beta, eps, act, slope, gate_dims, drop, *gate):
        y = as_prenorm(y)
        L.require_cuda(y, "instance_norm")
        res = as_cl(res) if res is not None else None
        if res is not None and res.shape != y.shape:
            raise ValueError(f"residual shape {tuple(res.shape)} != {tuple(y.shape)}")
        z, st = _norm_forward(y, res, gamma, beta, eps, act, slope, None, gate if gate else None, gate_dims, drop)
        ctx.save_for_backward(y, z if act else None)
        ctx.st = st
        return z

    @staticmethod
    def backward(ctx, dz):
        y, z = ctx.saved_tensors
        st = ctx.st
        dy, dres, dgamma, dbeta, ggate = _norm_backward(st, y, z, as_cl(dz), st.has_res and ctx.needs_input_grad[1])
        if y.dtype != dy.dtype:
            dy = dy.to(y.dtype)      # autograd wants the gradient in the input's dtype (stand-alone use only)
        ggate = ggate if ggate is not None else ()
        return (dy, dres, dgamma, dbeta, None, None, None, None, None, *ggate)


def instance_norm_act(y, res=None, gamma=None, beta=None, eps=1e-5, act=True, slope=LRELU_SLOPE_DEFAULT, drop=None):
    """z = [LeakyReLU]( [drop *] (InstanceNorm(y) [* gamma + beta]) [+ res] ); drop = per-sample factor [N] or None."""
    return _NormActFn.apply(y, res, gamma, beta, float(eps), bool(act), float(slope), "all", drop)


def instance_norm_se_act(y, res, gamma, beta, fc1_w, fc1_b, fc2_w, fc2_b, eps=1e-5, act=True,
                         slope=LRELU_SLOPE_DEFAULT, reduce_dims="all", drop=None):
    """z = [LeakyReLU]( SE([drop *] InstanceNorm(y)) + res ), SE(o) = o * sigmoid(fc2(relu(fc1(mean_dims(o)))))."""
    _se_per_w(reduce_dims)
    return _NormActFn.apply(y, res, gamma, beta, float(eps), bool(act), float(slope), reduce_dims, drop,
                            fc1_w, fc1_b, fc2_w, fc2_b)


# ------------------------------------------------------------------------------------------
# autograd: the fused unit the network is built from
#     z = act( [SE]( IN( conv(x [, x_cat]) ) ) + res )
# The pre-norm conv output stays internal (fp32, accumulator precision; its statistics come from the
# tcgen05 epilogue), so exactly one rounding to bf16 happens per stored activation.
# ------------------------------------------------------------------------------------------
class _ConvNormActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, cfg, x0, x1, res, gamma, beta, drop, *gate):
        stride, impl, eps, act, slope, gate_dims, stem = cfg
        if stem:
            src0, src1 = _stem_im2col(x0, tuple(weight.shape[2:])), None
            y, stats = _stem_forward(weight, impl, src0, out_f32=True, want_stats=True)
        else:
            src0 = as_cl(x0)
            src1 = as_cl(x1) if x1 is not None else None
            L.require_cuda(src0, "conv3d")
            y, stats = _conv_forward(weight, stride, impl, src0, src1, out_f32=True, want_stats=True)
        res = as_cl(res) if res is not None else None
        if res is not None and res.shape != y.shape:
            raise ValueError(f"residual shape {tuple(res.shape)} != {tuple(y.shape)}")
        z, st = _norm_forward(y, res, gamma, beta, eps, act, slope, stats, gate if gate else None, gate_dims, drop)
        ctx.save_for_backward(weight, src0, src1, y, z if act else None)
        ctx.st, ctx.cfg = st, cfg
        return z

    @staticmethod
    def backward(ctx, dz):
        weight, src0, src1, y, z = ctx.saved_tensors
        stride, impl, eps, act, slope, gate_dims, stem = ctx.cfg
        st = ctx.st
        need = ctx.needs_input_grad
        dy, dres, dgamma, dbeta, ggate = _norm_backward(st, y, z, as_cl(dz), st.has_res and need[4])
        if stem:
            gw = _stem_backward(tuple(weight.shape), src0, dy) if need[0] else None
            gx0 = gx1 = None
        else:
            gw, gx0, gx1 = _conv_backward(weight, stride, impl, src0, src1, dy, need[0], need[2], need[3])
        ggate = ggate if ggate is not None else ()
        return (gw, None, gx0, gx1, dres, dgamma, dbeta, None, *ggate)


def conv_norm_act(x, weight, stride=1, x_cat=None, res=None, gamma=None, beta=None, eps=1e-5, act=True,
                  slope=LRELU_SLOPE_DEFAULT, se=None, se_reduce_dims="all", stem=False, impl=None, drop=None):
    """act( [SE]( [drop *] InstanceNorm( conv3d(cat(x, x_cat), weight, stride) ) [*gamma + beta] ) + res ).
    se = (fc1_w, fc1_b, fc2_w, fc2_b) enables the squeeze-excitation gate; stem=True reads the raw NCDHW
    fp32 network input (any channel count); drop = per-sample stochastic-depth factor [N] fp32 (0 or 1/keep,
    the DropPath of resblocks.py:109-110) folded into the per-(n, c) scale / shift of the apply pass."""
    if _PreciseState.on:
        from . import precise
        return precise.conv_norm_act(x, weight, stride, x_cat, res, gamma, beta, eps, act, slope, se, se_reduce_dims,
                                     stem, impl or _PreciseState.impl, drop)
    if se is not None:
        _se_per_w(se_reduce_dims)
    if torch.compiler.is_compiling() or FORCE_CUSTOM_OPS:
        return custom_ops.conv_norm_act(x, weight, _triple(stride), x_cat, res, gamma, beta, float(eps), bool(act),
                                        float(slope), se, se_reduce_dims, bool(stem), impl, drop)
    cfg = (_triple(stride), impl, float(eps), bool(act), float(slope), se_reduce_dims, bool(stem))
    return _ConvNormActFn.apply(weight, cfg, x, x_cat, res, gamma, beta, drop, *(se or ()))


FUSE_HEAD = os.environ.get("RESENC_FUSE_HEAD", "1") != "0"


def can_fuse_head(conv_weight, head, se=None, drop=None) -> bool:
    """True when conv -> InstanceNorm -> LeakyReLU -> 1x1x1 head can run without storing the activation in between
    (`conv_norm_act_head`): autograd off (inference), no gate, no stochastic depth, channel count the kernel takes."""
    if head is None or not FUSE_HEAD or torch.is_grad_enabled() or _PreciseState.on or se is not None or drop is not None:
        return False
    if torch.compiler.is_compiling() or FORCE_CUSTOM_OPS:
        return False
    cg, k = conv_weight.shape[0] // 8, head[0].shape[0]
    return conv_weight.shape[0] % 8 == 0 and cg & (cg - 1) == 0 and cg <= 32 and k <= 8


def conv_norm_act_head(x, weight, stride, x_cat, res, gamma, beta, eps, act, slope, stem, head):
    """Inference tail of a task decoder (decoder.py:144-152): head( act( IN( conv(cat(x, x_cat)) ) [+ res] ) ) with the
    decoder's last activation never written to HBM - the normalise pass and the head are one kernel
    (`rb_norm_act_head_fwd`) that reads the pre-norm conv output once.  No autograd (see `can_fuse_head`);
    head = (weight [K, C, 1, 1, 1], bias [K] | None, activation None | 'sigmoid' | 'softmax').  Returns NCDHW fp32."""
    hw, hb, hact = head
    lib = L.load()
    if stem:
        src0 = _stem_im2col(x, tuple(weight.shape[2:]))
        y, stats = _stem_forward(weight, None, src0, out_f32=True, want_stats=True)
    else:
        src0 = as_cl(x)
        src1 = as_cl(x_cat) if x_cat is not None else None
        L.require_cuda(src0, "conv3d")
        y, stats = _conv_forward(weight, _triple(stride), None, src0, src1, out_f32=True, want_stats=True)
    res = as_cl(res) if res is not None else None
    if res is not None and res.shape != y.shape:
        raise ValueError(f"residual shape {tuple(res.shape)} != {tuple(y.shape)}")
    n, c, d, h, w = y.shape
    S = d * h * w
    sums = None if stats is not None else _plane_reduce(0, y, None, None, False, slope)
    small = torch.empty((4, n, c), dtype=torch.float32, device=y.device)     # mean, rstd, scale, shift
    L.check(lib.rb_in_finalize_fwd(L.ptr(sums), L.ptr(stats[0]) if stats else None, L.ptr(stats[1]) if stats else None,
                                   L.ptr(gamma), L.ptr(beta), small[0].data_ptr(), small[1].data_ptr(),
                                   small[2].data_ptr(), small[3].data_ptr(), n, c, float(S), float(eps),
                                   L.stream_ptr()), "rb_in_finalize_fwd")
    k = hw.shape[0]
    w2 = hw.detach().reshape(k, c).float().contiguous()
    b = hb.detach().float().contiguous() if hb is not None else None
    out = torch.empty((n, k, d, h, w), dtype=torch.float32, device=y.device)
    el = float(n * c * S)
    yb = 4 if y.dtype == torch.float32 else 2
    with KERNEL_TIMER.span("norm_apply", nbytes=el * (2 + (2 if res is not None else 0)) + 4.0 * n * k * S,
                           moved=el * (yb + (2 if res is not None else 0)) + 4.0 * n * k * S):
        rc = lib.rb_norm_act_head_fwd(y.data_ptr(), _YMODE[y.dtype], L.ptr(res), small[2].data_ptr(), small[3].data_ptr(),
                                      w2.data_ptr(), L.ptr(b), out.data_ptr(), n, S, c, k, 1 if act else 0, float(slope),
                                      _ACT[hact if hact is None else str(hact).lower()], L.stream_ptr())
    L.check(rc, "rb_norm_act_head_fwd")
    return out


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
def _permuted_bf16(weight, perm, shape):
    """bf16(weight.permute(perm)) laid out contiguously and viewed as `shape`, in ONE strided copy + cast kernel
    (`permute().reshape().to(bf16)` materialises the fp32 permutation first: two launches per pack, 20 per step for the
    transposed convs of the two decoders)."""
    src = weight.detach().permute(*perm)
    out = torch.empty(src.shape, dtype=BF16, device=weight.device)
    out.copy_(src)
    return out.view(shape)


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
        wpk = _cached_pack(weight, "t", lambda: _permuted_bf16(weight, (2, 3, 4, 1, 0), (1, npar * co, ci)))
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
        fork = None
        if ctx.needs_input_grad[0]:
            if CONCURRENT_CONVT_WGRAD and not KERNEL_TIMER.on and ctx.needs_input_grad[3] and dy.is_cuda:
                fork = (torch.cuda.current_stream(dy.device), _side_stream(dy.device))    # see _conv_backward
                fork[1].wait_stream(fork[0])
            with (torch.cuda.stream(fork[1]) if fork else contextlib.nullcontext()):
                dw = _launch_wgrad(x, dy, None, grid=in_dims, qdims=full, taps=stride, off=(0, 0, 0), istr=stride, impl=impl)
                gw = unpack_wgrad(dw, ci, co, (sd, sh, sw))
        if ctx.needs_input_grad[3]:
            gx = new_cl(n, ci, *in_dims, x.device)
            wpk = _permuted_bf16(weight, (2, 3, 4, 0, 1), (npar, ci, co))
            _launch_gather(dy, None, wpk, gx, None, in_dims=full, taps=stride, off=(0, 0, 0), istr=stride,
                           out_grid=in_dims, nout=ci, impl=impl)
        if fork:
            fork[0].wait_stream(fork[1])
        return gw, None, None, gx


def conv_transpose3d(x, weight, stride, impl=None, bias=None):
    """ConvTranspose3d with kernel_size == stride (builders/decoder.py:110-113,147), optional bias."""
    if _PreciseState.on:
        from . import precise
        return precise.conv_transpose3d(x, weight, stride, impl or _PreciseState.impl, bias)
    if torch.compiler.is_compiling() or FORCE_CUSTOM_OPS:
        up = custom_ops.conv_transpose3d(x, weight, _triple(stride), impl)
    else:
        up = _ConvT3dFn.apply(weight, _triple(stride), impl, x)
    if bias is not None:
        # rare configuration (conv_bias=True): per-channel add on the channels-last buffer, glue
        up = (up.permute(0, 2, 3, 4, 1) + bias.to(up.dtype)).permute(0, 4, 1, 2, 3)
    return up


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
    if _PreciseState.on:
        from . import precise
        return precise.avg_pool3d(x, stride)
    if torch.compiler.is_compiling() or FORCE_CUSTOM_OPS:
        return custom_ops.avg_pool3d(x, _triple(stride))
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
    if _PreciseState.on:
        from . import precise
        return precise.head_conv1x1(x, weight, bias, activation)
    if weight.shape[0] > 8:
        raise NotImplementedError("task heads with more than 8 output channels are not implemented")
    act = _ACT[activation if activation is None else str(activation).lower()]
    if torch.compiler.is_compiling() or FORCE_CUSTOM_OPS:
        return custom_ops.head_conv1x1(x, weight, bias, act)
    return _HeadFn.apply(x, weight, bias, act)


# ------------------------------------------------------------------------------------------
# Stem convolution on the raw NCDHW fp32 input (Cin not a multiple of 8): im2col + 1-tap GEMM
# ------------------------------------------------------------------------------------------
class _StemConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, impl):
        col = _stem_im2col(x, tuple(weight.shape[2:]))
        y, _ = _stem_forward(weight, impl, col)
        ctx.save_for_backward(col)
        ctx.wshape = tuple(weight.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        (col,) = ctx.saved_tensors
        gw = _stem_backward(ctx.wshape, col, as_cl(dy)) if ctx.needs_input_grad[1] else None
        return None, gw, None


def stem_conv3d(x, weight, impl=None):
    """Stride-1 'same' convolution of the raw network input (any Cin); pre-norm bf16 output."""
    if x.shape[1] != weight.shape[1]:
        raise ValueError(f"stem conv: weight expects {weight.shape[1]} input channels, got {x.shape[1]}")
    return _StemConvFn.apply(x, weight, impl)


# the torch.library registration of the same units (imported last: it builds on the primitives above)
from . import custom_ops  # noqa: E402
