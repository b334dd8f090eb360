"""Optimiser step of the training hot path: global-norm gradient clip + AdamW on the sm_100a kernels.

Reference: /root/reference/train.py:79-83 (`AdamW(model.parameters(), lr=initial_lr, weight_decay=weight_decay)`),
train.py:227-228 (`clip_grad_norm_(model.parameters(), 3)` then `optimizer.step()`).  `ClippedAdamW` is a
`torch.optim.Optimizer` with torch.optim.AdamW's hyper-parameters, state layout (`step`, `exp_avg`, `exp_avg_sq` per
parameter) and `state_dict()` format, so checkpoints move between the two (train.py:160,251); the difference is where
the work happens: two multi-tensor passes (`rb_grad_sumsq`, `rb_adamw_clip_step`) instead of PyTorch's three sweeps
(per-tensor norms, in-place scaling of every gradient, fused update), with the clip coefficient applied to each gradient
as the update reads it.  Gradients are therefore NOT rescaled in place (after `clip_grad_norm_` they would be).

Everything the step needs lives on the device (`lr`, the step count, the squared norm), so a step captured in a CUDA
graph follows LR schedules that write `param_group["lr"]` as a tensor, and never synchronises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


class _OptTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_longlong)]


class ClippedAdamW(torch.optim.Optimizer):
    """AdamW with decoupled weight decay (torch.optim.AdamW semantics) and an optional global-norm gradient clip
    (`max_grad_norm`, the `clip_grad_norm_(..., max_norm)` of train.py:227) folded into the update.

    fp32 CUDA parameters and gradients only; `amsgrad` / `maximize` are not implemented (the reference uses neither).
    `last_grad_norm()` returns the pre-clip global norm of the last step as a device tensor (what `clip_grad_norm_`
    returns)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None,
                 amsgrad=False, maximize=False, manage_packs=False):
        if amsgrad or maximize:
            raise NotImplementedError("ClippedAdamW: amsgrad / maximize are not implemented on the B200 path")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"invalid betas {betas}")
        if eps < 0.0 or weight_decay < 0.0:
            raise ValueError("eps and weight_decay must be non-negative")
        if max_grad_norm is not None and not max_grad_norm > 0:
            raise ValueError("max_grad_norm must be positive (or None: no clipping)")
        # the keys torch.optim.AdamW writes into param_groups, so state_dicts are interchangeable
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=True, differentiable=False, fused=True, decoupled_weight_decay=True)
        super().__init__(params, defaults)
        self.max_grad_norm = None if max_grad_norm is None else float(max_grad_norm)
        # manage_packs: conv weights that the network consumes as packed bf16 operands (ops.pack_conv_fprop marks them)
        # are updated by rb_adamw_clip_pack_step, which writes next step's operands too (ops.opt_packs).  Bit-identical
        # results; measured 0.3 ms per step SLOWER than the multi-tensor update + pack kernel on the 128^3 network
        # (csrc/optim.cuh), hence off by default
        self.manage_packs = bool(manage_packs)
        self._sumsq = None
        self._lr_dev = {}
        self._tables = {}

    # ---- state ------------------------------------------------------------------------------------------------
    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif not (torch.is_tensor(st["step"]) and st["step"].is_cuda and st["step"].dtype == torch.float32):
            # a checkpoint written by torch.optim.AdamW (python float / CPU tensor step)
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32, device=p.device)
        return st

    def _group_step(self, gi, params):
        """One device step counter per group: every parameter's `step` entry aliases it (state_dict() still writes one
        entry per parameter, as torch does)."""
        first = self.state[params[0]]["step"]
        for p in params[1:]:
            st = self.state[p]
            if st["step"] is not first:
                if float(st["step"]) != float(first):          # only right after a load / for late-activated parameters
                    raise RuntimeError("ClippedAdamW: parameters of one group carry different step counts")
                st["step"] = first
        return first

    def _lr_tensor(self, gi, group, device):
        lr = group["lr"]
        if torch.is_tensor(lr):
            if not lr.is_cuda or lr.dtype != torch.float32:
                lr = lr.to(device=device, dtype=torch.float32)
                group["lr"] = lr
            return lr.reshape(())
        t = self._lr_dev.get(gi)
        if t is None or t[0] != float(lr):
            if torch.cuda.is_current_stream_capturing() and t is not None:
                raise RuntimeError("ClippedAdamW: a python-float lr changed under CUDA graph capture; set "
                                   "param_group['lr'] to a CUDA tensor for schedules under graphs")
            dev = t[1] if t is not None else torch.empty((), dtype=torch.float32, device=device)
            dev.fill_(float(lr))
            t = (float(lr), dev)
            self._lr_dev[gi] = t
        return t[1]

    def last_grad_norm(self):
        return None if self._sumsq is None else self._sumsq.sqrt().float()

    # ---- step -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        active = []
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32 and p.is_contiguous()
                        and p.grad.is_contiguous() and not p.grad.is_sparse):
                    raise NotImplementedError("ClippedAdamW: contiguous fp32 CUDA parameters and dense gradients only")
                self._init_state(p)
            ps.sort(key=lambda q: -q.numel())               # launches of similar-sized tensors
            active.append((gi, group, ps))
        if not active:
            return loss
        device = active[0][2][0].device
        stream = L.stream_ptr()
        if self._sumsq is None:
            self._sumsq = torch.zeros((), dtype=torch.float64, device=device)
        tables = []
        for gi, group, ps in active:
            arr = (_OptTensor * len(ps))()
            for i, p in enumerate(ps):
                st = self.state[p]
                arr[i].p, arr[i].g = p.data_ptr(), p.grad.data_ptr()
                arr[i].m, arr[i].v, arr[i].n = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()
            tables.append(arr)
        if self.max_grad_norm is not None:
            # the norm runs over ALL parameter groups, like clip_grad_norm_(model.parameters(), ...)
            total = sum(len(t) for t in tables)
            flat = (_OptTensor * total)()
            k = 0
            for t in tables:
                for e in t:
                    flat[k] = e
                    k += 1
            L.check(lib.rb_grad_sumsq(flat, total, self._sumsq.data_ptr(), stream), "rb_grad_sumsq")
        sumsq_ptr = self._sumsq.data_ptr() if self.max_grad_norm is not None else None
        for (gi, group, ps), arr in zip(active, tables):
            step_t = self._group_step(gi, ps)
            step_t.add_(1.0)
            lr_t = self._lr_tensor(gi, group, device)
            b1, b2 = group["betas"]
            hyper = (lr_t.data_ptr(), step_t.data_ptr(), sumsq_ptr, float(self.max_grad_norm or 0.0), float(b1), float(b2),
                     float(group["eps"]), float(group["weight_decay"]), stream)
            if self.manage_packs:
                rest = (_OptTensor * len(ps))()
                k = 0
                for i, p in enumerate(ps):
                    if self._packable(p):
                        self._fused_pack_step(lib, p, arr[i], hyper)
                    else:
                        rest[k] = arr[i]
                        k += 1
                arr, n = rest, k
            else:
                n = len(ps)
            if n:
                L.check(lib.rb_adamw_clip_step(arr, n, *hyper), "rb_adamw_clip_step")
        return loss

    # ---- optimiser-managed operand packs ------------------------------------------------------------------------
    @staticmethod
    def _packable(p):
        return (getattr(p, "_rb_wants_fd", False) and p.dim() == 5 and p.shape[1] % 32 == 0
                and p.shape[2] * p.shape[3] * p.shape[4] <= 27 and p.data_ptr() % 16 == 0 and p.grad.data_ptr() % 16 == 0)

    @staticmethod
    def _pack_buffers(p):
        ent = getattr(p, "_rb_opt_packs", None)
        co, ci, kd, kh, kw = p.shape
        T = kd * kh * kw
        if ent is not None and ent["ptr"] == p.data_ptr() and tuple(ent["f"].shape) == (T, co, ci):
            return ent["f"], ent["d"]            # same buffers for the life of the parameter (captured graphs read them)
        f = torch.empty((T, co, ci), dtype=torch.bfloat16, device=p.device)
        d = torch.empty((T, ci, co), dtype=torch.bfloat16, device=p.device)
        return f, d

    def _fused_pack_step(self, lib, p, e, hyper):
        from . import ops
        f, d = self._pack_buffers(p)
        co, ci, kd, kh, kw = p.shape
        L.check(lib.rb_adamw_clip_pack_step(e.p, e.g, e.m, e.v, f.data_ptr(), d.data_ptr(), co, ci, kd * kh * kw, *hyper),
                "rb_adamw_clip_pack_step")
        # valid for the epoch the global post-step hook (ops._on_optimizer_step) is about to open
        ops.opt_packs_store(p, f, d, ops._PACK_EPOCH[0] + 1)

    @torch.no_grad()
    def refresh_packs(self):
        """Re-pack every managed weight from its CURRENT value into its persistent buffers (after parameters were
        modified behind the optimiser's back - copy_, load_state_dict, broadcast - while a captured graph that reads
        those buffers is still in use)."""
        from . import ops
        lib = L.load()
        for group in self.param_groups:
            for p in group["params"]:
                ent = getattr(p, "_rb_opt_packs", None)
                if ent is None or ent["ptr"] != p.data_ptr():
                    continue
                co, ci, kd, kh, kw = p.shape
                L.check(lib.rb_pack_conv_weights(p.data_ptr(), ent["f"].data_ptr(), ent["d"].data_ptr(), co, ci,
                                                 kd * kh * kw, L.stream_ptr()), "rb_pack_conv_weights")
                ops.opt_packs_store(p, ent["f"], ent["d"], ops._PACK_EPOCH[0])

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._lr_dev = {}
        for group in self.param_groups:                      # steps become device tensors again and re-alias lazily
            for p in group["params"]:
                if p in self.state and len(self.state[p]):
                    self._init_state(p)
