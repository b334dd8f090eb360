"""The two task losses of the BASELINE training step, as the reference trainer applies them to the network's output
dict (training/losses/losses.py): `BCEDiceLoss(alpha, beta)` for sigmoid targets (:307-318 = label-smoothed
BCE-with-logits :217-238 + `DiceLoss` :105-126 over `compute_per_channel_dice` :17-43) and `MaskedCosineLoss`
(:187-215) for the normals.

These are the *callers'* side of the drop-in boundary: plain PyTorch compositions that run on the
fp32 logits the head kernel writes, kept here so that `bench.py` and the tools time a training step without touching
the oracle.  Same class names, constructor arguments and reductions as the reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _channel_first_flat(t: torch.Tensor) -> torch.Tensor:
    """[N, C, ...] -> [C, N * spatial] (losses.py:321-332)."""
    return t.transpose(0, 1).reshape(t.shape[1], -1)


class DiceLoss(nn.Module):
    """1 - mean over channels of 2 * sum(p * t) / max(sum(p^2) + sum(t^2), eps), p = sigmoid(logits)."""

    def __init__(self, epsilon: float = 1e-6):
        super().__init__()
        self.epsilon = epsilon

    def forward(self, logits, target):
        p = _channel_first_flat(torch.sigmoid(logits))
        t = _channel_first_flat(target.float())
        inter = (p * t).sum(-1)
        den = (p * p).sum(-1) + (t * t).sum(-1)
        return 1.0 - (2.0 * inter / den.clamp(min=self.epsilon)).mean()


class BCEWithLogitsLossLabelSmoothing(nn.Module):
    """y in {0, 1} -> y * (1 - 2a) + a, then BCE-with-logits."""

    def __init__(self, smoothing: float = 0.1, reduction: str = "mean"):
        super().__init__()
        self.smoothing = smoothing
        self.reduction = reduction

    def forward(self, logits, targets):
        with torch.no_grad():
            smoothed = targets * (1.0 - 2.0 * self.smoothing) + self.smoothing
        return F.binary_cross_entropy_with_logits(logits, smoothed, reduction=self.reduction)


class BCEWithLogitsLossZSmooth(nn.Module):
    """Label smoothing that grows linearly with the distance from the centre z slice: alpha(z) = center +
    (edge - center) * |z - (D-1)/2| / (D // 2), target -> y * (1 - 2 alpha) + alpha, then BCE-with-logits
    (losses.py:240-304)."""

    def __init__(self, center_smoothing: float = 0.1, edge_smoothing: float = 0.4, reduction: str = "mean"):
        super().__init__()
        self.center_smoothing, self.edge_smoothing, self.reduction = center_smoothing, edge_smoothing, reduction

    def forward(self, logits, targets):
        if logits.shape != targets.shape or logits.dim() != 5:
            raise ValueError("BCEWithLogitsLossZSmooth expects matching [B, C, D, H, W] tensors")
        D = logits.shape[2]
        z = torch.arange(D, device=logits.device, dtype=logits.dtype)
        ratio = torch.abs(z - (D - 1) / 2.0) / (D // 2)
        alpha = (self.center_smoothing + (self.edge_smoothing - self.center_smoothing) * ratio).view(1, 1, D, 1, 1)
        smoothed = targets * (1.0 - 2.0 * alpha) + alpha
        return F.binary_cross_entropy_with_logits(logits, smoothed, reduction=self.reduction)


class BCEDiceLoss(nn.Module):
    def __init__(self, alpha: float, beta: float):
        super().__init__()
        self.alpha, self.beta = alpha, beta
        self.bce = BCEWithLogitsLossLabelSmoothing(smoothing=0.1, reduction="mean")
        self.dice = DiceLoss()

    def forward(self, input, target):
        return self.alpha * self.bce(input, target) + self.beta * self.dice(input, target)


class MaskedCosineLoss(nn.Module):
    """1 - mean cosine similarity between the unit-normalised prediction and the target over voxels whose target
    vector is non-zero."""

    def forward(self, pred, target):
        mask = (torch.norm(target, dim=1) > 1e-6).float()
        unit = pred / torch.norm(pred, dim=1, keepdim=True).clamp(min=1e-8)
        cos = F.cosine_similarity(unit, target, dim=1, eps=1e-8) * mask
        return 1.0 - cos.sum() / (mask.sum() + 1e-8)


# ------------------------------------------------------------------------------------------
# Fused CUDA versions (csrc/loss.cuh): one reduction pass + a few scalar ops forward, one pass that writes
# d(loss)/d(logits) backward.  Same values as the compositions above (tests/test_gpu_ops.py).
# ------------------------------------------------------------------------------------------
def _fused_ok(x, t):
    return x.is_cuda and t.is_cuda and x.dtype == torch.float32 and t.dtype == torch.float32 and x.dim() >= 3 \
        and x.shape == t.shape


class _BCEDiceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, alpha, beta, eps, smoothing):
        from . import _lib as L
        x, t = logits.contiguous(), target.contiguous()
        nb, c = x.shape[:2]
        s = x[0, 0].numel()
        stats = torch.zeros((c, 4), dtype=torch.float64, device=x.device)
        L.check(L.load().rb_loss_bce_dice_reduce(x.data_ptr(), t.data_ptr(), stats.data_ptr(), nb, c, s, smoothing,
                                                 L.stream_ptr()), "rb_loss_bce_dice_reduce")
        bce = stats[:, 0].sum() / float(nb * c * s)
        dice = 2.0 * stats[:, 1] / (stats[:, 2] + stats[:, 3]).clamp(min=eps)
        ctx.save_for_backward(x, t, stats)
        ctx.cfg = (alpha, beta, eps, smoothing)
        return (alpha * bce + beta * (1.0 - dice.mean())).float()

    @staticmethod
    def backward(ctx, g):
        from . import _lib as L
        x, t, stats = ctx.saved_tensors
        alpha, beta, eps, smoothing = ctx.cfg
        nb, c = x.shape[:2]
        dx = torch.empty_like(x)
        g = g.detach().float().contiguous()
        L.check(L.load().rb_loss_bce_dice_grad(x.data_ptr(), t.data_ptr(), stats.data_ptr(), g.data_ptr(), dx.data_ptr(), nb,
                                               c, x[0, 0].numel(), alpha, beta, eps, smoothing, L.stream_ptr()),
                "rb_loss_bce_dice_grad")
        return dx, None, None, None, None, None


class _MaskedCosineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        from . import _lib as L
        x, t = pred.contiguous(), target.contiguous()
        nb = x.shape[0]
        s = x[0, 0].numel()
        stats = torch.zeros(2, dtype=torch.float64, device=x.device)
        L.check(L.load().rb_loss_cosine_reduce(x.data_ptr(), t.data_ptr(), stats.data_ptr(), nb, s, L.stream_ptr()),
                "rb_loss_cosine_reduce")
        ctx.save_for_backward(x, t, stats)
        return (1.0 - stats[0] / (stats[1] + 1e-8)).float()

    @staticmethod
    def backward(ctx, g):
        from . import _lib as L
        x, t, stats = ctx.saved_tensors
        dx = torch.empty_like(x)
        g = g.detach().float().contiguous()
        L.check(L.load().rb_loss_cosine_grad(x.data_ptr(), t.data_ptr(), stats.data_ptr(), g.data_ptr(), dx.data_ptr(),
                                             x.shape[0], x[0, 0].numel(), L.stream_ptr()), "rb_loss_cosine_grad")
        return dx, None


class FusedBCEDiceLoss(BCEDiceLoss):
    """BCEDiceLoss through csrc/loss.cuh when logits and target are fp32 CUDA tensors (else the composition)."""

    def forward(self, input, target):
        if _fused_ok(input, target):
            return _BCEDiceFn.apply(input, target, float(self.alpha), float(self.beta), float(self.dice.epsilon),
                                    float(self.bce.smoothing))
        return super().forward(input, target)


class FusedMaskedCosineLoss(MaskedCosineLoss):
    def forward(self, pred, target):
        if _fused_ok(pred, target) and pred.shape[1] == 3:
            return _MaskedCosineFn.apply(pred, target)
        return super().forward(pred, target)


def task_losses(tasks, fused: bool = True):
    """name -> loss module, the pairing `bench.py` uses: MaskedCosineLoss for a 3-channel task called "normals",
    BCEDiceLoss(0.5, 0.5) otherwise; `fused` selects the CUDA kernels (same values)."""
    cos = FusedMaskedCosineLoss if fused else MaskedCosineLoss
    bd = FusedBCEDiceLoss if fused else BCEDiceLoss
    return {t: (cos() if t == "normals" and info.get("channels") == 3 else bd(0.5, 0.5)) for t, info in tasks.items()}


# name -> class, the table BaseTrainer._build_loss resolves `tasks[t]["loss_fn"]` against (train.py:47-56)
LOSS_FN_MAP = {
    "BCEDiceLoss": BCEDiceLoss,
    "BCEWithLogitsLossLabelSmoothing": BCEWithLogitsLossLabelSmoothing,
    "BCEWithLogitsLossZSmooth": BCEWithLogitsLossZSmooth,
    "BCEWithLogitsLoss": nn.BCEWithLogitsLoss,
    "BCELoss": nn.BCELoss,
    "CrossEntropyLoss": nn.CrossEntropyLoss,
    "MSELoss": nn.MSELoss,
    "MaskedCosineLoss": MaskedCosineLoss,
}
_FUSED = {"BCEDiceLoss": FusedBCEDiceLoss, "MaskedCosineLoss": FusedMaskedCosineLoss}


def build_task_losses(tasks, fused: bool = True):
    """BaseTrainer._build_loss (train.py:45-66): per task `loss_fn` (default "BCEDiceLoss") constructed with
    `loss_kwargs`; unknown names raise ValueError like the reference.  `fused` swaps in the CUDA versions of the two
    losses that have one (same values)."""
    out = {}
    for t, info in tasks.items():
        name = info.get("loss_fn", "BCEDiceLoss")
        if name not in LOSS_FN_MAP:
            raise ValueError(f"Loss function {name} not found in LOSS_FN_MAP. Add it to the mapping and try again.")
        cls = _FUSED.get(name, LOSS_FN_MAP[name]) if fused else LOSS_FN_MAP[name]
        out[t] = cls(**info.get("loss_kwargs", {}))
    return out
