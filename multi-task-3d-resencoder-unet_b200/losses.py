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


def task_losses(tasks):
    """name -> loss module, the pairing `bench.py` uses: MaskedCosineLoss for a 3-channel task called "normals",
    BCEDiceLoss(0.5, 0.5) otherwise."""
    return {t: (MaskedCosineLoss() if t == "normals" and info.get("channels") == 3 else BCEDiceLoss(0.5, 0.5))
            for t, info in tasks.items()}
