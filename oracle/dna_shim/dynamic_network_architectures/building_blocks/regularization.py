"""ORACLE SCAFFOLDING — parity UNPINNED for this file.

`SqueezeExcite` and `DropPath` (reference call sites builders/resblocks.py:11,81,86-87,
110,112) live in PyPI `dynamic-network-architectures`, which the reference neither vendors
nor pins.  This is a restatement of the published timm-derived algorithm (SURVEY.md
Appendix A).  `SqueezeExcite.reduce_dims` selects the squeeze: "all" = global average pool
(what the north star names), (2, 3) = timm's 2-D code path applied verbatim to 5-D input.
"""
import torch
from torch import nn

SE_REDUCE_DIMS = "all"   # module-level switch used by the oracle tests


def make_divisible(v, divisor=8, min_value=None, round_limit=.9):
    min_value = min_value or divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < round_limit * v:
        new_v += divisor
    return new_v


class SqueezeExcite(nn.Module):
    def __init__(self, channels, conv_op, rd_ratio=1. / 16, rd_channels=None, rd_divisor=8,
                 add_maxpool=False, act_layer=nn.ReLU, norm_layer=None, gate_layer=nn.Sigmoid):
        super().__init__()
        if not rd_channels:
            rd_channels = make_divisible(channels * rd_ratio, rd_divisor, round_limit=0.)
        self.fc1 = conv_op(channels, rd_channels, kernel_size=1, bias=True)
        self.bn = nn.Identity()
        self.act = act_layer(inplace=True)
        self.fc2 = conv_op(rd_channels, channels, kernel_size=1, bias=True)
        self.gate = gate_layer()

    def forward(self, x):
        dims = tuple(range(2, x.dim())) if SE_REDUCE_DIMS == "all" else tuple(SE_REDUCE_DIMS)
        x_se = x.mean(dims, keepdim=True)
        x_se = self.fc2(self.act(self.bn(self.fc1(x_se))))
        return x * self.gate(x_se)


class DropPath(nn.Module):
    def __init__(self, drop_prob=0., scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask
