"""ConvDropoutNormReLU / StackedConvBlocks — drop-in for the reference's
builders/simple_conv_blocks.py:13-138 (same constructor signatures, attribute names and
state_dict keys), with the forward pass running on the sm_100a kernels.

The torch.nn children (`conv`, `dropout`, `norm`, `nonlin`, `all_modules`) are kept as parameter
containers so that initialisation (same RNG stream as the reference), `.parameters()`,
`.state_dict()` and checkpoints are interchangeable with the reference; they are never called.
"""
from __future__ import annotations

import numpy as np
from torch import nn

from .. import ops
from .utils import maybe_convert_scalar_to_list


def _unsupported(what):
    raise NotImplementedError(f"{what} is not implemented on the B200 path (there is no PyTorch fallback)")


class ConvDropoutNormReLU(nn.Module):
    def __init__(self, conv_op, input_channels, output_channels, kernel_size, stride, conv_bias=False,
                 norm_op=None, norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None,
                 nonlin=None, nonlin_kwargs=None, nonlin_first=False):
        super().__init__()
        if conv_op is not nn.Conv3d:
            _unsupported(f"conv_op {conv_op} (only nn.Conv3d)")
        self.input_channels = input_channels
        self.output_channels = output_channels
        self.stride = maybe_convert_scalar_to_list(conv_op, stride)
        kernel_size = maybe_convert_scalar_to_list(conv_op, kernel_size)
        norm_op_kwargs = norm_op_kwargs or {}
        nonlin_kwargs = nonlin_kwargs or {}
        if any(k not in (1, 3) for k in kernel_size):
            _unsupported(f"kernel size {kernel_size} (1 and 3 per axis)")
        if any(s not in (1, 2) for s in self.stride):
            _unsupported(f"stride {self.stride} (1 and 2 per axis)")
        if norm_op is not None and norm_op is not nn.InstanceNorm3d:
            _unsupported(f"norm_op {norm_op} (only nn.InstanceNorm3d)")
        if norm_op is None:
            _unsupported("a conv block without normalisation")
        if norm_op_kwargs.get("track_running_stats", False):
            _unsupported("InstanceNorm3d(track_running_stats=True)")
        if nonlin is not None and nonlin is not nn.LeakyReLU:
            _unsupported(f"nonlin {nonlin} (only nn.LeakyReLU)")
        if nonlin_first and nonlin is not None:
            _unsupported("nonlin_first=True")
        if dropout_op is not None and float((dropout_op_kwargs or {}).get("p", 0.5)) != 0.0:
            _unsupported("dropout with p > 0")

        mods = []
        self.conv = conv_op(input_channels, output_channels, kernel_size, self.stride,
                            padding=[(k - 1) // 2 for k in kernel_size], dilation=1, bias=conv_bias)
        mods.append(self.conv)
        if dropout_op is not None:
            self.dropout = dropout_op(**dropout_op_kwargs)
            mods.append(self.dropout)
        self.norm = norm_op(output_channels, **norm_op_kwargs)
        mods.append(self.norm)
        if nonlin is not None:
            self.nonlin = nonlin(**nonlin_kwargs)
            mods.append(self.nonlin)
        self.all_modules = nn.Sequential(*mods)
        self._act = nonlin is not None
        self._slope = float(self.nonlin.negative_slope) if self._act else 0.0

    _raw_input = False      # StemConv: consumes the raw NCDHW fp32 network input

    def forward(self, x, x_cat=None, res=None, act=None, slope=None, se=None, se_reduce_dims="all", drop=None, head=None):
        """act( [SE]( IN( conv(cat(x, x_cat)) ) ) + res ) as ONE fused unit (ops.conv_norm_act).  Called with
        only `x` this is the reference's conv -> dropout(p=0) -> norm -> nonlin; residual blocks pass their
        tail (`res`, `act`, `se`) so the block needs no elementwise pass of its own.  A conv bias feeding
        InstanceNorm cancels exactly: it is not applied and receives the exact gradient, zero."""
        act = self._act if act is None else act
        slope = (self._slope or ops.LRELU_SLOPE_DEFAULT) if slope is None else slope
        if ops.can_fuse_head(self.conv.weight, head, se, drop):
            # inference: `head` = (weight, bias, activation) of the task's 1x1x1 seg layer consumes this unit's output and
            # nothing else does - norm + act + head run as one pass and the activation is never stored
            return ops.conv_norm_act_head(x, self.conv.weight, self.stride, x_cat, res, self.norm.weight, self.norm.bias,
                                          self.norm.eps, act, slope, self._raw_input, head)
        z = ops.conv_norm_act(x, self.conv.weight, self.stride, x_cat, res, self.norm.weight, self.norm.bias,
                              self.norm.eps, act, slope, se, se_reduce_dims, stem=self._raw_input, drop=drop)
        z = ops.attach_cancelled_bias(z, self.conv.bias)
        return z if head is None else ops.head_conv1x1(z, *head)

    def compute_conv_feature_map_size(self, input_size):
        assert len(input_size) == len(self.stride), "give the spatial size only, e.g. (x, y, z)"
        return np.prod([self.output_channels, *[i // j for i, j in zip(input_size, self.stride)]], dtype=np.int64)


class StemConv(ConvDropoutNormReLU):
    """First conv of the network: consumes the raw NCDHW fp32 input (any channel count)."""
    _raw_input = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if any(s != 1 for s in self.stride):
            _unsupported("a strided stem convolution")


class StackedConvBlocks(nn.Module):
    def __init__(self, num_convs, conv_op, input_channels, output_channels, kernel_size, initial_stride,
                 conv_bias=False, norm_op=None, norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None,
                 nonlin=None, nonlin_kwargs=None, nonlin_first=False, raw_input=False):
        super().__init__()
        if not isinstance(output_channels, (tuple, list)):
            output_channels = [output_channels] * num_convs
        common = (conv_bias, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs, nonlin_first)
        first = StemConv if raw_input else ConvDropoutNormReLU
        self.convs = nn.Sequential(
            first(conv_op, input_channels, output_channels[0], kernel_size, initial_stride, *common),
            *[ConvDropoutNormReLU(conv_op, output_channels[i - 1], output_channels[i], kernel_size, 1, *common)
              for i in range(1, num_convs)])
        self.output_channels = output_channels[-1]
        self.initial_stride = maybe_convert_scalar_to_list(conv_op, initial_stride)

    def forward(self, x, x_cat=None, head=None):
        """`head` (weight, bias, activation): the task head applied to the last conv unit's output (decoder tail)."""
        last = len(self.convs) - 1
        for i, blk in enumerate(self.convs):
            hd = head if i == last else None
            x = blk(x, x_cat, head=hd) if i == 0 else blk(x, head=hd)
        return x

    def compute_conv_feature_map_size(self, input_size):
        assert len(input_size) == len(self.initial_stride), "give the spatial size only, e.g. (x, y, z)"
        out = self.convs[0].compute_conv_feature_map_size(input_size)
        after = [i // j for i, j in zip(input_size, self.initial_stride)]
        for b in self.convs[1:]:
            out += b.compute_conv_feature_map_size(after)
        return out
