"""Residual blocks — drop-in for the reference's builders/resblocks.py (BasicBlockD :15-132,
BottleneckD :135-259, StackedResidualBlocks :262-353) plus the two pieces the reference imports
from the un-vendored `dynamic_network_architectures` package (SqueezeExcite, DropPath;
call sites resblocks.py:9-11,79-87,109-112).

Block forward on the B200 path (one kernel sequence, no elementwise pass of its own):
    r  = skip(x)                      identity | AvgPool [-> 1x1x1 conv -> IN]
    z1 = LReLU(IN(conv1(x)))          one fused unit (ops.conv_norm_act)
    out = LReLU( SE(IN(conv2(z1))) + r )   one fused unit: the last conv's normalise pass also applies the
                                      gate, the residual add and the activation
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .. import ops
from .simple_conv_blocks import ConvDropoutNormReLU, _unsupported
from .utils import get_matching_pool_op, maybe_convert_scalar_to_list

# How SqueezeExcite pools: "all" = global average over (D, H, W) (what the task statement names);
# (2, 3) = the timm 2-D code path applied verbatim to 5-D input (mean over D, H; one gate per W).
# The upstream package is neither vendored nor pinned by the reference, so this is configurable.
SE_REDUCE_DIMS = "all"


def make_divisible(v, divisor=8, min_value=None, round_limit=.9):
    min_value = min_value or divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < round_limit * v:
        new_v += divisor
    return new_v


class SqueezeExcite(nn.Module):
    """Parameter container + configuration of the SE gate; the arithmetic is fused into the
    block-tail kernel (ops.instance_norm_se_act)."""

    def __init__(self, channels, conv_op, rd_ratio=1. / 16, rd_channels=None, rd_divisor=8, add_maxpool=False,
                 act_layer=nn.ReLU, norm_layer=None, gate_layer=nn.Sigmoid, reduce_dims=None):
        super().__init__()
        if add_maxpool or norm_layer is not None or act_layer is not nn.ReLU or gate_layer is not nn.Sigmoid:
            _unsupported("SqueezeExcite variants other than mean -> fc -> ReLU -> fc -> sigmoid")
        if not rd_channels:
            rd_channels = make_divisible(channels * rd_ratio, rd_divisor, round_limit=0.)
        self.fc1 = conv_op(channels, rd_channels, kernel_size=1, bias=True)
        self.bn = nn.Identity()
        self.act = act_layer(inplace=True)
        self.fc2 = conv_op(rd_channels, channels, kernel_size=1, bias=True)
        self.gate = gate_layer()
        self.reduce_dims = reduce_dims

    def dims(self):
        return SE_REDUCE_DIMS if self.reduce_dims is None else self.reduce_dims

    def forward(self, x):
        raise RuntimeError("SqueezeExcite is fused into the residual block tail; it is not called on its own")


class DropPath(nn.Module):
    """Stochastic depth (DNA `DropPath`, reference call sites resblocks.py:79-81,109-110): in training every sample
    of the batch keeps its residual branch with probability 1 - drop_prob (scaled by 1 / keep).  The per-sample
    factor is O(N); it is folded into the per-(n, c) scale / shift of the block-tail kernel, so the module is
    never called on a full tensor."""

    def __init__(self, drop_prob=0., scale_by_keep=True):
        super().__init__()
        if not 0.0 <= float(drop_prob) <= 1.0:
            raise ValueError(f"drop_prob must lie in [0, 1], got {drop_prob}")
        self.drop_prob = float(drop_prob)
        self.scale_by_keep = scale_by_keep
        self.forced_factor = None      # parity tests replay the draws of a reference run through this

    def factor(self, n, device):
        """[n] fp32 tensor of 0 / (1/keep) drawn from torch's generator of `device`, or None (eval, p = 0)."""
        if self.drop_prob == 0. or not self.training:
            return None
        if self.forced_factor is not None:
            return self.forced_factor.to(device=device, dtype=torch.float32)
        keep = 1.0 - self.drop_prob
        f = torch.empty(n, dtype=torch.float32, device=device).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            f.div_(keep)
        return f

    def forward(self, x):
        raise RuntimeError("DropPath is fused into the residual block tail; it is not called on its own")


class _ResidualBlock(nn.Module):
    """Shared tail / skip logic of BasicBlockD and BottleneckD."""

    def _finish_init(self, conv_op, input_channels, output_channels, stride, norm_op, norm_op_kwargs, nonlin,
                     nonlin_kwargs, stochastic_depth_p, squeeze_excitation, rd_ratio, act_name):
        if nonlin is None:
            _unsupported("a residual block without nonlinearity")
        setattr(self, act_name, nonlin(**nonlin_kwargs))
        self._slope = float(getattr(self, act_name).negative_slope)
        self.apply_stochastic_depth = stochastic_depth_p != 0.0
        if self.apply_stochastic_depth:
            self.drop_path = DropPath(drop_prob=stochastic_depth_p)
        self.apply_se = squeeze_excitation
        if self.apply_se:
            self.squeeze_excitation = SqueezeExcite(self.output_channels, conv_op, rd_ratio=rd_ratio, rd_divisor=8)
        has_stride = any(i != 1 for i in stride)
        projects = input_channels != output_channels
        self._pool_stride = tuple(stride) if has_stride else None
        self._proj_index = None
        if has_stride or projects:
            mods = []
            if has_stride:
                mods.append(get_matching_pool_op(conv_op=conv_op, adaptive=False, pool_type='avg')(stride, stride))
            if projects:
                self._proj_index = len(mods)
                mods.append(ConvDropoutNormReLU(conv_op, input_channels, output_channels, 1, 1, False, norm_op,
                                                norm_op_kwargs, None, None, None, None))
            self.skip = nn.Sequential(*mods)
        else:
            self.skip = lambda x: x

    def _residual(self, x):
        r = x
        if self._pool_stride is not None:
            r = ops.avg_pool3d(r, self._pool_stride)
        if self._proj_index is not None:
            r = self.skip[self._proj_index](r)
        return r

    def _tail(self, last, x, r):
        """LReLU( [SE]( IN( last.conv(x) ) ) + r ): the block's last conv carries the whole tail."""
        se, dims = None, "all"
        if self.apply_se:
            m = self.squeeze_excitation
            se, dims = (m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias), m.dims()
        drop = self.drop_path.factor(x.shape[0], x.device) if self.apply_stochastic_depth else None
        return last(x, res=r, act=True, slope=self._slope, se=se, se_reduce_dims=dims, drop=drop)


class BasicBlockD(_ResidualBlock):
    def __init__(self, conv_op, input_channels, output_channels, kernel_size, stride, conv_bias=False, norm_op=None,
                 norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None, nonlin=None, nonlin_kwargs=None,
                 stochastic_depth_p=0.0, squeeze_excitation=False, squeeze_excitation_reduction_ratio=1. / 16):
        super().__init__()
        self.input_channels = input_channels
        self.output_channels = output_channels
        stride = maybe_convert_scalar_to_list(conv_op, stride)
        self.stride = stride
        kernel_size = maybe_convert_scalar_to_list(conv_op, kernel_size)
        norm_op_kwargs = norm_op_kwargs or {}
        nonlin_kwargs = nonlin_kwargs or {}
        self.conv1 = ConvDropoutNormReLU(conv_op, input_channels, output_channels, kernel_size, stride, conv_bias,
                                         norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs)
        self.conv2 = ConvDropoutNormReLU(conv_op, output_channels, output_channels, kernel_size, 1, conv_bias,
                                         norm_op, norm_op_kwargs, None, None, None, None)
        self._finish_init(conv_op, input_channels, output_channels, stride, norm_op, norm_op_kwargs, nonlin,
                          nonlin_kwargs, stochastic_depth_p, squeeze_excitation, squeeze_excitation_reduction_ratio,
                          "nonlin2")

    def forward(self, x, x_cat=None):
        if x_cat is not None:
            _unsupported("a residual block on a virtual concatenation")
        r = self._residual(x)
        return self._tail(self.conv2, self.conv1(x), r)

    def compute_conv_feature_map_size(self, input_size):
        assert len(input_size) == len(self.stride), "give the spatial size only, e.g. (x, y, z)"
        after = [i // j for i, j in zip(input_size, self.stride)]
        one = np.prod([self.output_channels, *after], dtype=np.int64)
        has_skip = isinstance(self.skip, nn.Sequential)
        return one + one + (one if has_skip else 0)


class BottleneckD(_ResidualBlock):
    def __init__(self, conv_op, input_channels, bottleneck_channels, output_channels, kernel_size, stride,
                 conv_bias=False, norm_op=None, norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None,
                 nonlin=None, nonlin_kwargs=None, stochastic_depth_p=0.0, squeeze_excitation=False,
                 squeeze_excitation_reduction_ratio=1. / 16):
        super().__init__()
        self.input_channels = input_channels
        self.output_channels = output_channels
        self.bottleneck_channels = bottleneck_channels
        stride = maybe_convert_scalar_to_list(conv_op, stride)
        self.stride = stride
        kernel_size = maybe_convert_scalar_to_list(conv_op, kernel_size)
        norm_op_kwargs = norm_op_kwargs or {}
        nonlin_kwargs = nonlin_kwargs or {}
        if bottleneck_channels % 8 != 0:
            _unsupported(f"bottleneck_channels={bottleneck_channels} (must be a multiple of 8)")
        self.conv1 = ConvDropoutNormReLU(conv_op, input_channels, bottleneck_channels, 1, 1, conv_bias, norm_op,
                                         norm_op_kwargs, None, None, nonlin, nonlin_kwargs)
        self.conv2 = ConvDropoutNormReLU(conv_op, bottleneck_channels, bottleneck_channels, kernel_size, stride,
                                         conv_bias, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin,
                                         nonlin_kwargs)
        self.conv3 = ConvDropoutNormReLU(conv_op, bottleneck_channels, output_channels, 1, 1, conv_bias, norm_op,
                                         norm_op_kwargs, None, None, None, None)
        self._finish_init(conv_op, input_channels, output_channels, stride, norm_op, norm_op_kwargs, nonlin,
                          nonlin_kwargs, stochastic_depth_p, squeeze_excitation, squeeze_excitation_reduction_ratio,
                          "nonlin3")

    def forward(self, x, x_cat=None):
        if x_cat is not None:
            _unsupported("a residual block on a virtual concatenation")
        r = self._residual(x)
        return self._tail(self.conv3, self.conv2(self.conv1(x)), r)

    def compute_conv_feature_map_size(self, input_size):
        assert len(input_size) == len(self.stride), "give the spatial size only, e.g. (x, y, z)"
        after = [i // j for i, j in zip(input_size, self.stride)]
        c1 = np.prod([self.bottleneck_channels, *input_size], dtype=np.int64)
        c2 = np.prod([self.bottleneck_channels, *after], dtype=np.int64)
        c3 = np.prod([self.output_channels, *after], dtype=np.int64)
        return c1 + c2 + c3 + (c3 if isinstance(self.skip, nn.Sequential) else 0)


class StackedResidualBlocks(nn.Module):
    def __init__(self, n_blocks, conv_op, input_channels, output_channels, kernel_size, initial_stride,
                 conv_bias=False, norm_op=None, norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None,
                 nonlin=None, nonlin_kwargs=None, block=BasicBlockD, bottleneck_channels=None,
                 stochastic_depth_p=0.0, squeeze_excitation=False, squeeze_excitation_reduction_ratio=1. / 16):
        super().__init__()
        assert n_blocks > 0, 'n_blocks must be > 0'
        assert block in [BasicBlockD, BottleneckD], 'block must be BasicBlockD or BottleneckD'
        if not isinstance(output_channels, (tuple, list)):
            output_channels = [output_channels] * n_blocks
        if not isinstance(bottleneck_channels, (tuple, list)):
            bottleneck_channels = [bottleneck_channels] * n_blocks
        tail = (conv_bias, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs,
                stochastic_depth_p, squeeze_excitation, squeeze_excitation_reduction_ratio)
        mods = []
        for n in range(n_blocks):
            cin = input_channels if n == 0 else output_channels[n - 1]
            st = initial_stride if n == 0 else 1
            if block is BasicBlockD:
                mods.append(block(conv_op, cin, output_channels[n], kernel_size, st, *tail))
            else:
                mods.append(block(conv_op, cin, bottleneck_channels[n], output_channels[n], kernel_size, st, *tail))
        self.blocks = nn.Sequential(*mods)
        self.initial_stride = maybe_convert_scalar_to_list(conv_op, initial_stride)
        self.output_channels = output_channels[-1]

    def forward(self, x, x_cat=None):
        if x_cat is not None:
            # decoder stages built from residual blocks see the concatenation (decoder.py:147);
            # materialise it once, the block then reads it twice (conv1 and the projection skip)
            if ops.precise_active():
                # a channel concatenation of two [hi | lo | hi] rows is not a split row of the concatenation
                _unsupported("residual decoder stages in the split-precision inference tier")
            x = torch.cat((ops.as_cl(x), ops.as_cl(x_cat)), 1)
        for blk in self.blocks:
            x = blk(x)
        return x

    def compute_conv_feature_map_size(self, input_size):
        assert len(input_size) == len(self.initial_stride), "give the spatial size only, e.g. (x, y, z)"
        out = self.blocks[0].compute_conv_feature_map_size(input_size)
        after = [i // j for i, j in zip(input_size, self.initial_stride)]
        for b in self.blocks[1:]:
            out += b.compute_conv_feature_map_size(after)
        return out
