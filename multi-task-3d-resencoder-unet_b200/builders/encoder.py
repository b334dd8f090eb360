"""Encoder — drop-in for the reference's builders/encoder.py:27-170 (same constructor signature,
attributes read by the decoder :134-146, and state_dict keys).

Differences that are deliberate:
  * block-type strings are compared by value (the reference uses `is`, encoder.py:74-79, which
    only works for interned literals); the reference's precedence is kept: `basic_block` decides
    when it names a block, `bottleneck_block` only matters otherwise.
  * the stem takes the raw NCDHW fp32 network input (any channel count) and emits the internal
    channels-last bf16 layout; every stage consumes and produces that layout.
"""
from __future__ import annotations

import numpy as np
from torch import nn

from .resblocks import BasicBlockD, BottleneckD, StackedResidualBlocks
from .simple_conv_blocks import StackedConvBlocks, _unsupported
from .utils import maybe_convert_scalar_to_list


class Encoder(nn.Module):
    def __init__(self, input_channels, basic_block, n_stages, features_per_stage, n_blocks_per_stage, conv_op,
                 strides, kernel_sizes, conv_bias, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin,
                 nonlin_kwargs, do_stem=True, stem_channels=None, squeeze_excitation=False,
                 squeeze_excitation_reduction_ratio=1. / 16, stochastic_depth_p=0.0, return_skips=False,
                 bottleneck_block=BasicBlockD, pool_type='conv', bottleneck_channels=None, n_conv_per_stage=None):
        super().__init__()
        if isinstance(kernel_sizes, int):
            kernel_sizes = [kernel_sizes] * n_stages
        if isinstance(features_per_stage, int):
            features_per_stage = [features_per_stage] * n_stages
        if isinstance(n_blocks_per_stage, int):
            n_blocks_per_stage = [n_blocks_per_stage] * n_stages
        if isinstance(strides, int):
            strides = [strides] * n_stages
        if bottleneck_channels is None or isinstance(bottleneck_channels, int):
            bottleneck_channels = [bottleneck_channels] * n_stages
        if pool_type != 'conv':
            _unsupported(f"pool_type={pool_type!r} (only strided-conv downsampling)")

        residual = basic_block in ('BasicBlockD', 'BottleneckBlockD')
        block = None
        if bottleneck_block == 'BottleneckBlockD' or bottleneck_block is BottleneckD:
            block = BottleneckD
        if basic_block == 'BasicBlockD':
            block = BasicBlockD
        if basic_block == 'ConvBlock':
            block = None
        if residual and block is None:
            raise ValueError(f"basic_block={basic_block!r} needs bottleneck_block='BottleneckBlockD'")

        raw = True     # the first conv of the network reads the raw NCDHW fp32 input
        if do_stem:
            if stem_channels is None:
                stem_channels = features_per_stage[0]
            self.stem = StackedConvBlocks(1, conv_op, input_channels, stem_channels, kernel_sizes[0], 1, conv_bias,
                                          norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs,
                                          raw_input=True)
            input_channels = stem_channels
            raw = False
        else:
            self.stem = None
        if raw and input_channels % 8 != 0:
            _unsupported("do_stem=False with an input channel count that is not a multiple of 8")

        stages = []
        for s in range(n_stages):
            if residual:
                stages.append(StackedResidualBlocks(
                    n_blocks_per_stage[s], conv_op, input_channels, features_per_stage[s], kernel_sizes[s], strides[s],
                    conv_bias, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs,
                    block=block, bottleneck_channels=bottleneck_channels[s], stochastic_depth_p=stochastic_depth_p,
                    squeeze_excitation=squeeze_excitation,
                    squeeze_excitation_reduction_ratio=squeeze_excitation_reduction_ratio))
            else:
                stages.append(nn.Sequential(StackedConvBlocks(
                    n_blocks_per_stage[s], conv_op, input_channels, features_per_stage[s], kernel_sizes[s], strides[s],
                    conv_bias, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs)))
            input_channels = features_per_stage[s]

        # what a decoder needs to know
        self.stages = nn.Sequential(*stages)
        self.output_channels = features_per_stage
        self.strides = [maybe_convert_scalar_to_list(conv_op, i) for i in strides]
        self.return_skips = return_skips
        self.conv_op = conv_op
        self.norm_op = norm_op
        self.norm_op_kwargs = norm_op_kwargs
        self.nonlin = nonlin
        self.nonlin_kwargs = nonlin_kwargs
        self.dropout_op = dropout_op
        self.dropout_op_kwargs = dropout_op_kwargs
        self.conv_bias = conv_bias
        self.kernel_sizes = kernel_sizes

    def forward(self, x):
        if self.stem is not None:
            x = self.stem(x)
        skips = []
        for stage in self.stages:
            x = stage(x)
            skips.append(x)
        return skips if self.return_skips else skips[-1]

    def compute_conv_feature_map_size(self, input_size):
        out = self.stem.compute_conv_feature_map_size(input_size) if self.stem is not None else np.int64(0)
        for s, stage in enumerate(self.stages):
            inner = stage[0] if isinstance(stage, nn.Sequential) and not hasattr(stage, "compute_conv_feature_map_size") else stage
            out += inner.compute_conv_feature_map_size(input_size)
            input_size = [i // j for i, j in zip(input_size, self.strides[s])]
        return out
