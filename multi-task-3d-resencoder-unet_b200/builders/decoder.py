"""Decoder — drop-in for the reference's builders/decoder.py:16-193 (one instance per task).

Per stage: ConvTranspose3d(k == stride) -> cat(up, skip) -> conv block(s); the last stage feeds
the 1x1x1 head.  On the B200 path the transposed conv is a GEMM with a pixel-shuffle store, the
concatenation is never materialised (the conv kernel reads `up` and `skip` as two K segments) and
the head writes NCDHW fp32 logits directly.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .. import ops
from .resblocks import StackedResidualBlocks
from .simple_conv_blocks import StackedConvBlocks, _unsupported
from .utils import get_matching_convtransp


class Decoder(nn.Module):
    def __init__(self, encoder, basic_block, num_classes, n_conv_per_stage, deep_supervision, nonlin_first=False,
                 norm_op=None, norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None, nonlin=None,
                 nonlin_kwargs=None, conv_bias=None):
        super().__init__()
        self.deep_supervision = deep_supervision
        self.encoder = encoder
        self.num_classes = num_classes
        n_enc = len(encoder.output_channels)
        if isinstance(n_conv_per_stage, int):
            n_conv_per_stage = [n_conv_per_stage] * (n_enc - 1)
        assert len(n_conv_per_stage) == n_enc - 1, \
            "n_conv_per_stage must have one entry per resolution stage below the top (n_stages - 1), here: %d" % n_enc
        if basic_block not in ('ResidualBlock', 'ConvBlock'):
            raise ValueError(f"basic_decoder_block must be 'ConvBlock' or 'ResidualBlock', got {basic_block!r}")
        if num_classes > 8:
            _unsupported("more than 8 output channels per task head")

        transp = get_matching_convtransp(conv_op=encoder.conv_op)
        conv_bias = encoder.conv_bias if conv_bias is None else conv_bias
        norm_op = encoder.norm_op if norm_op is None else norm_op
        norm_op_kwargs = encoder.norm_op_kwargs if norm_op_kwargs is None else norm_op_kwargs
        dropout_op = encoder.dropout_op if dropout_op is None else dropout_op
        dropout_op_kwargs = encoder.dropout_op_kwargs if dropout_op_kwargs is None else dropout_op_kwargs
        nonlin = encoder.nonlin if nonlin is None else nonlin
        nonlin_kwargs = encoder.nonlin_kwargs if nonlin_kwargs is None else nonlin_kwargs

        stages, ups, heads = [], [], []
        for s in range(1, n_enc):
            below = encoder.output_channels[-s]
            skip = encoder.output_channels[-(s + 1)]
            stride = encoder.strides[-s]
            # the reference passes encoder.conv_bias here for residual decoders and the (possibly
            # overridden) conv_bias for plain ones (decoder.py:76,112)
            up_bias = encoder.conv_bias if basic_block == 'ResidualBlock' else conv_bias
            ups.append(transp(below, skip, stride, stride, bias=up_bias))
            if basic_block == 'ResidualBlock':
                stages.append(StackedResidualBlocks(
                    n_blocks=n_conv_per_stage[s - 1], conv_op=encoder.conv_op, input_channels=2 * skip,
                    output_channels=skip, kernel_size=encoder.kernel_sizes[-(s + 1)], initial_stride=1,
                    conv_bias=conv_bias, norm_op=norm_op, norm_op_kwargs=norm_op_kwargs, dropout_op=dropout_op,
                    dropout_op_kwargs=dropout_op_kwargs, nonlin=nonlin, nonlin_kwargs=nonlin_kwargs))
            else:
                stages.append(StackedConvBlocks(
                    n_conv_per_stage[s - 1], encoder.conv_op, 2 * skip, skip, encoder.kernel_sizes[-(s + 1)], 1,
                    conv_bias, norm_op, norm_op_kwargs, dropout_op, dropout_op_kwargs, nonlin, nonlin_kwargs,
                    nonlin_first))
            # one head per stage is always built so checkpoints stay loadable (decoder.py:128-131);
            # only the last is used unless deep supervision is on
            heads.append(encoder.conv_op(skip, num_classes, 1, 1, 0, bias=True))

        self.stages = nn.ModuleList(stages)
        self.transpconvs = nn.ModuleList(ups)
        self.seg_layers = nn.ModuleList(heads)

    def _upsample(self, s, x):
        t = self.transpconvs[s]
        return ops.conv_transpose3d(x, t.weight, t.stride, bias=t.bias)

    def forward(self, skips, activation=None):
        """`skips` in encoder order (bottleneck last).  `activation` ("sigmoid" | "softmax" | None) is
        fused into the head kernel for the full-resolution output (eval mode of NetworkFromConfig)."""
        low = skips[-1]
        outs = []
        last = len(self.stages) - 1
        for s in range(len(self.stages)):
            up = self._upsample(s, low)
            if s == last and not self.deep_supervision and isinstance(self.stages[s], StackedConvBlocks):
                # the full-resolution activation feeds the head and nothing else: the last conv unit applies the head
                # itself (in inference as one pass that never stores the activation, ops.conv_norm_act_head)
                h = self.seg_layers[-1]
                outs.append(self.stages[s](up, skips[-(s + 2)], head=(h.weight, h.bias, activation)))
                break
            low = self.stages[s](up, skips[-(s + 2)])
            if self.deep_supervision:
                h = self.seg_layers[s]
                outs.append(ops.head_conv1x1(low, h.weight, h.bias, activation if s == last else None))
            elif s == last:
                h = self.seg_layers[-1]
                outs.append(ops.head_conv1x1(low, h.weight, h.bias, activation))
        outs = outs[::-1]
        return outs if self.deep_supervision else outs[0]

    def compute_conv_feature_map_size(self, input_size):
        """`input_size` is the ENCODER input size."""
        skip_sizes = []
        for s in range(len(self.encoder.strides) - 1):
            skip_sizes.append([i // j for i, j in zip(input_size, self.encoder.strides[s])])
            input_size = skip_sizes[-1]
        assert len(skip_sizes) == len(self.stages)
        out = np.int64(0)
        for s in range(len(self.stages)):
            size = skip_sizes[-(s + 1)]
            out += self.stages[s].compute_conv_feature_map_size(size)
            out += np.prod([self.encoder.output_channels[-(s + 2)], *size], dtype=np.int64)
            if self.deep_supervision or (s == (len(self.stages) - 1)):
                out += np.prod([self.num_classes, *size], dtype=np.int64)
        return out
