"""Host-side topology helpers of the drop-in `builders` package.

Semantics follow the reference's builders/utils.py (pooling schedule :334-402, pad_shape :405-426,
blocks per stage :428-445, operator matching :128-285); the arithmetic is integer-only and is
checked against the reference's own outputs in tests/golden/host_goldens.json.
"""
from __future__ import annotations

import math

import numpy as np
from torch import nn

_CONV_DIM = {nn.Conv1d: 1, nn.Conv2d: 2, nn.Conv3d: 3}


def convert_conv_op_to_dim(conv_op):
    try:
        return _CONV_DIM[conv_op]
    except KeyError:
        raise ValueError("Unknown dimension. Only 1d 2d and 3d conv are supported. got %s" % str(conv_op))


def convert_dim_to_conv_op(dimension):
    for op, d in _CONV_DIM.items():
        if d == dimension:
            return op
    raise ValueError("Unknown dimension. Only 1, 2 and 3 are supported")


def _dim_of(conv_op, dimension):
    assert not ((conv_op is not None) and (dimension is not None)), \
        "You MUST set EITHER conv_op OR dimension. Do not set both!"
    if conv_op is not None:
        dimension = convert_conv_op_to_dim(conv_op)
    assert dimension in (1, 2, 3), "Dimension must be 1, 2 or 3"
    return dimension


def get_matching_pool_op(conv_op=None, dimension=None, adaptive=False, pool_type="avg"):
    assert pool_type in ("avg", "max"), "pool_type must be either avg or max"
    d = _dim_of(conv_op, dimension)
    name = ("Adaptive" if adaptive else "") + ("Avg" if pool_type == "avg" else "Max") + f"Pool{d}d"
    return getattr(nn, name)


def get_matching_instancenorm(conv_op=None, dimension=None):
    return getattr(nn, f"InstanceNorm{_dim_of(conv_op, dimension)}d")


def get_matching_batchnorm(conv_op=None, dimension=None):
    return getattr(nn, f"BatchNorm{_dim_of(conv_op, dimension)}d")


def get_matching_convtransp(conv_op=None, dimension=None):
    return getattr(nn, f"ConvTranspose{_dim_of(conv_op, dimension)}d")


def get_matching_dropout(conv_op=None, dimension=None):
    d = _dim_of(conv_op, dimension)
    return nn.Dropout if d == 1 else getattr(nn, f"Dropout{d}d")


def maybe_convert_scalar_to_list(conv_op, scalar):
    """kernel_size=3 -> [3, 3, 3] for nn.Conv3d; sequences pass through unchanged."""
    if isinstance(scalar, (tuple, list, np.ndarray)):
        return scalar
    try:
        return [scalar] * _CONV_DIM[conv_op]
    except KeyError:
        raise RuntimeError("Invalid conv op: %s" % str(conv_op))


def pad_shape(shape, must_be_divisible_by):
    """Round every extent up to the next multiple (already-divisible extents are kept)."""
    if not isinstance(must_be_divisible_by, (tuple, list, np.ndarray)):
        must_be_divisible_by = [must_be_divisible_by] * len(shape)
    assert len(must_be_divisible_by) == len(shape)
    return tuple(int(s) + (-int(s)) % int(m) for s, m in zip(shape, must_be_divisible_by))


def get_pool_and_conv_props(spacing, patch_size, min_feature_map_size, max_numpool):
    """Pooling / kernel schedule: halve every axis that is still >= 2*min_feature_map_size and whose
    spacing is within 2x of the finest poolable axis; an axis gets kernel 3 once its spacing is
    within 2x of the finest.  Returns (num_pool_per_axis, strides, kernels, padded_patch, divisor)."""
    dim = len(spacing)
    cur_spacing = [float(s) for s in spacing]
    cur_size = [int(s) for s in patch_size]
    strides = [tuple([1] * dim)]
    kernels = []
    npool = [0] * dim
    ksize = [1] * dim
    while True:
        ok = [i for i in range(dim) if cur_size[i] >= 2 * min_feature_map_size]
        if not ok:
            break
        finest = min(cur_spacing[i] for i in ok)
        ok = [i for i in ok if cur_spacing[i] / finest < 2 and npool[i] < max_numpool]
        if not ok:
            break
        for a in range(dim):
            if ksize[a] != 3 and cur_spacing[a] / min(cur_spacing) < 2:
                ksize[a] = 3
        step = [1] * dim
        for a in ok:
            step[a] = 2
            npool[a] += 1
            cur_spacing[a] *= 2
            cur_size[a] = int(math.ceil(cur_size[a] / 2))
        strides.append(tuple(step))
        kernels.append(tuple(ksize))
    divisor = 2 ** np.array(npool)
    kernels.append(tuple([3] * dim))
    return npool, tuple(strides), tuple(kernels), tuple(pad_shape(patch_size, divisor)), divisor


def get_n_blocks_per_stage(num_stages):
    """1, 3, 4 then 6 residual blocks per stage."""
    table = (1, 3, 4)
    return [table[i] if i < len(table) else 6 for i in range(num_stages)]
