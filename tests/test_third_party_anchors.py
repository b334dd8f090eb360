"""SqueezeExcite / DropPath live in PyPI `dynamic-network-architectures` (timm-derived; reference call sites
builders/resblocks.py:11,81,86-87,109-112), which is neither vendored by the reference nor installed in this image, so
`oracle/dna_shim` restates them and the goldens of the SE / stochastic-depth fixtures are generated through that shim
(parity UNPINNED).  The closest independent implementations available offline are the copies of the same
timm algorithms inside `transformers` (RegNet's SE layer, ConvNeXt's drop_path, MobileNetV2's make_divisible): this
file anchors the shim - and through it the fixtures - on those.  It narrows the gap; it does not replace a check against
DNA itself."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "dna_shim"))

transformers = pytest.importorskip("transformers")


def _shim():
    from dynamic_network_architectures.building_blocks import regularization as R
    return R


def test_squeeze_excite_matches_an_independent_se_layer():
    """squeeze (global average pool) -> 1x1 conv -> ReLU -> 1x1 conv -> sigmoid -> scale: the shim's 3-D module against
    transformers' RegNetSELayer (2-D) on the same weights, the volume folded to [N, C, D*H, W]."""
    from transformers.models.regnet.modeling_regnet import RegNetSELayer
    R = _shim()
    torch.manual_seed(0)
    c = 64
    se = R.SqueezeExcite(c, torch.nn.Conv3d, rd_ratio=1. / 16, rd_divisor=8).double()
    rd = se.fc1.out_channels
    ref = RegNetSELayer(c, rd).double()
    with torch.no_grad():
        ref.attention[0].weight.copy_(se.fc1.weight.reshape(rd, c, 1, 1))
        ref.attention[0].bias.copy_(se.fc1.bias)
        ref.attention[2].weight.copy_(se.fc2.weight.reshape(c, rd, 1, 1))
        ref.attention[2].bias.copy_(se.fc2.bias)
    x = torch.randn(3, c, 4, 6, 5, dtype=torch.float64)
    old = R.SE_REDUCE_DIMS
    R.SE_REDUCE_DIMS = "all"
    try:
        mine = se(x)
    finally:
        R.SE_REDUCE_DIMS = old
    theirs = ref(x.reshape(3, c, 24, 5)).reshape_as(x)
    assert torch.allclose(mine, theirs, rtol=1e-12, atol=1e-14)


def test_reduced_channel_rule_matches_make_divisible():
    """timm's SE: rd_channels = make_divisible(channels * rd_ratio, 8, round_limit=0.) - the rounding to a multiple of 8
    with floor 8 is MobileNetV2's make_divisible wherever that one's 10 % rule does not fire; the drop-in's own module
    builds the same fc1 / fc2 shapes."""
    from transformers.models.mobilenet_v2.modeling_mobilenet_v2 import make_divisible as md
    R = _shim()
    import resenc_b200 as rb
    for c in (32, 64, 96, 128, 256, 320, 512, 1024):
        v = c / 16
        mine = R.make_divisible(v, 8, round_limit=0.)
        base = max(8, int(v + 4) // 8 * 8)
        assert mine == base
        if base >= 0.9 * v:
            assert mine == md(v, 8)
        se = rb.builders.resblocks.SqueezeExcite(c, torch.nn.Conv3d, rd_ratio=1. / 16, rd_divisor=8)
        assert se.fc1.out_channels == mine and se.fc2.in_channels == mine and se.fc2.out_channels == c


def test_drop_path_matches_the_timm_definition():
    """Per-sample Bernoulli(keep) mask, survivors scaled by 1 / keep, identity in eval mode: the shim, the drop-in's
    DropPath.factor and transformers' copy of timm's drop_path agree on the set of values and on the keep rate."""
    from transformers.models.convnext.modeling_convnext import drop_path
    R = _shim()
    import resenc_b200 as rb
    p, n = 0.3, 8192
    keep = 1 - p
    x = torch.ones(n, 2, 1, 1, 1)
    torch.manual_seed(1)
    dp = R.DropPath(p)
    dp.train()
    a = dp(x)[:, 0, 0, 0, 0]
    b = drop_path(x, p, training=True)[:, 0, 0, 0, 0]
    mine = rb.builders.resblocks.DropPath(p)
    mine.train()
    f = mine.factor(n, "cpu")
    for t in (a, b, f):
        vals = set(round(float(v), 6) for v in t.unique())
        assert vals == {0.0, round(1 / keep, 6)}
        rate = float((t > 0).float().mean())
        assert abs(rate - keep) < 4 * (keep * p / n) ** 0.5            # 4 sigma of the binomial
    # same generator state -> the drop-in draws exactly the shim's mask (the fixtures replay these draws)
    torch.manual_seed(2)
    a = dp(x)[:, 0, 0, 0, 0]
    torch.manual_seed(2)
    f = mine.factor(n, "cpu")
    assert torch.equal(a, f)
    dp.eval(); mine.eval()
    assert torch.equal(dp(x), x) and mine.factor(n, "cpu") is None
    assert torch.equal(drop_path(x, p, training=False), x)
