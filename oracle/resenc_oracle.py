"""ORACLE — test infrastructure only.  NEVER imported by the product package.

CPU/fp32 restatement of the reference hot path (bruniss/multi-task-3d-resencoder-unet).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / the timed CPU baseline.

Pinned against the reference's own code (imported in the build container through
`oracle/reference_loader.py`) by `oracle/make_golden.py` -> `tests/golden/*.npz`, and by
`tests/test_oracle_vs_golden.py`.  The SqueezeExcite / DropPath arithmetic comes from an
un-vendored, un-pinned third-party package (PyPI dynamic-network-architectures) and is
therefore "parity unpinned" (SURVEY.md section 8c); every other function below cites the reference
file:line it restates.

Everything here is written functionally over a flat ``state_dict`` with the reference's key
names, so that it shares no structure with the product's nn.Module mirror.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# topology  (builders/utils.py:334-445, builders/build_network_from_config.py:39-70)
# --------------------------------------------------------------------------------------
def pool_and_conv_props(patch_size, min_feature_map_size=4, max_numpool=999999):
    """builders/utils.py:334-402 with spacing == (1,1,1) (the only value the reference passes,
    build_network_from_config.py:49).  Returns (num_pool_per_axis, strides, kernel_sizes)."""
    dim = len(patch_size)
    size = [int(s) for s in patch_size]
    spacing = [1.0] * dim
    strides = [tuple([1] * dim)]
    kernels = []
    npool = [0] * dim
    ksz = [1] * dim
    while True:
        valid = [i for i in range(dim) if size[i] >= 2 * min_feature_map_size]
        if not valid:
            break
        mn = min(spacing[i] for i in valid)
        valid = [i for i in valid if spacing[i] / mn < 2]
        valid = [i for i in valid if npool[i] < max_numpool]
        if not valid:
            break
        for d in range(dim):
            if ksz[d] != 3 and spacing[d] / min(spacing) < 2:
                ksz[d] = 3
        pk = [1] * dim
        for v in valid:
            pk[v] = 2
            npool[v] += 1
            spacing[v] *= 2
            size[v] = int(math.ceil(size[v] / 2))
        strides.append(tuple(pk))
        kernels.append(tuple(ksz))
    kernels.append(tuple([3] * dim))
    return npool, tuple(strides), tuple(kernels)


def blocks_per_stage(n):
    """builders/utils.py:428-445."""
    return [1 if i == 0 else 3 if i == 1 else 4 if i == 2 else 6 for i in range(n)]


def autoconfig(patch_size):
    """build_network_from_config.py:39-70."""
    _, strides, kernels = pool_and_conv_props(patch_size)
    n = len(strides)
    return SimpleNamespace(
        n_stages=n,
        features=[min(32 * 2 ** i, 512) for i in range(n)],
        n_blocks=blocks_per_stage(n),
        strides=strides,
        kernels=kernels,
        n_conv_dec=[1] * (n - 1),
    )


def manual_topology(model_config):
    """build_network_from_config.py:81-148: the manual (autoconfigure=False) branch reads these keys verbatim."""
    mc = model_config
    return SimpleNamespace(
        n_stages=int(mc["num_stages"]),
        features=list(mc["features_per_stage"]),
        n_blocks=list(mc["n_blocks_per_stage"]),
        strides=tuple(tuple(s) for s in mc["strides"]),
        kernels=tuple(tuple(k) for k in mc["kernel_sizes"]),
        n_conv_dec=list(mc["n_conv_per_stage_decoder"]),
    )


def se_rd_channels(c, ratio=1.0 / 16, divisor=8):
    """timm/DNA make_divisible(c*ratio, 8, round_limit=0.) — parity unpinned."""
    v = c * ratio
    return max(divisor, int(v + divisor / 2) // divisor * divisor)


# --------------------------------------------------------------------------------------
# functional network forward over a reference-keyed state_dict
# --------------------------------------------------------------------------------------
def _inorm(x, sd, key, eps=1e-5):
    """nn.InstanceNorm3d(affine=False|True, eps=1e-5, track_running_stats=False)
    (build_network_from_config.py:172).  Biased variance, same in train and eval."""
    w = sd.get(key + ".weight")
    b = sd.get(key + ".bias")
    return F.instance_norm(x, None, None, w, b, True, 0.0, eps)


def _cdnr(x, sd, p, stride, norm=True, act=True):
    """ConvDropoutNormReLU.forward (simple_conv_blocks.py:43-72): conv(pad=(k-1)//2) ->
    dropout(p=0) -> InstanceNorm -> LeakyReLU(0.01)."""
    w = sd[p + ".conv.weight"]
    pad = [(k - 1) // 2 for k in w.shape[2:]]
    x = F.conv3d(x, w, sd.get(p + ".conv.bias"), stride=stride, padding=pad)
    if norm:
        x = _inorm(x, sd, p + ".norm")
    if act:
        x = F.leaky_relu(x, 0.01)
    return x


def _se(x, sd, p, reduce_dims):
    """DNA SqueezeExcite (not in the reference repo; call site resblocks.py:86-87,111-112)."""
    dims = (2, 3, 4) if reduce_dims == "all" else tuple(reduce_dims)
    s = x.mean(dims, keepdim=True)
    s = F.conv3d(s, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"])
    s = F.relu(s)
    s = F.conv3d(s, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"])
    return x * torch.sigmoid(s)


def _skip(x, sd, p, stride, cin, cout):
    """BasicBlockD.skip (resblocks.py:89-104): AvgPool(stride) if strided, then 1x1x1
    conv + norm (no act) if channels change; identity otherwise."""
    has_stride = any(s != 1 for s in stride)
    proj = cin != cout
    idx = 0
    if has_stride:
        x = F.avg_pool3d(x, stride, stride)
        idx = 1
    if proj:
        x = _cdnr(x, sd, f"{p}.skip.{idx}", 1, norm=True, act=False)
    return x


def _drop_path(o, drop, p):
    """DNA DropPath (not in the reference repo; call sites resblocks.py:79-81,109-110,234-235): per-sample factor
    0 or 1/keep.  `drop` maps block prefixes to the [N] factor the caller drew (so both sides share the draw)."""
    if drop is None or p not in drop:
        return o
    return o * drop[p].to(o.dtype).view(-1, *([1] * (o.dim() - 1)))


def _basic_block(x, sd, p, stride, cin, cout, se, reduce_dims, drop=None):
    """BasicBlockD.forward (resblocks.py:106-114)."""
    r = _skip(x, sd, p, stride, cin, cout)
    o = _cdnr(x, sd, p + ".conv1", stride)
    o = _cdnr(o, sd, p + ".conv2", 1, act=False)
    o = _drop_path(o, drop, p)
    if se:
        o = _se(o, sd, p + ".squeeze_excitation", reduce_dims)
    return F.leaky_relu(o + r, 0.01)


def _bottleneck_block(x, sd, p, stride, cin, cout, se, reduce_dims, drop=None):
    """BottleneckD.forward (resblocks.py:231-239)."""
    r = _skip(x, sd, p, stride, cin, cout)
    o = _cdnr(x, sd, p + ".conv1", 1)
    o = _cdnr(o, sd, p + ".conv2", stride)
    o = _cdnr(o, sd, p + ".conv3", 1, act=False)
    o = _drop_path(o, drop, p)
    if se:
        o = _se(o, sd, p + ".squeeze_excitation", reduce_dims)
    return F.leaky_relu(o + r, 0.01)


def net_forward(sd, topo, x, tasks, training=True, se=False, reduce_dims="all",
                block="basic", residual_encoder=True, residual_decoder=False, drop=None):
    """NetworkFromConfig.forward (build_network_from_config.py:312-326) = Encoder.forward
    (encoder.py:148-158) + one Decoder.forward per task (decoder.py:137-162).

    `sd` uses the reference's key names; `topo` is `autoconfig(...)` or an equivalent
    namespace; `tasks` is {name: {"channels": c, "activation": str}}."""
    e = "shared_encoder"
    x = _cdnr(x, sd, f"{e}.stem.convs.0", 1)
    skips = []
    cin = topo.features[0]
    for s in range(topo.n_stages):
        cout = topo.features[s]
        for b in range(topo.n_blocks[s]):
            st = topo.strides[s] if b == 0 else (1, 1, 1)
            if residual_encoder:
                p = f"{e}.stages.{s}.blocks.{b}"
                fn = _basic_block if block == "basic" else _bottleneck_block
                x = fn(x, sd, p, st, cin, cout, se, reduce_dims, drop if training else None)
            else:
                x = _cdnr(x, sd, f"{e}.stages.{s}.0.convs.{b}", st)
            cin = cout
        skips.append(x)
    out = {}
    for name, info in tasks.items():
        d = f"task_decoders.{name}"
        low = skips[-1]
        nst = topo.n_stages - 1
        for s in range(nst):
            stride = topo.strides[-(s + 1)]
            up = F.conv_transpose3d(low, sd[f"{d}.transpconvs.{s}.weight"],
                                    sd.get(f"{d}.transpconvs.{s}.bias"), stride=stride)
            cat = torch.cat((up, skips[-(s + 2)]), 1)
            c_skip = topo.features[-(s + 2)]
            if residual_decoder:
                for b in range(topo.n_conv_dec[s]):
                    cat = _basic_block(cat, sd, f"{d}.stages.{s}.blocks.{b}", (1, 1, 1),
                                       2 * c_skip if b == 0 else c_skip, c_skip, False, reduce_dims)
            else:
                for b in range(topo.n_conv_dec[s]):
                    cat = _cdnr(cat, sd, f"{d}.stages.{s}.convs.{b}", 1)
            low = cat
        logits = F.conv3d(low, sd[f"{d}.seg_layers.{nst - 1}.weight"], sd[f"{d}.seg_layers.{nst - 1}.bias"])
        act = str(info.get("activation", "none")).lower()
        if not training:
            if act == "sigmoid":
                logits = torch.sigmoid(logits)
            elif act == "softmax":
                logits = torch.softmax(logits, 1)
        out[name] = logits
    return out


# --------------------------------------------------------------------------------------
# losses used by the headline training config (training/losses/losses.py)
# --------------------------------------------------------------------------------------
def bce_dice_loss(logits, target, alpha=0.5, beta=0.5, smoothing=0.1, eps=1e-6):
    """BCEDiceLoss (losses.py:307-318) = alpha*BCEWithLogits(label-smoothed, :217-238) +
    beta*(1 - mean per-channel Dice of sigmoid(logits), :17-43,115-126)."""
    t = target * (1.0 - 2.0 * smoothing) + smoothing   # losses.py:234
    bce = F.binary_cross_entropy_with_logits(logits, t)
    p = torch.sigmoid(logits)
    c = p.shape[1]
    pf = p.transpose(0, 1).reshape(c, -1)
    tf = target.float().transpose(0, 1).reshape(c, -1)
    inter = (pf * tf).sum(-1)
    den = (pf * pf).sum(-1) + (tf * tf).sum(-1)
    dice = 2 * (inter / den.clamp(min=eps))
    return alpha * bce + beta * (1.0 - dice.mean())


def masked_cosine_loss(pred, target):
    """MaskedCosineLoss.forward (losses.py:191-215)."""
    mask = (torch.norm(target, dim=1) > 1e-6).float()
    pu = pred / torch.norm(pred, dim=1, keepdim=True).clamp(min=1e-8)
    cs = F.cosine_similarity(pu, target, dim=1, eps=1e-8) * mask
    return 1.0 - cs.sum() / (mask.sum() + 1e-8)


# --------------------------------------------------------------------------------------
# sliding window: enumeration, importance map, accumulate / finalise / cast
# --------------------------------------------------------------------------------------
def positions_1d(lo, hi, patch, step):
    """helpers.py:200-216 generate_positions (raises IndexError when hi-lo < patch, as the
    reference does at :213)."""
    pos = []
    p = lo
    while p + patch <= hi:
        pos.append(p)
        p += step
    last = hi - patch
    if last > pos[-1]:
        pos.append(last)
    return sorted(set(pos))


def all_positions(vol_shape, patch, overlap):
    """inference_dataset.py:38-56: step = int(round(p*(1-overlap))) (Python banker's
    rounding of a double product), z-major nested order."""
    steps = [int(round(p * (1 - overlap))) for p in patch]
    zs, ys, xs = (positions_1d(0, vol_shape[i], patch[i], steps[i]) for i in range(3))
    return [(z, y, x) for z in zs for y in ys for x in xs]


def _gauss_kernel1d(sigma, radius):
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def gaussian_map(tile):
    """inference/helpers.py:8-68 compute_gaussian_3d (sigma_scale=1/8, scaling 1): scipy
    gaussian_filter (truncate 4 sigma, mode constant/0, fp32 array => one fp32 rounding per
    axis pass) of a centred delta, peak rescaled to 1, zeros replaced by the min positive."""
    d = [int(t) for t in tile]
    acc = None
    for ax, n in enumerate(d):
        sigma = n * (1.0 / 8)
        radius = int(4.0 * sigma + 0.5)
        w = _gauss_kernel1d(sigma, radius)
        c = n // 2
        line = np.zeros(n, np.float64)
        for i in range(n):
            off = i - c
            if -radius <= off <= radius:
                line[i] = w[radius - off]   # symmetric kernel, so the flip is immaterial
        line32 = line.astype(np.float32)
        shp = [1, 1, 1]
        shp[ax] = n
        if acc is None:
            acc = line32.reshape(shp)
        else:
            # the next pass multiplies the fp32 plane by the double weights, rounds to fp32
            acc = (acc.astype(np.float64) * line.reshape(shp)).astype(np.float32)
    g = np.ascontiguousarray(np.broadcast_to(acc, d)).astype(np.float32)
    g /= (g.max() / 1.0)
    mn = g[g > 0].min()
    g[g == 0] = mn
    return g


def blend_reference(preds, positions, vol_shape, targets):
    """inference.py:135-157 restated over numpy arrays standing in for the zarr datasets.

    preds: {target: [n_patches, c, pz, py, px] fp32 (already activated, :124-133)}.
    Returns ({target: sum}, {target: count}) with c==1 squeezed (:143-144)."""
    sums, counts = {}, {}
    for t, info in targets.items():
        c = info["channels"]
        shp = tuple(vol_shape) if c == 1 else (c,) + tuple(vol_shape)
        sums[t] = np.zeros(shp, np.float32)
        counts[t] = np.zeros(tuple(vol_shape), np.float32)
    for i, (z0, y0, x0) in enumerate(positions):
        for t, info in targets.items():
            p = preds[t][i]
            if info["channels"] == 1 and p.shape[0] == 1:
                p = np.squeeze(p, 0)
            pz, py, px = p.shape[-3:]
            sums[t][..., z0:z0 + pz, y0:y0 + py, x0:x0 + px] += p
            counts[t][z0:z0 + pz, y0:y0 + py, x0:x0 + px] += 1
    return sums, counts


def finalize_reference(sums, counts, targets):
    """inference.py:166-210 (finalise) + :213-263 (cast), whole-volume instead of chunk-wise
    (the arithmetic is elementwise so chunking is immaterial).  Target named "normals" with
    c==3: vector re-normalise, no division by count; else sum/count where count>0.
    Returns {target: uint8|uint16 array}."""
    out = {}
    for t, info in targets.items():
        s = sums[t].copy()
        cnt = counts[t]
        mask = cnt > 0
        if t.lower() == "normals":
            if info["channels"] == 3:
                mag = np.sqrt(s[0] ** 2 + s[1] ** 2 + s[2] ** 2) + np.float32(1e-8)
                for k in range(3):
                    s[k][mask] /= mag[mask]
            v = (s + np.float32(1.0)) / np.float32(2.0)
            v *= np.float32(65535.0)
            np.clip(v, 0, 65535, out=v)
            out[t] = v.astype(np.uint16)
        else:
            s[..., mask] /= cnt[mask]
            v = s * np.float32(255.0)
            np.clip(v, 0, 255, out=v)
            out[t] = v.astype(np.uint8)
    return out


def blend_weighted_reference(preds, positions, vol_shape, targets, weight):
    """Gaussian-weighted variant the north star asks for (the reference ships the map,
    inference/helpers.py:8-68, but never applies it): sum += w*pred, wsum += w."""
    sums, wsum = {}, {}
    for t, info in targets.items():
        c = info["channels"]
        shp = tuple(vol_shape) if c == 1 else (c,) + tuple(vol_shape)
        sums[t] = np.zeros(shp, np.float32)
        wsum[t] = np.zeros(tuple(vol_shape), np.float32)
    for i, (z0, y0, x0) in enumerate(positions):
        for t, info in targets.items():
            p = preds[t][i]
            if info["channels"] == 1 and p.shape[0] == 1:
                p = np.squeeze(p, 0)
            pz, py, px = p.shape[-3:]
            sums[t][..., z0:z0 + pz, y0:y0 + py, x0:x0 + px] += p * weight
            wsum[t][z0:z0 + pz, y0:y0 + py, x0:x0 + px] += weight
    return sums, wsum


def standardize_patch(patch):
    """pytorch3dunet Standardize(channelwise=False) as used by inference_dataset.py:62-75
    (third-party, un-vendored => parity unpinned): (m - mean) / clip(std, 1e-10)."""
    m = patch.astype(np.float32)
    return (m - m.mean()) / np.clip(m.std(), 1e-10, None)
