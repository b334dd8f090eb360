"""ORACLE SCAFFOLDING — regenerates tests/golden/* from the UNMODIFIED reference code.

Run in the build container only (`python oracle/make_golden.py`); needs /root/reference.
The fixtures are small: network weights are not stored, they are re-derived on any machine
from `seeded_state()` (numpy PCG64, platform independent), only inputs/outputs/losses/gradient
digests are.  tests/test_oracle_vs_golden.py replays them through oracle/resenc_oracle.py
(CPU) and tests/test_gpu_parity.py through the CUDA product path.
"""
import hashlib
import json
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import reference_loader as rl          # noqa: E402
from oracle import resenc_oracle as O              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

NET_CASES = {
    # name: (patch, in_channels, tasks, model_config, se_reduce_dims, batch)
    "sheet_normals_16": ([16, 16, 16], 1,
                         {"sheet": {"channels": 1, "activation": "sigmoid"},
                          "normals": {"channels": 3, "activation": "none"}}, {}, "all", 2),
    "ink_se_16": ([16, 16, 16], 4, {"ink": {"channels": 1, "activation": "sigmoid"}},
                  {"squeeze_excitation": True, "conv_bias": True}, "all", 2),
    "ink_se23_16": ([16, 16, 16], 4, {"ink": {"channels": 1, "activation": "sigmoid"}},
                    {"squeeze_excitation": True}, (2, 3), 1),
    "aniso_8x32x32": ([8, 32, 32], 1, {"sheet": {"channels": 2, "activation": "softmax"}}, {}, "all", 1),
    "affine_16": ([16, 16, 16], 2, {"sheet": {"channels": 1, "activation": "none"}},
                  {"norm_op_kwargs": {"affine": True, "eps": 1e-5}}, "all", 1),
    # stochastic depth (DropPath, resblocks.py:79-81,109-110) with SE behind it; the per-block draws of the
    # reference run are stored in the fixture as `drop::<block>` so that every implementation replays them
    "droppath_se_16": ([16, 16, 16], 1, {"sheet": {"channels": 1, "activation": "sigmoid"}},
                       {"squeeze_excitation": True, "stochastic_depth_p": 0.3}, "all", 4),
    # --- round 2: block families / decoder variants that only a manual config reaches ---
    # BottleneckD (resblocks.py:135-239): 1x1 -> k^3 (stride) -> 1x1, bottleneck_channels = features // 4
    "bottleneck_16": ([16, 16, 16], 1,
                      {"sheet": {"channels": 1, "activation": "sigmoid"},
                       "normals": {"channels": 3, "activation": "none"}},
                      dict(features_per_stage=[32, 64, 128], num_stages=3, n_blocks_per_stage=[1, 2, 2],
                           kernel_sizes=[[3, 3, 3]] * 3, n_conv_per_stage_decoder=[1, 1],
                           strides=[[1, 1, 1], [2, 2, 2], [2, 2, 2]], basic_encoder_block="BottleneckBlockD",
                           bottleneck_block="BottleneckBlockD", basic_decoder_block="ConvBlock"), "all", 2),
    # residual decoder stages (decoder.py:68-95): StackedResidualBlocks on cat(up, skip), two blocks in the first stage
    "resdec_16": ([16, 16, 16], 1,
                  {"sheet": {"channels": 1, "activation": "sigmoid"},
                   "normals": {"channels": 3, "activation": "none"}},
                  dict(features_per_stage=[32, 64, 128], num_stages=3, n_blocks_per_stage=[1, 2, 2],
                       kernel_sizes=[[3, 3, 3]] * 3, n_conv_per_stage_decoder=[2, 1],
                       strides=[[1, 1, 1], [2, 2, 2], [2, 2, 2]], basic_encoder_block="BasicBlockD",
                       bottleneck_block="BasicBlockD", basic_decoder_block="ResidualBlock"), "all", 2),
    # default autoconfigured network at a geometry that selects the fast kernels (slab fprop / dgrad at W = 64,
    # weights-on-M h-major tiles, two-sided tap-stacked wgrad) with gradients; eval outputs are not stored (the
    # network has no train / eval difference besides the head activation) to keep the fixture small
    "default_32x64x64": ([32, 64, 64], 1,
                         {"sheet": {"channels": 1, "activation": "sigmoid"},
                          "normals": {"channels": 3, "activation": "none"}}, {}, "all", 1),
}
# per-case switches that do not fit the 6-tuple: autoconfigure flag, which oracle block family, what is stored
NET_EXTRA = {
    "bottleneck_16": {"autoconfigure": False, "block": "bottleneck"},
    "resdec_16": {"autoconfigure": False, "residual_decoder": True},
    "default_32x64x64": {"store_eval": False},
}
DROP_SEED = 4321


def oracle_kwargs(case):
    """Arguments of resenc_oracle.net_forward that select the block family / decoder variant of `case`."""
    ex = NET_EXTRA.get(case, {})
    return {"block": ex.get("block", "basic"), "residual_decoder": bool(ex.get("residual_decoder", False))}


def oracle_topology(case):
    patch, _, _, mc, _, _ = NET_CASES[case]
    if NET_EXTRA.get(case, {}).get("autoconfigure", True):
        return O.autoconfig(patch)
    return O.manual_topology(mc)


def seeded_state(named_shapes, seed):
    """Deterministic, platform-independent weights for a list of (name, shape)."""
    out = {}
    for name, shape in named_shapes:
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        shape = tuple(shape)
        if len(shape) > 1:
            fan_in = int(np.prod(shape[1:]))
            if "transpconvs" in name:
                fan_in = int(shape[0])
            b = 1.0 / np.sqrt(fan_in)
            a = rng.uniform(-b, b, size=shape)
        elif name.endswith("norm.weight"):
            a = 1.0 + rng.uniform(-0.2, 0.2, size=shape)
        else:
            a = rng.uniform(-0.1, 0.1, size=shape)
        out[name] = torch.from_numpy(a.astype(np.float32))
    return out


def seeded_inputs(case, seed=1234):
    patch, cin, tasks, _, _, batch = NET_CASES[case]
    rng = np.random.default_rng([seed, zlib.crc32(case.encode())])
    x = rng.random((batch, cin, *patch), dtype=np.float32)
    tgt = {}
    for t, info in tasks.items():
        c = info["channels"]
        if t == "normals":
            v = rng.standard_normal((batch, c, *patch)).astype(np.float32)
            v /= np.linalg.norm(v, axis=1, keepdims=True)
            tgt[t] = v
        else:
            tgt[t] = (rng.random((batch, c, *patch)) > 0.8).astype(np.float32)
    return x, tgt


def loss_for(task, pred, target):
    if task == "normals":
        return O.masked_cosine_loss(pred, target)
    return O.bce_dice_loss(pred, target)


def unique_named_params(model):
    return [(n, tuple(p.shape)) for n, p in model.named_parameters()]


def make_net_case(case):
    patch, cin, tasks, mc, rd, batch = NET_CASES[case]
    extra = NET_EXTRA.get(case, {})
    model = rl.build_reference(rl.make_mgr(patch, tasks, in_channels=cin, batch=batch, model_config=mc,
                                           autoconfigure=extra.get("autoconfigure", True)), se_reduce_dims=rd)
    names = unique_named_params(model)
    st = seeded_state(names, seed=7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.copy_(st[n])
    x, tgt = seeded_inputs(case)
    xt = torch.from_numpy(x)
    model.train()
    # record the stochastic-depth draws (factor per sample = 0 or 1 / keep) of every residual block
    drops = {}

    def _rec(name):
        def hook(mod, inp, outp):
            if not mod.training or name in drops:
                return
            a, b = inp[0].detach().flatten(1).abs().sum(1), outp.detach().flatten(1).abs().sum(1)
            f = torch.where(b > 0, torch.full_like(b, 1.0 / (1.0 - mod.drop_prob)), torch.zeros_like(b))
            assert torch.allclose(a * f, b, rtol=1e-4)
            drops.setdefault(name, f.numpy().astype(np.float32))
        return hook
    for n_, m_ in model.named_modules():
        if type(m_).__name__ == "DropPath":
            m_.register_forward_hook(_rec(n_[:-len(".drop_path")]))
    torch.manual_seed(DROP_SEED)
    out = model(xt)
    # reference losses through the reference's own loss classes (training/losses/losses.py)
    losses_mod = rl.reference_module("training/losses/losses.py", "ref_losses")
    total = 0.0
    per = {}
    for t in tasks:
        fn = losses_mod.MaskedCosineLoss() if t == "normals" else losses_mod.BCEDiceLoss(0.5, 0.5)
        l = fn(out[t], torch.from_numpy(tgt[t]))
        per[t] = float(l)
        total = total + l
    total.backward()
    grads = {n: p.grad for n, p in model.named_parameters()}
    gnorm = np.array([0.0 if grads[n] is None else float(grads[n].double().norm()) for n, _ in names])
    has_grad = np.array([grads[n] is not None for n, _ in names])
    model.eval()
    with torch.no_grad():
        ev = model(xt)
    # calibration of the bf16 tolerance: PyTorch's own bf16 autocast of the UNMODIFIED reference vs its fp32 output
    model.train()
    for p_ in model.parameters():
        p_.grad = None
    torch.manual_seed(DROP_SEED)      # same stochastic-depth draws as the fp32 pass
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ac = model(xt)
    ac_total = 0.0
    for t in tasks:
        fn = losses_mod.MaskedCosineLoss() if t == "normals" else losses_mod.BCEDiceLoss(0.5, 0.5)
        ac_total = ac_total + fn(ac[t].float(), torch.from_numpy(tgt[t]))
    ac_total.backward()
    ac_grads = {n: p_.grad for n, p_ in model.named_parameters()}
    ac = {t: v.detach() for t, v in ac.items()}
    model.eval()
    rec = {"x": x, "loss_total": np.float64(float(total)), "grad_norms": gnorm, "has_grad": has_grad,
           "param_names": np.array([n for n, _ in names])}
    # a few full gradients (first conv, a deep conv, a transposed conv, a head)
    picks = [n for n, _ in names if n.endswith("stem.convs.0.conv.weight")
             or n.endswith("stages.1.blocks.0.conv1.conv.weight")
             or n.endswith("transpconvs.0.weight") or "seg_layers" in n and n.endswith(".1.weight")
             or n.endswith("squeeze_excitation.fc1.weight") and ".stages.1.blocks.0." in n]
    for n in picks:
        if grads[n] is not None and grads[n].numel() <= 200000:
            rec["grad::" + n] = grads[n].numpy()
            if ac_grads[n] is not None:
                rec["autocast_bf16_gradrel::" + n] = np.float64(
                    float((ac_grads[n].float() - grads[n]).norm() / grads[n].norm()))
    gmax = float(gnorm.max())
    devs = [abs(float(ac_grads[n].double().norm()) - g) / g for (n, _), g in zip(names, gnorm)
            if ac_grads[n] is not None and g >= 1e-4 * gmax]
    rec["autocast_bf16_gradnorm_dev"] = np.float64(max(devs))
    for k_, v_ in drops.items():
        rec["drop::" + k_] = v_
    for t in tasks:
        rec["target::" + t] = tgt[t]
        rec["train::" + t] = out[t].detach().numpy()
        if extra.get("store_eval", True):
            rec["eval::" + t] = ev[t].numpy()
        rec["loss::" + t] = np.float64(per[t])
        a, b = ac[t].float(), out[t].detach()
        rec["autocast_bf16_rel::" + t] = np.float64(float((a - b).norm() / b.norm()))
    np.savez_compressed(os.path.join(GOLD, f"net_{case}.npz"), **rec)
    # state-dict key census (drop-in contract)
    keys = {k: list(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(GOLD, f"keys_{case}.json"), "w") as f:
        json.dump({"state_dict": keys, "parameters": [[n, list(s)] for n, s in names]}, f)
    print(case, "loss", float(total), {t: per[t] for t in per}, "params", len(names), "keys", len(keys))


def sha16(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def make_host_goldens():
    gen = rl.reference_function("helpers.py", "generate_positions")
    pos_cases = [((1024,) * 3, (128,) * 3, 0.5), ((1000,) * 3, (128,) * 3, 0.5),
                 ((300, 200, 130), (128,) * 3, 0.25), ((2000, 1500, 1750), (64, 192, 192), 0.05),
                 ((256,) * 3, (128,) * 3, 0.1), ((128, 128, 128), (128,) * 3, 0.5),
                 ((130, 257, 129), (64, 128, 64), 0.5), ((96, 96, 96), (32, 32, 32), 0.3)]
    pos = []
    for vol, patch, ov in pos_cases:
        # inference_dataset.py:44-56 executed literally with the reference's generate_positions
        steps = [int(round(p * (1 - ov))) for p in patch]
        axes = [gen(0, vol[i], patch[i], steps[i]) for i in range(3)]
        table = np.array([(z, y, x) for z in axes[0] for y in axes[1] for x in axes[2]], dtype=np.int64)
        pos.append({"vol": list(vol), "patch": list(patch), "overlap": ov, "steps": steps,
                    "axes": [list(map(int, a)) for a in axes], "count": int(len(table)),
                    "sha1": sha16(table)})
    ih = rl.reference_module("inference/helpers.py", "ref_inf_helpers")
    gau = []
    for tile in [(128, 128, 128), (64, 64, 64), (64, 192, 192), (32, 32, 32), (16, 24, 40), (14, 256, 256)]:
        g = ih.compute_gaussian_3d(tile).numpy()
        gau.append({"tile": list(tile), "sha1": sha16(g), "sum": float(g.astype(np.float64).sum()),
                    "min": float(g.min()), "max": float(g.max()),
                    "argmax": [int(i) for i in np.unravel_index(g.argmax(), g.shape)],
                    "edge_center": float(g[0, tile[1] // 2, tile[2] // 2])})
    np.save(os.path.join(GOLD, "gaussian_16x24x40.npy"), ih.compute_gaussian_3d((16, 24, 40)).numpy())
    topo = []
    for patch in [(64, 64, 64), (96, 96, 96), (128, 128, 128), (192, 192, 192), (14, 256, 256), (16, 16, 16),
                  (8, 32, 32), (32, 64, 160)]:
        um = rl.reference_module("builders/utils.py", "ref_utils")
        npa, strides, kernels, final, div = um.get_pool_and_conv_props((1.0, 1.0, 1.0), list(patch), 4, 999999)
        topo.append({"patch": list(patch), "num_pool": [int(v) for v in npa],
                     "strides": [list(map(int, s)) for s in strides],
                     "kernels": [list(map(int, k)) for k in kernels],
                     "final_patch": [int(v) for v in final],
                     "blocks": um.get_n_blocks_per_stage(len(strides))})
    with open(os.path.join(GOLD, "host_goldens.json"), "w") as f:
        json.dump({"positions": pos, "gaussian": gau, "topology": topo}, f, indent=1)
    print("host goldens:", len(pos), "position cases,", len(gau), "gaussian,", len(topo), "topology")


def make_loss_goldens():
    """Values and gradient norms of the reference trainer's loss table (train.py:47-56 over
    training/losses/losses.py) on seeded inputs -> tests/golden/loss_goldens.json."""
    L = rl.reference_module("training/losses/losses.py", "ref_losses")
    rng = np.random.default_rng(77)
    logits = torch.from_numpy(rng.standard_normal((2, 2, 5, 6, 7)).astype(np.float32) * 2)
    target = torch.from_numpy((rng.random((2, 2, 5, 6, 7)) > 0.7).astype(np.float32))
    vec_p = torch.from_numpy(rng.standard_normal((2, 3, 5, 6, 7)).astype(np.float32))
    vec_t = rng.standard_normal((2, 3, 5, 6, 7)).astype(np.float32)
    vec_t /= np.linalg.norm(vec_t, axis=1, keepdims=True)
    vec_t[:, :, :2] = 0                                # masked-out voxels
    vec_t = torch.from_numpy(vec_t)
    cases = [("BCEDiceLoss", {"alpha": 0.5, "beta": 0.5}, "bin"), ("BCEDiceLoss", {"alpha": 0.3, "beta": 1.0}, "bin"),
             ("BCEWithLogitsLossLabelSmoothing", {}, "bin"), ("BCEWithLogitsLossLabelSmoothing", {"smoothing": 0.25}, "bin"),
             ("BCEWithLogitsLossZSmooth", {}, "bin"), ("BCEWithLogitsLossZSmooth", {"center_smoothing": 0.05, "edge_smoothing": 0.3}, "bin"),
             ("MaskedCosineLoss", {}, "vec")]
    out = []
    for name, kw, kind in cases:
        fn = getattr(L, name)(**kw)
        x = (logits if kind == "bin" else vec_p).clone().requires_grad_(True)
        t = target if kind == "bin" else vec_t
        v = fn(x, t)
        v.backward()
        out.append({"loss_fn": name, "loss_kwargs": kw, "kind": kind, "value": float(v),
                    "grad_norm": float(x.grad.double().norm()), "grad_sum": float(x.grad.double().sum())})
    with open(os.path.join(GOLD, "loss_goldens.json"), "w") as f:
        json.dump({"seed": 77, "cases": out}, f, indent=1)
    print("loss goldens:", [(c["loss_fn"], round(c["value"], 5)) for c in out])


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:]                  # `python oracle/make_golden.py <case> ...` regenerates just those cases
    if not only:
        make_host_goldens()
        make_loss_goldens()
    if only == ["losses"]:
        make_loss_goldens()
        only = ["__none__"]
    for c in (only or NET_CASES):
        if c in NET_CASES:
            make_net_case(c)
