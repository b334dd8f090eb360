"""CPU tests of the product's host logic against the reference-generated goldens, the drop-in
surface (state_dict keys, constructor errors) and the C-ABI library's exports."""
import ctypes
import hashlib
import importlib
import json
import os
import re

import numpy as np
import pytest
import torch

from helpers import GOLDEN, NET_CASES, ROOT, case_mgr, load_keys, make_mgr, quiet_build

with open(os.path.join(GOLDEN, "host_goldens.json")) as f:
    HOST = json.load(f)


def sha16(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.mark.parametrize("rec", HOST["positions"], ids=lambda r: "x".join(map(str, r["vol"])))
def test_positions_bit_exact(rb, rec):
    inf = rb.inference
    assert inf.patch_steps(rec["patch"], rec["overlap"]) == rec["steps"]
    assert inf.axis_positions(rec["vol"], rec["patch"], rec["overlap"]) == rec["axes"]
    table = np.array(inf.all_positions(rec["vol"], rec["patch"], rec["overlap"]), dtype=np.int64)
    assert len(table) == rec["count"] and sha16(table) == rec["sha1"]


def test_positions_volume_smaller_than_patch(rb):
    with pytest.raises(ValueError):
        rb.inference.all_positions((100, 200, 200), (128, 128, 128), 0.5)


@pytest.mark.parametrize("rec", HOST["gaussian"], ids=lambda r: "x".join(map(str, r["tile"])))
def test_gaussian_bit_exact(rb, rec):
    g = rb.inference.compute_gaussian_3d(rec["tile"])
    assert g.dtype == np.float32 and sha16(g) == rec["sha1"]
    assert [int(i) for i in np.unravel_index(g.argmax(), g.shape)] == rec["argmax"]
    assert float(g.astype(np.float64).sum()) == pytest.approx(rec["sum"], rel=1e-12)


def test_gaussian_full_array(rb):
    ref = np.load(os.path.join(GOLDEN, "gaussian_16x24x40.npy"))
    assert np.array_equal(rb.inference.compute_gaussian_3d((16, 24, 40)), ref)


@pytest.mark.parametrize("rec", HOST["topology"], ids=lambda r: "x".join(map(str, r["patch"])))
def test_topology(rb, rec):
    u = rb.builders.utils if hasattr(rb.builders, "utils") else None
    from importlib import import_module
    u = import_module(rb.builders.__name__ + ".utils")
    npool, strides, kernels, final, div = u.get_pool_and_conv_props((1.0, 1.0, 1.0), rec["patch"], 4, 999999)
    assert [int(v) for v in npool] == rec["num_pool"]
    assert [list(s) for s in strides] == rec["strides"]
    assert [list(k) for k in kernels] == rec["kernels"]
    assert [int(v) for v in final] == rec["final_patch"]
    assert u.get_n_blocks_per_stage(len(strides)) == rec["blocks"]


@pytest.mark.parametrize("case", list(NET_CASES))
def test_state_dict_contract(rb, case):
    """Same keys, shapes and parameter order as the reference's NetworkFromConfig."""
    mgr, _ = case_mgr(case)
    model = quiet_build(rb.NetworkFromConfig, mgr)
    gold = load_keys(case)
    assert {k: list(v.shape) for k, v in model.state_dict().items()} == gold["state_dict"]
    assert [[n, list(p.shape)] for n, p in model.named_parameters()] == gold["parameters"]


def test_default_init_matches_reference_rng_stream(rb):
    """Construction order equals the reference's, so torch.manual_seed(s) gives the same weights:
    the fixture stores none, but two builds under one seed must agree and differ across seeds."""
    mgr, _ = case_mgr("sheet_normals_16")
    torch.manual_seed(3)
    a = quiet_build(rb.NetworkFromConfig, mgr).state_dict()
    torch.manual_seed(3)
    b = quiet_build(rb.NetworkFromConfig, mgr).state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_constructor_errors(rb):
    tasks = {"t": {"channels": 1, "activation": "none"}}
    with pytest.raises(ValueError):
        quiet_build(rb.NetworkFromConfig, make_mgr([16, 16, 16], tasks, autoconfigure=False, model_config={}))
    with pytest.raises(ValueError):
        quiet_build(rb.NetworkFromConfig, make_mgr([16, 16, 16], {"t": {"channels": 1, "activation": "tanh"}}))
    with pytest.raises(ValueError):
        quiet_build(rb.NetworkFromConfig, make_mgr([16], tasks))
    with pytest.raises(NotImplementedError):
        quiet_build(rb.NetworkFromConfig, make_mgr([16, 16], tasks))
    with pytest.raises(NotImplementedError):
        quiet_build(rb.NetworkFromConfig, make_mgr([16, 16, 16], tasks, model_config={"dropout_op_kwargs": {"p": 0.2}}))
    with pytest.raises(NotImplementedError):
        quiet_build(rb.NetworkFromConfig, make_mgr([16, 16, 16], tasks, model_config={"nonlin": "nn.ReLU"}))


def test_manual_config_variants(rb):
    """Manual configs with YAML-style (non-interned) strings: residual/bottleneck/plain encoders and a
    residual decoder build with the reference's key layout."""
    tasks = {"t": {"channels": 2, "activation": "softmax"}}
    base = dict(features_per_stage=[32, 64, 128], num_stages=3, n_blocks_per_stage=[1, 2, 2],
                kernel_sizes=[[3, 3, 3]] * 3, n_conv_per_stage_decoder=[1, 1], strides=[[1, 1, 1], [2, 2, 2], [2, 2, 2]])
    for enc, bott, dec, probe in [
        ("".join(["Basic", "BlockD"]), "BasicBlockD", "ConvBlock", "shared_encoder.stages.1.blocks.1.conv2.conv.weight"),
        ("BottleneckBlockD", "BottleneckBlockD", "ConvBlock", "shared_encoder.stages.1.blocks.0.conv3.conv.weight"),
        ("ConvBlock", "BasicBlockD", "ResidualBlock", "task_decoders.t.stages.0.blocks.0.conv1.conv.weight"),
        ("ResidualBlock", "BasicBlockD", "ConvBlock", "shared_encoder.stages.2.0.convs.1.conv.weight"),
    ]:
        mc = dict(base, basic_encoder_block=enc, bottleneck_block=bott, basic_decoder_block=dec)
        m = quiet_build(rb.NetworkFromConfig, make_mgr([16, 16, 16], tasks, autoconfigure=False, model_config=mc))
        assert probe in m.state_dict(), (enc, dec)


def test_cpu_tensor_fails_loudly(rb, built_lib):
    """No CPU fallback: a CPU input raises instead of silently running PyTorch."""
    mgr, _ = case_mgr("sheet_normals_16")
    model = quiet_build(rb.NetworkFromConfig, mgr)
    with pytest.raises(rb._lib.ResencLibraryError):
        model(torch.rand(1, 1, 16, 16, 16))


def test_c_abi_exports_every_declared_symbol(rb, built_lib):
    header = open(os.path.join(ROOT, "include", "resenc_b200.h")).read()
    declared = set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(rb._lib.SIGNATURES), (declared ^ set(rb._lib.SIGNATURES))
    lib = ctypes.CDLL(rb._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert built_lib.rb_version() >= 100
    assert built_lib.rb_launch_count() >= 0


def test_c_abi_rejects_bad_descriptors_without_gpu(rb, built_lib):
    """Argument validation happens on the host before any launch."""
    d = rb._lib.ConvDesc()
    rc = built_lib.rb_conv_gather(ctypes.byref(d), None, None, None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"nsrc" in built_lib.rb_last_error()
    assert built_lib.rb_plane_reduce(0, None, 0, None, None, None, None, None, 1, 8, 8, 1, 0, 0.01, None) == -1
    assert built_lib.rb_blend_finalize_cast(None, None, None, None, 8, 1, 0, None) == -1


def test_slab_exchange_plan(rb):
    inf = rb.inference
    zs = inf.generate_positions(0, 1024, 128, 64)
    runs = inf.shard_z_starts(zs, 8)
    assert [len(r) for r in runs] == [2, 2, 2, 2, 2, 2, 2, 1] and sum(runs, []) == zs
    pairs, own = inf.plan_slab_exchange(zs, 128, 1024, 8)
    assert own[0] == (0, 128) and own[-1] == (896, 1024)
    covered = sorted(own)
    assert covered[0][0] == 0 and covered[-1][1] == 1024
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    # every slab reaches 64 planes into its successor's range only
    assert all(dst == src + 1 and hi - lo == 64 for src, dst, lo, hi in pairs)
    # more ranks than z-starts: empty ranks own nothing and exchange nothing
    pairs2, own2 = inf.plan_slab_exchange([0, 64], 128, 192, 4)
    assert own2[2] == (0, 0) and own2[3] == (0, 0) and all(d < 2 for _, d, _, _ in pairs2)


@pytest.mark.parametrize("k,stride,dims", [((3, 3, 3), (2, 2, 2), (8, 8, 8)), ((1, 3, 3), (1, 2, 2), (4, 8, 8)),
                                           ((3, 3, 3), (1, 2, 2), (5, 4, 10)), ((1, 1, 1), (2, 2, 2), (4, 4, 4))])
def test_merged_strided_dgrad_pack(rb, k, stride, dims):
    """Host side of the one-launch strided data gradient (ops._merged_dgrad_plan / pack_conv_dgrad_merged): the
    zero-padded [window tap][(parity, ci)][co] weights, applied as a stride-1 gather over dy followed by a pixel
    shuffle, reproduce autograd's data gradient of the strided convolution (pure torch emulation, bf16 weights)."""
    import torch.nn.functional as F
    ops = rb.ops
    torch.manual_seed(0)
    co, ci = 16, 32
    w = torch.randn(co, ci, *k)
    x = torch.randn(2, ci, *dims, requires_grad=True)
    pad = tuple((kk - 1) // 2 for kk in k)
    y = F.conv3d(x, w.to(torch.bfloat16).float(), None, stride, pad)
    dy = torch.randn_like(y)
    y.backward(dy)
    od = tuple(y.shape[2:])
    axes = ops._merged_dgrad_plan(k, stride, pad, dims, od)
    assert axes is not None
    wp = ops.pack_conv_dgrad_merged(w, axes, stride).float()
    nt, off = [a[0] for a in axes], [a[1] for a in axes]
    lo = [max(0, -o) for o in off]
    dyp = F.pad(dy, (lo[2], nt[2], lo[1], nt[1], lo[0], nt[0]))
    out = torch.zeros(2, wp.shape[1], *od)
    t = 0
    for ud in range(nt[0]):
        for uh in range(nt[1]):
            for uw in range(nt[2]):
                s0 = [u + o + l for u, o, l in zip((ud, uh, uw), off, lo)]
                sl = dyp[:, :, s0[0]:s0[0] + od[0], s0[1]:s0[1] + od[1], s0[2]:s0[2] + od[2]]
                out += torch.einsum("bcdhw,nc->bndhw", sl, wp[t])
                t += 1
    o6 = out.view(2, stride[0], stride[1], stride[2], ci, *od)
    dx = o6.permute(0, 4, 5, 1, 6, 2, 7, 3).reshape(2, ci, *[o * s for o, s in zip(od, stride)])
    assert float((dx - x.grad).norm() / x.grad.norm()) < 1e-5
    # geometries that do not tile exactly keep the per-class path
    assert ops._merged_dgrad_plan((3, 3, 3), (2, 2, 2), (1, 1, 1), (7, 8, 8), (4, 4, 4)) is None
    assert ops._merged_dgrad_plan((3, 3, 3), (1, 1, 1), (1, 1, 1), (8, 8, 8), (8, 8, 8)) is None


def test_package_losses_match_oracle(rb):
    """losses.py of the package (what bench.py's GPU arm calls) against the oracle's restatement of
    training/losses/losses.py on random logits / targets, values and gradients."""
    import importlib
    from oracle import resenc_oracle as O   # checker only
    L = importlib.import_module(rb._pkg.__name__ + ".losses")
    torch.manual_seed(0)
    z = torch.randn(2, 1, 6, 7, 8, requires_grad=True)
    t = (torch.rand(2, 1, 6, 7, 8) > 0.8).float()
    a = L.BCEDiceLoss(0.5, 0.5)(z, t)
    a.backward()
    z2 = z.detach().clone().requires_grad_(True)
    b = O.bce_dice_loss(z2, t)
    b.backward()
    assert abs(float(a) - float(b)) < 1e-6 and torch.allclose(z.grad, z2.grad, atol=1e-8)
    n = torch.randn(2, 3, 6, 7, 8, requires_grad=True)
    tn = torch.nn.functional.normalize(torch.randn(2, 3, 6, 7, 8), dim=1)
    tn[:, :, :2] = 0          # masked-out voxels
    c = L.MaskedCosineLoss()(n, tn)
    c.backward()
    n2 = n.detach().clone().requires_grad_(True)
    d = O.masked_cosine_loss(n2, tn)
    d.backward()
    assert abs(float(c) - float(d)) < 1e-6 and torch.allclose(n.grad, n2.grad, atol=1e-8)
    crit = L.task_losses({"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}})
    assert isinstance(crit["sheet"], L.BCEDiceLoss) and isinstance(crit["normals"], L.MaskedCosineLoss)


def test_zarr_v2_final_writer_roundtrip(rb, tmp_path):
    """inference.FinalVolumeWriter: the reference's `<target>_final` layout (inference.py:213-263)
    as a zarr v2 directory - shapes / chunks / dtypes / fill value in .zarray, all-zero chunks not written, ragged
    edge chunks padded, z-ranges that do not align with the chunk grid merged without losing data."""
    import json
    inf = rb.inference
    targets = {"sheet": {"channels": 1}, "normals": {"channels": 3}}
    vol, patch = (37, 20, 26), (16, 8, 12)
    rng = np.random.default_rng(5)
    sheet = rng.integers(0, 256, vol, dtype=np.uint8)
    sheet[16:32, 8:16, 12:24] = 0                       # one all-zero chunk: must not exist on disk
    normals = rng.integers(0, 65536, (3, *vol), dtype=np.uint16)
    for comp in ("blosc", "zlib", None):
        root = str(tmp_path / f"out_{comp}.zarr")
        w = inf.FinalVolumeWriter(root, targets, vol, patch, compressor=comp, threads=4)
        for z0, z1 in [(0, 16), (16, 21), (21, 32), (32, 37)]:      # aligned, two partial pieces of one row, ragged tail
            w.submit(z0, {"sheet": torch.from_numpy(sheet[z0:z1]), "normals": normals[:, z0:z1]})
        w.close()
        assert json.load(open(os.path.join(root, ".zgroup"))) == {"zarr_format": 2}
        ms = json.load(open(os.path.join(root, "sheet_final", ".zarray")))
        mn = json.load(open(os.path.join(root, "normals_final", ".zarray")))
        assert ms["shape"] == list(vol) and ms["chunks"] == list(patch) and ms["dtype"] == "|u1" and ms["fill_value"] == 0
        assert mn["shape"] == [3, *vol] and mn["chunks"] == [3, *patch] and mn["dtype"] == "<u2" and mn["order"] == "C"
        assert (ms["compressor"] or {}).get("id") == comp
        assert not os.path.exists(os.path.join(root, "sheet_final", "1.1.1"))
        assert os.path.exists(os.path.join(root, "sheet_final", "2.2.2"))        # ragged corner chunk
        assert os.path.exists(os.path.join(root, "normals_final", "0.1.1.1"))
        assert np.array_equal(w.arrays["sheet"].read(), sheet)
        assert np.array_equal(w.arrays["normals"].read(), normals)
        # read back through the zarr-style reader (what DeviceVolume slices): whole array, z-ranges, sub-boxes, ints
        rs, rn = inf.open_zarr_array(root, "sheet_final"), inf.open_zarr_array(root, "normals_final")
        assert rs.shape == vol and rs.dtype == np.uint8 and rn.shape == (3, *vol) and rn.dtype == np.uint16
        assert np.array_equal(rs[:], sheet) and np.array_equal(rn[...], normals)
        assert np.array_equal(rs[5:33], sheet[5:33]) and np.array_equal(rs[17], sheet[17])
        assert np.array_equal(rn[..., 3:20, 7:19, 11:26], normals[..., 3:20, 7:19, 11:26])
        assert np.array_equal(rn[1, 30:, :, -3:], normals[1, 30:, :, -3:]) and rs[4:4].shape == (0, 20, 26)
        if comp is None:                                  # raw chunk = C-order bytes of the padded chunk
            raw = np.fromfile(os.path.join(root, "sheet_final", "2.2.2"), dtype=np.uint8).reshape(patch)
            assert np.array_equal(raw[:5, :4, :2], sheet[32:, 16:, 24:]) and not raw[5:].any()
        if comp == "blosc":                              # the reference's codec configuration (inference.py:92,224)
            assert ms["compressor"] == {"id": "blosc", "cname": "zstd", "clevel": 5, "shuffle": 2, "blocksize": 0}
            from importlib import import_module
            bc = import_module(inf.__name__ + ".blosc_codec")
            h = bc.header(open(os.path.join(root, "normals_final", "0.1.1.1"), "rb").read())
            assert h["typesize"] == 2 and h["nbytes"] == 3 * 16 * 8 * 12 * 2 and h["version"] == 2
    with pytest.raises(NotImplementedError):
        inf.FinalVolumeWriter(str(tmp_path / "x.zarr"), targets, vol, patch, compressor="lz4")


def _bitshuffle_literal(buf, ts):
    """bitshuffle's bshuf_trans_bit_elem spelled out bit by bit: out[(j*8 + b) * (n/8) + k] bit i = bit b of byte j of
    element 8k + i (blocks whose element count is not a multiple of 8 are copied, as c-blosc 1.x does)."""
    n = len(buf) // ts
    out = bytearray(buf)
    if n % 8:
        return bytes(out)                  # c-blosc 1.x copies such a block
    for j in range(ts):
        for b in range(8):
            for k in range(n // 8):
                v = 0
                for i in range(8):
                    v |= ((buf[(8 * k + i) * ts + j] >> b) & 1) << i
                out[(j * 8 + b) * (n // 8) + k] = v
    return bytes(out)


def test_blosc_zstd_bitshuffle_container(rb):
    """The reference's output codec, Blosc(cname='zstd', clevel=5, shuffle=BITSHUFFLE) (inference.py:92,224), written
    without libblosc: byte-level checks of the Blosc-1 header / block table, the bit-shuffle against a literal
    restatement, zstd frames that the system libzstd decodes, and round trips over the edge cases (empty, shorter
    than the 128-byte minimum, incompressible, ragged last block, typesize 1 / 2 / 4, every shuffle mode)."""
    import struct
    from importlib import import_module
    bc = import_module(rb.inference.__name__ + ".blosc_codec")
    rng = np.random.default_rng(11)
    # bit-shuffle == the literal definition, for element counts that are and are not multiples of 8
    for ts, n in ((1, 64), (2, 40), (4, 24), (4, 27), (2, 7)):
        raw = rng.integers(0, 256, ts * n, dtype=np.uint8)
        assert bc._bit_shuffle(raw, ts).tobytes() == _bitshuffle_literal(raw.tobytes(), ts)
        assert np.array_equal(bc._bit_unshuffle(bc._bit_shuffle(raw, ts), ts), raw)
        assert np.array_equal(bc._byte_unshuffle(bc._byte_shuffle(raw, ts), ts), raw)
    # a smooth uint16 volume compresses; header fields and block table are what c-blosc documents
    vol = (np.add.outer(np.arange(300), np.arange(500)) * 3).astype("<u2")
    buf = bc.compress(vol.tobytes(), typesize=2, clevel=5, shuffle=bc.BITSHUFFLE)
    ver, verlz, flags, ts, nbytes, bs, cbytes = struct.unpack("<BBBBiii", buf[:16])
    assert (ver, verlz, ts, nbytes, cbytes) == (2, 1, 2, vol.nbytes, len(buf)) and len(buf) < vol.nbytes // 4
    assert flags == 0x04 | (4 << 5)                     # bit-shuffle, codec zstd, blocks split per byte plane of the type
    nblocks = -(-nbytes // bs)
    bstarts = struct.unpack("<%di" % nblocks, buf[16:16 + 4 * nblocks])
    assert bstarts[0] == 16 + 4 * nblocks and list(bstarts) == sorted(bstarts) and nblocks > 1
    (cs0,) = struct.unpack("<i", buf[bstarts[0]:bstarts[0] + 4])
    (cs1,) = struct.unpack("<i", buf[bstarts[0] + 4 + cs0:bstarts[0] + 8 + cs0])
    assert bstarts[1] == bstarts[0] + 8 + cs0 + cs1      # two streams (typesize 2) per full block
    assert buf[bstarts[0] + 4:bstarts[0] + 8] == b"\x28\xb5\x2f\xfd"      # zstd frame magic
    shuffled = bc._bit_shuffle(np.frombuffer(vol.tobytes()[:bs], np.uint8), 2).tobytes()
    assert bc._zstd().decompress(buf[bstarts[0] + 4:bstarts[0] + 4 + cs0], bs // 2) == shuffled[:bs // 2]
    assert bc._zstd().decompress(buf[bstarts[0] + 8 + cs0:bstarts[0] + 8 + cs0 + cs1], bs // 2) == shuffled[bs // 2:]
    assert bc.decompress(buf) == vol.tobytes()
    # edge cases
    for data, ts in ((b"", 1), (bytes(range(100)), 1), (rng.integers(0, 256, 70001, dtype=np.uint8).tobytes(), 1),
                     (rng.integers(0, 256, 300000, dtype=np.uint8).tobytes(), 4), (bytes(1 << 20), 2),
                     (np.arange(100003, dtype="<u4").tobytes()[:-1], 4)):
        for sh in (bc.NOSHUFFLE, bc.SHUFFLE, bc.BITSHUFFLE):
            for blocksize in (0, 4096):
                b = bc.compress(data, typesize=ts, clevel=5, shuffle=sh, blocksize=blocksize)
                h = bc.header(b)
                assert h["nbytes"] == len(data) and h["cbytes"] == len(b) and len(b) <= len(data) + 16
                assert bc.decompress(b) == data
    assert bc.header(bc.compress(bytes(range(100)), 1))["memcpyed"]                     # < 128 bytes: stored
    assert bc.header(bc.compress(rng.integers(0, 256, 70001, dtype=np.uint8).tobytes(), 1))["memcpyed"]   # random: stored
    assert bc.header(bc.compress(bytes(1 << 20), 2, clevel=0))["memcpyed"]
    with pytest.raises(ValueError):
        bc.decompress(buf[:len(buf) // 2])


def test_weight_pack_cache_sees_fused_optimizer_updates(rb):
    """torch's fused optimisers update parameters without bumping Tensor._version; the pack cache must not serve the
    pre-update bf16 operand (found by the trainer test: eager fused-AdamW training otherwise learns on stale packs)."""
    ops = rb.ops
    w = torch.nn.Parameter(torch.randn(8, 8, 3, 3, 3))
    a = ops.pack_conv_fprop(w)
    assert ops.pack_conv_fprop(w) is a                                    # cached while nothing changes
    w.grad = torch.ones_like(w)
    opt = torch.optim.AdamW([w], lr=0.5, fused=True)
    v = w._version
    opt.step()
    b = ops.pack_conv_fprop(w)
    exp = w.detach().permute(2, 3, 4, 0, 1).reshape(27, 8, 8).to(torch.bfloat16)
    assert torch.equal(b, exp) and not torch.equal(a, b), f"stale pack (version {v} -> {w._version})"
    with torch.no_grad():
        w.add_(1.0)                                                       # ordinary in-place update: version bump
    assert torch.equal(ops.pack_conv_fprop(w), w.detach().permute(2, 3, 4, 0, 1).reshape(27, 8, 8).to(torch.bfloat16))


def test_split_precision_weight_rows(rb):
    """precise._pack3: [hi | hi | lo] rows reproduce the fp32 weight to 2^-16 relative when paired with activation
    rows [hi | lo | hi] (the dropped lo*lo term is second order), and the context manager restores the bf16 tier."""
    P = rb.precise
    torch.manual_seed(3)
    w = torch.randn(5, 7, 16) * 0.1
    pk = P._pack3(w)
    assert pk.dtype == torch.bfloat16 and tuple(pk.shape) == (5, 7, 48)
    hi, hi2, lo = pk[..., :16].float(), pk[..., 16:32].float(), pk[..., 32:].float()
    assert torch.equal(hi, hi2) and torch.equal(hi, w.to(torch.bfloat16).float())
    assert float(((hi + lo) - w).abs().max() / w.abs().max()) < 2.0 ** -16
    x = torch.randn(16)
    xh = x.to(torch.bfloat16).float()
    xl = (x - xh).to(torch.bfloat16).float()
    row = torch.cat((xh, xl, xh))
    got = (pk.float() * row).sum(-1)                       # hi*hi_w + lo*hi_w + hi*lo_w
    ref = (w.double() * x.double()).sum(-1)
    assert float((got.double() - ref).abs().max() / ref.abs().max()) < 1e-4
    plain = (w.to(torch.bfloat16).float() * xh).sum(-1)    # the bf16 tier's operands
    assert float((plain.double() - ref).abs().max()) > 10 * float((got.double() - ref).abs().max())
    assert not rb.ops.precise_active()
    with rb.ops.precise_inference(impl="mma"):
        assert rb.ops.precise_active() and not torch.is_grad_enabled()
        with pytest.raises(rb._lib.ResencLibraryError):
            P.avg_pool3d(torch.zeros(1, 24, 2, 2, 2), 2)     # CPU tensor: the tier has no CPU fallback either
    assert not rb.ops.precise_active() and torch.is_grad_enabled()


def test_topology_census_matches_survey(rb):
    """SURVEY 0.6 / 0.8 / 8(c) censuses measured on the reference, reproduced by the drop-in on the meta device:
    parameter counts, unique parameters vs state_dict keys (aliases), module counts, stage counts."""
    import contextlib
    import io
    from collections import Counter
    from types import SimpleNamespace
    two = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
    ink = {"ink": {"channels": 1, "activation": "sigmoid"}}

    def census(patch, tasks, cin=1, mc=None):
        mgr = SimpleNamespace(tasks=tasks, train_patch_size=list(patch), train_batch_size=2, in_channels=cin, vram_max=16.0,
                              autoconfigure=True, model_config=mc or {})
        with torch.device("meta"), contextlib.redirect_stdout(io.StringIO()):
            m = rb.NetworkFromConfig(mgr)
        c = Counter(type(x).__name__ for x in m.modules())
        return (round(sum(p.numel() for p in m.parameters()) / 1e6, 2), len(list(m.parameters())), len(m.state_dict()),
                c["Conv3d"], c["ConvTranspose3d"], c["InstanceNorm3d"], c["AvgPool3d"], m.num_stages)

    # 128^3 and 192^3: the same 6 stages, 235.53 M parameters; 69 convs used per forward + 8 unused deep-supervision heads
    assert census([128] * 3, two) == (235.53, 97, 392, 77, 10, 67, 5, 6)
    assert census([192] * 3, two) == census([128] * 3, two)
    # BASELINE config 4: 96^3, 4 input channels, one task, SE on
    assert census([96] * 3, ink, 4, {"squeeze_excitation": True})[0] == 112.26
    # 64^3, two tasks, SE on: 157 unique parameters behind 550 state_dict keys
    assert census([64] * 3, two, 1, {"squeeze_excitation": True})[1:3] == (157, 550)
    assert census([64] * 3, two)[0] == 118.09
    # ink.yaml's literal patch: 7 stages with anisotropic strides, 312 M parameters
    c = census([14, 256, 256], ink, 1, {"squeeze_excitation": True, "conv_bias": True})
    assert c[-1] == 7 and abs(c[0] - 312.27) < 0.01


def test_manual_config_reachability_matrix_matches_survey(rb):
    """SURVEY Appendix A reachability matrix (manual config, 6 stages to 320 features, SE on, two tasks), measured on
    the reference: which block classes a (encoder, bottleneck, decoder) triple builds and the parameter counts."""
    import contextlib
    import io
    from types import SimpleNamespace
    two = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
    base = dict(features_per_stage=[32, 64, 128, 256, 320, 320], num_stages=6, n_blocks_per_stage=[1, 3, 4, 6, 6, 6],
                kernel_sizes=[[3, 3, 3]] * 6, n_conv_per_stage_decoder=[1] * 5, strides=[[1, 1, 1]] + [[2, 2, 2]] * 5,
                squeeze_excitation=True)
    expect = [("BasicBlockD", "BasicBlockD", "ConvBlock", 114.6, "BasicBlockD", 26),
              ("BottleneckBlockD", "BottleneckBlockD", "ConvBlock", 28.2, "BottleneckD", 26),
              ("BasicBlockD", "BottleneckBlockD", "ConvBlock", 114.6, "BasicBlockD", 26),      # encoder.py:74-77
              ("BasicBlockD", "BasicBlockD", "ResidualBlock", 125.6, "BasicBlockD", 36),       # + 5 per decoder, no SE
              ("ConvBlock", "BasicBlockD", "ConvBlock", 68.3, "BasicBlockD", 0)]               # plain conv encoder
    for enc, bott, dec, mparams, cls, nblocks in expect:
        mc = dict(base, basic_encoder_block=enc, bottleneck_block=bott, basic_decoder_block=dec)
        mgr = SimpleNamespace(tasks=two, train_patch_size=[128] * 3, train_batch_size=2, in_channels=1, vram_max=16.0,
                              autoconfigure=False, model_config=mc)
        with torch.device("meta"), contextlib.redirect_stdout(io.StringIO()):
            m = rb.NetworkFromConfig(mgr)
        assert round(sum(p.numel() for p in m.parameters()) / 1e6, 1) == mparams, (enc, bott, dec)
        # block modules, counted once (each decoder holds the shared encoder as a child: modules() de-duplicates)
        assert sum(type(x).__name__ == cls for x in m.modules()) == nblocks, (enc, bott, dec)
        n_se = sum(type(x).__name__ == "SqueezeExcite" for x in m.modules())
        assert n_se == (26 if enc != "ConvBlock" else 0)


def test_nvtx_hooks_attach_and_detach(rb):
    """tracing.enable_nvtx: one push / pop pair per block-level module (the shared encoder once, although every
    decoder lists it as a child), removable without a trace."""
    mgr, _ = case_mgr("sheet_normals_16")
    model = quiet_build(rb.NetworkFromConfig, mgr)
    n_hooks = lambda: sum(len(m._forward_hooks) + len(m._forward_pre_hooks) + len(m._backward_hooks) + len(m._backward_pre_hooks)
                          for m in model.modules())
    assert n_hooks() == 0
    h = rb.tracing.enable_nvtx(model)
    assert "shared_encoder" in h.names and "task_decoders.sheet" in h.names
    assert "shared_encoder.stages.1.blocks.0" in h.names and "shared_encoder.stem.convs.0" in h.names
    assert not any(n.startswith("task_decoders.sheet.encoder") for n in h.names)       # aliases instrumented once
    assert len(h.names) == len(set(h.names)) and n_hooks() == 4 * len(h.names)
    h.remove()
    assert n_hooks() == 0


def test_loss_table_matches_reference_goldens(rb):
    """losses.LOSS_FN_MAP / build_task_losses against values and gradients of the reference trainer's loss classes
    (tests/golden/loss_goldens.json, written by oracle/make_golden.py from training/losses/losses.py)."""
    gold = json.load(open(os.path.join(GOLDEN, "loss_goldens.json")))
    rng = np.random.default_rng(gold["seed"])
    logits = torch.from_numpy(rng.standard_normal((2, 2, 5, 6, 7)).astype(np.float32) * 2)
    target = torch.from_numpy((rng.random((2, 2, 5, 6, 7)) > 0.7).astype(np.float32))
    vec_p = torch.from_numpy(rng.standard_normal((2, 3, 5, 6, 7)).astype(np.float32))
    vec_t = rng.standard_normal((2, 3, 5, 6, 7)).astype(np.float32)
    vec_t /= np.linalg.norm(vec_t, axis=1, keepdims=True)
    vec_t[:, :, :2] = 0
    vec_t = torch.from_numpy(vec_t)
    for fused in (False, True):          # on CPU tensors the fused classes take their composition path
        for c in gold["cases"]:
            fn = rb.losses.build_task_losses({"t": {"loss_fn": c["loss_fn"], "loss_kwargs": c["loss_kwargs"]}}, fused)["t"]
            x = (logits if c["kind"] == "bin" else vec_p).clone().requires_grad_(True)
            v = fn(x, target if c["kind"] == "bin" else vec_t)
            v.backward()
            assert abs(float(v) - c["value"]) < 1e-6, c
            assert abs(float(x.grad.double().norm()) - c["grad_norm"]) < 1e-7 + 1e-5 * c["grad_norm"], c
            assert abs(float(x.grad.double().sum()) - c["grad_sum"]) < 1e-6, c
    assert set(rb.losses.LOSS_FN_MAP) == {"BCEDiceLoss", "BCEWithLogitsLossLabelSmoothing", "BCEWithLogitsLossZSmooth",
                                          "BCEWithLogitsLoss", "BCELoss", "CrossEntropyLoss", "MSELoss", "MaskedCosineLoss"}
    with pytest.raises(ValueError):
        rb.losses.build_task_losses({"t": {"loss_fn": "FocalLoss"}})
    with pytest.raises(TypeError):       # the reference's default (BCEDiceLoss without kwargs) fails the same way
        rb.losses.build_task_losses({"t": {}})


def test_slab_exchange_plan_property(rb):
    """For random volumes / patches / overlaps / world sizes: after applying the planned plane transfers, every rank's
    own z-range holds exactly what a single-rank sweep accumulates there, the own ranges tile [0, Z) without overlap,
    and planes only ever travel forward (src < dst)."""
    from hypothesis import given, settings, strategies as st
    inf = rb.inference

    @settings(max_examples=150, deadline=None)
    @given(st.integers(8, 200), st.integers(4, 64), st.sampled_from([0.0, 0.1, 0.25, 0.5, 0.6, 0.75]), st.integers(1, 9))
    def check(vol_z, patch_z, overlap, world):
        if patch_z > vol_z:
            return
        step = max(1, int(round(patch_z * (1 - overlap))))
        zs = inf.generate_positions(0, vol_z, patch_z, step)
        full = np.zeros(vol_z, np.int64)
        for z in zs:
            full[z:z + patch_z] += 1
        runs = inf.shard_z_starts(zs, world)
        assert sum(runs, []) == list(zs)
        slabs = {}
        for r, run in enumerate(runs):
            if not run:
                continue
            lo, hi = run[0], run[-1] + patch_z
            a = np.zeros(hi - lo, np.int64)
            for z in run:
                a[z - lo:z - lo + patch_z] += 1
            slabs[r] = (lo, a)
        pairs, own = inf.plan_slab_exchange(zs, patch_z, vol_z, world)
        for src, dst, lo, hi in pairs:       # in order and in place, exactly as merge_slabs sends / receives / adds
            assert src < dst and lo < hi
            slo, sa = slabs[src]
            dlo, da = slabs[dst]
            assert own[dst][0] <= lo and hi <= own[dst][1]            # only planes the receiver owns travel
            da[lo - dlo:hi - dlo] += sa[lo - slo:hi - slo]
        covered = np.zeros(vol_z, np.int64)
        for r, (lo, hi) in enumerate(own):
            if hi <= lo:
                assert not runs[r]
                continue
            covered[lo:hi] += 1
            slo, sa = slabs[r]
            assert np.array_equal(sa[lo - slo:hi - slo], full[lo:hi]), (vol_z, patch_z, overlap, world, r)
        assert (covered == 1).all()

    check()


def test_enumeration_and_gaussian_match_oracle_on_random_shapes(rb):
    """Product enumeration / importance map against the oracle (itself pinned to the reference's goldens) on random
    geometries: bit-exact tables, bit-exact fp32 maps."""
    from hypothesis import given, settings, strategies as st
    from oracle import resenc_oracle as O
    inf = rb.inference

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 400), st.integers(1, 96), st.floats(0.0, 0.9))
    def positions(vol, patch, overlap):
        step = max(1, int(round(patch * (1 - overlap))))
        if patch > vol:
            with pytest.raises(ValueError):
                inf.generate_positions(0, vol, patch, step)
            return
        assert list(inf.generate_positions(0, vol, patch, step)) == list(O.positions_1d(0, vol, patch, step))

    @settings(max_examples=25, deadline=None)
    @given(st.tuples(st.integers(2, 40), st.integers(2, 40), st.integers(2, 40)))
    def gaussian(tile):
        a = inf.compute_gaussian_3d(tile)
        a = a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)
        b = O.gaussian_map(tile)
        assert a.dtype == np.float32 and a.tobytes() == np.asarray(b, np.float32).tobytes(), tile

    positions()
    gaussian()


def test_per_class_strided_dgrad_decomposition_property(rb):
    """Host side of the per-parity-class data gradient (ops._axis_classes / pack_conv_dgrad_class, the path for grids
    that do not tile exactly): every input coordinate i = s*j + r receives from outputs j + off + t through the class's
    tap list.  Emulated in torch on random kernel / stride / (odd) sizes and compared with autograd."""
    import itertools

    import torch.nn.functional as F
    from hypothesis import given, settings, strategies as st
    ops = rb.ops
    ax = st.tuples(st.sampled_from([1, 3]), st.sampled_from([1, 2]), st.integers(1, 9))

    @settings(max_examples=60, deadline=None)
    @given(ax, ax, ax)
    def check(a0, a1, a2):
        k, stride, dims = zip(a0, a1, a2)
        torch.manual_seed(1)
        co, ci = 8, 16
        w = torch.randn(co, ci, *k)
        x = torch.randn(1, ci, *dims, requires_grad=True)
        pad = tuple((kk - 1) // 2 for kk in k)
        y = F.conv3d(x, w.to(torch.bfloat16).float(), None, stride, pad)
        dy = torch.randn_like(y)
        y.backward(dy)
        od = tuple(y.shape[2:])
        classes = [ops._axis_classes(k[a], stride[a], pad[a], dims[a]) for a in range(3)]
        dx = torch.zeros(1, ci, *dims)
        for cd, ch, cw in itertools.product(*classes):
            if not (cd[2] and ch[2] and cw[2]):
                continue                                   # class without a kernel index: gradient stays zero
            wpk = ops.pack_conv_dgrad_class(w, cd[2], ch[2], cw[2]).float()      # [taps][ci][co]
            grid = (cd[1], ch[1], cw[1])
            acc = torch.zeros(1, ci, *grid)
            t = 0
            for td in range(len(cd[2])):
                for th in range(len(ch[2])):
                    for tw in range(len(cw[2])):
                        # gather dy at j + off + t with zero fill outside [0, od)
                        src = torch.zeros(1, co, *grid)
                        rng = []
                        for g, o, n in zip(grid, (cd[3] + td, ch[3] + th, cw[3] + tw), od):
                            lo, hi = max(0, -o), min(g, n - o)
                            rng.append((lo, hi, o))
                        if all(hi > lo for lo, hi, _ in rng):
                            (l0, h0, o0), (l1, h1, o1), (l2, h2, o2) = rng
                            src[:, :, l0:h0, l1:h1, l2:h2] = dy[:, :, l0 + o0:h0 + o0, l1 + o1:h1 + o1, l2 + o2:h2 + o2]
                        acc += torch.einsum("bcdhw,nc->bndhw", src, wpk[t])
                        t += 1
            dx[:, :, cd[0]::stride[0], ch[0]::stride[1], cw[0]::stride[2]] = acc
        den = float(x.grad.norm())
        assert float((dx - x.grad).norm()) <= 1e-5 * max(den, 1e-6) + 1e-6, (k, stride, dims)

    check()


def test_transposed_conv_formulations_property(rb):
    """Host formulation of ConvTranspose3d(kernel == stride) (ops._ConvT3dFn): forward = one 1-tap GEMM with N columns
    [(parity, co)] stored pixel-shuffled; data gradient = a stride-s, s^3-tap gather over dy.  Emulated in torch with
    the same packs and compared with F.conv_transpose3d / autograd on random strides and sizes."""
    import torch.nn.functional as F
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.tuples(st.sampled_from([1, 2]), st.sampled_from([1, 2]), st.sampled_from([1, 2])),
           st.tuples(st.integers(1, 5), st.integers(1, 5), st.integers(1, 5)))
    def check(stride, dims):
        torch.manual_seed(2)
        ci, co = 16, 8
        sd, sh, sw = stride
        w = torch.randn(ci, co, *stride)
        x = torch.randn(2, ci, *dims, requires_grad=True)
        wq = w.to(torch.bfloat16).float()
        ref = F.conv_transpose3d(x, wq, None, stride)
        g = torch.randn_like(ref)
        ref.backward(g)
        npar = sd * sh * sw
        # forward pack exactly as ops._ConvT3dFn builds it: rows = (parity, co), cols = ci
        wpk = w.detach().permute(2, 3, 4, 1, 0).reshape(1, npar * co, ci).to(torch.bfloat16).float()
        cols = torch.einsum("bcdhw,nc->bndhw", x.detach(), wpk[0])                     # [B, npar*co, d, h, w]
        c6 = cols.view(2, sd, sh, sw, co, *dims)
        y = c6.permute(0, 4, 5, 1, 6, 2, 7, 3).reshape(2, co, *[d * s for d, s in zip(dims, stride)])
        assert float((y - ref.detach()).norm()) <= 1e-5 * float(ref.detach().norm()) + 1e-6
        # data gradient pack: taps = parities, [tap][ci][co]; gather dy at s*i + p
        wd = w.detach().permute(2, 3, 4, 0, 1).reshape(npar, ci, co).to(torch.bfloat16).float()
        dx = torch.zeros(2, ci, *dims)
        t = 0
        for pd in range(sd):
            for ph in range(sh):
                for pw in range(sw):
                    dx += torch.einsum("bcdhw,nc->bndhw", g[:, :, pd::sd, ph::sh, pw::sw], wd[t])
                    t += 1
        assert float((dx - x.grad).norm()) <= 1e-5 * float(x.grad.norm()) + 1e-6

    check()


def test_gradient_buckets_skip_the_never_used_heads_of_the_real_network(rb):
    """the deep-supervision heads `task_decoders.<t>.seg_layers[:-1]` never receive a gradient
    (builders/decoder.py:128-131).  They are known up front, sit in no bucket and keep .grad None, so the buckets that
    become ready first in backward (the decoders') can launch their all-reduce during backward."""
    par = importlib.import_module(rb._pkg.__name__ + ".parallel")
    tasks = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
    model = quiet_build(rb.NetworkFromConfig, make_mgr([128, 128, 128], tasks, batch=2))
    unused = par.never_used_parameters(model)
    names = {id(p): n for n, p in model.named_parameters()}
    assert len(unused) == 16 and all(".seg_layers." in names[id(p)] and not names[id(p)].split(".seg_layers.")[1].startswith("4.")
                                     for p in unused)
    buckets = par.GradientBuckets(model, bucket_bytes=64 << 20)
    bucketed = {id(p) for b in buckets.buckets for p in b["params"]}
    assert not (bucketed & {id(p) for p in unused})
    assert len(bucketed) + 16 == len(list(model.parameters()))
    assert all(p.grad is None for p in model.parameters())
    # fp32 on the wire: 4 bytes per bucketed parameter (slots padded to 16-byte boundaries); bf16: 2
    ids = {id(p) for p in unused}
    n = sum((p.numel() + 3) & ~3 for p in model.parameters() if id(p) not in ids)
    assert buckets.bytes_per_step == 4 * n
    assert n - sum(p.numel() for p in model.parameters() if id(p) not in ids) < 64          # a handful of bias slots
    assert par.GradientBuckets(model, comm_dtype=torch.bfloat16).bytes_per_step == 2 * n
    for b in buckets.buckets:
        for p in b["params"]:
            assert (buckets._slots[id(p)].data_ptr() - b["flat"].data_ptr()) % 16 == 0


def test_transposed_conv_operand_packs_in_one_copy(rb):
    """ops._permuted_bf16 (one strided copy + cast) == the permute / reshape / cast / contiguous chain it replaces, for
    both operand layouts of the transposed conv (decoder.py:110-113)."""
    ops = rb.ops
    torch.manual_seed(12)
    for ci, co, s in ((64, 32, (2, 2, 2)), (16, 8, (1, 2, 2)), (8, 24, (2, 2, 1))):
        w = torch.randn(ci, co, *s)
        npar = s[0] * s[1] * s[2]
        a = ops._permuted_bf16(w, (2, 3, 4, 1, 0), (1, npar * co, ci))
        b = w.detach().permute(2, 3, 4, 1, 0).reshape(1, npar * co, ci).to(torch.bfloat16).contiguous()
        assert a.is_contiguous() and a.dtype == torch.bfloat16 and torch.equal(a, b)
        a = ops._permuted_bf16(w, (2, 3, 4, 0, 1), (npar, ci, co))
        b = w.detach().permute(2, 3, 4, 0, 1).reshape(npar, ci, co).to(torch.bfloat16).contiguous()
        assert a.is_contiguous() and torch.equal(a, b)


def test_blosc_container_roundtrip_property(rb):
    """Any byte string, element size, shuffle mode, block size and level: decompress(compress(x)) == x, the header
    describes the buffer, and the buffer never exceeds the input by more than the 16-byte header."""
    import importlib
    from hypothesis import given, settings, strategies as st
    bc = importlib.import_module(rb.inference.__name__ + ".blosc_codec")

    @settings(max_examples=60, deadline=None)
    @given(st.binary(min_size=0, max_size=5000), st.sampled_from([1, 2, 4, 8]), st.sampled_from([0, 1, 2]),
           st.sampled_from([0, 64, 256, 1000, 4096]), st.integers(0, 9))
    def check(data, typesize, shuffle, blocksize, clevel):
        if blocksize and typesize > 1:
            blocksize -= blocksize % typesize
        buf = bc.compress(data, typesize=typesize, clevel=clevel, shuffle=shuffle, blocksize=blocksize or 0)
        h = bc.header(buf)
        assert h["nbytes"] == len(data) and h["cbytes"] == len(buf) and h["typesize"] == typesize
        assert len(buf) <= len(data) + 16
        assert bc.decompress(buf) == data
        # repetitive input of a compressible size must actually shrink
        rep = (data[:7] or b"\x01") * 400
        assert len(bc.compress(rep, typesize=typesize, clevel=max(clevel, 1), shuffle=shuffle)) < len(rep)

    check()


def test_decoder_routes_the_head_through_its_last_conv_unit(rb, monkeypatch):
    """Host logic of the inference-tail fusion, without a GPU: `Decoder.forward` hands the task head to the LAST conv
    unit of its LAST stage only; that unit calls `ops.conv_norm_act_head` when `ops.can_fuse_head` agrees (autograd off)
    and `ops.conv_norm_act` + `ops.head_conv1x1` otherwise; deep supervision keeps one separate head call per stage."""
    ops = rb.ops
    tasks = {"sheet": {"channels": 1, "activation": "sigmoid"}}
    model = quiet_build(rb.NetworkFromConfig, make_mgr([16, 16, 16], tasks))
    dec = model.task_decoders["sheet"]
    calls = []

    def fake_cna(x, weight, stride=1, x_cat=None, res=None, *a, **k):
        calls.append(("unit", weight.shape[0]))
        n, d = x.shape[0], x.shape[2:]
        return torch.zeros(n, weight.shape[0], *d)

    def fake_fused(x, weight, stride, x_cat, res, gamma, beta, eps, act, slope, stem, head):
        calls.append(("fused", weight.shape[0], head[0].shape[0], head[2]))
        return torch.zeros(x.shape[0], head[0].shape[0], *x.shape[2:])

    def fake_head(x, weight, bias, activation=None):
        calls.append(("head", weight.shape[0], activation))
        return torch.zeros(x.shape[0], weight.shape[0], *x.shape[2:])

    def fake_up(x, weight, stride, impl=None, bias=None):
        calls.append(("up", weight.shape[1]))
        return torch.zeros(x.shape[0], weight.shape[1], *[i * s for i, s in zip(x.shape[2:], stride)])

    monkeypatch.setattr(ops, "conv_norm_act", fake_cna)
    monkeypatch.setattr(ops, "conv_norm_act_head", fake_fused)
    monkeypatch.setattr(ops, "head_conv1x1", fake_head)
    monkeypatch.setattr(ops, "conv_transpose3d", fake_up)
    monkeypatch.setattr(ops, "attach_cancelled_bias", lambda z, b: z)
    enc = model.shared_encoder
    chans, strides = enc.output_channels, enc.strides
    dims, skips = [16, 16, 16], []
    for c, s in zip(chans, strides):
        dims = [d // ss for d, ss in zip(dims, s)]
        skips.append(torch.zeros(1, c, *dims))
    n_stage = len(dec.stages)

    monkeypatch.setattr(ops, "can_fuse_head", lambda w, head, se=None, drop=None: head is not None)
    calls.clear()
    out = dec(skips, activation="sigmoid")
    assert out.shape == (1, 1, 16, 16, 16)
    assert [c[0] for c in calls].count("fused") == 1 and calls[-1][0] == "fused" and calls[-1][2:] == (1, "sigmoid")
    assert not any(c[0] == "head" for c in calls)
    assert [c[0] for c in calls].count("up") == n_stage

    monkeypatch.setattr(ops, "can_fuse_head", lambda w, head, se=None, drop=None: False)      # training / compile / gate
    calls.clear()
    dec(skips, activation=None)
    assert not any(c[0] == "fused" for c in calls)
    assert calls[-1] == ("head", 1, None) and [c[0] for c in calls].count("head") == 1
    assert calls[-2][0] == "unit"

    dec.deep_supervision = True
    calls.clear()
    outs = dec(skips, activation="sigmoid")
    assert len(outs) == n_stage and [c[0] for c in calls].count("head") == n_stage
    assert not any(c[0] == "fused" for c in calls)
    heads = [c for c in calls if c[0] == "head"]
    assert heads[-1][2] == "sigmoid" and all(h[2] is None for h in heads[:-1])       # activation on the full-res head only
    dec.deep_supervision = False

    # the real predicate: never with autograd on, never for gated / stochastic-depth units, only supported widths
    monkeypatch.undo()
    w32, head = torch.zeros(32, 32, 3, 3, 3), (torch.zeros(3, 32, 1, 1, 1), None, None)
    with torch.no_grad():
        assert ops.can_fuse_head(w32, head)
        assert not ops.can_fuse_head(w32, None)
        assert not ops.can_fuse_head(w32, head, se=(1, 2, 3, 4)) and not ops.can_fuse_head(w32, head, drop=torch.ones(1))
        assert not ops.can_fuse_head(torch.zeros(24, 32, 3, 3, 3), (torch.zeros(3, 24, 1, 1, 1), None, None))   # C/8 = 3
        assert not ops.can_fuse_head(torch.zeros(512, 32, 3, 3, 3), (torch.zeros(3, 512, 1, 1, 1), None, None))  # C/8 = 64
    assert not ops.can_fuse_head(w32, head)          # autograd on
