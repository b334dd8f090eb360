"""Output writer: finalised z-ranges -> chunked uint8 / uint16 arrays in the reference's zarr
layout (inference.py:213-263): one `<target>_final` array per target, shape [Z, Y, X] (c == 1) or [c, Z, Y, X],
chunks = the patch size (with all channels in one chunk), fill_value 0, all-zero chunks not written
(`write_empty_chunks=False`), C order.

`zarr` / `numcodecs` / Blosc are not installed in this image, so the zarr **v2** directory format is written
directly (`.zgroup`, `<array>/.zarray`, one file per chunk named "i.j.k").  The reference compresses with
Blosc(zstd, clevel 5, bitshuffle) (inference.py:92,224): compressor "blosc" writes exactly that configuration
(`blosc_codec.py`: the Blosc-1 container around zstd frames from the system libzstd) and records numcodecs' Blosc
config in `.zarray`; "zlib" (numcodecs id "zlib") and None are kept as codecs with no third-party format involved.
All three open unchanged with `zarr.open(path)`.  Chunk compression and file writes run on a host
thread pool while the GPU keeps sweeping / finalising the next z-range (`submit` returns immediately; `close` joins).
Each rank of a z-slab-sharded sweep writes the chunks of the z-range it owns; ranges are aligned to the chunk grid
by the caller or fall back to read-modify-write of the boundary chunks on one rank at a time.
"""
from __future__ import annotations

import json
import os
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import blosc_codec

_DTYPES = {np.dtype("uint8"): "|u1", np.dtype("uint16"): "<u2", np.dtype("float32"): "<f4"}


def _zarray_meta(shape, chunks, dtype, compressor):
    comp = None
    if compressor == "zlib":
        comp = {"id": "zlib", "level": 1}
    elif compressor == "blosc":                    # numcodecs.Blosc(cname='zstd', clevel=5, shuffle=BITSHUFFLE).get_config()
        if not blosc_codec.available():
            raise NotImplementedError("compressor 'blosc' needs a zstd implementation (libzstd.so.1 or pyarrow)")
        comp = {"id": "blosc", "cname": "zstd", "clevel": 5, "shuffle": 2, "blocksize": 0}
    elif compressor is not None:
        raise NotImplementedError(f"compressor {compressor!r}: 'blosc' (zstd, bit-shuffle), 'zlib' or None")
    return {"zarr_format": 2, "shape": list(shape), "chunks": list(chunks), "dtype": _DTYPES[np.dtype(dtype)],
            "compressor": comp, "fill_value": 0, "order": "C", "filters": None, "dimension_separator": "."}


class ZarrArrayWriter:
    """One zarr v2 array on disk, written region by region."""

    def __init__(self, root: str, name: str, shape: Sequence[int], chunks: Sequence[int], dtype, compressor="zlib",
                 pool: Optional[ThreadPoolExecutor] = None, create: bool = True):
        self.path = os.path.join(root, name)
        self.shape = tuple(int(s) for s in shape)
        self.chunks = tuple(int(c) for c in chunks)
        if len(self.shape) != len(self.chunks) or any(c <= 0 for c in self.chunks):
            raise ValueError(f"shape {self.shape} / chunks {self.chunks} mismatch")
        self.dtype = np.dtype(dtype)
        self.compressor = compressor
        self.pool = pool
        self._futures = []
        self.meta = _zarray_meta(self.shape, self.chunks, self.dtype, compressor)
        if create:
            os.makedirs(self.path, exist_ok=True)
            with open(os.path.join(self.path, ".zarray"), "w") as f:
                json.dump(self.meta, f, indent=1)

    # -- chunk codec -----------------------------------------------------------------------------
    def _encode(self, block: np.ndarray) -> bytes:
        raw = np.ascontiguousarray(block, dtype=self.dtype.newbyteorder("<") if self.dtype.itemsize > 1 else self.dtype).tobytes()
        if self.compressor == "blosc":
            return blosc_codec.compress(raw, typesize=self.dtype.itemsize, clevel=5, shuffle=blosc_codec.BITSHUFFLE)
        return zlib.compress(raw, 1) if self.compressor == "zlib" else raw

    def _decode(self, data: bytes) -> np.ndarray:
        raw = (blosc_codec.decompress(data) if self.compressor == "blosc" else
               zlib.decompress(data) if self.compressor == "zlib" else data)
        return np.frombuffer(raw, dtype=self.dtype).reshape(self.chunks).copy()

    def _chunk_file(self, idx: Tuple[int, ...]) -> str:
        return os.path.join(self.path, ".".join(str(i) for i in idx))

    def _write_chunk(self, idx, block: np.ndarray):
        """`block` is the full chunk (edge chunks padded with the fill value, as zarr stores them)."""
        fn = self._chunk_file(idx)
        if not block.any():                       # write_empty_chunks=False
            if os.path.exists(fn):
                os.remove(fn)
            return 0
        data = self._encode(block)
        tmp = fn + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(data)
        os.replace(tmp, fn)
        return len(data)

    def read_chunk(self, idx) -> np.ndarray:
        fn = self._chunk_file(idx)
        if not os.path.exists(fn):
            return np.zeros(self.chunks, self.dtype)
        with open(fn, "rb") as f:
            return self._decode(f.read())

    # -- regions ----------------------------------------------------------------------------------
    def write_z_range(self, z0: int, data: np.ndarray):
        """Store `data` = array[..., z0:z0+nz, :, :] (all other axes complete).  Chunks fully covered along z are
        written directly; a chunk that the range covers only partly is read, merged and re-written (single writer per
        chunk is the caller's responsibility: align rank boundaries to the chunk grid or serialise the ranks)."""
        data = np.asarray(data)
        zax = len(self.shape) - 3
        if data.shape[:zax] != self.shape[:zax] or data.shape[zax + 1:] != self.shape[zax + 1:]:
            raise ValueError(f"region shape {data.shape} does not span the array {self.shape} outside z")
        nz = data.shape[zax]
        if z0 < 0 or z0 + nz > self.shape[zax]:
            raise ValueError(f"z-range [{z0}, {z0 + nz}) outside the array")
        cz, cy, cx = self.chunks[-3:]
        lead = tuple(range(-(-self.shape[i] // self.chunks[i])) for i in range(zax))
        for kz in range(z0 // cz, -(-(z0 + nz) // cz)):
            a, b = max(z0, kz * cz), min(z0 + nz, (kz + 1) * cz)          # volume z covered by this chunk row
            whole = (a == kz * cz) and (b == min((kz + 1) * cz, self.shape[zax]))
            if not whole:
                self.flush()        # read-modify-write of a shared chunk row: never concurrent with earlier writes
            for ky in range(-(-self.shape[-2] // cy)):
                for kx in range(-(-self.shape[-1] // cx)):
                    for li in np.ndindex(*[len(r) for r in lead]) if lead else [()]:
                        idx = tuple(li) + (kz, ky, kx)
                        if whole:
                            self._submit(self._store, idx, data, z0, a, b, True)
                        else:
                            self._store(idx, data, z0, a, b, False)

    def _store(self, idx, data, z0, a, b, whole):
        zax = len(self.shape) - 3
        block = np.zeros(self.chunks, self.dtype) if whole else self.read_chunk(idx)
        src, dst = [], []
        for ax, k in enumerate(idx):
            lo = k * self.chunks[ax]
            hi = min(lo + self.chunks[ax], self.shape[ax])
            if ax == zax:
                src.append(slice(a - z0, b - z0))
                dst.append(slice(a - lo, b - lo))
            else:
                src.append(slice(lo, hi))
                dst.append(slice(0, hi - lo))
        block[tuple(dst)] = data[tuple(src)]
        return self._write_chunk(idx, block)

    def _submit(self, fn, *args):
        if self.pool is None:
            fn(*args)
        else:
            self._futures.append(self.pool.submit(fn, *args))

    def flush(self):
        for f in self._futures:
            f.result()                  # re-raises worker exceptions
        self._futures = []

    def read(self) -> np.ndarray:
        """Whole array back in memory (tests / small volumes)."""
        out = np.zeros(self.shape, self.dtype)
        grid = [range(-(-s // c)) for s, c in zip(self.shape, self.chunks)]
        for idx in np.ndindex(*[len(g) for g in grid]):
            blk = self.read_chunk(idx)
            sl, bl = [], []
            for ax, k in enumerate(idx):
                lo = k * self.chunks[ax]
                hi = min(lo + self.chunks[ax], self.shape[ax])
                sl.append(slice(lo, hi))
                bl.append(slice(0, hi - lo))
            out[tuple(sl)] = blk[tuple(bl)]
        return out


class FinalVolumeWriter:
    """`<target>_final` arrays of one inference run (inference.py:213-263), fed by `SlabBlender.finalize`.

        w = FinalVolumeWriter(path, targets, vol_shape, patch, threads=8)
        for z_from, z_to in ranges:                      # e.g. one chunk row of planes at a time
            w.submit(z_from, blender.finalize(z_from, z_to))   # D2H copy here, compression + I/O on the pool
        w.close()
    """

    def __init__(self, root: str, targets: Dict[str, dict], vol_shape: Sequence[int], patch: Sequence[int],
                 compressor="zlib", threads: int = 8, create: bool = True):
        self.root = root
        self.pool = ThreadPoolExecutor(max_workers=max(1, int(threads))) if threads and threads > 0 else None
        if create:
            os.makedirs(root, exist_ok=True)
            with open(os.path.join(root, ".zgroup"), "w") as f:
                json.dump({"zarr_format": 2}, f)
        Z, Y, X = (int(s) for s in vol_shape)
        pz, py, px = (int(p) for p in patch)
        self.arrays: Dict[str, ZarrArrayWriter] = {}
        for t, info in targets.items():
            c = int(info["channels"])
            dtype = np.uint16 if t.lower() == "normals" else np.uint8           # inference.py:218-221
            shape, chunks = ((Z, Y, X), (pz, py, px)) if c == 1 else ((c, Z, Y, X), (c, pz, py, px))   # :86-91
            self.arrays[t] = ZarrArrayWriter(root, f"{t}_final", shape, chunks, dtype, compressor, self.pool, create)
        self.bytes_in = 0

    def submit(self, z_from: int, finalized: Dict[str, "object"]):
        """`finalized` = {target: uint8 / uint16 tensor or array [(c,) nz, Y, X]} for planes starting at z_from."""
        for t, v in finalized.items():
            a = v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v)        # torch.uint16 -> numpy uint16
            self.bytes_in += a.nbytes
            self.arrays[t].write_z_range(int(z_from), a)

    def close(self):
        for w in self.arrays.values():
            w.flush()
        if self.pool is not None:
            self.pool.shutdown(wait=True)
            self.pool = None


class ZarrArrayReader:
    """Read-only view of a zarr v2 array directory with zarr's slicing interface (`.shape`, `.dtype`, `a[z0:z1]`,
    `a[..., z0:z1, y0:y1, x0:x1]`), enough for `DeviceVolume` / `SlidingWindowInferer` to consume a volume without the
    `zarr` package.  Codecs: none, zlib and Blosc with zstd frames (what `ZarrArrayWriter` writes and what the
    reference's own outputs use); other Blosc codecs (lz4, blosclz) raise `NotImplementedError`."""

    def __init__(self, path: str):
        with open(os.path.join(path, ".zarray")) as f:
            meta = json.load(f)
        if meta.get("zarr_format") != 2 or meta.get("order", "C") != "C" or meta.get("filters"):
            raise NotImplementedError("only unfiltered C-order zarr v2 arrays are supported")
        comp = meta.get("compressor")
        if comp is not None and comp.get("id") not in ("zlib", "blosc"):
            raise NotImplementedError(f"compressor {comp.get('id')!r}: only blosc (zstd) / zlib / none can be decoded here")
        self.path = path
        self.shape = tuple(meta["shape"])
        self.chunks = tuple(meta["chunks"])
        self.dtype = np.dtype(meta["dtype"])
        self.fill_value = meta.get("fill_value") or 0
        self.sep = meta.get("dimension_separator", ".")
        self._codec = comp.get("id") if comp is not None else None
        self.ndim = len(self.shape)

    def _chunk(self, idx):
        fn = os.path.join(self.path, self.sep.join(str(i) for i in idx))
        if not os.path.exists(fn):
            return np.full(self.chunks, self.fill_value, self.dtype)
        with open(fn, "rb") as f:
            raw = f.read()
        if self._codec == "zlib":
            raw = zlib.decompress(raw)
        elif self._codec == "blosc":
            raw = blosc_codec.decompress(raw)
        return np.frombuffer(raw, dtype=self.dtype).reshape(self.chunks)

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if any(k is Ellipsis for k in key):
            i = [k is Ellipsis for k in key].index(True)
            key = key[:i] + (slice(None),) * (self.ndim - len(key) + 1) + key[i + 1:]
        key = key + (slice(None),) * (self.ndim - len(key))
        if len(key) != self.ndim:
            raise IndexError(f"too many indices for a {self.ndim}-D array")
        ranges, squeeze = [], []
        for ax, k in enumerate(key):
            if isinstance(k, (int, np.integer)):
                k = int(k) + (self.shape[ax] if k < 0 else 0)
                if not 0 <= k < self.shape[ax]:
                    raise IndexError(f"index {k} out of range for axis {ax}")
                ranges.append((k, k + 1))
                squeeze.append(ax)
            elif isinstance(k, slice):
                lo, hi, stp = k.indices(self.shape[ax])
                if stp != 1:
                    raise NotImplementedError("strided slices are not supported")
                ranges.append((lo, max(lo, hi)))
            else:
                raise NotImplementedError(f"index type {type(k).__name__}")
        out = np.empty([hi - lo for lo, hi in ranges], self.dtype)
        grid = [range(lo // c, -(-hi // c)) if hi > lo else range(0) for (lo, hi), c in zip(ranges, self.chunks)]
        for idx in np.ndindex(*[len(g) for g in grid]):
            cidx = tuple(g[i] for g, i in zip(grid, idx))
            blk = self._chunk(cidx)
            src, dst = [], []
            for ax, ci in enumerate(cidx):
                c0 = ci * self.chunks[ax]
                lo, hi = max(ranges[ax][0], c0), min(ranges[ax][1], c0 + self.chunks[ax])
                src.append(slice(lo - c0, hi - c0))
                dst.append(slice(lo - ranges[ax][0], hi - ranges[ax][0]))
            out[tuple(dst)] = blk[tuple(src)]
        return out.squeeze(axis=tuple(squeeze)) if squeeze else out


def open_zarr_array(root: str, name: Optional[str] = None) -> ZarrArrayReader:
    """`open_zarr_array("out.zarr", "sheet_final")` or `open_zarr_array("vol.zarr/0")`."""
    return ZarrArrayReader(os.path.join(root, name) if name else root)
