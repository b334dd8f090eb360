"""Blosc-1 chunk container with the zstd codec and bit-shuffle: the compressor of the reference's output arrays,
`Blosc(cname='zstd', clevel=5, shuffle=Blosc.BITSHUFFLE)` (inference.py:92,224).

`numcodecs` / `blosc` are not installed in this image, so the container is written (and read back) directly from the
published c-blosc 1.x format; the zstd frames inside come from the system `libzstd.so.1` through ctypes (pyarrow's
bundled zstd as the fallback).  PARITY NOTE: there is no libblosc here to decode these buffers with, so the encoder is
checked by (a) an independent decoder in this file, (b) byte-level assertions on the header / block table and (c) the
bit-shuffle against a literal per-bit restatement (tests/test_host_logic.py) - not against libblosc itself.

Container layout (little endian), as c-blosc 1.21 writes it:
    byte 0   format version (2)            byte 1   codec format version (zstd: 1)
    byte 2   flags: 0x01 byte-shuffle, 0x02 memcpyed, 0x04 bit-shuffle, 0x10 blocks are NOT split per byte of the
             type (written clear here), bits 5..7 codec (0 blosclz, 1 lz4, 3 zlib, 4 zstd)
    byte 3   typesize
    4..7     nbytes (uncompressed)     8..11  blocksize     12..15  cbytes (whole buffer, header included)
    then     int32 bstarts[nblocks]    offset of every block from the start of the buffer
    block    `typesize` streams (one per byte plane of the shuffled block) when typesize <= 16 and the block holds
             >= 128 elements and is not the ragged last block, else one stream; each stream = int32 csize, then csize
             bytes: one zstd frame, or the stream itself when csize == its length (incompressible)
A buffer that does not shrink (or is shorter than 128 bytes) is stored with the memcpyed flag: header + raw bytes.

Bit-shuffle (bitshuffle's `bshuf_trans_bit_elem`, per block): out[(j * 8 + b) * (n / 8) + k] holds bit b of byte j
of elements 8k .. 8k+7, element 8k+i in bit i.  c-blosc 1.x - the library numcodecs bundles - shuffles a block only when
its element count n is a multiple of 8 and copies it unchanged otherwise (c-blosc2 shuffles the largest multiple of 8
instead); the 1.x rule is followed here because the flag in the header is all a decoder has to go by.  Trailing bytes
that do not fill an element are copied.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import struct

import numpy as np

BLOSC_VERSION_FORMAT = 2
BLOSC_ZSTD_VERSION_FORMAT = 1
BLOSC_ZSTD_FORMAT = 4
FLAG_SHUFFLE, FLAG_MEMCPYED, FLAG_BITSHUFFLE, FLAG_DONT_SPLIT = 0x01, 0x02, 0x04, 0x10
NOSHUFFLE, SHUFFLE, BITSHUFFLE = 0, 1, 2
MAX_OVERHEAD = 16
MIN_BUFFERSIZE = 128
MAX_SPLITS = 16
_L1 = 32 * 1024


# ---------------------------------------------------------------------------------------------
# zstd
# ---------------------------------------------------------------------------------------------
class _Zstd:
    def __init__(self):
        self.lib = None
        for name in ("libzstd.so.1", ctypes.util.find_library("zstd")):
            if not name:
                continue
            try:
                lib = ctypes.CDLL(name)
            except OSError:
                continue
            lib.ZSTD_compressBound.restype = ctypes.c_size_t
            lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
            lib.ZSTD_compress.restype = ctypes.c_size_t
            lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
            lib.ZSTD_decompress.restype = ctypes.c_size_t
            lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
            lib.ZSTD_isError.restype = ctypes.c_uint
            lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
            lib.ZSTD_maxCLevel.restype = ctypes.c_int
            self.lib = lib
            break
        self.pa = None
        if self.lib is None:
            try:
                import pyarrow as pa
                if pa.Codec.is_available("zstd"):
                    self.pa = pa
            except ImportError:
                pass
        if self.lib is None and self.pa is None:
            raise RuntimeError("no zstd implementation found (libzstd.so.1 / pyarrow): the Blosc-zstd codec cannot run")

    def compress(self, data: bytes, level: int) -> bytes:
        if self.lib is not None:
            cap = self.lib.ZSTD_compressBound(len(data))
            dst = ctypes.create_string_buffer(cap)
            n = self.lib.ZSTD_compress(dst, cap, data, len(data), int(level))
            if self.lib.ZSTD_isError(n):
                raise RuntimeError("ZSTD_compress failed")
            return dst.raw[:n]
        return self.pa.Codec("zstd", compression_level=int(level)).compress(data, asbytes=True)

    def decompress(self, data: bytes, size: int) -> bytes:
        if self.lib is not None:
            dst = ctypes.create_string_buffer(max(size, 1))
            n = self.lib.ZSTD_decompress(dst, size, data, len(data))
            if self.lib.ZSTD_isError(n) or n != size:
                raise ValueError("corrupt zstd frame inside a Blosc block")
            return dst.raw[:size]
        return self.pa.Codec("zstd").decompress(data, decompressed_size=size, asbytes=True)

    def max_level(self) -> int:
        return int(self.lib.ZSTD_maxCLevel()) if self.lib is not None else 22


_zstd_singleton = None


def _zstd() -> _Zstd:
    global _zstd_singleton
    if _zstd_singleton is None:
        _zstd_singleton = _Zstd()
    return _zstd_singleton


def available() -> bool:
    try:
        _zstd()
        return True
    except RuntimeError:
        return False


# ---------------------------------------------------------------------------------------------
# shuffles (per block)
# ---------------------------------------------------------------------------------------------
def _byte_shuffle(block: np.ndarray, ts: int) -> np.ndarray:
    n = block.size // ts
    out = block.copy()
    out[:n * ts] = block[:n * ts].reshape(n, ts).T.reshape(-1)
    return out


def _byte_unshuffle(block: np.ndarray, ts: int) -> np.ndarray:
    n = block.size // ts
    out = block.copy()
    out[:n * ts] = block[:n * ts].reshape(ts, n).T.reshape(-1)
    return out


def _bit_shuffle(block: np.ndarray, ts: int) -> np.ndarray:
    n = block.size // ts
    out = block.copy()
    if n == 0 or n % 8 != 0:                # c-blosc 1.x: not a multiple of 8 elements -> the block is copied
        return out
    by = block[:n * ts].reshape(n, ts).T                                     # [byte j][element]
    bits = np.unpackbits(by.reshape(ts, n // 8, 8, 1), axis=-1, bitorder="little")   # [j][k][i][b]
    out[:n * ts] = np.packbits(bits.transpose(0, 3, 1, 2), axis=-1, bitorder="little").reshape(-1)   # [j][b][k] <- bits i
    return out


def _bit_unshuffle(block: np.ndarray, ts: int) -> np.ndarray:
    n = block.size // ts
    out = block.copy()
    if n == 0 or n % 8 != 0:
        return out
    rows = block[:n * ts].reshape(ts, 8, n // 8, 1)                          # [j][b][k]
    bits = np.unpackbits(rows, axis=-1, bitorder="little")                   # [j][b][k][i]
    by = np.packbits(bits.transpose(0, 2, 3, 1), axis=-1, bitorder="little").reshape(ts, n)   # [j][element] <- bits b
    out[:n * ts] = by.T.reshape(-1)
    return out


def _auto_blocksize(nbytes: int, typesize: int, clevel: int) -> int:
    """c-blosc's automatic block size for a high-compression-ratio codec (zstd): L1 * 2 scaled by the level.  Any
    value decodes; this keeps the blocks close to what the reference's files contain."""
    if nbytes < _L1:
        bs = nbytes
    else:
        bs = _L1 * 2
        bs = {0: bs // 4, 1: bs // 2, 2: bs, 3: bs * 2, 4: bs * 4, 5: bs * 4, 6: bs * 4, 7: bs * 8, 8: bs * 8}.get(clevel, bs * 16)
    bs = min(bs, nbytes)
    if bs > typesize:
        bs -= bs % typesize
    return max(bs, 1)


# ---------------------------------------------------------------------------------------------
# container
# ---------------------------------------------------------------------------------------------
def compress(data, typesize: int = 1, clevel: int = 5, shuffle: int = BITSHUFFLE, blocksize: int = 0) -> bytes:
    """Blosc-1 buffer (cname 'zstd') of `data` (bytes-like).  `typesize` = itemsize of the array (numcodecs passes the
    chunk's dtype size), `shuffle` 0 none / 1 byte / 2 bit, `blocksize` 0 = automatic."""
    src = np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview, np.ndarray)) else data, dtype=np.uint8)
    src = src.reshape(-1)
    nbytes = src.size
    if not 1 <= typesize <= 255:
        typesize = 1
    if not 0 <= clevel <= 9:
        raise ValueError("clevel must lie in 0..9")
    if nbytes > 2 ** 31 - 1 - MAX_OVERHEAD:
        raise ValueError("Blosc-1 buffers hold at most 2 GiB")
    bs = int(blocksize) if blocksize else _auto_blocksize(nbytes, typesize, clevel)
    if shuffle == BITSHUFFLE and ((bs // typesize) % 8 != 0 or (nbytes // typesize) % 8 != 0):
        # a block whose element count is not a multiple of 8 is where c-blosc 1.x (copies the block) and c-blosc2
        # (shuffles the largest multiple of 8) disagree: such buffers - never a full zarr chunk of the reference's patch
        # shapes - are written unshuffled, which every decoder reads the same way
        shuffle = NOSHUFFLE
    # blocks are split into one stream per byte of the type (bit 4 clear) whenever c-blosc's rule allows it: decoders
    # older than the "don't split" flag (c-blosc < 1.15) always assume that rule, newer ones honour it when the flag is
    # clear, so every version reads the buffer the same way
    flags = BLOSC_ZSTD_FORMAT << 5
    if shuffle == SHUFFLE:
        flags |= FLAG_SHUFFLE
    elif shuffle == BITSHUFFLE:
        flags |= FLAG_BITSHUFFLE
    elif shuffle != NOSHUFFLE:
        raise ValueError("shuffle must be 0 (none), 1 (byte) or 2 (bit)")

    def memcpyed():
        hdr = struct.pack("<BBBBiii", BLOSC_VERSION_FORMAT, BLOSC_ZSTD_VERSION_FORMAT, flags | FLAG_MEMCPYED, typesize,
                          nbytes, max(bs, 0), nbytes + MAX_OVERHEAD)
        return hdr + src.tobytes()

    if clevel == 0 or nbytes < MIN_BUFFERSIZE:
        return memcpyed()
    z = _zstd()
    level = clevel * 2 - 1 if clevel < 9 else z.max_level()
    nblocks = -(-nbytes // bs)
    parts, bstarts = [], []
    pos = MAX_OVERHEAD + 4 * nblocks
    for b in range(nblocks):
        blk = src[b * bs:(b + 1) * bs]
        if shuffle == SHUFFLE and typesize > 1:
            blk = _byte_shuffle(blk, typesize)
        elif shuffle == BITSHUFFLE and blk.size >= typesize:
            blk = _bit_shuffle(blk, typesize)
        raw = blk.tobytes()
        leftover = len(raw) != bs
        nsplit = typesize if (not leftover and 1 < typesize <= MAX_SPLITS and len(raw) // typesize >= MIN_BUFFERSIZE) else 1
        ne = len(raw) // nsplit
        bstarts.append(pos)
        for j in range(nsplit):
            piece = raw[j * ne:(j + 1) * ne]
            comp = z.compress(piece, level)
            if len(comp) >= len(piece):     # incompressible stream: stored as is (csize == stream length)
                comp = piece
            parts.append(struct.pack("<i", len(comp)) + comp)
            pos += 4 + len(comp)
    if pos > nbytes + MAX_OVERHEAD:         # did not shrink: the whole buffer is stored raw
        return memcpyed()
    hdr = struct.pack("<BBBBiii", BLOSC_VERSION_FORMAT, BLOSC_ZSTD_VERSION_FORMAT, flags, typesize, nbytes, bs, pos)
    return hdr + struct.pack("<%di" % nblocks, *bstarts) + b"".join(parts)


def header(buf) -> dict:
    if len(buf) < MAX_OVERHEAD:
        raise ValueError("not a Blosc buffer (shorter than its header)")
    ver, verlz, flags, ts, nbytes, bs, cbytes = struct.unpack("<BBBBiii", bytes(buf[:MAX_OVERHEAD]))
    return {"version": ver, "versionlz": verlz, "flags": flags, "typesize": ts, "nbytes": nbytes, "blocksize": bs,
            "cbytes": cbytes, "codec": flags >> 5, "shuffle": SHUFFLE if flags & FLAG_SHUFFLE else BITSHUFFLE if flags & FLAG_BITSHUFFLE else NOSHUFFLE,
            "memcpyed": bool(flags & FLAG_MEMCPYED), "split": not (flags & FLAG_DONT_SPLIT)}


def decompress(buf) -> bytes:
    """Inverse of `compress` for zstd-coded (and memcpyed) Blosc-1 buffers, split or unsplit blocks."""
    buf = bytes(buf)
    h = header(buf)
    if h["version"] != BLOSC_VERSION_FORMAT:
        raise ValueError(f"Blosc format version {h['version']} is not supported")
    nbytes, bs, ts = h["nbytes"], h["blocksize"], h["typesize"]
    if h["cbytes"] > len(buf):
        raise ValueError("truncated Blosc buffer")
    if h["memcpyed"]:
        return buf[MAX_OVERHEAD:MAX_OVERHEAD + nbytes]
    if h["codec"] != BLOSC_ZSTD_FORMAT:
        raise NotImplementedError(f"Blosc codec id {h['codec']}: only zstd (4) is decoded here")
    if nbytes == 0:
        return b""
    z = _zstd()
    nblocks = -(-nbytes // bs)
    bstarts = struct.unpack("<%di" % nblocks, buf[MAX_OVERHEAD:MAX_OVERHEAD + 4 * nblocks])
    out = np.empty(nbytes, np.uint8)
    for b in range(nblocks):
        blen = min(bs, nbytes - b * bs)
        leftover = blen != bs
        nsplit = ts if (h["split"] and not leftover and ts <= MAX_SPLITS and blen // ts >= MIN_BUFFERSIZE) else 1
        ne = blen // nsplit
        pos = bstarts[b]
        chunks = []
        for _ in range(nsplit):
            (cs,) = struct.unpack("<i", buf[pos:pos + 4])
            pos += 4
            if cs < 0 or pos + cs > len(buf):
                raise ValueError("corrupt Blosc block table")
            chunks.append(buf[pos:pos + cs] if cs == ne else z.decompress(buf[pos:pos + cs], ne))
            pos += cs
        blk = np.frombuffer(b"".join(chunks), np.uint8)
        if h["shuffle"] == SHUFFLE and ts > 1:
            blk = _byte_unshuffle(blk, ts)
        elif h["shuffle"] == BITSHUFFLE and blen >= ts:
            blk = _bit_unshuffle(blk, ts)
        out[b * bs:b * bs + blen] = blk
    return out.tobytes()
