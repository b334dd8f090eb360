"""ctypes binding of the C ABI declared in include/resenc_b200.h.

The shared library is built in-tree (``csrc/libresenc_b200.so``) by ``build.py`` /
``__graft_entry__.build()``.  There is no fallback: if the library is missing or a call fails
this module raises, it never routes to PyTorch or the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libresenc_b200.so")

IMPL_AUTO, IMPL_MMA_SYNC, IMPL_TCGEN05, IMPL_TCGEN05_SPLITK, IMPL_TCGEN05_SLAB = 0, 1, 2, 3, 4
_IMPL_NAMES = {"auto": IMPL_AUTO, "mma": IMPL_MMA_SYNC, "mma_sync": IMPL_MMA_SYNC, "tc5": IMPL_TCGEN05,
               "tcgen05": IMPL_TCGEN05}


class ResencLibraryError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "nsrc", "srcC0", "srcC1", "NB", "ID", "IH", "IW", "tapD", "tapH", "tapW", "offD", "offH", "offW",
        "istrD", "istrH", "istrW", "OD", "OH", "OW", "Nout", "mode", "ostrD", "ostrH", "ostrW",
        "ooffD", "ooffH", "ooffW", "FD", "FH", "FW", "outC0", "outC1", "psC", "psD", "psH", "psW",
        "impl", "splitK", "outF32")]


class WgradDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "PC", "nq", "QC0", "QC1", "NB", "GD", "GH", "GW", "QD", "QH", "QW", "tapD", "tapH", "tapW",
        "offD", "offH", "offW", "istrD", "istrH", "istrW", "splits", "impl")]


class BlendTarget(C.Structure):
    _fields_ = [("pred", C.c_void_p), ("sum", C.c_void_p), ("C", C.c_int), ("activation", C.c_int)]


_P, _I, _LL, _F, _D, _SZ = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); every symbol include/resenc_b200.h declares
SIGNATURES = {
    "rb_last_error": (C.c_char_p, []),
    "rb_version": (_I, []),
    "rb_device_error": (_I, [_P]),
    "rb_launch_count": (_LL, []),
    "rb_debug_counters": (_I, [_P]),
    "rb_conv_gather_workspace": (_SZ, [C.POINTER(ConvDesc)]),
    "rb_conv_gather": (_I, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "rb_conv_gather_tc5_supported": (_I, [C.POINTER(ConvDesc)]),
    "rb_conv_gather_plan": (_I, [C.POINTER(ConvDesc)]),
    "rb_wgrad_gather": (_I, [C.POINTER(WgradDesc), _P, _P, _P, _P, _P]),
    "rb_plane_reduce": (_I, [_I, _P, _I, _P, _P, _P, _P, _P, _I, _LL, _I, _I, _I, _F, _P]),
    "rb_in_finalize_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _D, _D, _P]),
    "rb_in_finalize_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _D, _P]),
    "rb_norm_act_fwd": (_I, [_P, _I, _P, _P, _P, _P, _I, _LL, _I, _I, _I, _I, _F, _P]),
    "rb_norm_act_bwd": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _LL, _I, _I, _I, _I, _F, _P]),
    "rb_avgpool_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rb_avgpool_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rb_head_fwd": (_I, [_P, _P, _P, _P, _I, _LL, _I, _I, _I, _P]),
    "rb_head_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _LL, _I, _I, _P]),
    "rb_pack_conv_dgrad_merged": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "rb_grad_sumsq": (_I, [_P, _I, _P, _P]),
    "rb_adamw_clip_step": (_I, [_P, _I, _P, _P, _P, _F, _F, _F, _F, _F, _P]),
    "rb_adamw_clip_pack_step": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _F, _F, _F, _F, _F, _P]),
    "rb_norm_act_head_fwd": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _I, _LL, _I, _I, _I, _F, _I, _P]),
    "rb_stem_im2col": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rb_split_apply": (_I, [_P, _P, _P, _P, _P, _I, _LL, _I, _I, _F, _P]),
    "rb_avgpool_split": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rb_stem_im2col_split": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rb_pack_conv_weights": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "rb_unpack_wgrad": (_I, [_P, _P, _I, _I, _I, _P]),
    "rb_ncdhw_to_cl": (_I, [_P, _P, _I, _I, _LL, _P]),
    "rb_cl_to_ncdhw": (_I, [_P, _P, _I, _I, _LL, _P]),
    "rb_blend_accumulate": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rb_blend_finalize_cast": (_I, [_P, _P, _P, _P, _LL, _I, _I, _P]),
    "rb_blend_add": (_I, [_P, _P, _LL, _P]),
    "rb_blend_accumulate_multi": (_I, [C.POINTER(BlendTarget), _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rb_blend_finalize_cast2": (_I, [_P, _LL, _P, _P, _P, _LL, _I, _I, _P]),
    "rb_extract_patches": (_I, [_P, _I, _I, _I, _I, C.POINTER(C.c_int), _I, _I, _I, _I, _I, _P, _P, _P]),
    "rb_extract_patch": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "rb_loss_bce_dice_reduce": (_I, [_P, _P, _P, _I, _I, _LL, _F, _P]),
    "rb_loss_bce_dice_grad": (_I, [_P, _P, _P, _P, _P, _I, _I, _LL, _F, _F, _F, _F, _P]),
    "rb_loss_cosine_reduce": (_I, [_P, _P, _P, _I, _LL, _P]),
    "rb_loss_cosine_grad": (_I, [_P, _P, _P, _P, _P, _I, _LL, _P]),
}

_lib = None


def load():
    """Load the C-ABI library (once).  Raises ResencLibraryError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ResencLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().rb_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise ResencLibraryError(f"{what} failed ({rc}): {last_error()}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device=None) -> int:
    """cudaStream_t of torch's current stream on `device` (default: the current device).  `torch.cuda.current_stream()`
    builds a Python Stream object per call (~4 us, ~900 calls per eager training step); the raw accessor is a plain C call."""
    if _raw_stream is None:
        return torch.cuda.current_stream(device).cuda_stream
    if device is None:
        idx = torch._C._cuda_getDevice()
    elif isinstance(device, int):
        idx = device
    else:
        idx = torch.device(device).index
        if idx is None:
            idx = torch._C._cuda_getDevice()
    return _raw_stream(idx)


def ptr(t):
    return None if t is None else t.data_ptr()


def impl_code(name) -> int:
    if isinstance(name, int):
        return name
    return _IMPL_NAMES[str(name).lower()]


def default_impl() -> int:
    """Conv implementation selector; RESENC_CONV_IMPL=auto|mma|tc5 overrides (testing)."""
    return impl_code(os.environ.get("RESENC_CONV_IMPL", "auto"))


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise ResencLibraryError(
            f"{what}: tensor is on {t.device}; the B200 hot path runs on CUDA only (no CPU fallback)")


def device_error_check():
    check(load().rb_device_error(stream_ptr()), "rb_device_error")


def launch_count() -> int:
    return int(load().rb_launch_count())
