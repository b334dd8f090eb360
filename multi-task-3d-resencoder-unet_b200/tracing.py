"""Opt-in NVTX ranges around the network's modules (the reference has no tracing hooks).

    h = tracing.enable_nvtx(model)            # ranges named by module path: "shared_encoder.stages.2.blocks.1", ...
    ... run a step under `ncu --nvtx --nvtx-include "shared_encoder.stages.2/*" ...`
    h.remove()

Forward ranges come from module forward hooks; backward ranges carry the same names with a "bwd:" prefix (full
backward hooks on the block-level modules).  Nothing here touches the kernels or adds synchronisation; with the hooks
removed the model is byte-for-byte the same object.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

_BLOCK_TYPES = ("BasicBlockD", "BottleneckD", "ConvDropoutNormReLU", "StemConv", "Decoder", "Encoder")


class _Handles:
    def __init__(self):
        self.handles = []
        self.names = []

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []


def _push(name):
    def hook(*_):
        torch.cuda.nvtx.range_push(name)
    return hook


def _pop(*_):
    torch.cuda.nvtx.range_pop()


def enable_nvtx(model: torch.nn.Module, types: Optional[Iterable[str]] = None, backward: bool = True) -> _Handles:
    """Register push / pop hooks on every sub-module whose class name is in `types` (default: blocks, conv units,
    encoder, decoders).  Returns an object whose .remove() detaches them; .names lists the instrumented paths."""
    want = tuple(types) if types is not None else _BLOCK_TYPES
    out = _Handles()
    seen = set()
    for name, mod in model.named_modules():
        if type(mod).__name__ not in want or id(mod) in seen or not name:
            continue
        seen.add(id(mod))              # the decoders hold the shared encoder as a child: instrument it once
        out.names.append(name)
        out.handles.append(mod.register_forward_pre_hook(_push(name)))
        out.handles.append(mod.register_forward_hook(_pop))
        if backward:
            out.handles.append(mod.register_full_backward_pre_hook(_push("bwd:" + name)))
            out.handles.append(mod.register_full_backward_hook(_pop))
    return out
