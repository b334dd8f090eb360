"""ORACLE SCAFFOLDING — build-container only (needs /root/reference, which does not exist on
the GPU box).  Imports the UNMODIFIED reference `builders/` package through the 3-file
`dynamic_network_architectures` shim in oracle/dna_shim, so the
restatement in resenc_oracle.py and the golden fixtures can be pinned against the real code.
"""
import contextlib
import io
import os
import sys
import warnings
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("RESENC_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dna_shim")


_REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")


def use_ref_copy():
    """Point the loader at `oracle/_ref/reference` (the byte-for-byte copy made by oracle/build_ref.py, which travels
    to the GPU box) when /root/reference itself is absent.  Returns True when a reference tree is importable."""
    global REFERENCE_ROOT
    if not available() and os.path.isdir(os.path.join(_REF_COPY, "builders")):
        REFERENCE_ROOT = _REF_COPY
    return available()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "builders"))


def _ensure_path():
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    # the reference's top-level packages are called `builders`, `inference`, ...; if the
    # product's drop-in alias is installed under the same name, refuse rather than mix them.
    m = sys.modules.get("builders")
    if m is not None and not getattr(m, "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("a non-reference module named 'builders' is already imported")
    for p in (REFERENCE_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.append(p)


def make_mgr(patch, tasks, in_channels=1, batch=1, model_config=None, autoconfigure=True):
    return SimpleNamespace(tasks=tasks, train_patch_size=list(patch), train_batch_size=batch,
                           in_channels=in_channels, vram_max=16.0, autoconfigure=autoconfigure,
                           model_config=dict(model_config or {}))


def build_reference(mgr, se_reduce_dims="all", quiet=True):
    """NetworkFromConfig(mgr) from /root/reference/builders/build_network_from_config.py:20."""
    _ensure_path()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        from builders.build_network_from_config import NetworkFromConfig
        import dynamic_network_architectures.building_blocks.regularization as reg
    reg.SE_REDUCE_DIMS = se_reduce_dims
    ctx = contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()
    with ctx:
        return NetworkFromConfig(mgr)


def reference_function(relpath, name):
    """Extract one top-level function from a reference file whose module cannot be imported
    here (helpers.py imports zarr/fsspec at the top, helpers.py:4-5)."""
    import ast
    src = open(os.path.join(REFERENCE_ROOT, relpath)).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            ns = {}
            exec(compile(ast.Module([node], []), relpath, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def reference_module(relpath, modname):
    import importlib.util
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
