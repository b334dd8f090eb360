"""Importable alias of the package directory `multi-task-3d-resencoder-unet_b200/` (whose name is
not a Python identifier).  `import resenc_b200 as rb; rb.NetworkFromConfig(mgr)`;
`resenc_b200.install_as_builders()` registers the drop-in under the reference's module name
`builders`, so the reference's train.py / inference.py import it unchanged."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("multi-task-3d-resencoder-unet_b200")
sys.modules.setdefault("resenc_b200_pkg", _pkg)

NetworkFromConfig = _pkg.NetworkFromConfig
builders = _pkg.builders
inference = _pkg.inference
ops = _pkg.ops
losses = importlib.import_module(_pkg.__name__ + ".losses")
optim = importlib.import_module(_pkg.__name__ + ".optim")
parallel = importlib.import_module(_pkg.__name__ + ".parallel")
precise = importlib.import_module(_pkg.__name__ + ".precise")
training = importlib.import_module(_pkg.__name__ + ".training")
tracing = importlib.import_module(_pkg.__name__ + ".tracing")
_lib = _pkg._lib


def install_as_builders():
    """Make `import builders` / `from builders.build_network_from_config import NetworkFromConfig`
    resolve to the B200 drop-in (call before importing the reference's train.py)."""
    base = _pkg.__name__ + ".builders"
    sys.modules["builders"] = sys.modules[base]
    for sub in ("build_network_from_config", "encoder", "decoder", "resblocks", "simple_conv_blocks", "utils"):
        sys.modules["builders." + sub] = sys.modules[base + "." + sub]
    return sys.modules["builders"]
