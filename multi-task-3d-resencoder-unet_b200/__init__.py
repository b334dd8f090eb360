"""B200-native (sm_100a) implementation of the multi-task 3D ResEnc U-Net hot path.

The directory name contains hyphens, so import it with
    importlib.import_module("multi-task-3d-resencoder-unet_b200")
or through the `resenc_b200` alias module at the repository root.
"""
from . import _lib, ops  # noqa: F401
from . import builders, inference  # noqa: F401
from .builders import NetworkFromConfig  # noqa: F401

__all__ = ["NetworkFromConfig", "builders", "inference", "ops"]
