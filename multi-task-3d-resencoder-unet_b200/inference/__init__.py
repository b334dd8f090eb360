"""Sliding-window inference: enumeration, importance map, device blend, z-slab sharding."""
from .sliding_window import (DeviceVolume, SlabBlender, SlidingWindowInferer, all_positions, axis_positions,  # noqa: F401
                             compute_gaussian_3d, generate_positions, get_gaussian_map, merge_slabs, patch_steps,
                             plan_slab_exchange, shard_z_starts)
from .zarr_writer import FinalVolumeWriter, ZarrArrayReader, ZarrArrayWriter, open_zarr_array  # noqa: F401,E402
