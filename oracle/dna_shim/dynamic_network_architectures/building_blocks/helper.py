"""ORACLE SCAFFOLDING (test infrastructure, never imported by the product).

Shim for the un-vendored third-party package `dynamic_network_architectures`
that /root/reference/builders/resblocks.py:9 imports.  The reference vendors
twins of these two helpers in builders/utils.py:128,268 — re-export those.
"""
from builders.utils import maybe_convert_scalar_to_list, get_matching_pool_op  # noqa: F401
