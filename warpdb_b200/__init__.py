"""warpdb_b200 -- B200-native (sm_100a) execution core behind WarpDB's query API.

Layout: csrc/ (CUDA kernels + the C ABI of include/warpcore.h, built into libwarpcore.so),
_core.py (ctypes binding), table.py / db.py (host-side mirror of the reference's WarpDB class and
pywarpdb module), sharded.py (one process per GPU over torch.distributed).
"""
from . import _core  # noqa: F401

__all__ = ["_core"]
