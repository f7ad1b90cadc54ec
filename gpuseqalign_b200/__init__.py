"""gpuseqalign_b200 -- B200-native Needleman-Wunsch (linear gap) alignment engine.

The product is the C-ABI shared library ``libnwb200.so`` (``include/nwb200.h``) built from the
hand-written sm_100a kernels in ``csrc/``; this package is the thin Python host layer used by the
tests and the benchmark.  There is no CPU fallback: importing works anywhere, but creating an
engine without the built library or without a B200 raises.
"""
from .capi import NwStat, NwB200Error, Engine, HeaderInfo, Params, lib_path, load_library  # noqa: F401
from . import formats  # noqa: F401

__all__ = ["NwStat", "NwB200Error", "Engine", "HeaderInfo", "Params", "formats", "lib_path", "load_library"]
