"""galah_b200 -- B200-native implementation of Galah's two-stage dereplication hot path.

The product is libgalah_b200.so (hand-written sm_100a CUDA kernels + C ABI + C++ host side,
see include/galah_b200.h); this package is the thin Python face used by the tests, bench.py
and the multi-GPU driver.  There is no CPU fallback.
"""
from ._native import GalahB200Error, LIB_PATH, exported_symbols  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import PAIR_DTYPE, ROW_BLOCK  # noqa: F401
