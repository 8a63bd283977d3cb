"""ccs_b200 -- B200-native CCS consensus engine (per-ZMW hot path: draft -> Arrow polish).

The product is the C ABI in include/ccsgpu.h (libccsgpu.so: hand-written sm_100a CUDA
kernels + C++ host orchestration).  This package is the thin ctypes mirror used by the
tests and bench.py.
"""
from ._lib import lib, simlib, LIB_PATH, SIM_LIB_PATH, CcsLibraryMissing  # noqa: F401
