"""ctypes loader for libccsgpu.so (the C ABI declared in include/ccsgpu.h).

There is no Python or CPU fallback: if the shared library is missing this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libccsgpu.so")
_lib = None


class CcsLibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CcsLibraryMissing(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C ccs_b200/csrc`). ccs_b200 has no CPU fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


SIM_LIB_PATH = os.path.join(_HERE, "libccssim.so")
_simlib = None


def simlib():
    """The synthetic-data generator (include/ccssim.h): its own library, so that loading test inputs never maps the
    product (the reference arm of bench.py uses this and the CPU oracle only)."""
    global _simlib
    if _simlib is None:
        if not os.path.exists(SIM_LIB_PATH):
            raise CcsLibraryMissing(f"{SIM_LIB_PATH} not built: run `make -C ccs_b200/csrc`.")
        _simlib = ctypes.CDLL(SIM_LIB_PATH)
    return _simlib
