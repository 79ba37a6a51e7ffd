"""Loads the CUDA product library (rusterix_b200/librxcuda.so, built in-tree by build.py).
There is no CPU fallback: if the library is missing or does not export every symbol of
include/rxcuda.h this raises, and every entry point of the package fails with it."""
import ctypes
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RXC_LIB") or os.path.join(_HERE, "librxcuda.so")   # RXC_LIB: an experiment build (build.py, RX_BUILD_TAG)
_lib = None


class RxcError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{_abi.STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m rusterix_b200.build` "
                "(nvcc, sm_100a). rusterix_b200 has no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        _abi.bind(lib)
        if lib.rxc_abi_version() != _abi.RXC_ABI_VERSION:
            raise RuntimeError("librxcuda.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib
