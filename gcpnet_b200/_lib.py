"""Loader of the C-ABI library (gcpnet_b200/libgcpnet_b200.so, built by gcpnet_b200/build.py).

There is no fallback: if the library is missing or does not export every symbol of
include/gcpnet_b200.h, importing the product path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _cabi

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgcpnet_b200.so")
_lock = threading.Lock()
_lib = None


class GcpnetError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Return the loaded library (loading it on first use)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_LIB_PATH):
                raise GcpnetError(
                    f"{_LIB_PATH} not found: build it with `python -m gcpnet_b200.build` "
                    "(there is no CPU or PyTorch fallback for this path)")
            lib = C.CDLL(_LIB_PATH)
            missing = [s for s in _cabi.EXPORTS if not hasattr(lib, s)]
            if missing:
                raise GcpnetError(f"{_LIB_PATH} does not export {missing}")
            _cabi.declare(lib)
            _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gcpnet_last_error()
        raise GcpnetError(f"{what}: {msg.decode() if msg else 'unknown error'}")
