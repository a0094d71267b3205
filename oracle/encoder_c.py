"""ORACLE (test infrastructure) -- ctypes wrapper over ``oracle/encoder_c.c``.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs may import this."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_encoder.so")
_lib = None
IMG = 227


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "encoder_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle_encoder.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        for name in ("svo_encode_bits", "svo_encode_f32", "svo_encode_digest"):
            getattr(_lib, name).restype = None
            getattr(_lib, name).argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _lib.svo_set_threads.argtypes = [ctypes.c_int]
        _lib.svo_get_threads.restype = ctypes.c_int
    return _lib


def _rows(rows):
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    assert rows.ndim == 2 and rows.shape[1] == 12
    return rows


def encode_bits(rows) -> np.ndarray:
    rows = _rows(rows)
    out = np.empty((rows.shape[0], 3, IMG, IMG), dtype=np.uint8)
    lib().svo_encode_bits(rows.ctypes.data, rows.shape[0], out.ctypes.data)
    return out


def encode_f32(rows) -> np.ndarray:
    rows = _rows(rows)
    out = np.empty((rows.shape[0], IMG, IMG, 3), dtype=np.float32)
    lib().svo_encode_f32(rows.ctypes.data, rows.shape[0], out.ctypes.data)
    return out


def encode_digest(rows) -> np.ndarray:
    rows = _rows(rows)
    out = np.empty(rows.shape[0], dtype=np.uint64)
    lib().svo_encode_digest(rows.ctypes.data, rows.shape[0], out.ctypes.data)
    return out


def set_threads(t: int) -> None:
    lib().svo_set_threads(int(t))


def get_threads() -> int:
    return int(lib().svo_get_threads())
