"""Loader for oracle/c/oracle.c (plain-C restatement of the surrogate forward pass).

TEST INFRASTRUCTURE ONLY.  Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  ``build()`` compiles it with gcc (OpenMP) into oracle/_build/.
"""
from __future__ import annotations

import ctypes
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_void_p
from pathlib import Path
from typing import Sequence

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _DIR / "c" / "oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "-B", "_build/liboracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        h = ctypes.CDLL(str(_SO))
        h.oracle_encode.restype = c_int64
        h.oracle_encode.argtypes = [c_char_p, c_int64, c_int, c_char_p, c_int, c_void_p]
        h.oracle_cnn_forward.restype = None
        h.oracle_cnn_forward.argtypes = [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int,
                                         POINTER(c_void_p), c_void_p]
        h.oracle_mlp_forward.restype = None
        h.oracle_mlp_forward.argtypes = [c_void_p, c_int64, c_int, c_int, c_int, c_int, POINTER(c_void_p), c_void_p]
        _lib = h
    return _lib


def _ptrs(arrays):
    keep = [np.ascontiguousarray(np.asarray(a, dtype=np.float32)) for a in arrays]
    arr_t = c_void_p * len(keep)
    return keep, arr_t(*[a.ctypes.data_as(c_void_p).value for a in keep])


def encode(chars: bytes, n: int, length: int, alphabet: str) -> np.ndarray:
    out = np.zeros((n, length), dtype=np.uint8)
    bad = lib().oracle_encode(chars, n, length, alphabet.encode("latin-1"), len(alphabet), out.ctypes.data_as(c_void_p))
    if bad >= 0:
        raise ValueError(f"substring not found at flat position {bad}")
    return out


def cnn_forward(idx: np.ndarray, member_weights: Sequence[Sequence[np.ndarray]], kernel_size: int = 5) -> np.ndarray:
    """``idx`` uint8[N, L]; ``member_weights``: list (one per ensemble member) of the 12 Keras arrays."""
    idx = np.ascontiguousarray(idx, dtype=np.uint8)
    n, length = idx.shape
    w0 = member_weights[0]
    a, f, h = w0[0].shape[1], w0[0].shape[2], w0[6].shape[1]
    flat = [w for member in member_weights for w in member]
    keep, ptrs = _ptrs(flat)
    out = np.empty(n, dtype=np.float32)
    lib().oracle_cnn_forward(idx.ctypes.data_as(c_void_p), n, length, a, f, h, kernel_size, len(member_weights),
                             ptrs, out.ctypes.data_as(c_void_p))
    return out


def mlp_forward(idx: np.ndarray, member_weights: Sequence[Sequence[np.ndarray]]) -> np.ndarray:
    idx = np.ascontiguousarray(idx, dtype=np.uint8)
    n, length = idx.shape
    w0 = member_weights[0]
    h = w0[0].shape[1]
    a = w0[0].shape[0] // length
    flat = [w for member in member_weights for w in member]
    keep, ptrs = _ptrs(flat)
    out = np.empty(n, dtype=np.float32)
    lib().oracle_mlp_forward(idx.ctypes.data_as(c_void_p), n, length, a, h, len(member_weights), ptrs,
                             out.ctypes.data_as(c_void_p))
    return out
