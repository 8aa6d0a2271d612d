"""ctypes front-end of oracle/liblev_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module. The product package (stringwars_b200/) never does. Parity status: unpinned by the reference
(see the header of lev_oracle.c).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblev_oracle.so")
_lib: Optional[ctypes.CDLL] = None

WAGNER_FISCHER, RAPIDFUZZ_CLASS, UTF8 = 0, 1, 2
GAP_STRINGZILLA, GAP_BIO = 0, 1

_u8p = ctypes.POINTER(ctypes.c_uint8)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i8p = ctypes.POINTER(ctypes.c_int8)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "lev_oracle.c"), os.path.join(_HERE, "align_bruteforce.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B", "liblev_oracle.so"], check=True, capture_output=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.orc_lev_bytes.restype = ctypes.c_uint64
        L.orc_lev_bytes.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
        L.orc_rf_lev_bytes.restype = ctypes.c_uint64
        L.orc_rf_lev_bytes.argtypes = L.orc_lev_bytes.argtypes
        L.orc_lev_utf8.restype = ctypes.c_uint64
        L.orc_lev_utf8.argtypes = L.orc_lev_bytes.argtypes
        L.orc_align_score.restype = ctypes.c_int64
        L.orc_align_score.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                                      _u8p, _i8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.orc_align_bruteforce.restype = ctypes.c_int64
        L.orc_align_bruteforce.argtypes = L.orc_align_score.argtypes
        L.orc_lev_pairs.restype = None
        L.orc_lev_pairs.argtypes = [ctypes.c_int, _u8p, _u64p, _u8p, _u64p, ctypes.c_size_t, _u64p, ctypes.c_int]
        L.orc_lev_pairs_timed.restype = ctypes.c_double
        L.orc_lev_pairs_timed.argtypes = [ctypes.c_int, _u8p, _u64p, _u8p, _u64p, ctypes.c_size_t, _u64p, ctypes.c_int,
                                          ctypes.c_int]
        L.orc_lev_cross.restype = None
        L.orc_lev_cross.argtypes = [ctypes.c_int, _u8p, _u64p, ctypes.c_size_t, _u8p, _u64p, ctypes.c_size_t,
                                    _u64p, ctypes.c_size_t, ctypes.c_int]
        L.orc_align_pairs.restype = None
        L.orc_align_pairs.argtypes = [_u8p, _u64p, _u8p, _u64p, ctypes.c_size_t, _u8p, _i8p, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, _i64p, ctypes.c_int]
        L.orc_align_cross.restype = None
        L.orc_align_cross.argtypes = [_u8p, _u64p, ctypes.c_size_t, _u8p, _u64p, ctypes.c_size_t, _u8p, _i8p,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i64p,
                                      ctypes.c_size_t, ctypes.c_int]
        _lib = L
    return _lib


def _p(arr: np.ndarray, ctype):
    return arr.ctypes.data_as(ctypes.POINTER(ctype))


def _tape(data, offsets):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    if data.size == 0:
        data = np.zeros(1, dtype=np.uint8)
    return data, np.ascontiguousarray(offsets, dtype=np.uint64)


def levenshtein(a: bytes, b: bytes, kind: int = WAGNER_FISCHER) -> int:
    f = (lib().orc_lev_bytes, lib().orc_rf_lev_bytes, lib().orc_lev_utf8)[kind]
    return int(f(a, len(a), b, len(b)))


def levenshtein_pairs(a_data, a_offs, b_data, b_offs, kind: int = WAGNER_FISCHER, threads: int = 1) -> np.ndarray:
    a_data, a_offs = _tape(a_data, a_offs)
    b_data, b_offs = _tape(b_data, b_offs)
    n = len(a_offs) - 1
    assert len(b_offs) - 1 == n
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_lev_pairs(kind, _p(a_data, ctypes.c_uint8), _p(a_offs, ctypes.c_uint64), _p(b_data, ctypes.c_uint8),
                        _p(b_offs, ctypes.c_uint64), n, _p(out, ctypes.c_uint64), threads)
    return out


def levenshtein_pairs_timed(a_data, a_offs, b_data, b_offs, kind: int = RAPIDFUZZ_CLASS, threads: int = 1,
                            repeats: int = 1):
    """(seconds, results): wall time measured inside C around `threads` workers each doing `repeats` passes."""
    a_data, a_offs = _tape(a_data, a_offs)
    b_data, b_offs = _tape(b_data, b_offs)
    n = len(a_offs) - 1
    out = np.zeros(n, dtype=np.uint64)
    secs = lib().orc_lev_pairs_timed(kind, _p(a_data, ctypes.c_uint8), _p(a_offs, ctypes.c_uint64),
                                     _p(b_data, ctypes.c_uint8), _p(b_offs, ctypes.c_uint64), n,
                                     _p(out, ctypes.c_uint64), threads, repeats)
    return float(secs), out


def levenshtein_cross(q_data, q_offs, c_data, c_offs, kind: int = WAGNER_FISCHER, threads: int = 1) -> np.ndarray:
    q_data, q_offs = _tape(q_data, q_offs)
    c_data, c_offs = _tape(c_data, c_offs)
    nq, nc = len(q_offs) - 1, len(c_offs) - 1
    out = np.zeros((nq, nc), dtype=np.uint64)
    lib().orc_lev_cross(kind, _p(q_data, ctypes.c_uint8), _p(q_offs, ctypes.c_uint64), nq,
                        _p(c_data, ctypes.c_uint8), _p(c_offs, ctypes.c_uint64), nc, _p(out, ctypes.c_uint64),
                        nc, threads)
    return out


def unary_class_costs(match: int, mismatch: int):
    """Restatement of `unary_class_costs` (similarities/bench.rs:98-108)."""
    byte_to_class = (np.arange(256) % 32).astype(np.uint8)
    costs = np.full((32, 32), mismatch, dtype=np.int8)
    np.fill_diagonal(costs, match)
    return byte_to_class, costs


def align_score(a: bytes, b: bytes, byte_to_class, costs, open: int, extend: int,
                convention: int = GAP_STRINGZILLA, local: bool = False) -> int:
    btc = np.ascontiguousarray(byte_to_class, dtype=np.uint8)
    cst = np.ascontiguousarray(costs, dtype=np.int8)
    return int(lib().orc_align_score(a, len(a), b, len(b), _p(btc, ctypes.c_uint8), _p(cst, ctypes.c_int8),
                                     open, extend, convention, int(local)))


def align_bruteforce(a: bytes, b: bytes, byte_to_class, costs, open: int, extend: int,
                     convention: int = GAP_STRINGZILLA, local: bool = False) -> int:
    """Definition-level score: every alignment of two SHORT strings enumerated and scored (oracle/align_bruteforce.c).
    Shares no code or recurrence with `align_score`."""
    btc = np.ascontiguousarray(byte_to_class, dtype=np.uint8)
    cst = np.ascontiguousarray(costs, dtype=np.int8)
    return int(lib().orc_align_bruteforce(a, len(a), b, len(b), _p(btc, ctypes.c_uint8), _p(cst, ctypes.c_int8),
                                          open, extend, convention, int(local)))


def align_pairs(a_data, a_offs, b_data, b_offs, byte_to_class, costs, open: int, extend: int,
                convention: int = GAP_STRINGZILLA, local: bool = False, threads: int = 1) -> np.ndarray:
    a_data, a_offs = _tape(a_data, a_offs)
    b_data, b_offs = _tape(b_data, b_offs)
    btc = np.ascontiguousarray(byte_to_class, dtype=np.uint8)
    cst = np.ascontiguousarray(costs, dtype=np.int8)
    n = len(a_offs) - 1
    out = np.zeros(n, dtype=np.int64)
    lib().orc_align_pairs(_p(a_data, ctypes.c_uint8), _p(a_offs, ctypes.c_uint64), _p(b_data, ctypes.c_uint8),
                          _p(b_offs, ctypes.c_uint64), n, _p(btc, ctypes.c_uint8), _p(cst, ctypes.c_int8),
                          open, extend, convention, int(local), _p(out, ctypes.c_int64), threads)
    return out


def align_cross(q_data, q_offs, c_data, c_offs, byte_to_class, costs, open: int, extend: int,
                convention: int = GAP_STRINGZILLA, local: bool = False, threads: int = 1) -> np.ndarray:
    q_data, q_offs = _tape(q_data, q_offs)
    c_data, c_offs = _tape(c_data, c_offs)
    btc = np.ascontiguousarray(byte_to_class, dtype=np.uint8)
    cst = np.ascontiguousarray(costs, dtype=np.int8)
    nq, nc = len(q_offs) - 1, len(c_offs) - 1
    out = np.zeros((nq, nc), dtype=np.int64)
    lib().orc_align_cross(_p(q_data, ctypes.c_uint8), _p(q_offs, ctypes.c_uint64), nq, _p(c_data, ctypes.c_uint8),
                          _p(c_offs, ctypes.c_uint64), nc, _p(btc, ctypes.c_uint8), _p(cst, ctypes.c_int8),
                          open, extend, convention, int(local), _p(out, ctypes.c_int64), nc, threads)
    return out
