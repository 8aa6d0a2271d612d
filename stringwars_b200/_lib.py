"""ctypes binding of include/stringwars_b200.h. Loading fails loudly: there is no Python/CPU fallback."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libstringwars_b200.so")

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_i8p = ctypes.POINTER(ctypes.c_int8)
c_err = ctypes.POINTER(ctypes.c_char_p)

class TapeView(ctypes.Structure):
    """`swb_tape_view_t`: AnyBytesTape::View64 / View32 as a plain struct (bench.rs:141-142)."""
    _fields_ = [("data", ctypes.c_void_p), ("offsets", ctypes.c_void_p), ("n", ctypes.c_size_t), ("offset_bits", ctypes.c_int)]


c_viewp = ctypes.POINTER(TapeView)
RESULT_64, RESULT_32, RESULT_16 = 0, 1, 2

# every symbol include/stringwars_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "swb_version": (ctypes.c_int, []),
    "swb_device_count": (ctypes.c_int, []),
    "swb_multiprocessor_count": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int), c_err]),
    "swb_launch_count": (ctypes.c_uint64, []),
    "swb_scope_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), c_err]),
    "swb_scope_destroy": (None, [ctypes.c_void_p]),
    "swb_scope_set_transport": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_err]),
    "swb_scope_num_devices": (ctypes.c_int, [ctypes.c_void_p]),
    "swb_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), c_err]),
    "swb_host_free": (None, [ctypes.c_void_p]),
    "swb_unified_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), c_err]),
    "swb_unified_free": (None, [ctypes.c_void_p]),
    "swb_scope_create_devices": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), c_err]),
    "swb_compute_pairs_view": (ctypes.c_int, [ctypes.c_void_p, c_viewp, c_viewp, ctypes.c_void_p, ctypes.c_int, c_err]),
    "swb_compute_cross_view": (ctypes.c_int, [ctypes.c_void_p, c_viewp, c_viewp, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, c_err]),
    "swb_compute_pairs_device_view": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_viewp, c_viewp, ctypes.c_void_p, ctypes.c_int,
                                                     ctypes.c_void_p, c_err]),
    "swb_tape_upload": (ctypes.c_int, [ctypes.c_void_p, c_viewp, ctypes.POINTER(ctypes.c_void_p), c_err]),
    "swb_tape_destroy": (None, [ctypes.c_void_p]),
    "swb_tape_len": (ctypes.c_size_t, [ctypes.c_void_p]),
    "swb_tape_device_view": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, c_viewp, c_err]),
    "swb_compute_pairs_tape": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                              ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, c_err]),
    "swb_compute_cross_tape": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p,
                                              ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, c_err]),
    "swb_compute_pairs_peer": (ctypes.c_int, [ctypes.c_void_p, c_viewp, c_viewp, ctypes.c_void_p, ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_double), c_err]),
    "swb_compute_pairs_scatter": (ctypes.c_int, [ctypes.c_void_p, c_viewp, c_viewp, ctypes.c_void_p, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_double), c_err]),
    "swb_levenshtein_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_void_p), c_err]),
    "swb_levenshtein_utf8_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), c_err]),
    "swb_needleman_wunsch_create": (ctypes.c_int, [ctypes.c_void_p, c_u8p, c_i8p, ctypes.c_int, ctypes.c_int,
                                                   ctypes.POINTER(ctypes.c_void_p), c_err]),
    "swb_smith_waterman_create": (ctypes.c_int, [ctypes.c_void_p, c_u8p, c_i8p, ctypes.c_int, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_void_p), c_err]),
    "swb_engine_destroy": (None, [ctypes.c_void_p]),
    "swb_engine_set_algorithm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_err]),
    "swb_engine_set_gap_convention": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_err]),
    "swb_engine_last_kernel": (ctypes.c_char_p, [ctypes.c_void_p]),
    "swb_engine_set_profiling": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_err]),
    "swb_engine_kernel_time": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                              ctypes.POINTER(ctypes.c_uint64), c_err]),
    "swb_compute_pairs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, c_err]),
    "swb_compute_cross": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                         ctypes.c_size_t, c_err]),
    "swb_levenshtein_pair": (ctypes.c_uint64, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p,
                                               ctypes.c_size_t, c_err]),
    "swb_compute_pairs_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                                ctypes.c_void_p, c_err]),
    "swb_compute_cross_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, c_err]),
    "swb_load_phase_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                             ctypes.c_void_p, c_err]),
}

STATUS_NAMES = {0: "OK", 1: "BAD_ALLOC", 2: "INVALID_ARG", 3: "NO_DEVICE", 4: "CUDA_ERROR", 5: "TOO_LONG",
                6: "UNSUPPORTED", 7: "NCCL_ERROR"}

_lib: Optional[ctypes.CDLL] = None


class Error(RuntimeError):
    """A non-zero `swb_status_t`, carrying the library's message (reference: `Result<_, E: Display>`)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"stringwars_b200: {STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
        self.status_name = STATUS_NAMES.get(status, str(status))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m stringwars_b200.build` "
                "(stringwars_b200 has no CPU or pure-Python fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here == header/library mismatch; let it surface
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def check(status: int, err: ctypes.c_char_p) -> None:
    if status != 0:
        raise Error(status, (err.value or b"").decode("utf-8", "replace"))
