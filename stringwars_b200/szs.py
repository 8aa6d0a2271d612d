"""Host-side mirror of the engine API the reference harness drives (`stringzilla::szs`).

Same names, argument meaning and error behaviour as the call sites in StringWars similarities/bench.rs and
similarities/bench.py, so the parity tests read like the reference's own usage:

    scope  = DeviceScope.gpu_device(0)                         # bench.rs:379
    engine = LevenshteinDistances(scope, 0, 1, 1, 1)           # bench.rs:390-393
    matrix = engine.compute(scope, queries, candidates)        # bench.rs:587-589  (cross product, allocates)
    engine.compute_into(scope, queries, candidates, matrix)    # bench.rs:599-603  (preallocated)
    out    = engine.compute_pairs(scope, a, b)                 # element-wise pairs (BASELINE.json configs;
                                                               #   cuDF precedent bench.py:596-604)

Everything executes in the native CUDA library through its C ABI (include/stringwars_b200.h); this module holds
no compute and no fallback. Errors surface as `stringwars_b200.Error` carrying the native message.
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Optional

import numpy as np

from . import _lib
from .tape import BytesTape, TapeLike, as_tape, unified_mat

_RESULT_TYPE = {8: _lib.RESULT_64, 4: _lib.RESULT_32, 2: _lib.RESULT_16}

ALGO_AUTO, ALGO_WAVEFRONT, ALGO_BITPARALLEL = 0, 1, 2
GAP_STRINGZILLA, GAP_BIO = 0, 1
TRANSPORT_NCCL, TRANSPORT_COPY_ENGINE = 0, 1


def device_count() -> int:
    return int(_lib.lib().swb_device_count())


def multiprocessor_count(device: int = 0) -> int:
    """`gpu_multiprocessor_count(device)` (utils.rs:826-836, utils.py:212-237)."""
    out = ctypes.c_int()
    err = ctypes.c_char_p()
    _lib.check(_lib.lib().swb_multiprocessor_count(device, ctypes.byref(out), ctypes.byref(err)), err)
    return out.value


def launch_count() -> int:
    return int(_lib.lib().swb_launch_count())


class DeviceScope:
    """`szs::DeviceScope`. Only GPU scopes exist; `cpu_cores` raises (no CPU path in this library)."""

    def __init__(self, first_device: int = 0, num_devices: int = 1, devices=None):
        self._h = ctypes.c_void_p()
        err = ctypes.c_char_p()
        if devices is not None:
            devices = [int(d) for d in devices]
            arr = (ctypes.c_int * len(devices))(*devices)
            _lib.check(_lib.lib().swb_scope_create_devices(arr, len(devices), ctypes.byref(self._h), ctypes.byref(err)), err)
            first_device, num_devices = (devices[0] if devices else 0), len(devices)
        else:
            _lib.check(_lib.lib().swb_scope_create(first_device, num_devices, ctypes.byref(self._h), ctypes.byref(err)), err)
        self.first_device, self.num_devices = first_device, num_devices
        self._engines = weakref.WeakSet()  # engines hold device state owned by the scope: they are closed first
        self._tapes = weakref.WeakSet()

    @classmethod
    def gpu_device(cls, device: int = 0) -> "DeviceScope":
        return cls(device, 1)

    @classmethod
    def gpu_devices(cls, first_device: int, num_devices: int) -> "DeviceScope":
        return cls(first_device, num_devices)

    @classmethod
    def device_list(cls, devices) -> "DeviceScope":
        """Explicit ordinals; an ordinal may repeat (one GPU then plays several devices of the sharding path)."""
        return cls(devices=list(devices))

    def set_transport(self, transport: int) -> "DeviceScope":
        """What carries the chunks of `compute_pairs_scatter`: TRANSPORT_NCCL (default) or TRANSPORT_COPY_ENGINE."""
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_scope_set_transport(self._h, transport, ctypes.byref(err)), err)
        return self

    def upload(self, tape: TapeLike) -> "DeviceTape":
        """Resident copy of a tape on every device of the scope (what `UnifiedAlloc` gives the reference after the
        first touch): later `compute_*` calls on it move no input bytes."""
        return DeviceTape(self, as_tape(tape))

    @classmethod
    def cpu_cores(cls, _count: int) -> "DeviceScope":
        raise _lib.Error(6, "stringwars_b200 has no CPU engines; use a GPU scope")

    def close(self) -> None:
        if self._h:
            for engine in list(self._engines):
                engine.close()
            for tape in list(self._tapes):
                tape.close()
            _lib.lib().swb_scope_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(arr: np.ndarray) -> int:
    return arr.ctypes.data


class DeviceTape:
    """`swb_tape_t`: a tape uploaded once to every device of a scope. `subview(lo, hi)` is a zero-copy range of it."""

    def __init__(self, scope: DeviceScope, tape: BytesTape, _parent: "DeviceTape" = None, _lo: int = 0, _hi: int = None):
        self._scope = scope
        if _parent is not None:
            self._root, self._h, self.lo, self.hi = _parent._root, _parent._h, _lo, _hi
            return
        self._root = self
        self._h = ctypes.c_void_p()
        err = ctypes.c_char_p()
        view = tape.view()
        _lib.check(_lib.lib().swb_tape_upload(scope._h, ctypes.byref(view), ctypes.byref(self._h), ctypes.byref(err)), err)
        self.lo, self.hi = 0, len(tape)
        scope._tapes.add(self)

    def __len__(self) -> int:
        return self.hi - self.lo

    def subview(self, lo: int, hi: int) -> "DeviceTape":
        if not (0 <= lo <= hi <= len(self)):
            raise IndexError(f"subview [{lo}, {hi}) outside tape of {len(self)}")
        return DeviceTape(self._scope, None, _parent=self, _lo=self.lo + lo, _hi=self.lo + hi)

    def device_view(self, device_index: int = 0) -> "_lib.TapeView":
        view = _lib.TapeView()
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_tape_device_view(self._root._h, device_index, self.lo, len(self), ctypes.byref(view), ctypes.byref(err)), err)
        return view

    def close(self) -> None:
        if self._root is self and self._h:
            _lib.lib().swb_tape_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check_out(out: np.ndarray, base_dtype, n: int) -> int:
    """Results may be narrower than usize / isize (`swb_result_type_t`): same signedness, 8 / 4 / 2 bytes."""
    kind = np.dtype(base_dtype).kind
    if out.dtype.kind != kind or out.dtype.itemsize not in _RESULT_TYPE or out.size < n or not out.flags.c_contiguous:
        raise ValueError(f"out must be a contiguous {'unsigned' if kind == 'u' else 'signed'} 64/32/16-bit array of >= {n} elements")
    return _RESULT_TYPE[out.dtype.itemsize]


class _Engine:
    out_dtype = np.uint64

    def __init__(self, scope: DeviceScope):
        self._scope = scope  # keeps the scope alive for the engine's lifetime
        self._h = ctypes.c_void_p()
        scope._engines.add(self)

    # ---- options
    def set_algorithm(self, algorithm: int) -> "_Engine":
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_engine_set_algorithm(self._h, algorithm, ctypes.byref(err)), err)
        return self

    def set_gap_convention(self, convention: int) -> "_Engine":
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_engine_set_gap_convention(self._h, convention, ctypes.byref(err)), err)
        return self

    def set_profiling(self, enabled: bool) -> "_Engine":
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_engine_set_profiling(self._h, int(enabled), ctypes.byref(err)), err)
        return self

    def kernel_time(self):
        """(summed compute-kernel device time in ms, launches) since the last call; see the C header."""
        ms, n = ctypes.c_double(), ctypes.c_uint64()
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_engine_kernel_time(self._h, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(err)), err)
        return ms.value, int(n.value)

    @property
    def last_kernel(self) -> str:
        return (_lib.lib().swb_engine_last_kernel(self._h) or b"").decode()

    # ---- host tapes
    def compute_pairs(self, scope: DeviceScope, a, b, out: Optional[np.ndarray] = None) -> np.ndarray:
        """out[i] = score(a[i], b[i]) — the batched form of `(a:&[u8], b:&[u8]) -> usize` (bench.rs:416-419).
        `a` / `b`: BytesTape (any memory, u64 or u32 offsets), sequences of strings, or DeviceTape (resident)."""
        resident = isinstance(a, DeviceTape) or isinstance(b, DeviceTape)
        if resident and not (isinstance(a, DeviceTape) and isinstance(b, DeviceTape)):
            raise ValueError("either both tapes are resident (DeviceTape) or neither")
        ta, tb = (a, b) if resident else (as_tape(a), as_tape(b))
        if len(ta) != len(tb):
            raise ValueError(f"tapes differ in length: {len(ta)} vs {len(tb)}")
        n = len(ta)
        if out is None:
            out = np.zeros(n, dtype=self.out_dtype)
        rtype = _check_out(out, self.out_dtype, n)
        err = ctypes.c_char_p()
        if resident:
            _lib.check(_lib.lib().swb_compute_pairs_tape(self._h, ta._root._h, ta.lo, tb._root._h, tb.lo, n, _ptr(out), rtype,
                                                         ctypes.byref(err)), err)
        else:
            va, vb = ta.view(), tb.view()
            _lib.check(_lib.lib().swb_compute_pairs_view(self._h, ctypes.byref(va), ctypes.byref(vb), _ptr(out), rtype, ctypes.byref(err)), err)
        return out

    def compute_into(self, scope: DeviceScope, queries, candidates, matrix: np.ndarray) -> None:
        """`engine.compute_into(&scope, queries, Some(candidates), &mut matrix)` (bench.rs:599-603)."""
        resident = isinstance(queries, DeviceTape)
        tq = queries if resident else as_tape(queries)
        tc = tq if candidates is None else (candidates if isinstance(candidates, DeviceTape) else as_tape(candidates))
        if resident != isinstance(tc, DeviceTape):
            raise ValueError("either both tapes are resident (DeviceTape) or neither")
        nq, nc = len(tq), len(tc)
        kind = np.dtype(self.out_dtype).kind
        if matrix.dtype.kind != kind or matrix.dtype.itemsize not in _RESULT_TYPE or matrix.ndim != 2 or matrix.shape[0] < nq or matrix.shape[1] < nc:
            raise ValueError(f"matrix must be a 64/32/16-bit {'unsigned' if kind == 'u' else 'signed'} array with shape >= ({nq}, {nc})")
        if matrix.strides[1] != matrix.itemsize or matrix.strides[0] % matrix.itemsize:
            raise ValueError("matrix rows must be contiguous")
        rtype = _RESULT_TYPE[matrix.dtype.itemsize]
        stride = matrix.strides[0] // matrix.itemsize
        err = ctypes.c_char_p()
        if resident:
            _lib.check(_lib.lib().swb_compute_cross_tape(self._h, tq._root._h, tq.lo, nq, tc._root._h, tc.lo, nc, _ptr(matrix), stride, rtype,
                                                         ctypes.byref(err)), err)
        else:
            vq, vc = tq.view(), tc.view()
            _lib.check(_lib.lib().swb_compute_cross_view(self._h, ctypes.byref(vq), ctypes.byref(vc), _ptr(matrix), stride, rtype,
                                                         ctypes.byref(err)), err)

    def compute_pairs_peer(self, a_view: "_lib.TapeView", b_view: "_lib.TapeView", out_ptr: int, result_type: int = _lib.RESULT_64):
        """The sharded call with no scatter step: every device's kernels read the root's tapes over NVLink (peer loads) and
        write their results into `out` on the root (swb_compute_pairs_peer). Returns {remote_read_bytes, remote_written_bytes, ms}."""
        stats = (ctypes.c_double * 4)()
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_compute_pairs_peer(self._h, ctypes.byref(a_view), ctypes.byref(b_view), out_ptr, result_type, stats,
                                                     ctypes.byref(err)), err)
        return {"remote_read_bytes": int(stats[0]), "remote_written_bytes": int(stats[1]), "ms": float(stats[2])}

    def compute_pairs_scatter(self, a_view: "_lib.TapeView", b_view: "_lib.TapeView", out_ptr: int, result_type: int = _lib.RESULT_64):
        """`LevenshteinDistances<{n}gpu>` over a batch resident on the scope's first device: one NCCL scatter, per-device
        kernels, one NCCL gather (swb_compute_pairs_scatter). Returns {sent_bytes, gathered_bytes, device_ms, chunks}."""
        stats = (ctypes.c_double * 4)()
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_compute_pairs_scatter(self._h, ctypes.byref(a_view), ctypes.byref(b_view), out_ptr, result_type, stats,
                                                        ctypes.byref(err)), err)
        return {"sent_bytes": int(stats[0]), "gathered_bytes": int(stats[1]), "device_ms": float(stats[2]), "chunks": int(stats[3])}

    def compute(self, scope: DeviceScope, queries: TapeLike, candidates: Optional[TapeLike] = None) -> np.ndarray:
        """`engine.compute(&scope, &queries, &candidates) -> UnifiedMat` (bench.rs:587-589): allocates the matrix."""
        tq = queries if isinstance(queries, DeviceTape) else as_tape(queries)
        tc = tq if candidates is None else (candidates if isinstance(candidates, DeviceTape) else as_tape(candidates))
        matrix = unified_mat(len(tq), len(tc), self.out_dtype)
        self.compute_into(scope, tq, tc, matrix)
        return matrix

    def __call__(self, queries, candidates=None, scope=None, out=None):
        """Python-harness call shape `engine(queries, candidates, scope, out=matrix)` (bench.py:399-422)."""
        if out is None:
            return self.compute(scope, queries, candidates)
        self.compute_into(scope, queries, candidates, out)
        return out

    # ---- device-resident tapes (torch tensors or raw pointers), used by bench.py and dist.py
    def compute_pairs_device(self, a_data: int, a_offsets: int, b_data: int, b_offsets: int, n_pairs: int, out: int,
                             stream: int = 0, device_index: int = 0) -> None:
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_compute_pairs_device(self._h, device_index, a_data, a_offsets, b_data, b_offsets,
                                                       n_pairs, out, stream, ctypes.byref(err)), err)

    def compute_pairs_device_view(self, a_view: "_lib.TapeView", b_view: "_lib.TapeView", out: int, result_type: int = _lib.RESULT_64,
                                  stream: int = 0, device_index: int = 0) -> None:
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_compute_pairs_device_view(self._h, device_index, ctypes.byref(a_view), ctypes.byref(b_view), out, result_type,
                                                            stream, ctypes.byref(err)), err)

    def compute_cross_device(self, q_data: int, q_offsets: int, n_q: int, c_data: int, c_offsets: int, n_c: int,
                             out: int, row_stride: int, stream: int = 0, device_index: int = 0) -> None:
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_compute_cross_device(self._h, device_index, q_data, q_offsets, n_q, c_data, c_offsets,
                                                       n_c, out, row_stride, stream, ctypes.byref(err)), err)

    def load_phase_device(self, a_data: int, a_offsets: int, b_data: int, b_offsets: int, n_pairs: int, out: int,
                          stream: int = 0, device_index: int = 0) -> None:
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_load_phase_device(self._h, device_index, a_data, a_offsets, b_data, b_offsets,
                                                    n_pairs, out, stream, ctypes.byref(err)), err)

    def close(self) -> None:
        if self._h:
            _lib.lib().swb_engine_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LevenshteinDistances(_Engine):
    """`LevenshteinDistances::new(&scope, match, mismatch, open, extend)` (bench.rs:382-393); byte symbols."""

    def __init__(self, scope: DeviceScope, match: int = 0, mismatch: int = 1, open: int = 1, extend: int = 1):
        super().__init__(scope)
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_levenshtein_create(scope._h, match, mismatch, open, extend, ctypes.byref(self._h),
                                                     ctypes.byref(err)), err)

    def distance(self, a: bytes, b: bytes) -> int:
        """The literal `(a:&[u8], b:&[u8]) -> usize` contender surface (bench.rs:416-419); a batch of one."""
        err = ctypes.c_char_p()
        d = _lib.lib().swb_levenshtein_pair(self._h, a, len(a), b, len(b), ctypes.byref(err))
        if d == 0xFFFFFFFFFFFFFFFF:
            raise _lib.Error(4, (err.value or b"").decode("utf-8", "replace"))
        return int(d)


class LevenshteinDistancesUtf8(_Engine):
    """`LevenshteinDistancesUtf8::new(&scope, 0, 1, 1, 1)` (bench.rs:386-399); Unicode scalar values."""

    def __init__(self, scope: DeviceScope, match: int = 0, mismatch: int = 1, open: int = 1, extend: int = 1):
        super().__init__(scope)
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_levenshtein_utf8_create(scope._h, match, mismatch, open, extend,
                                                          ctypes.byref(self._h), ctypes.byref(err)), err)


class _ScoreEngine(_Engine):
    out_dtype = np.int64
    _create = ""

    def __init__(self, scope: DeviceScope, byte_to_class, class_costs, open: int, extend: int):
        super().__init__(scope)
        btc = np.ascontiguousarray(byte_to_class, dtype=np.uint8)
        costs = np.ascontiguousarray(class_costs, dtype=np.int8)
        if btc.shape != (256,) or costs.shape != (32, 32):
            raise ValueError("byte_to_class must have 256 entries and class_costs must be 32x32")
        err = ctypes.c_char_p()
        fn = getattr(_lib.lib(), self._create)
        _lib.check(fn(scope._h, btc.ctypes.data_as(_lib.c_u8p), costs.ctypes.data_as(_lib.c_i8p), open, extend,
                      ctypes.byref(self._h), ctypes.byref(err)), err)


class NeedlemanWunschScores(_ScoreEngine):
    """`NeedlemanWunschScores::new(&scope, &byte_to_class, &class_costs, open, extend)` (bench.rs:658-662)."""
    _create = "swb_needleman_wunsch_create"


class SmithWatermanScores(_ScoreEngine):
    """`SmithWatermanScores::new(..)` (bench.rs:663-666)."""
    _create = "swb_smith_waterman_create"


def unary_class_costs(match_cost: int, mismatch_cost: int):
    """`unary_class_costs` (bench.rs:98-108; bench.py twin): bytes folded into 32 classes by `byte % 32`."""
    byte_to_class = (np.arange(256) % 32).astype(np.uint8)
    class_costs = np.full((32, 32), mismatch_cost, dtype=np.int8)
    np.fill_diagonal(class_costs, match_cost)
    return byte_to_class, class_costs
