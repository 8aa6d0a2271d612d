"""Arrow-style string tapes: one contiguous byte buffer + N+1 u64 offsets.

Host-side mirror of the `stringtape` containers the reference harness feeds its engines
(`BytesTape<u64, UnifiedAlloc>` / `BytesTapeView<u64>::subview`, StringWars similarities/bench.rs:62, 292-306,
134-139). `pinned=True` places both buffers in page-locked host memory obtained from the native library
(`swb_host_alloc`), the B200-native stand-in for the reference's `UnifiedAlloc`.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Sequence, Union

import numpy as np

from . import _lib


class PinnedBuffer:
    """Page-locked host allocation exposed as a numpy array (freed with the object)."""

    def __init__(self, nbytes: int):
        ptr = ctypes.c_void_p()
        err = ctypes.c_char_p()
        _lib.check(_lib.lib().swb_host_alloc(max(int(nbytes), 1), ctypes.byref(ptr), ctypes.byref(err)), err)
        self._ptr = ptr
        self.nbytes = int(nbytes)
        buf = (ctypes.c_uint8 * max(self.nbytes, 1)).from_address(ptr.value)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes)

    def view(self, dtype) -> np.ndarray:
        return self.array.view(dtype)

    def __del__(self):
        try:
            if self._ptr:
                _lib.lib().swb_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


def unified_mat(rows: int, cols: int, dtype=np.uint64) -> np.ndarray:
    """`UnifiedMat<usize>::new(rows, cols)` (bench.rs:130, 466-468): a result matrix the device can write without a
    staging copy. Here: page-locked host memory from `swb_host_alloc`, returned as a (rows, cols) numpy array that
    keeps its allocation alive."""
    buf = PinnedBuffer(int(rows) * int(cols) * np.dtype(dtype).itemsize)
    mat = buf.view(dtype).reshape(int(rows), int(cols))
    mat[...] = 0
    return _Owned(mat, buf)


class _Owned(np.ndarray):
    """ndarray view that holds a reference to the PinnedBuffer behind it."""

    def __new__(cls, arr: np.ndarray, owner):
        obj = arr.view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


class BytesTape:
    """`data[offsets[i]:offsets[i+1]]` is string i. Views share memory (`subview` is zero-copy)."""

    def __init__(self, data: np.ndarray, offsets: np.ndarray, _keepalive=None):
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if self.offsets.ndim != 1 or self.offsets.size < 1:
            raise ValueError("offsets must hold N+1 entries")
        self._keepalive = _keepalive

    @classmethod
    def from_strings(cls, strings: Iterable[Union[bytes, str]], pinned: bool = False) -> "BytesTape":
        items = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in strings]
        lengths = np.fromiter((len(s) for s in items), dtype=np.uint64, count=len(items))
        offsets = np.zeros(len(items) + 1, dtype=np.uint64)
        np.cumsum(lengths, out=offsets[1:])
        data = np.frombuffer(b"".join(items), dtype=np.uint8)
        return cls.from_arrays(data, offsets, pinned=pinned)

    @classmethod
    def from_arrays(cls, data: np.ndarray, offsets: np.ndarray, pinned: bool = False) -> "BytesTape":
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if not pinned:
            return cls(data, offsets)
        dbuf, obuf = PinnedBuffer(data.nbytes + 64), PinnedBuffer(offsets.nbytes)
        dview = dbuf.array[: data.nbytes]
        dview[:] = data
        oview = obuf.view(np.uint64)[: offsets.size]
        oview[:] = offsets
        return cls(dview, oview, _keepalive=(dbuf, obuf))

    def __len__(self) -> int:
        return int(self.offsets.size - 1)

    def __getitem__(self, i: int) -> bytes:
        return self.data[int(self.offsets[i]): int(self.offsets[i + 1])].tobytes()

    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets)

    def subview(self, lo: int, hi: int) -> "BytesTape":
        """Zero-copy `[lo, hi)` view (`BytesTapeView::subview`, bench.rs:134-139). Offsets stay absolute."""
        if not (0 <= lo <= hi <= len(self)):
            raise IndexError(f"subview [{lo}, {hi}) outside tape of {len(self)}")
        return BytesTape(self.data, self.offsets[lo: hi + 1], _keepalive=self._keepalive)


TapeLike = Union[BytesTape, Sequence[Union[bytes, str]]]


def as_tape(x: TapeLike) -> BytesTape:
    return x if isinstance(x, BytesTape) else BytesTape.from_strings(x)
