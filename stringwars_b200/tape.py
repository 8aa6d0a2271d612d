"""Arrow-style string tapes: one contiguous byte buffer + N+1 offsets (u64 or u32).

Host-side mirror of the `stringtape` containers the reference harness feeds its engines
(`BytesTape<u64, UnifiedAlloc>` / `BytesTapeView<u64>::subview` / `AnyBytesTape::View64`, StringWars
similarities/bench.rs:62, 292-306, 134-142). `memory=` picks where both buffers live:
  "pageable"  ordinary numpy memory
  "pinned"    page-locked host memory (`swb_host_alloc`): host tapes copy at full PCIe rate
  "managed"   CUDA managed memory (`swb_unified_alloc`) — the reference's own `UnifiedAlloc`; the native library
              uses such tapes in place, without a staging copy
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Sequence, Union

import numpy as np

from . import _lib


class _NativeBuffer:
    """A native allocation exposed as a numpy array (freed with the object)."""
    _alloc, _free = "", ""

    def __init__(self, nbytes: int):
        ptr = ctypes.c_void_p()
        err = ctypes.c_char_p()
        _lib.check(getattr(_lib.lib(), self._alloc)(max(int(nbytes), 1), ctypes.byref(ptr), ctypes.byref(err)), err)
        self._ptr = ptr
        self.nbytes = int(nbytes)
        buf = (ctypes.c_uint8 * max(self.nbytes, 1)).from_address(ptr.value)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes)

    @property
    def address(self) -> int:
        return int(self._ptr.value)

    def view(self, dtype) -> np.ndarray:
        return self.array.view(dtype)

    def __del__(self):
        try:
            if self._ptr:
                getattr(_lib.lib(), self._free)(self._ptr)
                self._ptr = None
        except Exception:
            pass


class PinnedBuffer(_NativeBuffer):
    """Page-locked host allocation (`swb_host_alloc`)."""
    _alloc, _free = "swb_host_alloc", "swb_host_free"


class ManagedBuffer(_NativeBuffer):
    """CUDA managed allocation (`swb_unified_alloc`): `UnifiedAlloc` (bench.rs:292)."""
    _alloc, _free = "swb_unified_alloc", "swb_unified_free"


_BUFFERS = {"pinned": PinnedBuffer, "managed": ManagedBuffer}


def unified_mat(rows: int, cols: int, dtype=np.uint64, memory: str = "pinned") -> np.ndarray:
    """`UnifiedMat<usize>::new(rows, cols)` (bench.rs:130, 466-468): a result matrix the device can fill without a
    pageable staging copy. memory="pinned": page-locked host memory (results arrive by DMA); memory="managed": CUDA
    managed memory like the reference's, written by the kernels in place. Returned as a (rows, cols) numpy array
    that keeps its allocation alive."""
    buf = _BUFFERS[memory](int(rows) * int(cols) * np.dtype(dtype).itemsize)
    mat = buf.view(dtype).reshape(int(rows), int(cols))
    mat[...] = 0
    return _Owned(mat, buf)


class _Owned(np.ndarray):
    """ndarray view that holds a reference to the native buffer behind it."""

    def __new__(cls, arr: np.ndarray, owner):
        obj = arr.view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


class BytesTape:
    """`data[offsets[i]:offsets[i+1]]` is string i. Views share memory (`subview` is zero-copy). Offsets are u64
    (`BytesTapeView<u64>`) or u32 (`BytesTapeView<u32>`)."""

    def __init__(self, data: np.ndarray, offsets: np.ndarray, _keepalive=None):
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.asarray(offsets)
        if offsets.dtype not in (np.dtype(np.uint64), np.dtype(np.uint32)):
            offsets = offsets.astype(np.uint64)
        self.offsets = np.ascontiguousarray(offsets)
        if self.offsets.ndim != 1 or self.offsets.size < 1:
            raise ValueError("offsets must hold N+1 entries")
        self._keepalive = _keepalive

    @property
    def offset_bits(self) -> int:
        return 8 * self.offsets.dtype.itemsize

    @classmethod
    def from_strings(cls, strings: Iterable[Union[bytes, str]], pinned: bool = False, memory: str = None,
                     offset_dtype=np.uint64) -> "BytesTape":
        items = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in strings]
        lengths = np.fromiter((len(s) for s in items), dtype=np.uint64, count=len(items))
        offsets = np.zeros(len(items) + 1, dtype=np.uint64)
        np.cumsum(lengths, out=offsets[1:])
        data = np.frombuffer(b"".join(items), dtype=np.uint8)
        return cls.from_arrays(data, offsets.astype(offset_dtype), pinned=pinned, memory=memory)

    @classmethod
    def from_arrays(cls, data: np.ndarray, offsets: np.ndarray, pinned: bool = False, memory: str = None) -> "BytesTape":
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.asarray(offsets)
        if offsets.dtype not in (np.dtype(np.uint64), np.dtype(np.uint32)):
            offsets = offsets.astype(np.uint64)
        offsets = np.ascontiguousarray(offsets)
        memory = memory or ("pinned" if pinned else "pageable")
        if memory == "pageable":
            return cls(data, offsets)
        dbuf, obuf = _BUFFERS[memory](data.nbytes + 64), _BUFFERS[memory](offsets.nbytes)
        dview = dbuf.array[: data.nbytes]
        dview[:] = data
        oview = obuf.view(offsets.dtype)[: offsets.size]
        oview[:] = offsets
        return cls(dview, oview, _keepalive=(dbuf, obuf))

    def with_offsets(self, dtype) -> "BytesTape":
        """Same bytes, offsets converted to `dtype` (u32 tapes must span less than 4 GiB)."""
        if int(self.offsets[-1]) >= (1 << 32) and np.dtype(dtype).itemsize == 4:
            raise ValueError("tape too large for u32 offsets")
        return BytesTape(self.data, self.offsets.astype(dtype), _keepalive=self._keepalive)

    def __len__(self) -> int:
        return int(self.offsets.size - 1)

    def __getitem__(self, i: int) -> bytes:
        return self.data[int(self.offsets[i]): int(self.offsets[i + 1])].tobytes()

    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets.astype(np.uint64))

    def subview(self, lo: int, hi: int) -> "BytesTape":
        """Zero-copy `[lo, hi)` view (`BytesTapeView::subview`, bench.rs:134-139). Offsets stay absolute."""
        if not (0 <= lo <= hi <= len(self)):
            raise IndexError(f"subview [{lo}, {hi}) outside tape of {len(self)}")
        return BytesTape(self.data, self.offsets[lo: hi + 1], _keepalive=self._keepalive)

    def view(self) -> "_lib.TapeView":
        """The C ABI's `swb_tape_view_t` over this tape's buffers (valid while the tape is alive)."""
        data_ptr = self.data.ctypes.data if self.data.size else _EMPTY.ctypes.data
        return _lib.TapeView(data_ptr, self.offsets.ctypes.data, len(self), self.offset_bits)


_EMPTY = np.zeros(16, dtype=np.uint8)

TapeLike = Union[BytesTape, Sequence[Union[bytes, str]]]


def as_tape(x: TapeLike) -> BytesTape:
    return x if isinstance(x, BytesTape) else BytesTape.from_strings(x)
