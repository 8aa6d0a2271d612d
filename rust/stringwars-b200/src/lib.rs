//! Safe wrapper over the C ABI of the B200 edit-distance engine (include/stringwars_b200.h).
//!
//! Mirrors the shape of `stringzilla::szs` as used by StringWars `similarities/bench.rs`:
//! `DeviceScope::gpu_device(0)`, `LevenshteinDistances::new(&scope, 0, 1, 1, 1)`,
//! `engine.compute_into(&scope, queries, candidates, &mut matrix)`, plus element-wise `compute_pairs`, resident tapes
//! and the scalar `(a: &[u8], b: &[u8]) -> usize` surface the rapidfuzz baseline is timed behind (bench.rs:416-419).
//!
//! Lifetimes: engines and resident tapes borrow their `DeviceScope`, so the borrow checker keeps the scope alive for as
//! long as anything created on it (the C library additionally tolerates the opposite order of destruction).
//! Not compiled in the build image of this repository (no Rust toolchain there); the C++ twin under `harness/` and the
//! ctypes binding drive the very same entry points in the GPU tests.
use std::ffi::{c_char, c_int, c_void, CStr};
use std::fmt;
use std::marker::PhantomData;

#[repr(C)]
pub struct RawScope(c_void);
#[repr(C)]
pub struct RawEngine(c_void);
#[repr(C)]
pub struct RawTape(c_void);

/// `swb_tape_view_t`: the plain-struct form of `AnyBytesTape::View64 / View32` (bench.rs:141-142).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct RawTapeView {
    pub data: *const u8,
    pub offsets: *const c_void,
    pub n: usize,
    pub offset_bits: c_int,
}

extern "C" {
    fn swb_scope_create(first_device: c_int, num_devices: c_int, out: *mut *mut RawScope, err: *mut *const c_char) -> c_int;
    fn swb_scope_destroy(scope: *mut RawScope);
    fn swb_scope_num_devices(scope: *const RawScope) -> c_int;
    fn swb_levenshtein_create(scope: *mut RawScope, m: c_int, x: c_int, o: c_int, e: c_int, out: *mut *mut RawEngine, err: *mut *const c_char) -> c_int;
    fn swb_levenshtein_utf8_create(scope: *mut RawScope, m: c_int, x: c_int, o: c_int, e: c_int, out: *mut *mut RawEngine, err: *mut *const c_char) -> c_int;
    fn swb_needleman_wunsch_create(scope: *mut RawScope, byte_to_class: *const u8, class_costs: *const i8, open: c_int, extend: c_int,
                                   out: *mut *mut RawEngine, err: *mut *const c_char) -> c_int;
    fn swb_smith_waterman_create(scope: *mut RawScope, byte_to_class: *const u8, class_costs: *const i8, open: c_int, extend: c_int,
                                 out: *mut *mut RawEngine, err: *mut *const c_char) -> c_int;
    fn swb_engine_destroy(engine: *mut RawEngine);
    fn swb_engine_set_gap_convention(engine: *mut RawEngine, convention: c_int, err: *mut *const c_char) -> c_int;
    fn swb_compute_pairs_view(engine: *mut RawEngine, a: *const RawTapeView, b: *const RawTapeView, out: *mut c_void, out_type: c_int,
                              err: *mut *const c_char) -> c_int;
    fn swb_compute_cross_view(engine: *mut RawEngine, q: *const RawTapeView, c: *const RawTapeView, out: *mut c_void, row_stride: usize,
                              out_type: c_int, err: *mut *const c_char) -> c_int;
    fn swb_levenshtein_pair(engine: *mut RawEngine, a: *const u8, la: usize, b: *const u8, lb: usize, err: *mut *const c_char) -> u64;
    fn swb_tape_upload(scope: *mut RawScope, view: *const RawTapeView, out: *mut *mut RawTape, err: *mut *const c_char) -> c_int;
    fn swb_tape_destroy(tape: *mut RawTape);
    fn swb_tape_len(tape: *const RawTape) -> usize;
    fn swb_tape_device_view(tape: *const RawTape, device_index: c_int, first: usize, n: usize, out: *mut RawTapeView, err: *mut *const c_char) -> c_int;
    fn swb_compute_pairs_tape(engine: *mut RawEngine, a: *const RawTape, a_first: usize, b: *const RawTape, b_first: usize, n_pairs: usize,
                              out: *mut c_void, out_type: c_int, err: *mut *const c_char) -> c_int;
    fn swb_compute_cross_tape(engine: *mut RawEngine, q: *const RawTape, q_first: usize, n_q: usize, c: *const RawTape, c_first: usize, n_c: usize,
                              out: *mut c_void, row_stride: usize, out_type: c_int, err: *mut *const c_char) -> c_int;
    fn swb_compute_pairs_scatter(engine: *mut RawEngine, a: *const RawTapeView, b: *const RawTapeView, out: *mut c_void, out_type: c_int,
                                 stats: *mut f64, err: *mut *const c_char) -> c_int;
    fn swb_multiprocessor_count(device: c_int, out: *mut c_int, err: *mut *const c_char) -> c_int;
    fn swb_host_alloc(bytes: usize, out: *mut *mut c_void, err: *mut *const c_char) -> c_int;
    fn swb_host_free(ptr: *mut c_void);
    fn swb_unified_alloc(bytes: usize, out: *mut *mut c_void, err: *mut *const c_char) -> c_int;
    fn swb_unified_free(ptr: *mut c_void);
}

#[derive(Debug)]
pub struct Error { pub status: i32, pub message: String }
impl fmt::Display for Error {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result { write!(f, "stringwars_b200 status {}: {}", self.status, self.message) }
}
impl std::error::Error for Error {}

fn check(status: c_int, err: *const c_char) -> Result<(), Error> {
    if status == 0 { return Ok(()); }
    let message = if err.is_null() { String::new() } else { unsafe { CStr::from_ptr(err) }.to_string_lossy().into_owned() };
    Err(Error { status, message })
}

/// What a result element may be: `usize` / `isize` like the reference's `UnifiedMat`, or a narrower integer of the same
/// signedness (`swb_result_type_t`): a Levenshtein distance below 65,536 fits `u16`, a score fits `i32`.
pub trait ResultElement: Copy + Default { const TYPE: c_int; const SIGNED: bool; }
impl ResultElement for u64 { const TYPE: c_int = 0; const SIGNED: bool = false; }
impl ResultElement for usize { const TYPE: c_int = 0; const SIGNED: bool = false; } // 64-bit targets only (the harness's)
impl ResultElement for u32 { const TYPE: c_int = 1; const SIGNED: bool = false; }
impl ResultElement for u16 { const TYPE: c_int = 2; const SIGNED: bool = false; }
impl ResultElement for i64 { const TYPE: c_int = 0; const SIGNED: bool = true; }
impl ResultElement for isize { const TYPE: c_int = 0; const SIGNED: bool = true; }
impl ResultElement for i32 { const TYPE: c_int = 1; const SIGNED: bool = true; }
impl ResultElement for i16 { const TYPE: c_int = 2; const SIGNED: bool = true; }

/// Offsets of a tape: `u64` (`BytesTapeView<u64>`) or `u32`.
pub trait Offset: Copy { const BITS: c_int; }
impl Offset for u64 { const BITS: c_int = 64; }
impl Offset for u32 { const BITS: c_int = 32; }

/// A borrowed Arrow-style tape: `offsets.len() == n + 1`, string i is `data[offsets[i]..offsets[i+1]]`; offsets are
/// absolute into `data`, so a `subview` of a `stringtape::BytesTapeView` maps to the same `data` and a sub-slice of the
/// offsets. The memory may be ordinary, page-locked or CUDA-managed (`UnifiedAlloc`): the library looks at the pointer.
#[derive(Clone, Copy)]
pub struct TapeView<'a, O: Offset = u64> { pub data: &'a [u8], pub offsets: &'a [O] }
impl<'a, O: Offset> TapeView<'a, O> {
    pub fn len(&self) -> usize { self.offsets.len().saturating_sub(1) }
    pub fn is_empty(&self) -> bool { self.len() == 0 }
    pub fn subview(&self, lo: usize, hi: usize) -> Self { Self { data: self.data, offsets: &self.offsets[lo..=hi] } }
    fn raw(&self) -> RawTapeView {
        RawTapeView { data: self.data.as_ptr(), offsets: self.offsets.as_ptr() as *const c_void, n: self.len(), offset_bits: O::BITS }
    }
}

pub struct DeviceScope(*mut RawScope);
impl DeviceScope {
    /// `DeviceScope::gpu_device(0)` (bench.rs:379).
    pub fn gpu_device(device: usize) -> Result<Self, Error> { Self::gpu_devices(device, 1) }
    /// n GPUs of one box: host-tape calls shard the batch across them (`<{n}gpu>` variants, bench.rs:498-501).
    pub fn gpu_devices(first: usize, count: usize) -> Result<Self, Error> {
        let (mut out, mut err) = (std::ptr::null_mut(), std::ptr::null());
        check(unsafe { swb_scope_create(first as c_int, count as c_int, &mut out, &mut err) }, err)?;
        Ok(Self(out))
    }
    pub fn num_devices(&self) -> usize { unsafe { swb_scope_num_devices(self.0) as usize } }
    /// Upload once, compute many times: what `BytesTape<u64, UnifiedAlloc>` gives the reference after the first touch.
    pub fn upload<'s, O: Offset>(&'s self, tape: TapeView<O>) -> Result<DeviceTape<'s>, Error> {
        let (mut out, mut err) = (std::ptr::null_mut(), std::ptr::null());
        let raw = tape.raw();
        check(unsafe { swb_tape_upload(self.0, &raw, &mut out, &mut err) }, err)?;
        Ok(DeviceTape { raw: out, first: 0, n: tape.len(), owner: true, _scope: PhantomData })
    }
}
impl Drop for DeviceScope { fn drop(&mut self) { unsafe { swb_scope_destroy(self.0) } } }

/// A tape resident on every device of its scope; `subview` is a zero-copy range (`BytesTapeView::subview`, bench.rs:134-139).
pub struct DeviceTape<'s> { raw: *mut RawTape, first: usize, n: usize, owner: bool, _scope: PhantomData<&'s DeviceScope> }
impl<'s> DeviceTape<'s> {
    pub fn len(&self) -> usize { self.n }
    pub fn is_empty(&self) -> bool { self.n == 0 }
    pub fn total_len(&self) -> usize { unsafe { swb_tape_len(self.raw) } }
    pub fn subview(&self, lo: usize, hi: usize) -> DeviceTape<'_> {
        assert!(lo <= hi && hi <= self.n, "subview [{lo}, {hi}) outside a tape of {}", self.n);
        DeviceTape { raw: self.raw, first: self.first + lo, n: hi - lo, owner: false, _scope: PhantomData }
    }
    /// Device pointers of this range on scope device `device_index` (for `compute_pairs_scatter` and custom kernels).
    pub fn device_view(&self, device_index: usize) -> Result<RawTapeView, Error> {
        let mut view = RawTapeView { data: std::ptr::null(), offsets: std::ptr::null(), n: 0, offset_bits: 64 };
        let mut err = std::ptr::null();
        check(unsafe { swb_tape_device_view(self.raw, device_index as c_int, self.first, self.n, &mut view, &mut err) }, err)?;
        Ok(view)
    }
}
impl Drop for DeviceTape<'_> { fn drop(&mut self) { if self.owner { unsafe { swb_tape_destroy(self.raw) } } } }

/// Row-major result matrix the device fills without a pageable staging copy: `UnifiedMat<usize>` / `UnifiedMat<isize>`
/// (bench.rs:130, 186-211, 466-468). `new` allocates CUDA managed memory like the reference (the kernels write it in
/// place and it stays on the device until the host reads it); `pinned` allocates page-locked host memory (results
/// arrive by DMA inside every call).
pub struct UnifiedMat<T: ResultElement> { ptr: *mut T, rows: usize, cols: usize, managed: bool }
impl<T: ResultElement> UnifiedMat<T> {
    pub fn new(rows: usize, cols: usize) -> Result<Self, Error> { Self::alloc(rows, cols, true) }
    pub fn pinned(rows: usize, cols: usize) -> Result<Self, Error> { Self::alloc(rows, cols, false) }
    fn alloc(rows: usize, cols: usize, managed: bool) -> Result<Self, Error> {
        let bytes = rows * cols * std::mem::size_of::<T>();
        let (mut p, mut err) = (std::ptr::null_mut(), std::ptr::null());
        check(unsafe { if managed { swb_unified_alloc(bytes, &mut p, &mut err) } else { swb_host_alloc(bytes, &mut p, &mut err) } }, err)?;
        unsafe { std::ptr::write_bytes(p as *mut u8, 0, bytes) };
        Ok(Self { ptr: p as *mut T, rows, cols, managed })
    }
    pub fn rows(&self) -> usize { self.rows }
    pub fn cols(&self) -> usize { self.cols }
    pub fn as_slice(&self) -> &[T] { unsafe { std::slice::from_raw_parts(self.ptr, self.rows * self.cols) } }
    pub fn as_mut_slice(&mut self) -> &mut [T] { unsafe { std::slice::from_raw_parts_mut(self.ptr, self.rows * self.cols) } }
}
impl<T: ResultElement> Drop for UnifiedMat<T> {
    fn drop(&mut self) { unsafe { if self.managed { swb_unified_free(self.ptr as *mut c_void) } else { swb_host_free(self.ptr as *mut c_void) } } }
}

/// What every engine can do (the `szs` engines share one call shape, bench.rs:123-148, 186-211).
struct Engine<'s> { raw: *mut RawEngine, _scope: PhantomData<&'s DeviceScope> }
impl<'s> Engine<'s> {
    fn pairs<O: Offset, P: Offset, T: ResultElement>(&self, a: TapeView<O>, b: TapeView<P>, out: &mut [T]) -> Result<(), Error> {
        assert!(a.len() == b.len(), "tapes differ in length: {} vs {}", a.len(), b.len());
        assert!(out.len() >= a.len(), "out holds {} elements, {} pairs", out.len(), a.len());
        let (ra, rb, mut err) = (a.raw(), b.raw(), std::ptr::null());
        check(unsafe { swb_compute_pairs_view(self.raw, &ra, &rb, out.as_mut_ptr() as *mut c_void, T::TYPE, &mut err) }, err)
    }
    fn cross<O: Offset, P: Offset, T: ResultElement>(&self, q: TapeView<O>, c: TapeView<P>, matrix: &mut [T], row_stride: usize) -> Result<(), Error> {
        assert!(row_stride >= c.len(), "row_stride {row_stride} < {} candidates", c.len());
        assert!(q.is_empty() || matrix.len() >= (q.len() - 1) * row_stride + c.len(), "matrix too small for {} x {}", q.len(), c.len());
        let (rq, rc, mut err) = (q.raw(), c.raw(), std::ptr::null());
        check(unsafe { swb_compute_cross_view(self.raw, &rq, &rc, matrix.as_mut_ptr() as *mut c_void, row_stride, T::TYPE, &mut err) }, err)
    }
    fn pairs_resident<T: ResultElement>(&self, a: &DeviceTape, b: &DeviceTape, out: &mut [T]) -> Result<(), Error> {
        assert!(a.len() == b.len() && out.len() >= a.len());
        let mut err = std::ptr::null();
        check(unsafe { swb_compute_pairs_tape(self.raw, a.raw, a.first, b.raw, b.first, a.n, out.as_mut_ptr() as *mut c_void, T::TYPE, &mut err) }, err)
    }
    fn cross_resident<T: ResultElement>(&self, q: &DeviceTape, c: &DeviceTape, matrix: &mut [T], row_stride: usize) -> Result<(), Error> {
        assert!(row_stride >= c.len());
        assert!(q.is_empty() || matrix.len() >= (q.len() - 1) * row_stride + c.len());
        let mut err = std::ptr::null();
        check(unsafe { swb_compute_cross_tape(self.raw, q.raw, q.first, q.n, c.raw, c.first, c.n, matrix.as_mut_ptr() as *mut c_void, row_stride, T::TYPE, &mut err) }, err)
    }
}
impl Drop for Engine<'_> { fn drop(&mut self) { unsafe { swb_engine_destroy(self.raw) } } }

macro_rules! engine_methods {
    ($name:ident, $signed:expr) => {
        impl<'s> $name<'s> {
            /// out[i] = score(a[i], b[i]) — the batched form of the scalar contender surface (bench.rs:416-419).
            pub fn compute_pairs<O: Offset, P: Offset, T: ResultElement>(&self, _scope: &DeviceScope, a: TapeView<O>, b: TapeView<P>, out: &mut [T]) -> Result<(), Error> {
                assert!(T::SIGNED == $signed, "result signedness does not match the engine");
                self.0.pairs(a, b, out)
            }
            /// `engine.compute_into(&scope, queries, Some(candidates), &mut matrix)` (bench.rs:599-603, 186-211).
            pub fn compute_into<O: Offset, P: Offset, T: ResultElement>(&self, _scope: &DeviceScope, queries: TapeView<O>, candidates: TapeView<P>,
                                                                       matrix: &mut UnifiedMat<T>) -> Result<(), Error> {
                assert!(T::SIGNED == $signed, "result signedness does not match the engine");
                assert!(matrix.rows() >= queries.len() && matrix.cols() >= candidates.len(), "matrix smaller than {} x {}", queries.len(), candidates.len());
                let stride = matrix.cols();
                self.0.cross(queries, candidates, matrix.as_mut_slice(), stride)
            }
            /// `compute_into` over a caller-owned slice with an explicit row stride.
            pub fn compute_into_slice<O: Offset, P: Offset, T: ResultElement>(&self, _scope: &DeviceScope, queries: TapeView<O>, candidates: TapeView<P>,
                                                                             matrix: &mut [T], row_stride: usize) -> Result<(), Error> {
                assert!(T::SIGNED == $signed, "result signedness does not match the engine");
                self.0.cross(queries, candidates, matrix, row_stride)
            }
            /// Same calls over resident tapes (`DeviceScope::upload`): no input bytes move.
            pub fn compute_pairs_resident<T: ResultElement>(&self, a: &DeviceTape, b: &DeviceTape, out: &mut [T]) -> Result<(), Error> {
                assert!(T::SIGNED == $signed);
                self.0.pairs_resident(a, b, out)
            }
            pub fn compute_into_resident<T: ResultElement>(&self, queries: &DeviceTape, candidates: &DeviceTape, matrix: &mut UnifiedMat<T>) -> Result<(), Error> {
                assert!(T::SIGNED == $signed);
                assert!(matrix.rows() >= queries.len() && matrix.cols() >= candidates.len());
                let stride = matrix.cols();
                self.0.cross_resident(queries, candidates, matrix.as_mut_slice(), stride)
            }
        }
    };
}

/// `LevenshteinDistances::new(&scope, 0, 1, 1, 1)` (bench.rs:382-393): byte symbols, `usize` distances.
pub struct LevenshteinDistances<'s>(Engine<'s>);
impl<'s> LevenshteinDistances<'s> {
    pub fn new(scope: &'s DeviceScope, m: i32, x: i32, open: i32, extend: i32) -> Result<Self, Error> {
        let (mut out, mut err) = (std::ptr::null_mut(), std::ptr::null());
        check(unsafe { swb_levenshtein_create(scope.0, m, x, open, extend, &mut out, &mut err) }, err)?;
        Ok(Self(Engine { raw: out, _scope: PhantomData }))
    }
    /// The scalar contender surface (bench.rs:416-419). Launch-latency bound: for tests, not for timing.
    pub fn distance(&self, a: &[u8], b: &[u8]) -> Result<usize, Error> {
        let mut err = std::ptr::null();
        let d = unsafe { swb_levenshtein_pair(self.0.raw, a.as_ptr(), a.len(), b.as_ptr(), b.len(), &mut err) };
        if d == u64::MAX { check(4, err)?; }
        Ok(d as usize)
    }
    /// `LevenshteinDistances<{n}gpu>` for a batch resident on the scope's first device: one NCCL scatter, per-device
    /// kernels, one NCCL gather (`swb_compute_pairs_scatter`). `a`, `b`: device views (`DeviceTape::device_view(0)`),
    /// `out`: device memory of the first device. Returns (bytes scattered, bytes gathered, device milliseconds).
    ///
    /// # Safety
    /// `out` must be device (or managed) memory with room for `a.n` elements of `T`.
    pub unsafe fn compute_pairs_scatter<T: ResultElement>(&self, a: &RawTapeView, b: &RawTapeView, out: *mut T) -> Result<(u64, u64, f64), Error> {
        let mut stats = [0f64; 4];
        let mut err = std::ptr::null();
        check(swb_compute_pairs_scatter(self.0.raw, a, b, out as *mut c_void, T::TYPE, stats.as_mut_ptr(), &mut err), err)?;
        Ok((stats[0] as u64, stats[1] as u64, stats[2]))
    }
}
engine_methods!(LevenshteinDistances, false);

/// `LevenshteinDistancesUtf8::new(&scope, 0, 1, 1, 1)` (bench.rs:386-399): Unicode scalar values of valid UTF-8.
pub struct LevenshteinDistancesUtf8<'s>(Engine<'s>);
impl<'s> LevenshteinDistancesUtf8<'s> {
    pub fn new(scope: &'s DeviceScope, m: i32, x: i32, open: i32, extend: i32) -> Result<Self, Error> {
        let (mut out, mut err) = (std::ptr::null_mut(), std::ptr::null());
        check(unsafe { swb_levenshtein_utf8_create(scope.0, m, x, open, extend, &mut out, &mut err) }, err)?;
        Ok(Self(Engine { raw: out, _scope: PhantomData }))
    }
}
engine_methods!(LevenshteinDistancesUtf8, false);

/// What a length-k gap costs: `Stringzilla` = open + (k-1)·extend, `Bio` = open + k·extend.
#[derive(Clone, Copy)]
pub enum GapConvention { Stringzilla = 0, Bio = 1 }

macro_rules! score_engine {
    ($name:ident, $create:ident, $doc:expr) => {
        #[doc = $doc]
        pub struct $name<'s>(Engine<'s>);
        impl<'s> $name<'s> {
            pub fn new(scope: &'s DeviceScope, byte_to_class: &[u8; 256], class_costs: &[[i8; 32]; 32], open: i8, extend: i8) -> Result<Self, Error> {
                let (mut out, mut err) = (std::ptr::null_mut(), std::ptr::null());
                check(unsafe { $create(scope.0, byte_to_class.as_ptr(), class_costs.as_ptr() as *const i8, open as c_int, extend as c_int, &mut out, &mut err) }, err)?;
                Ok(Self(Engine { raw: out, _scope: PhantomData }))
            }
            pub fn set_gap_convention(&self, convention: GapConvention) -> Result<(), Error> {
                let mut err = std::ptr::null();
                check(unsafe { swb_engine_set_gap_convention(self.0.raw, convention as c_int, &mut err) }, err)
            }
        }
        engine_methods!($name, true);
    };
}
score_engine!(NeedlemanWunschScores, swb_needleman_wunsch_create,
              "`NeedlemanWunschScores::new(&scope, &byte_to_class, &class_costs, open, extend)` (bench.rs:658-662): global score, `isize`.");
score_engine!(SmithWatermanScores, swb_smith_waterman_create,
              "`SmithWatermanScores::new(..)` (bench.rs:663-666, 982-985): best local score (>= 0), `isize`.");

/// `unary_class_costs` (bench.rs:98-108): bytes folded into 32 classes by `byte % 32`, match / mismatch costs.
pub fn unary_class_costs(match_cost: i8, mismatch_cost: i8) -> ([u8; 256], [[i8; 32]; 32]) {
    let mut byte_to_class = [0u8; 256];
    for (b, class) in byte_to_class.iter_mut().enumerate() { *class = (b % 32) as u8; }
    let mut costs = [[mismatch_cost; 32]; 32];
    for (i, row) in costs.iter_mut().enumerate() { row[i] = match_cost; }
    (byte_to_class, costs)
}

/// `gpu_multiprocessor_count(device)` (utils.rs:826-836).
pub fn multiprocessor_count(device: usize) -> Option<usize> {
    let (mut n, mut err) = (0 as c_int, std::ptr::null());
    if unsafe { swb_multiprocessor_count(device as c_int, &mut n, &mut err) } == 0 { Some(n as usize) } else { None }
}
