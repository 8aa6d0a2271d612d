//! Safe wrapper over the C ABI of the B200 edit-distance engine (include/stringwars_b200.h).
//!
//! Mirrors the shape of `stringzilla::szs` as used by StringWars `similarities/bench.rs`:
//! `DeviceScope::gpu_device(0)`, `LevenshteinDistances::new(&scope, 0, 1, 1, 1)`,
//! `engine.compute_into(&scope, queries, candidates, &mut matrix)`, plus element-wise `compute_pairs` and the
//! scalar `(a: &[u8], b: &[u8]) -> usize` surface the rapidfuzz baseline is timed behind (bench.rs:416-419).
use std::ffi::{c_char, c_int, c_void, CStr};
use std::fmt;

#[repr(C)]
pub struct RawScope(c_void);
#[repr(C)]
pub struct RawEngine(c_void);

extern "C" {
    fn swb_scope_create(first_device: c_int, num_devices: c_int, out: *mut *mut RawScope, err: *mut *const c_char) -> c_int;
    fn swb_scope_destroy(scope: *mut RawScope);
    fn swb_levenshtein_create(scope: *mut RawScope, m: c_int, x: c_int, o: c_int, e: c_int, out: *mut *mut RawEngine, err: *mut *const c_char) -> c_int;
    fn swb_needleman_wunsch_create(scope: *mut RawScope, byte_to_class: *const u8, class_costs: *const i8, open: c_int, extend: c_int,
                                   out: *mut *mut RawEngine, err: *mut *const c_char) -> c_int;
    fn swb_engine_destroy(engine: *mut RawEngine);
    fn swb_compute_pairs(engine: *mut RawEngine, a_data: *const u8, a_offsets: *const u64, b_data: *const u8, b_offsets: *const u64,
                         n_pairs: usize, out: *mut u64, err: *mut *const c_char) -> c_int;
    fn swb_compute_cross(engine: *mut RawEngine, q_data: *const u8, q_offsets: *const u64, n_q: usize, c_data: *const u8,
                         c_offsets: *const u64, n_c: usize, out: *mut u64, row_stride: usize, err: *mut *const c_char) -> c_int;
    fn swb_levenshtein_pair(engine: *mut RawEngine, a: *const u8, la: usize, b: *const u8, lb: usize, err: *mut *const c_char) -> u64;
    fn swb_multiprocessor_count(device: c_int, out: *mut c_int, err: *mut *const c_char) -> c_int;
    fn swb_host_alloc(bytes: usize, out: *mut *mut c_void, err: *mut *const c_char) -> c_int;
    fn swb_host_free(ptr: *mut c_void);
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

/// A borrowed Arrow-style tape: `offsets.len() == n + 1`, string i is `data[offsets[i]..offsets[i+1]]`.
/// `stringtape::BytesTapeView<u64>` exposes exactly these two slices.
#[derive(Clone, Copy)]
pub struct TapeView<'a> { pub data: &'a [u8], pub offsets: &'a [u64] }
impl<'a> TapeView<'a> { pub fn len(&self) -> usize { self.offsets.len().saturating_sub(1) } }

pub struct DeviceScope(*mut RawScope);
impl DeviceScope {
    /// `DeviceScope::gpu_device(0)` (bench.rs:379).
    pub fn gpu_device(device: usize) -> Result<Self, Error> { Self::gpu_devices(device, 1) }
    /// n GPUs of one box: host-tape calls shard the batch across them (`<{n}gpu>` variants).
    pub fn gpu_devices(first: usize, count: usize) -> Result<Self, Error> {
        let (mut out, mut err) = (std::ptr::null_mut(), std::ptr::null());
        check(unsafe { swb_scope_create(first as c_int, count as c_int, &mut out, &mut err) }, err)?;
        Ok(Self(out))
    }
}
impl Drop for DeviceScope { fn drop(&mut self) { unsafe { swb_scope_destroy(self.0) } } }

pub struct LevenshteinDistances(*mut RawEngine);
impl LevenshteinDistances {
    /// `LevenshteinDistances::new(&scope, 0, 1, 1, 1)` (bench.rs:390-393).
    pub fn new(scope: &DeviceScope, m: i32, x: i32, open: i32, extend: i32) -> Result<Self, Error> {
        let (mut out, mut err) = (std::ptr::null_mut(), std::ptr::null());
        check(unsafe { swb_levenshtein_create(scope.0, m, x, open, extend, &mut out, &mut err) }, err)?;
        Ok(Self(out))
    }
    /// The scalar contender surface (bench.rs:416-419). Launch-latency bound: for tests, not for timing.
    pub fn distance(&self, a: &[u8], b: &[u8]) -> Result<usize, Error> {
        let mut err = std::ptr::null();
        let d = unsafe { swb_levenshtein_pair(self.0, a.as_ptr(), a.len(), b.as_ptr(), b.len(), &mut err) };
        if d == u64::MAX { check(4, err)?; }
        Ok(d as usize)
    }
    /// out[i] = distance(a[i], b[i]).
    pub fn compute_pairs(&self, _scope: &DeviceScope, a: TapeView, b: TapeView, out: &mut [u64]) -> Result<(), Error> {
        assert!(a.len() == b.len() && out.len() >= a.len());
        let mut err = std::ptr::null();
        check(unsafe { swb_compute_pairs(self.0, a.data.as_ptr(), a.offsets.as_ptr(), b.data.as_ptr(), b.offsets.as_ptr(), a.len(), out.as_mut_ptr(), &mut err) }, err)
    }
    /// `engine.compute_into(&scope, queries, Some(candidates), &mut matrix)` (bench.rs:599-603); row-major, `row_stride` elements per row.
    pub fn compute_into(&self, _scope: &DeviceScope, q: TapeView, c: TapeView, matrix: &mut [u64], row_stride: usize) -> Result<(), Error> {
        assert!(row_stride >= c.len() && matrix.len() >= q.len().saturating_sub(1) * row_stride + c.len());
        let mut err = std::ptr::null();
        check(unsafe { swb_compute_cross(self.0, q.data.as_ptr(), q.offsets.as_ptr(), q.len(), c.data.as_ptr(), c.offsets.as_ptr(), c.len(), matrix.as_mut_ptr(), row_stride, &mut err) }, err)
    }
}
impl Drop for LevenshteinDistances { fn drop(&mut self) { unsafe { swb_engine_destroy(self.0) } } }

pub struct NeedlemanWunschScores(*mut RawEngine);
impl NeedlemanWunschScores {
    /// `NeedlemanWunschScores::new(&scope, &byte_to_class, &class_costs, open, extend)` (bench.rs:658-662).
    pub fn new(scope: &DeviceScope, byte_to_class: &[u8; 256], class_costs: &[[i8; 32]; 32], open: i32, extend: i32) -> Result<Self, Error> {
        let (mut out, mut err) = (std::ptr::null_mut(), std::ptr::null());
        check(unsafe { swb_needleman_wunsch_create(scope.0, byte_to_class.as_ptr(), class_costs.as_ptr() as *const i8, open, extend, &mut out, &mut err) }, err)?;
        Ok(Self(out))
    }
    pub fn compute_into(&self, _scope: &DeviceScope, q: TapeView, c: TapeView, matrix: &mut [i64], row_stride: usize) -> Result<(), Error> {
        let mut err = std::ptr::null();
        check(unsafe { swb_compute_cross(self.0, q.data.as_ptr(), q.offsets.as_ptr(), q.len(), c.data.as_ptr(), c.offsets.as_ptr(), c.len(), matrix.as_mut_ptr() as *mut u64, row_stride, &mut err) }, err)
    }
}
impl Drop for NeedlemanWunschScores { fn drop(&mut self) { unsafe { swb_engine_destroy(self.0) } } }

/// Row-major result matrix in page-locked host memory (`swb_host_alloc`): the stand-in for `UnifiedMat<usize>`
/// (bench.rs:130, 466-468). The engine's result copy into it is a DMA at full PCIe rate; a `Vec<u64>` works too,
/// at about a third of that.
pub struct UnifiedMat { ptr: *mut u64, rows: usize, cols: usize }
impl UnifiedMat {
    pub fn new(rows: usize, cols: usize) -> Result<Self, Error> {
        let (mut p, mut err) = (std::ptr::null_mut(), std::ptr::null());
        check(unsafe { swb_host_alloc(rows * cols * 8, &mut p, &mut err) }, err)?;
        unsafe { std::ptr::write_bytes(p as *mut u8, 0, rows * cols * 8) };
        Ok(Self { ptr: p as *mut u64, rows, cols })
    }
    pub fn rows(&self) -> usize { self.rows }
    pub fn cols(&self) -> usize { self.cols }
    pub fn as_slice(&self) -> &[u64] { unsafe { std::slice::from_raw_parts(self.ptr, self.rows * self.cols) } }
    pub fn as_mut_slice(&mut self) -> &mut [u64] { unsafe { std::slice::from_raw_parts_mut(self.ptr, self.rows * self.cols) } }
}
impl Drop for UnifiedMat { fn drop(&mut self) { unsafe { swb_host_free(self.ptr as *mut c_void) } } }

/// `gpu_multiprocessor_count(device)` (utils.rs:826-836).
pub fn multiprocessor_count(device: usize) -> Option<usize> {
    let (mut n, mut err) = (0 as c_int, std::ptr::null());
    if unsafe { swb_multiprocessor_count(device as c_int, &mut n, &mut err) } == 0 { Some(n as usize) } else { None }
}
