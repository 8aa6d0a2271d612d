/*
 * stringwars_b200.h — C ABI of the B200-native batched edit-distance engine.
 *
 * This is the drop-in boundary for ONE path of ashvardanian/StringWars: the batched similarity engines
 * that `similarities/bench.rs` times. Each entry point names the reference interface it replaces
 * (file:line relative to the StringWars checkout). The reference reaches its engines through the
 * `stringzilla::szs` Rust types; a Rust (or ctypes / cgo) binding over THIS header gives the same shape —
 * see INTEGRATION.md for the binding a StringWars maintainer would add.
 *
 * Conventions
 *  - Plain C: opaque handles, raw pointers and sizes; no C++/torch types cross this boundary.
 *  - Every fallible call returns `swb_status_t` (0 = OK) and, when `err` is non-NULL, stores a pointer to a
 *    thread-local, NUL-terminated message valid until the next call on the same thread. Nothing aborts or
 *    throws across the boundary (reference behaviour: `Result<_, E: Display>`, bench.rs:390-399, 599-603).
 *  - Strings arrive as "tapes": one contiguous byte buffer plus N+1 monotonically increasing u64 offsets
 *    (Arrow LargeBinary layout == `BytesTapeView<u64>`, bench.rs:62, 292-306). Tapes are read-only.
 *  - The caller owns inputs and the preallocated result buffer (bench.rs:130, 466-468); the library never
 *    frees caller memory and performs no allocation in the steady state of a repeated call.
 *  - Host-pointer calls are blocking: results are host-visible on return (bench.rs:145 `black_box(&matrix)`).
 *  - One in-flight call per engine handle; handles are not thread-safe (the harness is single-threaded).
 *  - There is NO CPU fallback. Without a CUDA device `swb_scope_create` fails with SWB_NO_DEVICE.
 */
#ifndef STRINGWARS_B200_H_
#define STRINGWARS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SWB_API
#else
#define SWB_API __attribute__((visibility("default")))
#endif

typedef enum swb_status_t {
    SWB_OK = 0,
    SWB_BAD_ALLOC = 1,     /* host or device allocation failed */
    SWB_INVALID_ARG = 2,   /* NULL pointer, non-monotonic offsets, unsupported cost model, ... */
    SWB_NO_DEVICE = 3,     /* no usable CUDA device / device ordinal out of range */
    SWB_CUDA_ERROR = 4,    /* a CUDA runtime call or kernel failed; message carries cudaGetErrorString */
    SWB_TOO_LONG = 5,      /* an input string exceeds what the selected kernel supports */
    SWB_UNSUPPORTED = 6,   /* feature not built into this library */
    SWB_NCCL_ERROR = 7     /* libnccl could not be loaded, or an NCCL call of the multi-GPU scatter / gather failed */
} swb_status_t;

/* Kernel family selection, see DESIGN.md. AUTO picks the fastest exact kernel for the cost model. */
typedef enum swb_algorithm_t {
    SWB_ALGO_AUTO = 0,
    SWB_ALGO_WAVEFRONT = 1, /* anti-diagonal integer DP, sub-warp group per pair, min passed by warp shuffle */
    SWB_ALGO_BITPARALLEL = 2 /* Myers/Hyyrö bit-vector recurrence (unit costs only), nibble tables in smem */
} swb_algorithm_t;

/* What a length-k gap costs in the NW/SW engines (the reference's two contenders disagree). */
typedef enum swb_gap_convention_t {
    SWB_GAP_STRINGZILLA = 0, /* open + (k-1)*extend — `NeedlemanWunschScores::new(.., open, extend)`, bench.rs:658 */
    SWB_GAP_BIO = 1          /* open + k*extend     — `Aligner::with_capacity(.., open, extend, ..)`, bench.rs:746 */
} swb_gap_convention_t;

typedef struct swb_scope swb_scope_t;   /* replaces `szs::DeviceScope`            (bench.rs:376-379) */
typedef struct swb_engine swb_engine_t; /* replaces the `szs` Distances / Scores engines (bench.rs:382-399, 658-674) */

/* ---- library / device queries ------------------------------------------------------------------------- */

/* Library version, 10000*major + 100*minor + patch. */
SWB_API int swb_version(void);

/* Number of visible CUDA devices (0 when there is no driver / device). */
SWB_API int swb_device_count(void);

/* Streaming-multiprocessor count of `device`; replaces `gpu_multiprocessor_count` (utils.rs:826-836), which
 * the harness uses to size the GPU batch (bench.rs:286-289). */
SWB_API swb_status_t swb_multiprocessor_count(int device, int *out, const char **err);

/* Total kernels this library has launched in this process (bench.py's `gpu_launches` evidence). */
SWB_API uint64_t swb_launch_count(void);

/* ---- scopes -------------------------------------------------------------------------------------------- */

/* `DeviceScope::gpu_device(0)` (bench.rs:379). `num_devices` > 1 makes every host-pointer call shard its
 * batch over devices [first_device, first_device + num_devices) by contiguous pair ranges balanced on cells. */
SWB_API swb_status_t swb_scope_create(int first_device, int num_devices, swb_scope_t **out, const char **err);
/* Same, over an explicit list of device ordinals. An ordinal may repeat: every entry gets its own streams and
 * scratch, so a one-GPU box exercises the whole multi-device sharding path (`<{n}gpu>` variants, bench.rs:498-501). */
SWB_API swb_status_t swb_scope_create_devices(const int *devices, int num_devices, swb_scope_t **out, const char **err);
SWB_API void swb_scope_destroy(swb_scope_t *scope);
SWB_API int swb_scope_num_devices(const swb_scope_t *scope);

/* ---- pinned host memory ("UnifiedAlloc" replacement) ----------------------------------------------------
 * The reference keeps its tape in CUDA managed memory (`BytesTape<u64, UnifiedAlloc>`, bench.rs:292). The
 * B200-native equivalent is page-locked host memory + explicit async copies; tapes and result buffers
 * allocated here make the host-pointer calls copy at full PCIe rate. Ordinary malloc'd memory also works. */
SWB_API swb_status_t swb_host_alloc(size_t bytes, void **out, const char **err);
SWB_API void swb_host_free(void *ptr);
/* The reference's allocator itself: CUDA managed memory (`UnifiedAlloc`, bench.rs:292; `UnifiedMat`, bench.rs:130).
 * Tapes and matrices allocated here are used by the `_view` calls IN PLACE (no staging copy): pages migrate to the
 * device on first touch and stay there until the host touches them again. */
SWB_API swb_status_t swb_unified_alloc(size_t bytes, void **out, const char **err);
SWB_API void swb_unified_free(void *ptr);

/* ---- engines ------------------------------------------------------------------------------------------- */

/* `LevenshteinDistances::new(&scope, match, mismatch, open, extend)` (bench.rs:382-393). Only the cost model
 * the harness uses, (0, 1, 1, 1), is accepted; anything else returns SWB_INVALID_ARG. Byte-level symbols. */
SWB_API swb_status_t swb_levenshtein_create(swb_scope_t *scope, int match, int mismatch, int open, int extend,
                                            swb_engine_t **out, const char **err);

/* `LevenshteinDistancesUtf8::new(&scope, 0, 1, 1, 1)` (bench.rs:386-399): symbols are Unicode scalar values of
 * valid UTF-8 input (`rapidfuzz ... distance(a.chars(), b.chars())`, bench.rs:437). */
SWB_API swb_status_t swb_levenshtein_utf8_create(swb_scope_t *scope, int match, int mismatch, int open,
                                                 int extend, swb_engine_t **out, const char **err);

/* `NeedlemanWunschScores::new(&scope, &byte_to_class, &class_costs, open, extend)` (bench.rs:658-662):
 * `byte_to_class` has 256 entries < 32, `class_costs` is 32x32 row-major i8 (bench.rs:98-108). */
SWB_API swb_status_t swb_needleman_wunsch_create(swb_scope_t *scope, const uint8_t *byte_to_class,
                                                 const int8_t *class_costs, int open, int extend,
                                                 swb_engine_t **out, const char **err);

/* `SmithWatermanScores::new(..)` (bench.rs:663-666): best local alignment score (>= 0). */
SWB_API swb_status_t swb_smith_waterman_create(swb_scope_t *scope, const uint8_t *byte_to_class,
                                               const int8_t *class_costs, int open, int extend,
                                               swb_engine_t **out, const char **err);

SWB_API void swb_engine_destroy(swb_engine_t *engine);

/* Options (default AUTO / SWB_GAP_STRINGZILLA). */
SWB_API swb_status_t swb_engine_set_algorithm(swb_engine_t *engine, swb_algorithm_t algorithm, const char **err);
SWB_API swb_status_t swb_engine_set_gap_convention(swb_engine_t *engine, swb_gap_convention_t convention,
                                                   const char **err);
/* Name of the compute kernel the last call dispatched to (static string). */
SWB_API const char *swb_engine_last_kernel(const swb_engine_t *engine);

/* Kernel timing for roofline reports: while enabled, every compute-kernel launch (not the schedule kernels) is
 * bracketed by CUDA events on the stream it is launched on. `swb_engine_kernel_time` synchronises those
 * events, returns the summed device time in milliseconds and the number of launches since the last call, and
 * clears the record. */
SWB_API swb_status_t swb_engine_set_profiling(swb_engine_t *engine, int enabled, const char **err);
SWB_API swb_status_t swb_engine_kernel_time(swb_engine_t *engine, double *total_ms, uint64_t *launches,
                                            const char **err);

/* ---- compute: host tapes ------------------------------------------------------------------------------- */

/* Element-wise pairs: out[i] = score(a[i], b[i]) for i < n_pairs. This is the shape BASELINE.json's configs
 * use and the batched form of the scalar contender surface `(a:&[u8], b:&[u8]) -> usize` (bench.rs:416-419);
 * precedent for "one result per row pair" is the cuDF row (bench.py:596-604). Distances are written as u64
 * (`usize`, bench.rs:130); NW/SW scores as i64 (`isize`, bench.rs:193) — pass the i64 buffer cast to u64*. */
SWB_API swb_status_t swb_compute_pairs(swb_engine_t *engine, const uint8_t *a_data, const uint64_t *a_offsets,
                                       const uint8_t *b_data, const uint64_t *b_offsets, size_t n_pairs,
                                       uint64_t *out, const char **err);

/* Cross product: out[q * row_stride + c] = score(queries[q], candidates[c]) — `engine.compute_into(&scope,
 * queries_tape, Some(candidates_tape), &mut matrix)` (bench.rs:599-603, 123-148). */
SWB_API swb_status_t swb_compute_cross(swb_engine_t *engine, const uint8_t *q_data, const uint64_t *q_offsets,
                                       size_t n_queries, const uint8_t *c_data, const uint64_t *c_offsets,
                                       size_t n_candidates, uint64_t *out, size_t row_stride, const char **err);

/* Scalar convenience, a batch of one through the same kernels: the literal `(a:&[u8], b:&[u8]) -> usize`
 * surface of bench.rs:416-419. Launch-latency bound; for tests and API compatibility, not for benchmarking.
 * Returns UINT64_MAX on failure. */
SWB_API uint64_t swb_levenshtein_pair(swb_engine_t *engine, const uint8_t *a, size_t a_length, const uint8_t *b,
                                      size_t b_length, const char **err);

/* ---- compute: device-resident tapes -------------------------------------------------------------------- */

/* Same as swb_compute_pairs but every pointer is device memory on scope device `device_index` and the work
 * is enqueued on `cuda_stream` (a cudaStream_t; NULL = the legacy default stream) without host
 * synchronisation, except for one small internal readback when the batch needs length bucketing. Offsets
 * need not start at 0. This is what multi-process callers (one rank per GPU) and the kernel-only bench use.
 * Tape bytes are READ in whole aligned 16-byte pieces: up to 15 bytes before the first and after the last byte
 * of a string may be read (never written, never used). Any cudaMalloc'ed tape satisfies this. */
SWB_API swb_status_t swb_compute_pairs_device(swb_engine_t *engine, int device_index, const uint8_t *a_data,
                                              const uint64_t *a_offsets, const uint8_t *b_data,
                                              const uint64_t *b_offsets, size_t n_pairs, uint64_t *out,
                                              void *cuda_stream, const char **err);

SWB_API swb_status_t swb_compute_cross_device(swb_engine_t *engine, int device_index, const uint8_t *q_data,
                                              const uint64_t *q_offsets, size_t n_queries, const uint8_t *c_data,
                                              const uint64_t *c_offsets, size_t n_candidates, uint64_t *out,
                                              size_t row_stride, void *cuda_stream, const char **err);

/* ---- typed tape views, result widths, resident tapes ----------------------------------------------------
 * The reference hands its engines `AnyBytesTape::View64(view)` (bench.rs:141-142): an enum over offset widths,
 * over a tape that lives in CUDA managed memory (`BytesTape<u64, UnifiedAlloc>`, bench.rs:292-295), and a
 * `UnifiedMat` result matrix (bench.rs:130) that the device writes in place. `swb_tape_view_t` is that enum as
 * a plain struct, and the `_view` calls below accept ANY memory for every pointer — each is classified with
 * cudaPointerGetAttributes:
 *   pageable / page-locked host   staged: chunk-pipelined H2D copies (tape) / device staging + D2H (results);
 *   managed (cudaMallocManaged)   used in place, no copy: pages migrate on first touch and then stay resident,
 *                                 exactly like the reference's tape after its first call;
 *   device (cudaMalloc)           used in place on the owning device (it must be a device of the scope).
 * Results may be narrower than the ABI's usize / isize: a Levenshtein distance below 65,536 fits u16, an
 * alignment score fits i32; four times fewer bytes cross PCIe when the matrix lives in host memory. Values that
 * do not fit are truncated to the low bits (two's complement), so pick a width that holds them. */
typedef struct swb_tape_view_t {
    const uint8_t *data;  /* base of the byte buffer the offsets index into */
    const void *offsets;  /* n + 1 monotonically non-decreasing offsets, u64 or u32 (offset_bits) */
    size_t n;             /* number of strings */
    int offset_bits;      /* 64 or 32 */
} swb_tape_view_t;

typedef enum swb_result_type_t {
    SWB_RESULT_64 = 0, /* u64 distances / i64 scores: `UnifiedMat<usize>` / `UnifiedMat<isize>` (bench.rs:130, 193) */
    SWB_RESULT_32 = 1, /* u32 / i32 */
    SWB_RESULT_16 = 2  /* u16 / i16 */
} swb_result_type_t;

/* swb_compute_pairs / swb_compute_cross over typed views, any memory kind, any result width (blocking). */
SWB_API swb_status_t swb_compute_pairs_view(swb_engine_t *engine, const swb_tape_view_t *a, const swb_tape_view_t *b,
                                            void *out, swb_result_type_t out_type, const char **err);
SWB_API swb_status_t swb_compute_cross_view(swb_engine_t *engine, const swb_tape_view_t *queries,
                                            const swb_tape_view_t *candidates, void *out, size_t row_stride,
                                            swb_result_type_t out_type, const char **err);

/* swb_compute_pairs_device (device memory of scope device `device_index`, enqueued on `cuda_stream`, no final
 * synchronisation) over typed views and any result width. */
SWB_API swb_status_t swb_compute_pairs_device_view(swb_engine_t *engine, int device_index, const swb_tape_view_t *a,
                                                   const swb_tape_view_t *b, void *out, swb_result_type_t out_type,
                                                   void *cuda_stream, const char **err);

/* Resident tapes: upload once, compute many times — what `BytesTape<u64, UnifiedAlloc>` gives the reference after
 * the first touch (bench.rs:292-306; the timing loop at bench.rs:587-605 never copies). The tape is copied to
 * every device of the scope (a cross product needs all candidates everywhere; pairs are sharded by range).
 * `first` / `n` select a sub-range like `BytesTapeView::subview(lo, hi)` (bench.rs:134-139). */
typedef struct swb_tape swb_tape_t;
SWB_API swb_status_t swb_tape_upload(swb_scope_t *scope, const swb_tape_view_t *view, swb_tape_t **out, const char **err);
SWB_API void swb_tape_destroy(swb_tape_t *tape);
SWB_API size_t swb_tape_len(const swb_tape_t *tape);
/* Device view of strings [first, first + n) of the tape's copy on scope device `device_index` (device pointers). */
SWB_API swb_status_t swb_tape_device_view(const swb_tape_t *tape, int device_index, size_t first, size_t n,
                                          swb_tape_view_t *out, const char **err);
SWB_API swb_status_t swb_compute_pairs_tape(swb_engine_t *engine, const swb_tape_t *a, size_t a_first, const swb_tape_t *b,
                                            size_t b_first, size_t n_pairs, void *out, swb_result_type_t out_type,
                                            const char **err);
SWB_API swb_status_t swb_compute_cross_tape(swb_engine_t *engine, const swb_tape_t *queries, size_t q_first, size_t n_queries,
                                            const swb_tape_t *candidates, size_t c_first, size_t n_candidates, void *out,
                                            size_t row_stride, swb_result_type_t out_type, const char **err);

/* ---- multi-GPU: one NCCL scatter, per-device kernels, one NCCL gather (BASELINE.json north_star) ---------
 * `LevenshteinDistances<{n}gpu>` (multi-GPU variant naming, bench.rs:498-501) for a batch that is RESIDENT ON THE
 * SCOPE'S FIRST DEVICE: both views and `out` are device memory of scope device 0 (the root). The library owns a
 * single-process NCCL communicator over the scope's devices (ncclCommInitAll, created on first use; libnccl is
 * loaded at run time). The pair range is cut into contiguous shards balanced by bytes; every shard travels in
 * chunks — grouped ncclSend / ncclRecv of the tape slice and of u32 offsets rebuilt relative to the slice — and each
 * chunk's kernel starts as soon as it has landed, while the next chunk is still on NVLink; results return to `out`
 * on the root the same way. The root's own shard never moves. Blocking.
 * `stats`, when non-NULL, receives {scatter bytes sent by the root, gather bytes received by the root,
 * device ms from first send to last result (root clock), 0}. */
SWB_API swb_status_t swb_compute_pairs_scatter(swb_engine_t *engine, const swb_tape_view_t *a, const swb_tape_view_t *b,
                                               void *out, swb_result_type_t out_type, double stats[4], const char **err);

/* What carries the chunks of swb_compute_pairs_scatter between GPUs. NCCL (default) is what the north_star names: grouped
 * ncclSend / ncclRecv, SM-driven copies. COPY_ENGINE moves the very same chunks with cudaMemcpyPeerAsync over NVLink (DMA
 * engines, no SMs, no communicator), one stream per destination; everything else — cuts, chunking, overlap with the
 * kernels, the gather — is identical. */
typedef enum swb_transport_t { SWB_TRANSPORT_NCCL = 0, SWB_TRANSPORT_COPY_ENGINE = 1 } swb_transport_t;
SWB_API swb_status_t swb_scope_set_transport(swb_scope_t *scope, swb_transport_t transport, const char **err);

/* The same sharded call with NO scatter step: the kernels of every device read the root's tapes in place over NVLink
 * (peer loads through the NVSwitch; cudaDeviceEnablePeerAccess, enabled on first use) and store their results straight
 * into `out` on the root. Transfer and compute overlap load by load — the fused form of "scatter, then compute", with no
 * staging buffers, no chunking and no collective. Same arguments and the same `stats` layout as
 * swb_compute_pairs_scatter ({bytes read over NVLink, result bytes written over NVLink, milliseconds, 0}). */
SWB_API swb_status_t swb_compute_pairs_peer(swb_engine_t *engine, const swb_tape_view_t *a, const swb_tape_view_t *b,
                                            void *out, swb_result_type_t out_type, double stats[4], const char **err);

/* Pair-load phase only: stages every pair's bytes through shared memory with the same coalesced 16-byte
 * loads the compute kernels use and writes a byte-sum per pair. Exists so the HBM-read rate of the load
 * phase can be measured on its own (BASELINE.json: ">= 60% of HBM peak on the pair-load phase"). */
SWB_API swb_status_t swb_load_phase_device(swb_engine_t *engine, int device_index, const uint8_t *a_data,
                                           const uint64_t *a_offsets, const uint8_t *b_data,
                                           const uint64_t *b_offsets, size_t n_pairs, uint64_t *out,
                                           void *cuda_stream, const char **err);

#ifdef __cplusplus
}
#endif
#endif /* STRINGWARS_B200_H_ */
