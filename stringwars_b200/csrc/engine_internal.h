// engine_internal.h — host-side objects and launcher declarations shared by the translation units of the library.
//
// The library is built from several .cu files so that nvcc compiles the kernel families in parallel:
//   engine.cu        C ABI part 1: errors, queries, scopes, engines, options, profiling
//   pipeline.cu      C ABI part 2: schedule / UTF-8 launches, device-resident calls, host pipelines, resident tapes
//   launch_lev.cu    bit-vector Levenshtein kernels (lev_bitpar*, lev_short)
//   launch_align.cu  scalar wavefront kernels (align_wavefront)
//   launch_packed.cu packed s16x2 wavefront kernels
//   multi_gpu.cu     NCCL scatter / gather inside the boundary
// Nothing here crosses the C ABI.
#pragma once
#include "../../include/stringwars_b200.h"

#include <atomic>
#include <cstdint>
#include <set>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "align_costs.h" // AlignCosts, AlignTables (plain structs)
#include "common.cuh"
#include "schedule.cuh"

namespace swb {

// ------------------------------------------------------------------------------------------ errors
swb_status_t fail(const char **err, swb_status_t st, const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define SWB_CUDA(call)                                                                                              \
    do {                                                                                                            \
        cudaError_t e_ = (call);                                                                                    \
        if (e_ != cudaSuccess)                                                                                      \
            return ::swb::fail(err, e_ == cudaErrorMemoryAllocation ? SWB_BAD_ALLOC : SWB_CUDA_ERROR, "%s failed: %s", \
                               #call, cudaGetErrorString(e_));                                                      \
    } while (0)

// ------------------------------------------------------------------------------------------ objects
struct DeviceBuffer {
    void *ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr, cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr, cap = 0;
    }
};

// Page-locked host staging (narrow results on their way to the caller's matrix).
struct PinnedBufferHost {
    void *ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr, cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaHostAlloc(&ptr, want, cudaHostAllocPortable);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr, cap = 0;
    }
};

struct DeviceCtx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    int concurrent_managed = 0; // the host may touch managed memory without synchronising the device first
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
    cudaStream_t side_stream = nullptr; // schedule kernel of speculative device-resident calls
    cudaEvent_t ev_inputs = nullptr;    // "caller's stream has reached this call": inputs are ready
    // The length classes of a ragged batch run beside each other: class kernels have tails (a wave of the longest pairs)
    // and small classes do not fill the machine; the largest class stays on the caller's stream.
    static constexpr int CLASS_STREAMS = 3;
    cudaStream_t class_streams[CLASS_STREAMS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_class[CLASS_STREAMS] = {nullptr, nullptr, nullptr};
    ScheduleStats *d_stats = nullptr;
    ScheduleStats *h_stats = nullptr;     // pinned + mapped: the schedule kernel publishes into it directly
    ScheduleStats *h_stats_dev = nullptr; // device alias of h_stats
    uint32_t stats_seq = 0;               // sequence number of the last schedule launch
    SpecReport *d_spec = nullptr;         // device: counters of a self-validating speculative launch
    SpecPublished *h_spec = nullptr;      // pinned + mapped: what its last block publishes
    SpecPublished *h_spec_dev = nullptr;  // device alias of h_spec
    uint64_t *h_words = nullptr;          // pinned: 8 words for small synchronous readbacks (UTF-8 survey header, tape spans)
    uint32_t spec_seq = 0;
    cudaStream_t stats_stream = nullptr;  // stream that launch ran on (polled for errors while waiting)
    cudaEvent_t ev_offsets = nullptr;
    std::vector<cudaEvent_t> ev_chunks;
    DeviceBuffer order, a_data, a_offs, b_data, b_offs, out;
    DeviceBuffer sym_a, sym_a_offs, sym_b, sym_b_offs, scan_tmp, flags; // UTF-8 engine scratch
    DeviceBuffer ren_a, ren_a_offs, ren_b, ren_b_offs;                   // UTF-8 engine: pairs renamed to byte alphabets
    DeviceBuffer batch_dict;                                             // UTF-8 engine: the batch-wide dictionary of a cross product
    DeviceBuffer hscratch; // per-column delta bits / boundary rows between row strips of the long-string kernels
    // multi_gpu.cu: receive buffers of a scattered shard, streams / events of the NCCL exchange
    DeviceBuffer rx_a, rx_b, rx_a_offs, rx_b_offs, rx_out;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
};

} // namespace swb

struct swb_engine;

struct swb_tape {
    swb_scope *scope = nullptr;                 // nullptr once the scope has been destroyed under the tape
    std::vector<swb::DeviceBuffer> data, offs; // one copy per scope device (every device may need every string in cross mode)
    uint64_t n = 0;                             // strings
    uint64_t bytes = 0;                         // offsets[n] - offsets[0]
    uint64_t base = 0;                          // offsets[0]: device data pointers are biased by -base
    int offset_bits = 64;
    std::vector<uint64_t> h_offs64;             // host copy of the offsets (shard / chunk cuts), widened
};

struct swb_scope {
    std::vector<swb::DeviceCtx> devs;
    std::set<swb_engine *> engines; // live engines (their device state is released with the scope)
    std::set<swb_tape *> tapes;
    void *nccl = nullptr;           // multi_gpu.cu: NcclState (lazily created by the first scatter call)
    void *peer = nullptr;           // multi_gpu.cu: PeerState (lazily created by the first peer-load call)
    swb_transport_t transport = SWB_TRANSPORT_NCCL; // what carries the chunks of swb_compute_pairs_scatter
    bool peers_enabled = false;     // direct peer access root <-> every device has been enabled
};

enum EngineKind { ENGINE_LEVENSHTEIN = 0, ENGINE_LEVENSHTEIN_UTF8 = 1, ENGINE_NW = 2, ENGINE_SW = 3 };

struct ProfEvents {
    cudaEvent_t start = nullptr, stop = nullptr;
    int device = -1;
};

struct swb_engine {
    swb_scope *scope = nullptr; // nullptr once the scope has been destroyed under the engine
    EngineKind kind = ENGINE_LEVENSHTEIN;
    swb_algorithm_t algorithm = SWB_ALGO_AUTO;
    swb_gap_convention_t convention = SWB_GAP_STRINGZILLA;
    int open = 0, extend = 0;
    bool unary = true;
    int match = 0, mismatch = 0;
    swb::AlignTables tables;
    std::vector<swb::AlignTables *> d_tables; // per device
    const char *last_kernel = "none";
    uint32_t spec_cross_q = 0, spec_cross_n = 0; // cross products: longest query / candidate of the last call that took the cross kernel (0: do not speculate)
    uint32_t spec_max_n = 0; // longest text of that batch: the speculative kernel skips anything longer (unvalidated lengths)
    int spec_class = 0; // kernel code (launch_lev_bitpar's K) of the last uniform batch on the device-resident path; 0: do not speculate
    bool profiling = false;
    std::vector<ProfEvents> prof_events; // pool
    size_t prof_used = 0;
};

namespace swb {

// RAII bracket around one compute-kernel launch when the engine is profiling. Events belong to the device they
// were created on: a pool slot last used on another device is re-created.
struct ProfileBracket {
    swb_engine *eng;
    cudaStream_t stream;
    cudaEvent_t stop = nullptr;
    ProfileBracket(swb_engine *e, int device, cudaStream_t s) : eng(e), stream(s) {
        if (!eng->profiling) return;
        if (eng->prof_used == eng->prof_events.size()) eng->prof_events.push_back(ProfEvents{});
        ProfEvents &pe = eng->prof_events[eng->prof_used];
        if (pe.device != device) {
            if (pe.start) cudaEventDestroy(pe.start);
            if (pe.stop) cudaEventDestroy(pe.stop);
            pe.start = pe.stop = nullptr;
            if (cudaEventCreate(&pe.start) != cudaSuccess || cudaEventCreate(&pe.stop) != cudaSuccess) {
                pe.device = -1;
                return;
            }
            pe.device = device;
        }
        eng->prof_used++;
        cudaEventRecord(pe.start, stream);
        stop = pe.stop;
    }
    ~ProfileBracket() {
        if (stop) cudaEventRecord(stop, stream);
    }
};

// Opt a kernel in to `bytes` of dynamic shared memory (> 48 KB needs the attribute).
cudaError_t allow_smem(const void *func, int device, size_t bytes);

static inline uint32_t round16(uint32_t x) { return (x + 15u) & ~15u; }

// ------------------------------------------------------------------------------------------ pipeline.cu
// Run a fully device-resident PairSpace on scope device `dev_index`: schedule, (scatter), per-class kernels.
// `completed` (optional) is set when the call has already SEEN the batch finish (a self-validating launch whose report
// arrived): results are written and visible device-wide, no stream synchronisation is needed to know it.
swb_status_t run_device_batch(swb_engine *eng, int dev_index, PairSpace ps, cudaStream_t stream, const char **err, bool *completed = nullptr);
constexpr uint64_t SPAN_UNKNOWN = ~0ull; // run_device_batch_utf8: the caller cannot see the offsets (device / managed memory); the engine reads them
// Same for the code-point engine (decodes both tapes first). n_a / n_b: strings per tape; *_bytes: their byte spans.
swb_status_t run_device_batch_utf8(swb_engine *eng, int dev_index, PairSpace ps, uint64_t n_a, uint64_t n_b, uint64_t a_bytes,
                                   uint64_t b_bytes, cudaStream_t stream, const char **err);
inline PairSpace make_pair_space(const uint8_t *a_data, const void *a_offs, const uint8_t *b_data, const void *b_offs, uint64_t n_pairs,
                                 uint64_t n_c, uint64_t stride, void *out, uint32_t out_narrow, uint32_t offs32) {
    PairSpace ps;
    ps.a_data = a_data, ps.a_offs = (const uint64_t *)a_offs, ps.b_data = b_data, ps.b_offs = (const uint64_t *)b_offs;
    ps.n_pairs = n_pairs, ps.n_c = n_c, ps.stride = stride, ps.order = nullptr, ps.out = (uint64_t *)out;
    ps.len_shift = 0, ps.out_narrow = out_narrow, ps.offs32 = offs32, ps.idx_shift = 0;
    return ps;
}
static inline size_t result_bytes(uint32_t out_narrow) { return out_narrow == 0 ? 8 : (out_narrow == 1 ? 4 : 2); }
swb_status_t check_engine_alive(swb_engine *engine, const char **err);

// ------------------------------------------------------------------------------------------ launchers
// launch_lev.cu: one length class of byte Levenshtein by the bit-vector kernels (any pattern length).
// min_m: the class's shortest pattern when the caller knows it (0: unknown) — a class whose every pattern fills its K words
// exactly can take the software-pipelined kernel
swb_status_t launch_lev_class(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint64_t begin, uint64_t count, uint32_t max_m,
                              uint32_t max_n, cudaStream_t stream, const char **err, uint32_t min_m = 0);
// thread-per-pair kernel of word count K = 1, 2, 4, 8 (or the register kernel of -K pattern words) over the whole pair
// space — the speculative launch: pairs that are malformed, do not fit, or whose longer string exceeds n_cap are skipped and
// reported through `sa`
swb_status_t launch_lev_bitpar(swb_engine *eng, DeviceCtx &ctx, int K, const PairSpace &ps, uint64_t begin, uint64_t count, uint32_t n_cap,
                               const SpecArgs &sa, cudaStream_t stream, const char **err);
// cross product with per-query table reuse; max_q <= 256
swb_status_t launch_lev_cross(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint32_t n_queries, uint32_t cands_per_block,
                              uint32_t max_q, uint32_t n_cap, const SpecArgs &sa, cudaStream_t stream, const char **err);

// launch_align.cu / launch_packed.cu: wavefront kernels (Levenshtein in wavefront mode, UTF-8 symbols, NW, SW).
swb_status_t launch_wavefront(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin, uint64_t count,
                              uint32_t max_m, uint32_t max_n, cudaStream_t stream, const char **err, bool same_shape);
// SWB_TOO_LONG from the packed launchers means "not applicable / does not fit": the caller falls back to the scalar kernel.
swb_status_t launch_wavefront_packed_any(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin,
                                         uint64_t count, const AlignCosts &cs, bool local, bool affine, uint32_t max_m, uint32_t max_n,
                                         cudaStream_t stream, const char **err);

} // namespace swb
