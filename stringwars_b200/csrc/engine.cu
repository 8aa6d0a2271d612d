// engine.cu — C ABI (include/stringwars_b200.h) and host orchestration of the sm_100a kernels.
//
// Host design: one scope owns one context per device (streams, scratch, schedule counters);
// host-pointer calls shard the pair range over the scope's devices by bytes, then per device pipeline
// offsets H2D -> schedule counters -> chunked data H2D overlapped with compute -> results D2H. There is no
// CPU compute path anywhere in this file: without a device every entry point fails with SWB_NO_DEVICE.
#include "../../include/stringwars_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <utility>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "align_wavefront.cuh"
#include "align_wavefront_packed.cuh"
#include "align_wavefront_packed_general.cuh"
#include "lev_bitpar.cuh"
#include "lev_bitpar_coop.cuh"
#include "lev_bitpar_cross.cuh"
#include "schedule.cuh"
#include "utf8.cuh"

using namespace swb;

// ------------------------------------------------------------------------------------------ errors

static thread_local char g_err[512];
static std::atomic<uint64_t> g_launches{0};

static swb_status_t fail(const char **err, swb_status_t st, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (err) *err = g_err;
    return st;
}

#define SWB_CUDA(call)                                                                                         \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess)                                                                                 \
            return fail(err, e_ == cudaErrorMemoryAllocation ? SWB_BAD_ALLOC : SWB_CUDA_ERROR, "%s failed: %s", \
                        #call, cudaGetErrorString(e_));                                                        \
    } while (0)

// ------------------------------------------------------------------------------------------ tracing
// SWB_TRACE=1 prints host-side phase timestamps of the host-tape pipeline to stderr (diagnostics only).
#include <chrono>
static bool trace_enabled() {
    static const bool on = getenv("SWB_TRACE") && getenv("SWB_TRACE")[0] == '1';
    return on;
}
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

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

struct DeviceCtx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
    cudaStream_t side_stream = nullptr; // schedule kernel of speculative device-resident calls
    cudaEvent_t ev_inputs = nullptr;    // "caller's stream has reached this call": inputs are ready
    ScheduleStats *d_stats = nullptr;
    ScheduleStats *h_stats = nullptr;     // pinned + mapped: the schedule kernel publishes into it directly
    ScheduleStats *h_stats_dev = nullptr; // device alias of h_stats
    uint32_t stats_seq = 0;               // sequence number of the last schedule launch
    cudaStream_t stats_stream = nullptr;  // stream that launch ran on (polled for errors while waiting)
    cudaEvent_t ev_offsets = nullptr;
    std::vector<cudaEvent_t> ev_chunks;
    DeviceBuffer order, a_data, a_offs, b_data, b_offs, out;
    DeviceBuffer sym_a, sym_a_offs, sym_b, sym_b_offs, scan_tmp, flags; // UTF-8 engine scratch
    DeviceBuffer hscratch; // per-column delta bits between row strips of the lane-group bit-vector kernel
};

struct swb_scope {
    std::vector<DeviceCtx> devs;
};

enum EngineKind { ENGINE_LEVENSHTEIN = 0, ENGINE_LEVENSHTEIN_UTF8 = 1, ENGINE_NW = 2, ENGINE_SW = 3 };

struct swb_engine {
    swb_scope *scope = nullptr;
    EngineKind kind = ENGINE_LEVENSHTEIN;
    swb_algorithm_t algorithm = SWB_ALGO_AUTO;
    swb_gap_convention_t convention = SWB_GAP_STRINGZILLA;
    int open = 0, extend = 0;
    bool unary = true;
    int match = 0, mismatch = 0;
    AlignTables tables;
    std::vector<AlignTables *> d_tables; // per device
    const char *last_kernel = "none";
    int spec_class = -1; // pattern class of the last uniform batch on the device-resident path (-1: do not speculate)
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events; // pool
    std::vector<int> prof_device;                                  // device ordinal of each recorded pair
    size_t prof_used = 0;
};

// RAII bracket around one compute-kernel launch when the engine is profiling.
struct ProfileBracket {
    swb_engine *eng;
    cudaStream_t stream;
    cudaEvent_t stop = nullptr;
    ProfileBracket(swb_engine *e, int device, cudaStream_t s) : eng(e), stream(s) {
        if (!eng->profiling) return;
        if (eng->prof_used == eng->prof_events.size()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            eng->prof_events.push_back({a, b});
            eng->prof_device.push_back(device);
        }
        eng->prof_device[eng->prof_used] = device;
        auto &pr = eng->prof_events[eng->prof_used++];
        cudaEventRecord(pr.first, stream);
        stop = pr.second;
    }
    ~ProfileBracket() {
        if (stop) cudaEventRecord(stop, stream);
    }
};

static std::mutex g_attr_mutex;
static std::map<std::pair<const void *, int>, size_t> g_attr_done;

// Opt a kernel in to `bytes` of dynamic shared memory (> 48 KB needs the attribute). The attribute is raised to
// exactly what has been asked for so far: kernels with static shared memory cannot take the full 227 KB.
static cudaError_t allow_smem(const void *func, int device, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    auto it = g_attr_done.find({func, device});
    if (it != g_attr_done.end() && it->second >= bytes) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) g_attr_done[{func, device}] = bytes;
    return e;
}

// ------------------------------------------------------------------------------------------ queries

extern "C" int swb_version(void) { return 100; }

extern "C" int swb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" swb_status_t swb_multiprocessor_count(int device, int *out, const char **err) {
    if (!out) return fail(err, SWB_INVALID_ARG, "out is NULL");
    if (device < 0 || device >= swb_device_count()) return fail(err, SWB_NO_DEVICE, "no CUDA device %d", device);
    SWB_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, device));
    return SWB_OK;
}

extern "C" uint64_t swb_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------ scope

static void destroy_ctx(DeviceCtx &c) {
    cudaSetDevice(c.device);
    if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
    if (c.compute_stream) cudaStreamDestroy(c.compute_stream);
    if (c.side_stream) cudaStreamDestroy(c.side_stream);
    if (c.ev_inputs) cudaEventDestroy(c.ev_inputs);
    if (c.d_stats) cudaFree(c.d_stats);
    if (c.h_stats) cudaFreeHost(c.h_stats);
    if (c.ev_offsets) cudaEventDestroy(c.ev_offsets);
    for (auto e : c.ev_chunks) cudaEventDestroy(e);
    c.order.release(), c.a_data.release(), c.a_offs.release(), c.b_data.release(), c.b_offs.release(), c.out.release();
    c.sym_a.release(), c.sym_a_offs.release(), c.sym_b.release(), c.sym_b_offs.release(), c.scan_tmp.release(), c.flags.release(), c.hscratch.release();
}

extern "C" swb_status_t swb_scope_create(int first_device, int num_devices, swb_scope_t **out, const char **err) {
    if (!out) return fail(err, SWB_INVALID_ARG, "out is NULL");
    *out = nullptr;
    const int available = swb_device_count();
    if (available == 0) return fail(err, SWB_NO_DEVICE, "no CUDA device is visible (this library has no CPU fallback)");
    if (num_devices < 1 || first_device < 0 || first_device + num_devices > available)
        return fail(err, SWB_NO_DEVICE, "devices [%d, %d) requested but %d visible", first_device,
                    first_device + num_devices, available);
    swb_scope *scope = new swb_scope();
    scope->devs.resize(num_devices);
    for (int i = 0; i < num_devices; ++i) {
        DeviceCtx &c = scope->devs[i];
        c.device = first_device + i;
        cudaError_t e = cudaSetDevice(c.device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, c.device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c.compute_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) {
            // highest priority: the (tiny) schedule kernel must get SM slots while a compute kernel launched just
            // before it on another stream still has blocks waiting, otherwise it only runs in that kernel's tail
            int least = 0, greatest = 0;
            cudaDeviceGetStreamPriorityRange(&least, &greatest);
            e = cudaStreamCreateWithPriority(&c.side_stream, cudaStreamNonBlocking, greatest);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c.ev_inputs, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&c.d_stats, sizeof(ScheduleStats));
        if (e == cudaSuccess) e = cudaMemset(c.d_stats, 0, sizeof(ScheduleStats));
        if (e == cudaSuccess) e = cudaMemset(&c.d_stats->min_m, 0xFF, 2 * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaHostAlloc(&c.h_stats, sizeof(ScheduleStats), cudaHostAllocMapped);
        if (e == cudaSuccess) {
            memset(c.h_stats, 0, sizeof(ScheduleStats));
            e = cudaHostGetDevicePointer((void **)&c.h_stats_dev, c.h_stats, 0);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c.ev_offsets, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            for (auto &d : scope->devs) destroy_ctx(d);
            delete scope;
            return fail(err, SWB_CUDA_ERROR, "scope setup on device %d failed: %s", first_device + i, cudaGetErrorString(e));
        }
    }
    *out = scope;
    return SWB_OK;
}

extern "C" void swb_scope_destroy(swb_scope_t *scope) {
    if (!scope) return;
    for (auto &d : scope->devs) destroy_ctx(d);
    delete scope;
}

extern "C" int swb_scope_num_devices(const swb_scope_t *scope) { return scope ? (int)scope->devs.size() : 0; }

extern "C" swb_status_t swb_host_alloc(size_t bytes, void **out, const char **err) {
    if (!out) return fail(err, SWB_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (swb_device_count() == 0) return fail(err, SWB_NO_DEVICE, "no CUDA device is visible");
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(err, SWB_BAD_ALLOC, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return SWB_OK;
}

extern "C" void swb_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

// ------------------------------------------------------------------------------------------ engines

static swb_status_t engine_finish(swb_engine *e, swb_engine_t **out, const char **err) {
    e->d_tables.assign(e->scope->devs.size(), nullptr);
    for (size_t i = 0; i < e->scope->devs.size(); ++i) {
        cudaError_t ce = cudaSetDevice(e->scope->devs[i].device);
        if (ce == cudaSuccess) ce = cudaMalloc(&e->d_tables[i], sizeof(AlignTables));
        if (ce == cudaSuccess) ce = cudaMemcpy(e->d_tables[i], &e->tables, sizeof(AlignTables), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) {
            swb_engine_destroy(e);
            return fail(err, SWB_CUDA_ERROR, "engine table upload failed: %s", cudaGetErrorString(ce));
        }
    }
    *out = e;
    return SWB_OK;
}

static void identity_tables(AlignTables &t, int match, int mismatch) {
    for (int i = 0; i < 256; ++i) t.byte_to_class[i] = (uint8_t)i; // byte engines compare raw bytes (unary path only)
    for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) t.costs[i * 32 + j] = t.costs_t[i * 32 + j] = (int8_t)(i == j ? match : mismatch);
}

static swb_status_t lev_create(swb_scope_t *scope, int match, int mismatch, int open, int extend, EngineKind kind,
                               swb_engine_t **out, const char **err) {
    if (!out) return fail(err, SWB_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!scope) return fail(err, SWB_INVALID_ARG, "scope is NULL");
    if (match != 0 || mismatch != 1 || open != 1 || extend != 1)
        return fail(err, SWB_INVALID_ARG, "only the unit cost model (match=0, mismatch=1, open=1, extend=1) is supported, got (%d,%d,%d,%d)",
                    match, mismatch, open, extend);
    swb_engine *e = new swb_engine();
    e->scope = scope, e->kind = kind;
    e->match = 0, e->mismatch = -1, e->open = -1, e->extend = -1, e->unary = true; // maximising form
    identity_tables(e->tables, 0, -1);
    return engine_finish(e, out, err);
}

extern "C" swb_status_t swb_levenshtein_create(swb_scope_t *scope, int match, int mismatch, int open, int extend,
                                               swb_engine_t **out, const char **err) {
    return lev_create(scope, match, mismatch, open, extend, ENGINE_LEVENSHTEIN, out, err);
}

extern "C" swb_status_t swb_levenshtein_utf8_create(swb_scope_t *scope, int match, int mismatch, int open, int extend,
                                                    swb_engine_t **out, const char **err) {
    return lev_create(scope, match, mismatch, open, extend, ENGINE_LEVENSHTEIN_UTF8, out, err);
}

static swb_status_t align_create(swb_scope_t *scope, const uint8_t *byte_to_class, const int8_t *class_costs, int open,
                                 int extend, EngineKind kind, swb_engine_t **out, const char **err) {
    if (!out) return fail(err, SWB_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!scope || !byte_to_class || !class_costs) return fail(err, SWB_INVALID_ARG, "NULL argument");
    if (open > 0 || extend > 0) return fail(err, SWB_INVALID_ARG, "gap costs must be <= 0 (scores are maximised), got open=%d extend=%d", open, extend);
    swb_engine *e = new swb_engine();
    e->scope = scope, e->kind = kind, e->open = open, e->extend = extend;
    for (int i = 0; i < 256; ++i) {
        if (byte_to_class[i] >= 32) {
            delete e;
            return fail(err, SWB_INVALID_ARG, "byte_to_class[%d] = %d is not < 32", i, byte_to_class[i]);
        }
        e->tables.byte_to_class[i] = byte_to_class[i];
    }
    bool unary = true;
    for (int i = 0; i < 32; ++i)
        for (int j = 0; j < 32; ++j) {
            e->tables.costs[i * 32 + j] = class_costs[i * 32 + j];
            e->tables.costs_t[j * 32 + i] = class_costs[i * 32 + j];
            if (i == j && class_costs[i * 32 + j] != class_costs[0]) unary = false;
            if (i != j && class_costs[i * 32 + j] != class_costs[1]) unary = false;
        }
    e->unary = unary, e->match = class_costs[0], e->mismatch = class_costs[1];
    return engine_finish(e, out, err);
}

extern "C" swb_status_t swb_needleman_wunsch_create(swb_scope_t *scope, const uint8_t *byte_to_class,
                                                    const int8_t *class_costs, int open, int extend, swb_engine_t **out,
                                                    const char **err) {
    return align_create(scope, byte_to_class, class_costs, open, extend, ENGINE_NW, out, err);
}

extern "C" swb_status_t swb_smith_waterman_create(swb_scope_t *scope, const uint8_t *byte_to_class,
                                                  const int8_t *class_costs, int open, int extend, swb_engine_t **out,
                                                  const char **err) {
    return align_create(scope, byte_to_class, class_costs, open, extend, ENGINE_SW, out, err);
}

extern "C" void swb_engine_destroy(swb_engine_t *engine) {
    if (!engine) return;
    for (size_t i = 0; i < engine->d_tables.size(); ++i)
        if (engine->d_tables[i]) {
            cudaSetDevice(engine->scope->devs[i].device);
            cudaFree(engine->d_tables[i]);
        }
    for (auto &pr : engine->prof_events) cudaEventDestroy(pr.first), cudaEventDestroy(pr.second);
    delete engine;
}

extern "C" swb_status_t swb_engine_set_algorithm(swb_engine_t *engine, swb_algorithm_t algorithm, const char **err) {
    if (!engine) return fail(err, SWB_INVALID_ARG, "engine is NULL");
    if (algorithm != SWB_ALGO_AUTO && algorithm != SWB_ALGO_WAVEFRONT && algorithm != SWB_ALGO_BITPARALLEL)
        return fail(err, SWB_INVALID_ARG, "unknown algorithm %d", (int)algorithm);
    if (algorithm == SWB_ALGO_BITPARALLEL && engine->kind != ENGINE_LEVENSHTEIN)
        return fail(err, SWB_INVALID_ARG, "the bit-parallel kernel only exists for unit-cost byte Levenshtein");
    engine->algorithm = algorithm;
    return SWB_OK;
}

extern "C" swb_status_t swb_engine_set_gap_convention(swb_engine_t *engine, swb_gap_convention_t convention,
                                                      const char **err) {
    if (!engine) return fail(err, SWB_INVALID_ARG, "engine is NULL");
    if (convention != SWB_GAP_STRINGZILLA && convention != SWB_GAP_BIO) return fail(err, SWB_INVALID_ARG, "unknown gap convention %d", (int)convention);
    engine->convention = convention;
    return SWB_OK;
}

extern "C" const char *swb_engine_last_kernel(const swb_engine_t *engine) { return engine ? engine->last_kernel : "none"; }

extern "C" swb_status_t swb_engine_set_profiling(swb_engine_t *engine, int enabled, const char **err) {
    if (!engine) return fail(err, SWB_INVALID_ARG, "engine is NULL");
    engine->profiling = enabled != 0;
    engine->prof_used = 0;
    return SWB_OK;
}

extern "C" swb_status_t swb_engine_kernel_time(swb_engine_t *engine, double *total_ms, uint64_t *launches, const char **err) {
    if (!engine || !total_ms || !launches) return fail(err, SWB_INVALID_ARG, "NULL argument");
    double total = 0;
    for (size_t i = 0; i < engine->prof_used; ++i) {
        float ms = 0;
        SWB_CUDA(cudaSetDevice(engine->prof_device[i]));
        SWB_CUDA(cudaEventSynchronize(engine->prof_events[i].second));
        SWB_CUDA(cudaEventElapsedTime(&ms, engine->prof_events[i].first, engine->prof_events[i].second));
        total += ms;
    }
    *total_ms = total, *launches = engine->prof_used;
    engine->prof_used = 0;
    return SWB_OK;
}

// ------------------------------------------------------------------------------------------ kernel dispatch

static inline uint32_t round16(uint32_t x) { return (x + 15u) & ~15u; }

template <int K>
static swb_status_t launch_bitpar(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint64_t begin, uint64_t count,
                                  cudaStream_t stream, const char **err) {
    constexpr int T = BitparConfig<K>::THREADS;
    const size_t smem = (size_t)32 * T * K * sizeof(uint32_t);
    SWB_CUDA(allow_smem((const void *)lev_bitpar_kernel<K>, ctx.device, smem));
    const uint64_t blocks = (count + T - 1) / T;
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        lev_bitpar_kernel<K><<<(unsigned)blocks, T, smem, stream>>>(ps, begin, count);
    }
    g_launches++;
    static const char *names[] = {"", "lev_bitpar<K=1>", "lev_bitpar<K=2>", "", "lev_bitpar<K=4>", "", "", "", "lev_bitpar<K=8>"};
    eng->last_kernel = names[K];
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

// Lane-group bit-vector kernel: text streamed in shared-memory chunks, patterns beyond 32*G*K rows in row strips
// through a global scratch row per pair — no length limit.
template <int G, int K>
static swb_status_t launch_bitpar_coop(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint64_t begin, uint64_t count,
                                       uint32_t max_m, uint32_t max_n, const char *name, cudaStream_t stream, const char **err) {
    constexpr int PPB = BPC_THREADS / G;
    constexpr uint32_t CAP = 32u * G * K;
    const uint32_t chunk_cols = std::min<uint32_t>(1024u, std::max<uint32_t>(16u, round16(max_n)));
    const size_t smem = (size_t)32 * BPC_THREADS * K * 4 + (size_t)PPB * (CAP + 32u + chunk_cols + 32u);
    uint8_t *scratch = nullptr;
    uint64_t scratch_stride = 0;
    if (max_m > CAP) { // row strips: one byte of horizontal-delta bits per text column and pair
        scratch_stride = round16(max_n) + 16u;
        SWB_CUDA(ctx.hscratch.reserve(count * scratch_stride));
        scratch = (uint8_t *)ctx.hscratch.ptr;
    }
    const uint64_t blocks = (count + PPB - 1) / PPB;
    if (scratch) {
        SWB_CUDA(allow_smem((const void *)lev_bitpar_coop_kernel<G, K, true>, ctx.device, smem));
        ProfileBracket bracket(eng, ctx.device, stream);
        lev_bitpar_coop_kernel<G, K, true><<<(unsigned)blocks, BPC_THREADS, smem, stream>>>(ps, begin, count, chunk_cols, scratch, scratch_stride);
    } else {
        SWB_CUDA(allow_smem((const void *)lev_bitpar_coop_kernel<G, K, false>, ctx.device, smem));
        ProfileBracket bracket(eng, ctx.device, stream);
        lev_bitpar_coop_kernel<G, K, false><<<(unsigned)blocks, BPC_THREADS, smem, stream>>>(ps, begin, count, chunk_cols, scratch, scratch_stride);
    }
    g_launches++;
    eng->last_kernel = name;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

template <int G, int R, bool LOCAL, bool AFFINE, bool UNARY, typename BoundT, typename SymT = uint8_t>
static swb_status_t launch_wavefront_t(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin,
                                       uint64_t count, const AlignCosts &cs, uint32_t max_m, uint32_t max_n,
                                       cudaStream_t stream, const char **err) {
    constexpr int PPB = WF_THREADS / G;
    constexpr bool SHIFTED = !LOCAL && !AFFINE && UNARY;
    const bool single_strip = max_m <= (uint32_t)(G * R);
    const uint32_t slot_rows = round16(max_m * (uint32_t)sizeof(SymT) + 16u) + 16u, slot_cols = round16(max_n * (uint32_t)sizeof(SymT) + 16u) + 16u;
    const uint32_t bound_cap = (SHIFTED && single_strip) ? 0u : max_n + 1u; // shifted single-strip cells need no boundary row
    const uint32_t bound_bytes = round16(bound_cap * (uint32_t)sizeof(BoundT) * (AFFINE ? 2u : 1u));
    const size_t smem = 2304 + (size_t)PPB * (slot_rows + slot_cols + bound_bytes);
    if (smem > (size_t)ctx.max_smem_optin)
        return fail(err, SWB_TOO_LONG, "pair of %u x %u bytes needs %zu B of shared memory per block (limit %d)", max_m, max_n,
                    smem, ctx.max_smem_optin);
    auto kernel = align_wavefront_kernel<G, R, LOCAL, AFFINE, UNARY, BoundT, SymT>;
    SWB_CUDA(allow_smem((const void *)kernel, ctx.device, smem));
    const uint64_t blocks = (count + PPB - 1) / PPB;
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        kernel<<<(unsigned)blocks, WF_THREADS, smem, stream>>>(ps, begin, count, cs, eng->d_tables[dev_index], slot_rows, slot_cols, bound_cap);
    }
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

// Two pairs per lane group in packed s16x2 registers (align_wavefront_packed.cuh). SWB_TOO_LONG = does not fit smem.
template <int G, int R>
static swb_status_t launch_wavefront_packed(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin, uint64_t count,
                                            const AlignCosts &cs, uint32_t max_m, uint32_t max_n, cudaStream_t stream, const char **err) {
    constexpr int GPB = WF_THREADS / G;
    const bool single_strip = max_m <= (uint32_t)(G * R);
    const uint32_t slot_rows = round16(max_m + 16u) + 16u, slot_cols = round16(max_n + 16u) + 16u;
    const uint32_t bound_cap = single_strip ? 0u : max_n + 1u;
    const size_t smem = 2304 + (size_t)GPB * (2u * slot_rows + 2u * slot_cols + round16(bound_cap * 4u));
    if (smem > (size_t)ctx.max_smem_optin) return fail(err, SWB_TOO_LONG, "packed wavefront: %zu B of shared memory per block", smem);
    SWB_CUDA(allow_smem((const void *)align_wavefront_packed_kernel<G, R>, ctx.device, smem));
    const uint64_t blocks = ((count + 1) / 2 + GPB - 1) / GPB;
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        align_wavefront_packed_kernel<G, R><<<(unsigned)blocks, WF_THREADS, smem, stream>>>(ps, begin, count, cs, eng->d_tables[dev_index], slot_rows, slot_cols, bound_cap);
    }
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

template <int G, int R, bool LOCAL, bool AFFINE>
static swb_status_t launch_wavefront_packed_general(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin,
                                                    uint64_t count, const AlignCosts &cs, uint32_t max_m, uint32_t max_n, cudaStream_t stream,
                                                    const char **err) {
    constexpr int GPB = WF_THREADS / G;
    const uint32_t slot_rows = round16(max_m + 16u) + 16u, slot_cols = round16(max_n + 16u) + 16u;
    const uint32_t bound_cap = max_n + 1u;
    const size_t smem = 2304 + (size_t)GPB * (2u * slot_rows + 2u * slot_cols + round16(bound_cap * 4u) * (AFFINE ? 2u : 1u));
    if (smem > (size_t)ctx.max_smem_optin) return fail(err, SWB_TOO_LONG, "packed wavefront: %zu B of shared memory per block", smem);
    SWB_CUDA(allow_smem((const void *)align_wavefront_packed_general_kernel<G, R, LOCAL, AFFINE>, ctx.device, smem));
    const uint64_t blocks = ((count + 1) / 2 + GPB - 1) / GPB;
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        align_wavefront_packed_general_kernel<G, R, LOCAL, AFFINE><<<(unsigned)blocks, WF_THREADS, smem, stream>>>(
            ps, begin, count, cs, eng->d_tables[dev_index], slot_rows, slot_cols, bound_cap);
    }
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

template <int G, int R>
static swb_status_t launch_wavefront_gr(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin,
                                        uint64_t count, const AlignCosts &cs, bool local, bool affine, bool unary, bool narrow,
                                        uint32_t max_m, uint32_t max_n, cudaStream_t stream, const char **err) {
#define SWB_WF(L, A, U, T) return launch_wavefront_t<G, R, L, A, U, T>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err)
    if (ps.len_shift == 2) { // decoded UTF-8: 32-bit symbols, unit costs only
        if (narrow) return launch_wavefront_t<G, R, false, false, true, int16_t, uint32_t>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);
        return launch_wavefront_t<G, R, false, false, true, int32_t, uint32_t>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);
    }
    if (!unary) { // general 32x32 substitution table: int32 boundary rows only
        if (!local && !affine) SWB_WF(false, false, false, int32_t);
        if (!local && affine) SWB_WF(false, true, false, int32_t);
        if (local && !affine) SWB_WF(true, false, false, int32_t);
        SWB_WF(true, true, false, int32_t);
    }
    if (!local && !affine && narrow) SWB_WF(false, false, true, int16_t);
    if (!local && !affine && !narrow) SWB_WF(false, false, true, int32_t);
    if (!local && affine && narrow) SWB_WF(false, true, true, int16_t);
    if (!local && affine && !narrow) SWB_WF(false, true, true, int32_t);
    if (local && !affine && narrow) SWB_WF(true, false, true, int16_t);
    if (local && !affine && !narrow) SWB_WF(true, false, true, int32_t);
    if (local && affine && narrow) SWB_WF(true, true, true, int16_t);
    SWB_WF(true, true, true, int32_t);
#undef SWB_WF
}

static swb_status_t launch_wavefront(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin,
                                     uint64_t count, uint32_t max_m, uint32_t max_n, cudaStream_t stream, const char **err, bool same_shape = false) {
    AlignCosts cs;
    const bool is_lev = eng->kind == ENGINE_LEVENSHTEIN || eng->kind == ENGINE_LEVENSHTEIN_UTF8;
    cs.match = eng->match, cs.mismatch = eng->mismatch;
    cs.ident = is_lev ? 1 : 0;
    cs.negate = is_lev ? 1 : 0;
    cs.first = eng->convention == SWB_GAP_BIO && !is_lev ? eng->open + eng->extend : eng->open;
    cs.next = eng->extend;
    const bool local = eng->kind == ENGINE_SW;
    const bool affine = cs.first != cs.next;
    // int16 boundary rows when every reachable |score| (also in the shifted H - gap*(i+j) form) stays inside int16
    int maxabs = std::max(std::abs(cs.first), std::abs(cs.next));
    if (eng->unary) maxabs = std::max(maxabs, std::max(std::abs(cs.match), std::abs(cs.mismatch)));
    else maxabs = std::max(maxabs, 128);
    const bool narrow = (uint64_t)(max_m + max_n + 2u) * 3u * (uint64_t)maxabs < 30000u;
    // Packed s16x2 variant (two pairs per lane group, half the ALU instructions per cell): global, linear gaps,
    // unary costs, byte symbols, every pair of the same shape, shifted costs and scores inside 15 bits.
    static const bool packed_off = getenv("SWB_NO_PACKED") != nullptr;
    const int cm_s = cs.match - 2 * cs.next, cx_s = cs.mismatch - 2 * cs.next;
    if (!packed_off && same_shape && !local && !affine && eng->unary && ps.len_shift == 0 && count >= 2 && cx_s >= 0 && cm_s >= cx_s &&
        (uint64_t)(std::min(max_m, max_n) + 1u) * (uint64_t)cm_s < 32000u) {
        swb_status_t pst = SWB_TOO_LONG;
#define SWB_PK(G, NAME)                                                                                                   \
        if (pst == SWB_TOO_LONG && (max_m <= (uint32_t)(G) * 16u || (G) == 32)) {                                           \
            eng->last_kernel = NAME;                                                                                      \
            pst = launch_wavefront_packed<G, 16>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);   \
        }
        SWB_PK(1, "align_wavefront_packed<G=1,R=16>")
        SWB_PK(2, "align_wavefront_packed<G=2,R=16>")
        SWB_PK(4, "align_wavefront_packed<G=4,R=16>")
        SWB_PK(8, "align_wavefront_packed<G=8,R=16>")
        SWB_PK(16, "align_wavefront_packed<G=16,R=16>")
        SWB_PK(32, "align_wavefront_packed<G=32,R=16>")
#undef SWB_PK
        if (pst != SWB_TOO_LONG) return pst;
    }
    // Packed s16x2 variant for affine gaps and / or local alignment (align_wavefront_packed_general.cuh): unary costs with
    // mismatch <= 0 <= match, gaps <= 0, every pair of the same shape, every reachable score inside int16 with room
    // for the gap tracks' floor: H >= 3*first + (m+n)*min(next, mismatch), H <= min(m,n)*match.
    if (!packed_off && same_shape && (local || affine) && eng->unary && ps.len_shift == 0 && count >= 2 && cs.mismatch <= 0 && cs.match >= 0 &&
        cs.first <= 0 && cs.next <= 0 &&
        3 * (int64_t)cs.first + (int64_t)(max_m + max_n) * std::min(cs.next, cs.mismatch) > -28000 &&
        (int64_t)std::min(max_m, max_n) * cs.match < 32000) {
        swb_status_t pst = SWB_TOO_LONG;
#define SWB_PG(G, NAME)                                                                                                                        \
        if (pst == SWB_TOO_LONG && (max_m <= (uint32_t)(G) * 16u || (G) == 32)) {                                                                \
            eng->last_kernel = NAME;                                                                                                           \
            if (local && affine) pst = launch_wavefront_packed_general<G, 16, true, true>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);        \
            else if (local) pst = launch_wavefront_packed_general<G, 16, true, false>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);            \
            else pst = launch_wavefront_packed_general<G, 16, false, true>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);                        \
        }
        SWB_PG(1, "align_wavefront_packed_general<G=1,R=16>")
        SWB_PG(2, "align_wavefront_packed_general<G=2,R=16>")
        SWB_PG(4, "align_wavefront_packed_general<G=4,R=16>")
        SWB_PG(8, "align_wavefront_packed_general<G=8,R=16>")
        SWB_PG(16, "align_wavefront_packed_general<G=16,R=16>")
        SWB_PG(32, "align_wavefront_packed_general<G=32,R=16>")
#undef SWB_PG
        if (pst != SWB_TOO_LONG) return pst;
    }
    // Narrowest group that covers the pattern in one strip; if the class's longest text makes the block's staged
    // strings overflow shared memory (lopsided pairs), widen the group: fewer pairs per block, same kernel.
    swb_status_t st = SWB_TOO_LONG;
#define SWB_GR(G, R, NAME)                                                                                                              \
    if (st == SWB_TOO_LONG && (max_m <= (uint32_t)(G) * (R) || (G) == 32)) {                                                              \
        eng->last_kernel = NAME;                                                                                                        \
        st = launch_wavefront_gr<G, R>(eng, ctx, dev_index, ps, begin, count, cs, local, affine, eng->unary, narrow, max_m, max_n, stream, err); \
    }
    SWB_GR(1, 16, "align_wavefront<G=1,R=16>")
    SWB_GR(2, 16, "align_wavefront<G=2,R=16>")
    SWB_GR(4, 16, "align_wavefront<G=4,R=16>")
    SWB_GR(8, 16, "align_wavefront<G=8,R=16>")
    SWB_GR(16, 16, "align_wavefront<G=16,R=16>")
    SWB_GR(32, 16, "align_wavefront<G=32,R=16>")
#undef SWB_GR
    return st;
}

// Cross product with table reuse: one thread per query, candidates swept (lev_bitpar_cross.cuh).
template <int K>
static swb_status_t launch_bitpar_cross(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint32_t n_queries, uint32_t cands_per_block,
                                        const char *name, cudaStream_t stream, const char **err) {
    constexpr int T = CrossConfig<K>::THREADS;
    const uint32_t chunk = 1024;
    const size_t smem = (size_t)32 * T * K * 4 + (((size_t)T * (CROSS_TILE + 1) * 4 + 15) & ~(size_t)15) + chunk + 32;
    SWB_CUDA(allow_smem((const void *)lev_bitpar_cross_kernel<K>, ctx.device, smem));
    const dim3 grid((n_queries + T - 1) / T, (unsigned)((ps.n_c + cands_per_block - 1) / cands_per_block));
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        lev_bitpar_cross_kernel<K><<<grid, T, smem, stream>>>(ps, n_queries, cands_per_block, chunk);
    }
    g_launches++;
    eng->last_kernel = name;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

// One length class of one batch on one device.
static swb_status_t launch_class(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin,
                                 uint64_t count, uint32_t max_m, uint32_t max_n, cudaStream_t stream, const char **err, bool same_shape = false) {
    if (count == 0) return SWB_OK;
    const bool lev_bytes = (eng->kind == ENGINE_LEVENSHTEIN || eng->kind == ENGINE_LEVENSHTEIN_UTF8) && ps.len_shift == 0;
    if (lev_bytes && eng->algorithm != SWB_ALGO_WAVEFRONT) {
        // SWB_EXP (bit mask, experiments only): 1: <=256 B via lane groups <4,2>; 2: <=1024 B via <8,4>; 4: <=512 B via <4,4>
        static const int exp_mask = getenv("SWB_EXP") ? atoi(getenv("SWB_EXP")) : 0;
        if ((exp_mask & 1) && max_m > 128 && max_m <= 256) return launch_bitpar_coop<4, 2>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=4,K=2>", stream, err);
        if ((exp_mask & 4) && max_m > 256 && max_m <= 512) return launch_bitpar_coop<4, 4>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=4,K=4>", stream, err);
        if ((exp_mask & 2) && max_m > 512 && max_m <= 1024) return launch_bitpar_coop<8, 4>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=8,K=4>", stream, err);
        // thread-per-pair bit-vector kernels up to 256 pattern bytes, lane-group kernels up to 4096
        if (max_m <= 32) return launch_bitpar<1>(eng, ctx, ps, begin, count, stream, err);
        if (max_m <= 64) return launch_bitpar<2>(eng, ctx, ps, begin, count, stream, err);
        if (max_m <= 128) return launch_bitpar<4>(eng, ctx, ps, begin, count, stream, err);
        if (max_m <= 256) return launch_bitpar<8>(eng, ctx, ps, begin, count, stream, err);
        if (max_m <= 512) return launch_bitpar_coop<8, 2>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=8,K=2>", stream, err);
        if (max_m <= 1024) return launch_bitpar_coop<16, 2>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=16,K=2>", stream, err);
        if (max_m <= 2048) return launch_bitpar_coop<32, 2>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=32,K=2>", stream, err);
        return launch_bitpar_coop<32, 4>(eng, ctx, ps, begin, count, max_m, max_n, max_m <= 4096 ? "lev_bitpar_coop<G=32,K=4>" : "lev_bitpar_coop<G=32,K=4,strips>", stream, err);
    }
    return launch_wavefront(eng, ctx, dev_index, ps, begin, count, max_m, max_n, stream, err, same_shape);
}

// Schedule counters for a device-resident PairSpace: enqueue the count kernel on `stream`; it publishes into
// mapped host memory and re-arms the device counters itself (no memset / D2H copy / event on the critical path).
static swb_status_t enqueue_schedule_count(DeviceCtx &ctx, const PairSpace &ps, cudaStream_t stream, const char **err) {
    const unsigned blocks = (unsigned)std::min<uint64_t>((ps.n_pairs + 1023) / 1024, (uint64_t)ctx.sm_count * 4);
    ctx.stats_seq++;
    if (ctx.stats_seq == 0 || ctx.stats_seq == 0xFFFFFFFFu) ctx.stats_seq = 1;
    ctx.stats_stream = stream;
    schedule_count_kernel<<<blocks, 256, 0, stream>>>(ps, ctx.d_stats, ctx.h_stats_dev, ctx.stats_seq);
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

// Wait until the counters of the last enqueue_schedule_count are visible on the host. Spins on the sequence
// flag (the write lands ~1-2 us after the kernel's last block); checks the stream now and then so that a failed
// launch surfaces as an error instead of a hang.
static swb_status_t wait_schedule_count(DeviceCtx &ctx, const char **err) {
    volatile uint32_t *seq = &ctx.h_stats->seq;
    for (uint64_t spins = 0; *seq != ctx.stats_seq; ++spins) {
        if ((spins & 0x3FFF) == 0x3FFF) {
            cudaError_t q = cudaStreamQuery(ctx.stats_stream);
            if (q != cudaSuccess && q != cudaErrorNotReady)
                return fail(err, SWB_CUDA_ERROR, "schedule kernel failed: %s", cudaGetErrorString(q));
            if (q == cudaSuccess && *seq != ctx.stats_seq) {
                // the stream drained but the flag never arrived: treat as a device fault rather than spin forever
                if (spins > (1ull << 26)) return fail(err, SWB_CUDA_ERROR, "schedule counters were never published");
            }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return SWB_OK;
}

struct Plan {
    bool identity = true;        // single key: keep input order
    bool same_shape = false;     // every pair has the same (shorter, longer) lengths
    int n_classes = 0;
    int cls[SCHED_MCLASSES];     // pattern-length class index of each entry
    uint64_t begin[SCHED_MCLASSES], count[SCHED_MCLASSES];
    uint32_t max_m[SCHED_MCLASSES], max_n[SCHED_MCLASSES];
};

// Turn the counters (host copy, already synchronised) into a launch plan; fills cursor[] for the scatter.
static Plan make_plan(ScheduleStats &st) {
    Plan plan;
    int nonempty_keys = 0;
    uint64_t running = 0;
    for (int c = 0; c < SCHED_MCLASSES; ++c) {
        uint64_t class_count = 0;
        const uint64_t class_begin = running;
        for (int b = 0; b < SCHED_NBINS; ++b) {
            const int key = c * SCHED_NBINS + b;
            st.cursor[key] = (uint32_t)running;
            running += st.counts[key];
            class_count += st.counts[key];
            nonempty_keys += st.counts[key] != 0;
        }
        if (class_count) {
            const int i = plan.n_classes++;
            plan.cls[i] = c;
            plan.begin[i] = class_begin, plan.count[i] = class_count;
            plan.max_m[i] = st.max_m[c], plan.max_n[i] = st.max_n[c];
        }
    }
    // one pattern class and at most two text-length bins: the indirection would cost more than the divergence
    plan.identity = plan.n_classes <= 1 && nonempty_keys <= 2;
    plan.same_shape = plan.n_classes == 1 && st.min_m == plan.max_m[0] && st.min_n == plan.max_n[0];
    return plan;
}

// Run a fully device-resident batch: schedule, (scatter), per-class kernels. Synchronises `stream` once for
// the counters readback.
static swb_status_t run_device_batch(swb_engine *eng, int dev_index, PairSpace ps, cudaStream_t stream, const char **err) {
    DeviceCtx &ctx = eng->scope->devs[dev_index];
    if (ps.n_pairs == 0) return SWB_OK;
    if (ps.n_pairs >= 0xFFFFFFFFull) return fail(err, SWB_INVALID_ARG, "more than 2^32-1 pairs in one call");
    ps.order = nullptr;
    // Cross products large enough to amortise a per-query table build: statistics of the QUERY tape alone decide
    // the kernel (queries are the patterns), then one thread per query sweeps its share of the candidates.
    if (ps.n_c >= 32 && ps.len_shift == 0 && (eng->kind == ENGINE_LEVENSHTEIN || eng->kind == ENGINE_LEVENSHTEIN_UTF8) &&
        eng->algorithm != SWB_ALGO_WAVEFRONT) {
        const uint64_t n_q = ps.n_pairs / ps.n_c;
        PairSpace qs = ps; // degenerate pair space (q_i, q_i): its counters are the query-length statistics
        qs.b_data = ps.a_data, qs.b_offs = ps.a_offs, qs.n_pairs = n_q, qs.n_c = 0;
        swb_status_t qst = enqueue_schedule_count(ctx, qs, stream, err);
        if (qst != SWB_OK) return qst;
        if ((qst = wait_schedule_count(ctx, err)) != SWB_OK) return qst;
        uint32_t max_q = 0;
        for (int c = 0; c < SCHED_MCLASSES; ++c) max_q = std::max(max_q, ctx.h_stats->max_m[c]);
        // candidates per block: enough blocks to fill the machine twice, at least 16 candidates per table build
        const uint64_t query_blocks = (n_q + 127) / 128;
        uint64_t cpb = std::max<uint64_t>(16, ps.n_c * query_blocks / std::max<uint64_t>(1, (uint64_t)ctx.sm_count * 8));
        cpb = std::min<uint64_t>(cpb, ps.n_c);
        const bool worth = max_q <= 256 && query_blocks * ((ps.n_c + cpb - 1) / cpb) >= (uint64_t)ctx.sm_count;
        if (worth) {
            if (max_q <= 32) return launch_bitpar_cross<1>(eng, ctx, ps, (uint32_t)n_q, (uint32_t)cpb, "lev_bitpar_cross<K=1>", stream, err);
            if (max_q <= 64) return launch_bitpar_cross<2>(eng, ctx, ps, (uint32_t)n_q, (uint32_t)cpb, "lev_bitpar_cross<K=2>", stream, err);
            if (max_q <= 128) return launch_bitpar_cross<4>(eng, ctx, ps, (uint32_t)n_q, (uint32_t)cpb, "lev_bitpar_cross<K=4>", stream, err);
            return launch_bitpar_cross<8>(eng, ctx, ps, (uint32_t)n_q, (uint32_t)cpb, "lev_bitpar_cross<K=8>", stream, err);
        }
    }
    // Speculate that this batch has the shape of the previous one. A thread-per-pair bit-vector kernel is correct
    // for every pair that fits its K words and skips the others, and its launch geometry depends on nothing but
    // the pair count, so it goes onto the caller's stream IMMEDIATELY while the schedule kernel runs beside it on
    // a side stream: the host round trip and the schedule kernel leave the critical path. If the counters then
    // show a different shape, the batch is simply processed again the regular way (results overwritten in order).
    const bool lev_bytes = (eng->kind == ENGINE_LEVENSHTEIN || eng->kind == ENGINE_LEVENSHTEIN_UTF8) && ps.len_shift == 0;
    const int spec = (lev_bytes && eng->algorithm != SWB_ALGO_WAVEFRONT && eng->spec_class >= 0 && eng->spec_class <= 3) ? eng->spec_class : -1;
    swb_status_t st;
    if (spec >= 0) {
        SWB_CUDA(cudaEventRecord(ctx.ev_inputs, stream));
        SWB_CUDA(cudaStreamWaitEvent(ctx.side_stream, ctx.ev_inputs, 0));
        if (spec == 0) st = launch_bitpar<1>(eng, ctx, ps, 0, ps.n_pairs, stream, err);
        else if (spec == 1) st = launch_bitpar<2>(eng, ctx, ps, 0, ps.n_pairs, stream, err);
        else if (spec == 2) st = launch_bitpar<4>(eng, ctx, ps, 0, ps.n_pairs, stream, err);
        else st = launch_bitpar<8>(eng, ctx, ps, 0, ps.n_pairs, stream, err);
        if (st != SWB_OK) return st;
        if ((st = enqueue_schedule_count(ctx, ps, ctx.side_stream, err)) != SWB_OK) return st;
    } else if ((st = enqueue_schedule_count(ctx, ps, stream, err)) != SWB_OK) {
        return st;
    }
    if ((st = wait_schedule_count(ctx, err)) != SWB_OK) return st;
    Plan plan = make_plan(*ctx.h_stats);
    const bool uniform = plan.identity && plan.n_classes == 1;
    eng->spec_class = (uniform && lev_bytes) ? plan.cls[0] : -1;
    if (spec >= 0 && uniform && plan.cls[0] == spec) return SWB_OK; // the bet held: that kernel is the batch
    if (!plan.identity) {
        SWB_CUDA(ctx.order.reserve(ps.n_pairs * sizeof(uint32_t)));
        SWB_CUDA(cudaMemcpyAsync(ctx.d_stats->cursor, ctx.h_stats->cursor, sizeof(ctx.h_stats->cursor), cudaMemcpyHostToDevice, stream));
        const unsigned blocks = (unsigned)std::min<uint64_t>((ps.n_pairs + 255) / 256, (uint64_t)ctx.sm_count * 8);
        schedule_scatter_kernel<<<blocks, 256, 0, stream>>>(ps, ctx.d_stats, (uint32_t *)ctx.order.ptr);
        g_launches++;
        SWB_CUDA(cudaGetLastError());
        ps.order = (const uint32_t *)ctx.order.ptr;
    }
    for (int i = 0; i < plan.n_classes; ++i) {
        st = launch_class(eng, ctx, dev_index, ps, plan.begin[i], plan.count[i], plan.max_m[i], plan.max_n[i], stream, err, plan.same_shape);
        if (st != SWB_OK) return st;
    }
    return SWB_OK;
}

// ------------------------------------------------------------------------------------------ UTF-8 engine

// Split one device byte tape into 32-bit symbols (count -> scan -> pack). `offs` points at n+1 offsets valid
// with `data` (biased pointers are fine); `total_bytes` bounds the symbol count.
static swb_status_t decode_tape(DeviceCtx &ctx, const uint8_t *data, const uint64_t *offs, uint64_t n, uint64_t total_bytes,
                                DeviceBuffer &sym, DeviceBuffer &sym_offs, uint32_t *d_ascii_flag, cudaStream_t stream, const char **err) {
    SWB_CUDA(sym.reserve(total_bytes * 4 + 64));
    SWB_CUDA(sym_offs.reserve((n + 1) * sizeof(uint64_t)));
    size_t tmp_bytes = 0;
    uint64_t *so = (uint64_t *)sym_offs.ptr;
    SWB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, so, so, (int)(n + 1), stream));
    SWB_CUDA(ctx.scan_tmp.reserve(tmp_bytes + 16));
    const unsigned blocks = (unsigned)((n * 32 + 255) / 256);
    utf8_count_kernel<<<blocks, 256, 0, stream>>>(data, offs, n, so, d_ascii_flag);
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    SWB_CUDA(cub::DeviceScan::InclusiveSum(ctx.scan_tmp.ptr, tmp_bytes, so, so, (int)(n + 1), stream));
    utf8_pack_kernel<<<blocks, 256, 0, stream>>>(data, offs, n, so, (uint32_t *)sym.ptr);
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

static swb_status_t run_device_batch(swb_engine *eng, int dev_index, PairSpace ps, cudaStream_t stream, const char **err);

// Codepoint-level batch: decode both tapes, then either the byte kernels (all-ASCII input: bytes ARE code
// points) or the wavefront kernel over 32-bit symbols. n_a / n_b are the string counts of the two tapes.
static swb_status_t run_device_batch_utf8(swb_engine *eng, int dev_index, PairSpace ps, uint64_t n_a, uint64_t n_b,
                                          uint64_t a_bytes, uint64_t b_bytes, cudaStream_t stream, const char **err) {
    DeviceCtx &ctx = eng->scope->devs[dev_index];
    SWB_CUDA(ctx.flags.reserve(64));
    uint32_t *d_flag = (uint32_t *)ctx.flags.ptr;
    SWB_CUDA(cudaMemsetAsync(d_flag, 0x01, sizeof(uint32_t), stream));
    swb_status_t st;
    if ((st = decode_tape(ctx, ps.a_data, ps.a_offs, n_a, a_bytes, ctx.sym_a, ctx.sym_a_offs, d_flag, stream, err)) != SWB_OK) return st;
    const bool same_tape = ps.a_data == ps.b_data && ps.a_offs == ps.b_offs && n_a == n_b;
    if (!same_tape)
        if ((st = decode_tape(ctx, ps.b_data, ps.b_offs, n_b, b_bytes, ctx.sym_b, ctx.sym_b_offs, d_flag, stream, err)) != SWB_OK) return st;
    uint32_t h_flag = 0;
    SWB_CUDA(cudaMemcpyAsync(&h_flag, d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    SWB_CUDA(cudaStreamSynchronize(stream));
    if (h_flag != 0) return run_device_batch(eng, dev_index, ps, stream, err); // pure ASCII
    PairSpace sym = ps;
    sym.a_data = (const uint8_t *)ctx.sym_a.ptr, sym.a_offs = (const uint64_t *)ctx.sym_a_offs.ptr;
    sym.b_data = same_tape ? sym.a_data : (const uint8_t *)ctx.sym_b.ptr;
    sym.b_offs = same_tape ? sym.a_offs : (const uint64_t *)ctx.sym_b_offs.ptr;
    sym.len_shift = 2;
    return run_device_batch(eng, dev_index, sym, stream, err);
}

// Total bytes spanned by n strings of a DEVICE offsets array (one small synchronous readback).
static swb_status_t device_span(const uint64_t *offs, uint64_t n, cudaStream_t stream, uint64_t *bytes, const char **err) {
    uint64_t ends[2] = {0, 0};
    SWB_CUDA(cudaMemcpyAsync(&ends[0], offs, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    SWB_CUDA(cudaMemcpyAsync(&ends[1], offs + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    SWB_CUDA(cudaStreamSynchronize(stream));
    *bytes = ends[1] - ends[0];
    return SWB_OK;
}

// ------------------------------------------------------------------------------------------ device-resident API

static swb_status_t check_engine(swb_engine_t *engine, int device_index, const char **err) {
    if (!engine) return fail(err, SWB_INVALID_ARG, "engine is NULL");
    if (device_index < 0 || device_index >= (int)engine->scope->devs.size())
        return fail(err, SWB_INVALID_ARG, "device_index %d outside the scope's %zu devices", device_index, engine->scope->devs.size());
    return SWB_OK;
}

extern "C" swb_status_t swb_compute_pairs_device(swb_engine_t *engine, int device_index, const uint8_t *a_data,
                                                 const uint64_t *a_offsets, const uint8_t *b_data, const uint64_t *b_offsets,
                                                 size_t n_pairs, uint64_t *out, void *cuda_stream, const char **err) {
    swb_status_t st = check_engine(engine, device_index, err);
    if (st != SWB_OK) return st;
    if (n_pairs == 0) return SWB_OK;
    if (!a_data || !a_offsets || !b_data || !b_offsets || !out) return fail(err, SWB_INVALID_ARG, "NULL tape or output pointer");
    SWB_CUDA(cudaSetDevice(engine->scope->devs[device_index].device));
    PairSpace ps{a_data, a_offsets, b_data, b_offsets, n_pairs, 0, 0, nullptr, out};
    if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) {
        uint64_t a_bytes = 0, b_bytes = 0;
        if ((st = device_span(a_offsets, n_pairs, (cudaStream_t)cuda_stream, &a_bytes, err)) != SWB_OK) return st;
        if ((st = device_span(b_offsets, n_pairs, (cudaStream_t)cuda_stream, &b_bytes, err)) != SWB_OK) return st;
        return run_device_batch_utf8(engine, device_index, ps, n_pairs, n_pairs, a_bytes, b_bytes, (cudaStream_t)cuda_stream, err);
    }
    return run_device_batch(engine, device_index, ps, (cudaStream_t)cuda_stream, err);
}

extern "C" swb_status_t swb_compute_cross_device(swb_engine_t *engine, int device_index, const uint8_t *q_data,
                                                 const uint64_t *q_offsets, size_t n_queries, const uint8_t *c_data,
                                                 const uint64_t *c_offsets, size_t n_candidates, uint64_t *out,
                                                 size_t row_stride, void *cuda_stream, const char **err) {
    swb_status_t st = check_engine(engine, device_index, err);
    if (st != SWB_OK) return st;
    if (n_queries == 0 || n_candidates == 0) return SWB_OK;
    if (!q_data || !q_offsets || !c_data || !c_offsets || !out) return fail(err, SWB_INVALID_ARG, "NULL tape or output pointer");
    if (row_stride < n_candidates) return fail(err, SWB_INVALID_ARG, "row_stride %zu < n_candidates %zu", row_stride, n_candidates);
    SWB_CUDA(cudaSetDevice(engine->scope->devs[device_index].device));
    PairSpace ps{q_data, q_offsets, c_data, c_offsets, (uint64_t)n_queries * n_candidates, n_candidates, row_stride, nullptr, out};
    if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) {
        uint64_t a_bytes = 0, b_bytes = 0;
        if ((st = device_span(q_offsets, n_queries, (cudaStream_t)cuda_stream, &a_bytes, err)) != SWB_OK) return st;
        if ((st = device_span(c_offsets, n_candidates, (cudaStream_t)cuda_stream, &b_bytes, err)) != SWB_OK) return st;
        return run_device_batch_utf8(engine, device_index, ps, n_queries, n_candidates, a_bytes, b_bytes, (cudaStream_t)cuda_stream, err);
    }
    return run_device_batch(engine, device_index, ps, (cudaStream_t)cuda_stream, err);
}

extern "C" swb_status_t swb_load_phase_device(swb_engine_t *engine, int device_index, const uint8_t *a_data,
                                              const uint64_t *a_offsets, const uint8_t *b_data, const uint64_t *b_offsets,
                                              size_t n_pairs, uint64_t *out, void *cuda_stream, const char **err) {
    swb_status_t st = check_engine(engine, device_index, err);
    if (st != SWB_OK) return st;
    if (n_pairs == 0) return SWB_OK;
    if (!a_data || !a_offsets || !b_data || !b_offsets || !out) return fail(err, SWB_INVALID_ARG, "NULL tape or output pointer");
    DeviceCtx &ctx = engine->scope->devs[device_index];
    SWB_CUDA(cudaSetDevice(ctx.device));
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    PairSpace ps{a_data, a_offsets, b_data, b_offsets, n_pairs, 0, 0, nullptr, out};
    st = enqueue_schedule_count(ctx, ps, stream, err);
    if (st != SWB_OK) return st;
    if ((st = wait_schedule_count(ctx, err)) != SWB_OK) return st;
    uint32_t max_len = 0;
    for (int c = 0; c < SCHED_MCLASSES; ++c) max_len = std::max(max_len, ctx.h_stats->max_n[c]);
    // SWB_LOAD_VARIANT: 1 = one tile per CTA (default; measured faster, profiles/README.md), 0 = persistent CTAs with a
    // two-stage TMA pipeline
    static const int variant = getenv("SWB_LOAD_VARIANT") ? atoi(getenv("SWB_LOAD_VARIANT")) : 1;
    static const uint32_t tile_override = getenv("SWB_LOAD_TILE") ? (uint32_t)atoi(getenv("SWB_LOAD_TILE")) : 0u;
    const uint32_t stages = variant == 0 ? 2u : 1u;
    // pairs per tile: as many as the threads of a block (256) unless the staged spans overflow shared memory
    const uint32_t budget = (uint32_t)(ctx.max_smem_optin - 1024) / (2u * stages);
    uint32_t tile_pairs = tile_override ? std::min<uint32_t>(tile_override, LOAD_PAIRS) : (uint32_t)LOAD_PAIRS;
    while (tile_pairs > 1 && tile_pairs * max_len + 160u > budget) tile_pairs /= 2;
    const uint32_t cap = (round16(tile_pairs * max_len + 32u) + 127u) & ~127u;
    const size_t smem = 2 * stages * (size_t)cap;
    if (smem > (size_t)ctx.max_smem_optin - 1024) return fail(err, SWB_TOO_LONG, "load-phase kernel supports strings up to %u bytes", budget - 160u);
    const uint64_t tiles = (n_pairs + tile_pairs - 1) / tile_pairs;
    if (variant == 0) {
        SWB_CUDA(allow_smem((const void *)load_phase_kernel, ctx.device, smem));
        // persistent grid: as many CTAs as fit (shared memory bound), never more than there are tiles
        const unsigned per_sm = (unsigned)std::max<size_t>(1, std::min<size_t>(8, (size_t)(ctx.max_smem_optin - 1024) / (smem + 1024)));
        const unsigned grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)ctx.sm_count * per_sm);
        ProfileBracket bracket(engine, ctx.device, stream);
        load_phase_kernel<<<grid, LOAD_THREADS, smem, stream>>>(ps, cap, cap, tile_pairs);
    } else {
        SWB_CUDA(allow_smem((const void *)load_phase_simple_kernel, ctx.device, smem));
        ProfileBracket bracket(engine, ctx.device, stream);
        load_phase_simple_kernel<<<(unsigned)tiles, LOAD_THREADS, smem, stream>>>(ps, cap, cap, tile_pairs);
    }
    g_launches++;
    engine->last_kernel = "load_phase";
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

// ------------------------------------------------------------------------------------------ host-tape API

// Copy offsets[lo..hi] and bytes [offsets[lo], offsets[hi]) of one tape to the device (copy stream). The
// device data pointer is biased so that kernels can index it with the ORIGINAL offsets.
struct TapeOnDevice {
    const uint8_t *data_biased = nullptr;
    const uint64_t *offs_biased = nullptr;
};

static swb_status_t upload_offsets(DeviceCtx &ctx, DeviceBuffer &buf, const uint64_t *offs, uint64_t lo, uint64_t hi,
                                   const uint64_t **biased, const char **err) {
    SWB_CUDA(buf.reserve((hi - lo + 1) * sizeof(uint64_t)));
    SWB_CUDA(cudaMemcpyAsync(buf.ptr, offs + lo, (hi - lo + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx.copy_stream));
    *biased = (const uint64_t *)buf.ptr - lo;
    return SWB_OK;
}

static swb_status_t validate_offsets(const uint64_t *offs, size_t n, const char *name, const char **err) {
    // O(1) sanity only: a full monotonicity scan would cost more than the GPU call it protects.
    if (offs[n] < offs[0]) return fail(err, SWB_INVALID_ARG, "%s offsets are not monotonic (last %llu < first %llu)", name,
                                       (unsigned long long)offs[n], (unsigned long long)offs[0]);
    return SWB_OK;
}

extern "C" swb_status_t swb_compute_pairs(swb_engine_t *engine, const uint8_t *a_data, const uint64_t *a_offsets,
                                          const uint8_t *b_data, const uint64_t *b_offsets, size_t n_pairs, uint64_t *out,
                                          const char **err) {
    if (!engine) return fail(err, SWB_INVALID_ARG, "engine is NULL");
    if (n_pairs == 0) return SWB_OK;
    if (!a_data || !a_offsets || !b_data || !b_offsets || !out) return fail(err, SWB_INVALID_ARG, "NULL tape or output pointer");
    if (n_pairs >= 0xFFFFFFFFull) return fail(err, SWB_INVALID_ARG, "more than 2^32-1 pairs in one call");
    swb_status_t st;
    if ((st = validate_offsets(a_offsets, n_pairs, "a", err)) != SWB_OK) return st;
    if ((st = validate_offsets(b_offsets, n_pairs, "b", err)) != SWB_OK) return st;
    swb_scope *scope = engine->scope;
    const int D = (int)scope->devs.size();

    // Shard [0, n_pairs) over the devices by bytes (bytes track cells for similar-length pairs).
    std::vector<uint64_t> cut(D + 1, 0);
    cut[D] = n_pairs;
    const uint64_t total_bytes = (a_offsets[n_pairs] - a_offsets[0]) + (b_offsets[n_pairs] - b_offsets[0]);
    for (int d = 1; d < D; ++d) {
        const uint64_t target = total_bytes / D * d;
        uint64_t lo = cut[d - 1], hi = n_pairs;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) / 2;
            const uint64_t bytes = (a_offsets[mid] - a_offsets[0]) + (b_offsets[mid] - b_offsets[0]);
            if (bytes < target) lo = mid + 1;
            else hi = mid;
        }
        cut[d] = lo;
    }

    struct Shard {
        uint64_t lo, hi;
        PairSpace ps;
        std::vector<uint64_t> chunk_cut;
    };
    std::vector<Shard> shards(D);

    const double t_begin = trace_enabled() ? now_ms() : 0;
    double t_phase1 = 0, t_stats = 0, t_phase2 = 0;
    // Phase 1: per device, enqueue offsets H2D, schedule counters, then chunked data H2D.
    for (int d = 0; d < D; ++d) {
        DeviceCtx &ctx = scope->devs[d];
        Shard &sh = shards[d];
        sh.lo = cut[d], sh.hi = cut[d + 1];
        const uint64_t n = sh.hi - sh.lo;
        if (n == 0) continue;
        SWB_CUDA(cudaSetDevice(ctx.device));
        const uint64_t *d_a_offs, *d_b_offs;
        if ((st = upload_offsets(ctx, ctx.a_offs, a_offsets, sh.lo, sh.hi, &d_a_offs, err)) != SWB_OK) return st;
        if ((st = upload_offsets(ctx, ctx.b_offs, b_offsets, sh.lo, sh.hi, &d_b_offs, err)) != SWB_OK) return st;
        SWB_CUDA(cudaEventRecord(ctx.ev_offsets, ctx.copy_stream));
        const uint64_t a_lo = a_offsets[sh.lo], a_hi = a_offsets[sh.hi], b_lo = b_offsets[sh.lo], b_hi = b_offsets[sh.hi];
        SWB_CUDA(ctx.a_data.reserve(a_hi - a_lo + 64));
        SWB_CUDA(ctx.b_data.reserve(b_hi - b_lo + 64));
        SWB_CUDA(ctx.out.reserve(n * sizeof(uint64_t)));
        // keep (device address mod 16) == (tape offset mod 16) so aligned tapes stay aligned on the device
        uint8_t *d_a = (uint8_t *)ctx.a_data.ptr + (a_lo & 15u), *d_b = (uint8_t *)ctx.b_data.ptr + (b_lo & 15u);
        sh.ps = PairSpace{d_a - a_lo, d_a_offs + sh.lo, d_b - b_lo, d_b_offs + sh.lo, n, 0, 0, nullptr, (uint64_t *)ctx.out.ptr};
        // PairSpace indexes pairs from 0: shift the offset arrays so that local pair k is global pair lo + k
        // (d_a_offs is already biased by -lo, so +lo lands on the first uploaded element).
        SWB_CUDA(cudaStreamWaitEvent(ctx.compute_stream, ctx.ev_offsets, 0));
        if ((st = enqueue_schedule_count(ctx, sh.ps, ctx.compute_stream, err)) != SWB_OK) return st;
        // data in ~16 MiB chunks so compute on chunk i overlaps the copy of chunk i+1
        const uint64_t bytes = (a_hi - a_lo) + (b_hi - b_lo);
        static const uint64_t chunk_mb = getenv("SWB_CHUNK_MB") ? std::max(1, atoi(getenv("SWB_CHUNK_MB"))) : 16; // measured 2/4/8/16/32 MiB: 1.22/1.38/1.45/1.49/1.45 TCUPS e2e at 64x64
        int chunks = (int)std::min<uint64_t>(64, std::max<uint64_t>(1, bytes / (chunk_mb << 20)));
        sh.chunk_cut.assign(chunks + 1, sh.lo);
        sh.chunk_cut[chunks] = sh.hi;
        for (int c = 1; c < chunks; ++c) sh.chunk_cut[c] = sh.lo + n * c / chunks;
        while ((int)ctx.ev_chunks.size() < chunks) {
            cudaEvent_t ev;
            SWB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx.ev_chunks.push_back(ev);
        }
        for (int c = 0; c < chunks; ++c) {
            const uint64_t p0 = sh.chunk_cut[c], p1 = sh.chunk_cut[c + 1];
            const uint64_t ca0 = a_offsets[p0], ca1 = a_offsets[p1], cb0 = b_offsets[p0], cb1 = b_offsets[p1];
            if (ca1 > ca0) SWB_CUDA(cudaMemcpyAsync(d_a + (ca0 - a_lo), a_data + ca0, ca1 - ca0, cudaMemcpyHostToDevice, ctx.copy_stream));
            if (cb1 > cb0) SWB_CUDA(cudaMemcpyAsync(d_b + (cb0 - b_lo), b_data + cb0, cb1 - cb0, cudaMemcpyHostToDevice, ctx.copy_stream));
            SWB_CUDA(cudaEventRecord(ctx.ev_chunks[c], ctx.copy_stream));
        }
    }

    if (trace_enabled()) t_phase1 = now_ms();
    // Phase 2: per device, read the counters, then launch per chunk (identity) or per class (bucketed).
    for (int d = 0; d < D; ++d) {
        DeviceCtx &ctx = scope->devs[d];
        Shard &sh = shards[d];
        const uint64_t n = sh.hi - sh.lo;
        if (n == 0) continue;
        SWB_CUDA(cudaSetDevice(ctx.device));
        if ((st = wait_schedule_count(ctx, err)) != SWB_OK) return st;
        if (trace_enabled() && d == 0) t_stats = now_ms();
        Plan plan = make_plan(*ctx.h_stats);
        const int chunks = (int)sh.chunk_cut.size() - 1;
        if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) {
            // decode needs the whole shard resident: no chunk pipelining on this path
            SWB_CUDA(cudaStreamWaitEvent(ctx.compute_stream, ctx.ev_chunks[chunks - 1], 0));
            if ((st = run_device_batch_utf8(engine, d, sh.ps, n, n, a_offsets[sh.hi] - a_offsets[sh.lo], b_offsets[sh.hi] - b_offsets[sh.lo],
                                            ctx.compute_stream, err)) != SWB_OK) return st;
            SWB_CUDA(cudaMemcpyAsync(out + sh.lo, ctx.out.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx.compute_stream));
        } else if (plan.identity) {
            for (int c = 0; c < chunks; ++c) {
                const uint64_t p0 = sh.chunk_cut[c] - sh.lo, p1 = sh.chunk_cut[c + 1] - sh.lo;
                SWB_CUDA(cudaStreamWaitEvent(ctx.compute_stream, ctx.ev_chunks[c], 0));
                if ((st = launch_class(engine, ctx, d, sh.ps, p0, p1 - p0, plan.max_m[0], plan.max_n[0], ctx.compute_stream, err, plan.same_shape)) != SWB_OK) return st;
                SWB_CUDA(cudaMemcpyAsync(out + sh.lo + p0, (uint64_t *)ctx.out.ptr + p0, (p1 - p0) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx.compute_stream));
            }
        } else {
            SWB_CUDA(cudaStreamWaitEvent(ctx.compute_stream, ctx.ev_chunks[chunks - 1], 0));
            SWB_CUDA(ctx.order.reserve(n * sizeof(uint32_t)));
            SWB_CUDA(cudaMemcpyAsync(ctx.d_stats->cursor, ctx.h_stats->cursor, sizeof(ctx.h_stats->cursor), cudaMemcpyHostToDevice, ctx.compute_stream));
            const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx.sm_count * 8);
            schedule_scatter_kernel<<<blocks, 256, 0, ctx.compute_stream>>>(sh.ps, ctx.d_stats, (uint32_t *)ctx.order.ptr);
            g_launches++;
            SWB_CUDA(cudaGetLastError());
            PairSpace ps = sh.ps;
            ps.order = (const uint32_t *)ctx.order.ptr;
            for (int i = 0; i < plan.n_classes; ++i)
                if ((st = launch_class(engine, ctx, d, ps, plan.begin[i], plan.count[i], plan.max_m[i], plan.max_n[i], ctx.compute_stream, err)) != SWB_OK) return st;
            SWB_CUDA(cudaMemcpyAsync(out + sh.lo, ctx.out.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx.compute_stream));
        }
    }
    if (trace_enabled()) t_phase2 = now_ms();
    // Phase 3: results are host-visible on return.
    for (int d = 0; d < D; ++d) {
        if (shards[d].hi == shards[d].lo) continue;
        SWB_CUDA(cudaSetDevice(scope->devs[d].device));
        SWB_CUDA(cudaStreamSynchronize(scope->devs[d].compute_stream));
    }
    if (trace_enabled())
        fprintf(stderr, "[swb trace] pairs=%zu enqueue=%.3f ms stats_wait=%.3f ms launch=%.3f ms drain=%.3f ms total=%.3f ms\n", n_pairs,
                t_phase1 - t_begin, t_stats - t_phase1, t_phase2 - t_stats, now_ms() - t_phase2, now_ms() - t_begin);
    return SWB_OK;
}

extern "C" swb_status_t swb_compute_cross(swb_engine_t *engine, const uint8_t *q_data, const uint64_t *q_offsets,
                                          size_t n_queries, const uint8_t *c_data, const uint64_t *c_offsets,
                                          size_t n_candidates, uint64_t *out, size_t row_stride, const char **err) {
    if (!engine) return fail(err, SWB_INVALID_ARG, "engine is NULL");
    if (n_queries == 0 || n_candidates == 0) return SWB_OK;
    if (!q_data || !q_offsets || !c_data || !c_offsets || !out) return fail(err, SWB_INVALID_ARG, "NULL tape or output pointer");
    if (row_stride < n_candidates) return fail(err, SWB_INVALID_ARG, "row_stride %zu < n_candidates %zu", row_stride, n_candidates);
    swb_status_t st;
    if ((st = validate_offsets(q_offsets, n_queries, "queries", err)) != SWB_OK) return st;
    if ((st = validate_offsets(c_offsets, n_candidates, "candidates", err)) != SWB_OK) return st;
    swb_scope *scope = engine->scope;
    const int D = (int)scope->devs.size();
    // Shard query rows over devices; every device gets all candidates.
    for (int d = 0; d < D; ++d) {
        DeviceCtx &ctx = scope->devs[d];
        const uint64_t q_lo = (uint64_t)n_queries * d / D, q_hi = (uint64_t)n_queries * (d + 1) / D;
        if (q_hi == q_lo) continue;
        SWB_CUDA(cudaSetDevice(ctx.device));
        const uint64_t *d_q_offs, *d_c_offs;
        if ((st = upload_offsets(ctx, ctx.a_offs, q_offsets, q_lo, q_hi, &d_q_offs, err)) != SWB_OK) return st;
        if ((st = upload_offsets(ctx, ctx.b_offs, c_offsets, 0, n_candidates, &d_c_offs, err)) != SWB_OK) return st;
        const uint64_t a_lo = q_offsets[q_lo], a_hi = q_offsets[q_hi], b_lo = c_offsets[0], b_hi = c_offsets[n_candidates];
        SWB_CUDA(ctx.a_data.reserve(a_hi - a_lo + 64));
        SWB_CUDA(ctx.b_data.reserve(b_hi - b_lo + 64));
        // Results go to a device staging matrix and are copied out after the kernels. Writing a page-locked, mapped
        // result matrix from inside the kernels (coalesced 128-byte segments) was measured at 23.5 GB/s on this pool,
        // half the copy engine's 52 GB/s, so it stays an experiment (SWB_CROSS_ZEROCOPY=1).
        uint64_t *mapped_out = nullptr;
        static const bool zero_copy = getenv("SWB_CROSS_ZEROCOPY") && atoi(getenv("SWB_CROSS_ZEROCOPY")) != 0;
        if (zero_copy) {
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
                mapped_out = (uint64_t *)attr.devicePointer;
            else
                (void)cudaGetLastError();
        }
        if (!mapped_out) SWB_CUDA(ctx.out.reserve((q_hi - q_lo) * n_candidates * sizeof(uint64_t)));
        uint8_t *d_a = (uint8_t *)ctx.a_data.ptr + (a_lo & 15u), *d_b = (uint8_t *)ctx.b_data.ptr + (b_lo & 15u);
        if (a_hi > a_lo) SWB_CUDA(cudaMemcpyAsync(d_a, q_data + a_lo, a_hi - a_lo, cudaMemcpyHostToDevice, ctx.copy_stream));
        if (b_hi > b_lo) SWB_CUDA(cudaMemcpyAsync(d_b, c_data + b_lo, b_hi - b_lo, cudaMemcpyHostToDevice, ctx.copy_stream));
        SWB_CUDA(cudaEventRecord(ctx.ev_offsets, ctx.copy_stream));
        SWB_CUDA(cudaStreamWaitEvent(ctx.compute_stream, ctx.ev_offsets, 0));
        PairSpace ps{d_a - a_lo, d_q_offs + q_lo, d_b - b_lo, d_c_offs, (q_hi - q_lo) * n_candidates, n_candidates, n_candidates, nullptr, (uint64_t *)ctx.out.ptr};
        if (mapped_out) ps.stride = row_stride, ps.out = mapped_out + q_lo * row_stride;
        if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) {
            if ((st = run_device_batch_utf8(engine, d, ps, q_hi - q_lo, n_candidates, a_hi - a_lo, b_hi - b_lo, ctx.compute_stream, err)) != SWB_OK) return st;
        } else if ((st = run_device_batch(engine, d, ps, ctx.compute_stream, err)) != SWB_OK) return st;
        if (!mapped_out)
            SWB_CUDA(cudaMemcpy2DAsync(out + q_lo * row_stride, row_stride * sizeof(uint64_t), ctx.out.ptr, n_candidates * sizeof(uint64_t),
                                       n_candidates * sizeof(uint64_t), q_hi - q_lo, cudaMemcpyDeviceToHost, ctx.compute_stream));
    }
    for (int d = 0; d < D; ++d) {
        SWB_CUDA(cudaSetDevice(scope->devs[d].device));
        SWB_CUDA(cudaStreamSynchronize(scope->devs[d].compute_stream));
    }
    return SWB_OK;
}

extern "C" uint64_t swb_levenshtein_pair(swb_engine_t *engine, const uint8_t *a, size_t a_length, const uint8_t *b,
                                         size_t b_length, const char **err) {
    const uint8_t dummy = 0;
    const uint64_t a_offs[2] = {0, a_length}, b_offs[2] = {0, b_length};
    uint64_t out = UINT64_MAX;
    if (swb_compute_pairs(engine, a ? a : &dummy, a_offs, b ? b : &dummy, b_offs, 1, &out, err) != SWB_OK) return UINT64_MAX;
    return out;
}
