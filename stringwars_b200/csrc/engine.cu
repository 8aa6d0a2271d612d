// engine.cu — C ABI (include/stringwars_b200.h), part 1: errors, device queries, scopes, engines, options, profiling.
//
// Host design: one scope owns one context per device entry (streams, scratch, schedule counters);
// engines and resident tapes register with their scope, so destroying the scope first is safe (their device state
// goes with it and the handles turn into inert shells). There is no CPU compute path anywhere in this library:
// without a device every entry point fails with SWB_NO_DEVICE.
#include "engine_internal.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>

using namespace swb;

// ------------------------------------------------------------------------------------------ errors

static thread_local char g_err[512];
std::atomic<uint64_t> swb::g_launches{0};

swb_status_t swb::fail(const char **err, swb_status_t st, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (err) *err = g_err;
    return st;
}

static std::mutex g_attr_mutex;
static std::map<std::pair<const void *, int>, size_t> g_attr_done;

// The attribute is raised to exactly what has been asked for so far: kernels with static shared memory cannot take
// the full 227 KB.
cudaError_t swb::allow_smem(const void *func, int device, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    auto it = g_attr_done.find({func, device});
    if (it != g_attr_done.end() && it->second >= bytes) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) g_attr_done[{func, device}] = bytes;
    return e;
}

swb_status_t swb::check_engine_alive(swb_engine *engine, const char **err) {
    if (!engine) return fail(err, SWB_INVALID_ARG, "engine is NULL");
    if (!engine->scope) return fail(err, SWB_INVALID_ARG, "the engine's scope has been destroyed");
    return SWB_OK;
}

// ------------------------------------------------------------------------------------------ queries

extern "C" int swb_version(void) { return 200; }

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
    if (c.comm_stream) cudaStreamDestroy(c.comm_stream);
    if (c.ev_inputs) cudaEventDestroy(c.ev_inputs);
    for (auto st : c.class_streams)
        if (st) cudaStreamDestroy(st);
    for (auto ev : c.ev_class)
        if (ev) cudaEventDestroy(ev);
    if (c.ev_t0) cudaEventDestroy(c.ev_t0);
    if (c.ev_t1) cudaEventDestroy(c.ev_t1);
    if (c.d_stats) cudaFree(c.d_stats);
    if (c.h_stats) cudaFreeHost(c.h_stats);
    if (c.d_spec) cudaFree(c.d_spec);
    if (c.h_spec) cudaFreeHost(c.h_spec);
    if (c.h_words) cudaFreeHost(c.h_words);
    if (c.ev_offsets) cudaEventDestroy(c.ev_offsets);
    for (auto e : c.ev_chunks) cudaEventDestroy(e);
    for (DeviceBuffer *b : {&c.order, &c.a_data, &c.a_offs, &c.b_data, &c.b_offs, &c.out, &c.sym_a, &c.sym_a_offs, &c.sym_b, &c.sym_b_offs,
                            &c.scan_tmp, &c.flags, &c.hscratch, &c.ren_a, &c.ren_a_offs, &c.ren_b, &c.ren_b_offs, &c.batch_dict, &c.rx_a, &c.rx_b, &c.rx_a_offs, &c.rx_b_offs, &c.rx_out})
        b->release();
}

namespace swb {
void nccl_state_destroy(swb_scope *scope); // multi_gpu.cu
void peer_state_destroy(swb_scope *scope); // multi_gpu.cu
}

static void release_engine_device_state(swb_engine *e) {
    if (e->scope)
        for (size_t i = 0; i < e->d_tables.size(); ++i)
            if (e->d_tables[i]) {
                cudaSetDevice(e->scope->devs[i].device);
                cudaFree(e->d_tables[i]);
            }
    e->d_tables.clear();
    for (auto &pe : e->prof_events) {
        if (pe.start) cudaEventDestroy(pe.start);
        if (pe.stop) cudaEventDestroy(pe.stop);
    }
    e->prof_events.clear();
    e->prof_used = 0;
}

static void release_tape_device_state(swb_tape *t) {
    if (t->scope)
        for (size_t i = 0; i < t->data.size(); ++i) {
            cudaSetDevice(t->scope->devs[i].device);
            t->data[i].release();
            t->offs[i].release();
        }
    t->data.clear(), t->offs.clear();
}

extern "C" swb_status_t swb_scope_create_devices(const int *devices, int num_devices, swb_scope_t **out, const char **err) {
    if (!out) return fail(err, SWB_INVALID_ARG, "out is NULL");
    *out = nullptr;
    const int available = swb_device_count();
    if (available == 0) return fail(err, SWB_NO_DEVICE, "no CUDA device is visible (this library has no CPU fallback)");
    if (!devices || num_devices < 1) return fail(err, SWB_INVALID_ARG, "empty device list");
    for (int i = 0; i < num_devices; ++i)
        if (devices[i] < 0 || devices[i] >= available)
            return fail(err, SWB_NO_DEVICE, "device %d requested but %d visible", devices[i], available);
    swb_scope *scope = new swb_scope();
    scope->devs.resize(num_devices);
    for (int i = 0; i < num_devices; ++i) {
        DeviceCtx &c = scope->devs[i];
        c.device = devices[i];
        cudaError_t e = cudaSetDevice(c.device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, c.device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c.concurrent_managed, cudaDevAttrConcurrentManagedAccess, c.device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c.compute_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) {
            // highest priority: the (tiny) schedule kernel must get SM slots while a compute kernel launched just
            // before it on another stream still has blocks waiting, otherwise it only runs in that kernel's tail. Same for
            // the NCCL kernels of the scatter: a compute kernel that fills the machine would otherwise hold every later
            // chunk's send back until its own tail (measured: the exchange then serialises behind the root's shard).
            int least = 0, greatest = 0;
            cudaDeviceGetStreamPriorityRange(&least, &greatest);
            e = cudaStreamCreateWithPriority(&c.side_stream, cudaStreamNonBlocking, greatest);
            if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c.comm_stream, cudaStreamNonBlocking, greatest);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c.ev_inputs, cudaEventDisableTiming);
        for (int k = 0; k < DeviceCtx::CLASS_STREAMS && e == cudaSuccess; ++k) {
            e = cudaStreamCreateWithFlags(&c.class_streams[k], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c.ev_class[k], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventCreate(&c.ev_t0);
        if (e == cudaSuccess) e = cudaEventCreate(&c.ev_t1);
        if (e == cudaSuccess) e = cudaMalloc(&c.d_stats, sizeof(ScheduleStats));
        if (e == cudaSuccess) e = cudaMemset(c.d_stats, 0, sizeof(ScheduleStats));
        if (e == cudaSuccess) e = cudaMemset(&c.d_stats->min_m, 0xFF, 2 * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaHostAlloc(&c.h_stats, sizeof(ScheduleStats), cudaHostAllocMapped);
        if (e == cudaSuccess) {
            memset(c.h_stats, 0, sizeof(ScheduleStats));
            e = cudaHostGetDevicePointer((void **)&c.h_stats_dev, c.h_stats, 0);
        }
        if (e == cudaSuccess) e = cudaMalloc(&c.d_spec, sizeof(SpecReport));
        if (e == cudaSuccess) {
            const SpecReport armed{0u, 0u, 0u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u};
            e = cudaMemcpy(c.d_spec, &armed, sizeof(armed), cudaMemcpyHostToDevice);
        }
        if (e == cudaSuccess) e = cudaHostAlloc(&c.h_spec, sizeof(SpecPublished), cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaHostAlloc(&c.h_words, 8 * sizeof(uint64_t), cudaHostAllocDefault);
        if (e == cudaSuccess) {
            memset(c.h_spec, 0, sizeof(SpecPublished));
            e = cudaHostGetDevicePointer((void **)&c.h_spec_dev, c.h_spec, 0);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c.ev_offsets, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            for (auto &d : scope->devs) destroy_ctx(d);
            delete scope;
            return fail(err, SWB_CUDA_ERROR, "scope setup on device %d failed: %s", devices[i], cudaGetErrorString(e));
        }
    }
    *out = scope;
    return SWB_OK;
}

extern "C" swb_status_t swb_scope_create(int first_device, int num_devices, swb_scope_t **out, const char **err) {
    if (!out) return fail(err, SWB_INVALID_ARG, "out is NULL");
    *out = nullptr;
    const int available = swb_device_count();
    if (available == 0) return fail(err, SWB_NO_DEVICE, "no CUDA device is visible (this library has no CPU fallback)");
    if (num_devices < 1 || first_device < 0 || first_device + num_devices > available)
        return fail(err, SWB_NO_DEVICE, "devices [%d, %d) requested but %d visible", first_device, first_device + num_devices, available);
    std::vector<int> list(num_devices);
    for (int i = 0; i < num_devices; ++i) list[i] = first_device + i;
    return swb_scope_create_devices(list.data(), num_devices, out, err);
}

extern "C" void swb_scope_destroy(swb_scope_t *scope) {
    if (!scope) return;
    // engines and tapes created on this scope outlive it as inert handles: their device state is released here
    for (swb_engine *e : scope->engines) {
        release_engine_device_state(e);
        e->scope = nullptr;
    }
    for (swb_tape *t : scope->tapes) {
        release_tape_device_state(t);
        t->scope = nullptr;
    }
    nccl_state_destroy(scope);
    peer_state_destroy(scope);
    for (auto &d : scope->devs) destroy_ctx(d);
    delete scope;
}

extern "C" swb_status_t swb_scope_set_transport(swb_scope_t *scope, swb_transport_t transport, const char **err) {
    if (!scope) return fail(err, SWB_INVALID_ARG, "scope is NULL");
    if (transport != SWB_TRANSPORT_NCCL && transport != SWB_TRANSPORT_COPY_ENGINE) return fail(err, SWB_INVALID_ARG, "unknown transport %d", (int)transport);
    scope->transport = transport;
    return SWB_OK;
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

extern "C" swb_status_t swb_unified_alloc(size_t bytes, void **out, const char **err) {
    if (!out) return fail(err, SWB_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (swb_device_count() == 0) return fail(err, SWB_NO_DEVICE, "no CUDA device is visible");
    cudaError_t e = cudaMallocManaged(out, bytes ? bytes : 1, cudaMemAttachGlobal);
    if (e != cudaSuccess) return fail(err, SWB_BAD_ALLOC, "cudaMallocManaged(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return SWB_OK;
}

extern "C" void swb_unified_free(void *ptr) {
    if (ptr) cudaFree(ptr);
}

// ------------------------------------------------------------------------------------------ engines

static swb_status_t engine_finish(swb_engine *e, swb_engine_t **out, const char **err) {
    e->d_tables.assign(e->scope->devs.size(), nullptr);
    e->scope->engines.insert(e);
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
    // the reference's engines take i8 gap costs (`error_cost_t`): anything wider is a caller bug, and the packed s16x2
    // kernels rely on |gap| staying far from the int16 floor of their gap tracks
    if (open < -128 || extend < -128) return fail(err, SWB_INVALID_ARG, "gap costs must fit int8, got open=%d extend=%d", open, extend);
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
    release_engine_device_state(engine);
    if (engine->scope) engine->scope->engines.erase(engine);
    delete engine;
}

extern "C" swb_status_t swb_engine_set_algorithm(swb_engine_t *engine, swb_algorithm_t algorithm, const char **err) {
    if (!engine) return fail(err, SWB_INVALID_ARG, "engine is NULL");
    if (algorithm != SWB_ALGO_AUTO && algorithm != SWB_ALGO_WAVEFRONT && algorithm != SWB_ALGO_BITPARALLEL)
        return fail(err, SWB_INVALID_ARG, "unknown algorithm %d", (int)algorithm);
    if (algorithm == SWB_ALGO_BITPARALLEL && engine->kind != ENGINE_LEVENSHTEIN && engine->kind != ENGINE_LEVENSHTEIN_UTF8)
        return fail(err, SWB_INVALID_ARG, "the bit-parallel kernel only exists for unit-cost Levenshtein");
    engine->algorithm = algorithm;
    engine->spec_class = 0;
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
        const ProfEvents &pe = engine->prof_events[i];
        SWB_CUDA(cudaSetDevice(pe.device));
        SWB_CUDA(cudaEventSynchronize(pe.stop));
        SWB_CUDA(cudaEventElapsedTime(&ms, pe.start, pe.stop));
        total += ms;
    }
    *total_ms = total, *launches = engine->prof_used;
    engine->prof_used = 0;
    return SWB_OK;
}
