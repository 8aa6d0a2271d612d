// pipeline.cu — C ABI (include/stringwars_b200.h), part 2: schedule / UTF-8 launches, device-resident calls, the host
// pipelines over typed tape views (any memory kind), resident tapes.
//
// Host-pointer calls shard the pair range over the scope's devices, then per device pipeline
//   offsets H2D -> schedule counters (also the offsets validator) -> chunked data H2D overlapped with compute -> results D2H.
// Every caller pointer is classified once per call (cudaPointerGetAttributes): managed and device memory is used in
// place — no staging copy in either direction — which is what the reference's `BytesTape<u64, UnifiedAlloc>` /
// `UnifiedMat` (similarities/bench.rs:292-295, 130) rely on; only plain host memory is staged.
#define SWB_SCHEDULE_KERNELS
#include "engine_internal.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cub/device/device_scan.cuh>

#include "utf8.cuh"

using namespace swb;

// ------------------------------------------------------------------------------------------ tracing
// SWB_TRACE=1 prints host-side phase timestamps of the host-tape pipeline to stderr (diagnostics only; it changes
// nothing about what is launched).
static bool trace_enabled() {
    static const bool on = getenv("SWB_TRACE") && getenv("SWB_TRACE")[0] == '1';
    return on;
}
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------ schedule

// One length class of one batch on one device.
static swb_status_t launch_class(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin,
                                 uint64_t count, uint32_t max_m, uint32_t max_n, cudaStream_t stream, const char **err, bool same_shape = false) {
    if (count == 0) return SWB_OK;
    const bool lev = eng->kind == ENGINE_LEVENSHTEIN || eng->kind == ENGINE_LEVENSHTEIN_UTF8;
    if (lev && ps.len_shift == 0 && eng->algorithm != SWB_ALGO_WAVEFRONT)
        return launch_lev_class(eng, ctx, ps, begin, count, max_m, max_n, stream, err, same_shape ? max_m : 0u);
    return launch_wavefront(eng, ctx, dev_index, ps, begin, count, max_m, max_n, stream, err, same_shape);
}

// Schedule counters for a device-resident PairSpace: enqueue the count kernel on `stream`; it publishes into
// mapped host memory and re-arms the device counters itself (no memset / D2H copy / event on the critical path).
// `extra_*`: a second offsets array validated in the same pass (the candidates of a cross product).
static swb_status_t enqueue_schedule_count(DeviceCtx &ctx, const PairSpace &ps, cudaStream_t stream, const char **err,
                                           const void *extra_offs = nullptr, uint64_t extra_n = 0, uint32_t extra32 = 0) {
    const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((std::max(ps.n_pairs, extra_n) + 1023) / 1024, (uint64_t)ctx.sm_count * 4));
    ctx.stats_seq++;
    if (ctx.stats_seq == 0 || ctx.stats_seq == 0xFFFFFFFFu) ctx.stats_seq = 1;
    ctx.stats_stream = stream;
    schedule_count_kernel<<<blocks, 256, 0, stream>>>(ps, ctx.d_stats, ctx.h_stats_dev, ctx.stats_seq, extra_offs, extra_n, extra32);
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

static swb_status_t refuse_bad_offsets(DeviceCtx &ctx, const char **err) {
    if (ctx.h_stats->bad == 0) return SWB_OK;
    return fail(err, SWB_INVALID_ARG, "offsets are not monotonic (or a string spans 4 GiB or more): %u bad strings", ctx.h_stats->bad);
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
        // longest texts first: the blocks that take longest start first and the class's tail is made of short ones
        for (int b = SCHED_NBINS - 1; b >= 0; --b) {
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

static swb_status_t enqueue_scatter(DeviceCtx &ctx, PairSpace &ps, cudaStream_t stream, const char **err) {
    SWB_CUDA(ctx.order.reserve(ps.n_pairs * sizeof(uint32_t)));
    SWB_CUDA(cudaMemcpyAsync(ctx.d_stats->cursor, ctx.h_stats->cursor, sizeof(ctx.h_stats->cursor), cudaMemcpyHostToDevice, stream));
    const unsigned blocks = (unsigned)std::min<uint64_t>((ps.n_pairs + 255) / 256, (uint64_t)ctx.sm_count * 8);
    schedule_scatter_kernel<<<blocks, 256, 0, stream>>>(ps, ctx.d_stats, (uint32_t *)ctx.order.ptr);
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    ps.order = (const uint32_t *)ctx.order.ptr;
    return SWB_OK;
}

// The speculative kernel family of a uniform byte-Levenshtein batch: its launch geometry depends on nothing but the
// pair count. Classes 0..3 are the table kernels of 1, 2, 4, 8 words; short patterns use the register kernels.
static int spec_kernel_of(const Plan &plan) {
    const uint32_t m = plan.max_m[0], n = plan.max_n[0];
    if (plan.same_shape && m == 64) return 102; // every pattern fills two words exactly: the software-pipelined kernel
    if (n <= 8) return -12;
    if (n <= 16) return -14;
    if (n <= 24) return -16;
    if (m <= 8) return -2;
    if (m <= 16) return -4;
    if (m <= 24) return -6;
    if (m <= 32) return 1;
    if (m <= 64) return 2;
    if (m <= 128) return 4;
    if (m <= 256) return 8;
    return 0;
}

// A self-validating speculative launch (common.cuh, spec_finish): arm the report, and wait for what the kernel's last
// block publishes. The wait spins on one mapped word and checks the stream now and then for a failed launch.
static SpecArgs arm_spec(DeviceCtx &ctx) {
    ctx.spec_seq++;
    if (ctx.spec_seq == 0) ctx.spec_seq = 1;
    SpecArgs sa;
    sa.report = ctx.d_spec, sa.published = ctx.h_spec_dev, sa.seq = ctx.spec_seq;
    return sa;
}
static swb_status_t wait_spec(DeviceCtx &ctx, cudaStream_t stream, const char **err) {
    volatile uint32_t *seq = &ctx.h_spec->seq;
    for (uint64_t spins = 0; *seq != ctx.spec_seq; ++spins) {
        if ((spins & 0x3FFF) == 0x3FFF) {
            cudaError_t q = cudaStreamQuery(stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) return fail(err, SWB_CUDA_ERROR, "speculative kernel failed: %s", cudaGetErrorString(q));
            if (q == cudaSuccess && *seq != ctx.spec_seq && spins > (1ull << 26)) return fail(err, SWB_CUDA_ERROR, "speculative kernel never reported");
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return SWB_OK;
}

// The kernel a uniform byte-Levenshtein batch of these extrema would get from the regular path (launch_lev_bitpar's code),
// 0 when the batch is not uniform enough to be worth a bet (one pattern class, text lengths within two 32-byte bins).
static int spec_kernel_for(uint32_t min_m, uint32_t max_m, uint32_t min_n, uint32_t max_n) {
    auto cls = [](uint32_t m) { return m <= 32 ? 0 : (m <= 64 ? 1 : (m <= 128 ? 2 : (m <= 256 ? 3 : 4))); };
    if (max_m > 256 || cls(std::max(min_m, 1u)) != cls(std::max(max_m, 1u))) return 0;
    if ((max_n + 31u) / 32u > (min_n + 31u) / 32u + 1u) return 0;
    if (max_n <= 8) return -12;
    if (max_n <= 16) return -14;
    if (max_n <= 24) return -16;
    if (max_m <= 8) return -2;
    if (max_m <= 16) return -4;
    if (max_m <= 24) return -6;
    if (max_m <= 32) return 1;
    if (max_m <= 64) return 2;
    if (max_m <= 128) return 4;
    return 8;
}

// Batches up to this many pairs bet with ONE self-validating launch and wait for its report (the call is latency-bound:
// a second launch and a statistics pass would be most of it). Larger batches bet with the kernel on the caller's stream
// and the statistics kernel beside it on a side stream, and return as soon as the statistics are in — the compute kernel
// keeps running, so consecutive calls (and the devices of a multi-device scope) overlap.
constexpr uint64_t SPEC_SINGLE_LAUNCH_PAIRS = 1u << 17;

swb_status_t swb::run_device_batch(swb_engine *eng, int dev_index, PairSpace ps, cudaStream_t stream, const char **err, bool *completed) {
    DeviceCtx &ctx = eng->scope->devs[dev_index];
    if (completed) *completed = false;
    if (ps.n_pairs == 0) return SWB_OK;
    if (ps.n_pairs >= 0xFFFFFFFFull) return fail(err, SWB_INVALID_ARG, "more than 2^32-1 pairs in one call");
    ps.order = nullptr;
    const bool lev = eng->kind == ENGINE_LEVENSHTEIN || eng->kind == ENGINE_LEVENSHTEIN_UTF8;
    const bool lev_bytes = lev && ps.len_shift == 0;
    swb_status_t st;
    // Cross products large enough to amortise a per-query table build: statistics of the QUERY tape alone decide
    // the kernel (queries are the patterns), then one thread per query sweeps its share of the candidates.
    // candidates per block: enough blocks to fill the machine twice, at least 16 candidates per table build; a cross
    // product too small to fill the SMs with (query block, candidate range) tiles runs the pair kernels instead
    const uint64_t n_q_all = ps.n_c ? ps.n_pairs / ps.n_c : 0, query_blocks = (n_q_all + 127) / 128;
    uint64_t cpb = std::max<uint64_t>(16, ps.n_c * query_blocks / std::max<uint64_t>(1, (uint64_t)ctx.sm_count * 8));
    cpb = std::min<uint64_t>(cpb, std::max<uint64_t>(ps.n_c, 1));
    const bool enough_blocks = ps.n_c >= 32 && query_blocks * ((ps.n_c + cpb - 1) / cpb) >= (uint64_t)ctx.sm_count;
    if (enough_blocks && lev_bytes && eng->algorithm != SWB_ALGO_WAVEFRONT) {
        const uint64_t n_q = n_q_all;
        PairSpace qs = ps; // degenerate pair space (q_i, q_i): its counters are the query-length statistics
        qs.b_data = ps.a_data, qs.b_offs = ps.a_offs, qs.n_pairs = n_q, qs.n_c = 0;
        qs.offs32 = (ps.offs32 & 1u) ? 3u : 0u;
        // Speculate like the pair path does: the harness calls with the same shape over and over (bench.rs:587-605), so the
        // cross kernel of the previous call's class goes out at once and the statistics / offsets check runs beside it.
        // Queries that do not fit the speculated K are skipped by the kernel and candidates are never read past the
        // previous call's longest, so a wrong bet costs a re-run, never a wrong or out-of-bounds access.
        const uint32_t spec_q = eng->spec_cross_q, spec_n = eng->spec_cross_n;
        if (spec_q) {
            SWB_CUDA(cudaEventRecord(ctx.ev_inputs, stream));
            SWB_CUDA(cudaStreamWaitEvent(ctx.side_stream, ctx.ev_inputs, 0));
            if ((st = launch_lev_cross(eng, ctx, ps, (uint32_t)n_q, (uint32_t)cpb, spec_q, spec_n, SpecArgs{}, stream, err)) != SWB_OK) return st;
            if ((st = enqueue_schedule_count(ctx, qs, ctx.side_stream, err, ps.b_offs, ps.n_c, (ps.offs32 >> 1) & 1u)) != SWB_OK) return st;
        } else if ((st = enqueue_schedule_count(ctx, qs, stream, err, ps.b_offs, ps.n_c, (ps.offs32 >> 1) & 1u)) != SWB_OK) {
            return st;
        }
        if ((st = wait_schedule_count(ctx, err)) != SWB_OK) return st;
        eng->spec_cross_q = 0;
        if ((st = refuse_bad_offsets(ctx, err)) != SWB_OK) return st;
        uint32_t max_q = 0;
        for (int c = 0; c < SCHED_MCLASSES; ++c) max_q = std::max(max_q, ctx.h_stats->max_m[c]);
        const uint32_t max_c = ctx.h_stats->extra_max;
        auto words_of = [](uint32_t q) { return q <= 32 ? 1 : (q <= 64 ? 2 : (q <= 128 ? 4 : 8)); };
        if (max_q <= 256) {
            eng->spec_cross_q = std::max(max_q, 1u), eng->spec_cross_n = max_c;
            if (spec_q && words_of(max_q) == words_of(spec_q) && max_c <= spec_n) return SWB_OK; // the bet held
            return launch_lev_cross(eng, ctx, ps, (uint32_t)n_q, (uint32_t)cpb, max_q, 0xFFFFFFFFu, SpecArgs{}, stream, err);
        }
    }
    // Element-wise pairs (and cross products too small for the cross kernel): bet that this batch has the shape of the
    // previous one. A thread-per-pair kernel is correct for every pair that fits its words and skips the others (malformed
    // offsets and strings longer than the previous batch's longest included), and its launch geometry depends on nothing
    // but the pair count, so it goes onto the caller's stream IMMEDIATELY.
    const int spec = (lev_bytes && eng->algorithm != SWB_ALGO_WAVEFRONT) ? eng->spec_class : 0;
    bool have_stats = false;
    if (spec != 0 && ps.n_pairs <= SPEC_SINGLE_LAUNCH_PAIRS) {
        // small batch: the kernel validates as it goes and its last block reports misses and the batch's length extrema
        eng->spec_class = 0;
        const SpecArgs sa = arm_spec(ctx);
        if ((st = launch_lev_bitpar(eng, ctx, spec, ps, 0, ps.n_pairs, eng->spec_max_n, sa, stream, err)) != SWB_OK) return st;
        if ((st = wait_spec(ctx, stream, err)) != SWB_OK) return st;
        const SpecPublished &r = *ctx.h_spec;
        if (r.misses == 0) { // that kernel was the batch; the extrema decide the next bet (same kernel, a narrower one, or none)
            eng->spec_class = r.max_n ? spec_kernel_for(r.min_m, r.max_m, r.min_n, r.max_n) : spec;
            if (r.max_n) eng->spec_max_n = r.max_n;
            if (completed) *completed = true;
            return SWB_OK;
        }
    } else if (spec != 0) {
        // large batch: the statistics kernel runs beside the bet on a side stream; the host returns when the statistics are in
        SWB_CUDA(cudaEventRecord(ctx.ev_inputs, stream));
        SWB_CUDA(cudaStreamWaitEvent(ctx.side_stream, ctx.ev_inputs, 0));
        if ((st = launch_lev_bitpar(eng, ctx, spec, ps, 0, ps.n_pairs, eng->spec_max_n, SpecArgs{}, stream, err)) != SWB_OK) return st;
        if ((st = enqueue_schedule_count(ctx, ps, ctx.side_stream, err)) != SWB_OK) return st;
        have_stats = true;
    }
    // The regular way: statistics pass (also where malformed offsets become SWB_INVALID_ARG), bucketing, per-class kernels;
    // after a lost bet the results are simply overwritten in order.
    if (!have_stats && (st = enqueue_schedule_count(ctx, ps, stream, err)) != SWB_OK) return st;
    if ((st = wait_schedule_count(ctx, err)) != SWB_OK) return st;
    eng->spec_class = 0;
    if ((st = refuse_bad_offsets(ctx, err)) != SWB_OK) return st;
    Plan plan = make_plan(*ctx.h_stats);
    const bool uniform = plan.identity && plan.n_classes == 1;
    const int kernel_now = (uniform && lev_bytes) ? spec_kernel_of(plan) : 0;
    const uint32_t spec_cap = eng->spec_max_n;
    eng->spec_class = kernel_now, eng->spec_max_n = plan.n_classes ? plan.max_n[0] : 0u;
    if (have_stats && kernel_now == spec && eng->spec_max_n <= spec_cap) return SWB_OK; // the bet held: that kernel is the batch
    if (!plan.identity && (st = enqueue_scatter(ctx, ps, stream, err)) != SWB_OK) return st;
    // Byte-Levenshtein classes run beside each other (they share nothing but read-only tapes; only the topmost class can
    // need the strip scratch): the class with the most cells on the caller's stream, the others forked from it and joined
    // back. Alignment classes share scratch rows and stay in line.
    int on_side = 0;
    if (plan.n_classes > 1 && lev_bytes && eng->algorithm != SWB_ALGO_WAVEFRONT) {
        int biggest = 0;
        auto cells = [&](int i) { return (double)plan.count[i] * plan.max_m[i] * plan.max_n[i]; };
        for (int i = 1; i < plan.n_classes; ++i)
            if (cells(i) > cells(biggest)) biggest = i;
        SWB_CUDA(cudaEventRecord(ctx.ev_inputs, stream));
        st = launch_class(eng, ctx, dev_index, ps, plan.begin[biggest], plan.count[biggest], plan.max_m[biggest], plan.max_n[biggest], stream, err, plan.same_shape);
        if (st != SWB_OK) return st;
        const char *main_kernel = eng->last_kernel;
        for (int i = 0; i < plan.n_classes; ++i) {
            if (i == biggest || plan.count[i] == 0) continue;
            const int slot = on_side % DeviceCtx::CLASS_STREAMS;
            cudaStream_t side = ctx.class_streams[slot];
            if (on_side < DeviceCtx::CLASS_STREAMS) SWB_CUDA(cudaStreamWaitEvent(side, ctx.ev_inputs, 0));
            st = launch_class(eng, ctx, dev_index, ps, plan.begin[i], plan.count[i], plan.max_m[i], plan.max_n[i], side, err, plan.same_shape);
            if (st != SWB_OK) return st;
            ++on_side;
        }
        for (int slot = 0; slot < std::min(on_side, (int)DeviceCtx::CLASS_STREAMS); ++slot) {
            SWB_CUDA(cudaEventRecord(ctx.ev_class[slot], ctx.class_streams[slot]));
            SWB_CUDA(cudaStreamWaitEvent(stream, ctx.ev_class[slot], 0));
        }
        eng->last_kernel = main_kernel;
        return SWB_OK;
    }
    for (int i = 0; i < plan.n_classes; ++i) {
        st = launch_class(eng, ctx, dev_index, ps, plan.begin[i], plan.count[i], plan.max_m[i], plan.max_n[i], stream, err, plan.same_shape);
        if (st != SWB_OK) return st;
    }
    return SWB_OK;
}

// ------------------------------------------------------------------------------------------ UTF-8 engine

// Symbol counts of one device byte tape: sym_offs[i + 1] = 4 x (symbols of string i), and the ASCII flag is cleared when a
// byte >= 0x80 is seen. `offs` points at n+1 offsets (u64, or u32 when offs32) valid with `data` (biased pointers are fine).
static swb_status_t count_symbols(DeviceCtx &ctx, const uint8_t *data, const void *offs, uint32_t offs32, uint64_t n, DeviceBuffer &sym_offs,
                                  uint32_t *d_ascii_flag, cudaStream_t stream, const char **err) {
    SWB_CUDA(sym_offs.reserve((n + 1) * sizeof(uint64_t)));
    const unsigned blocks = (unsigned)((n * 32 + 255) / 256);
    utf8_count_kernel<<<blocks, 256, 0, stream>>>(data, offs, offs32, n, (uint64_t *)sym_offs.ptr, d_ascii_flag);
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}
// Counts -> offsets (inclusive scan in place): sym_offs[i] = 4 x (symbols before string i).
static swb_status_t scan_symbols(DeviceCtx &ctx, uint64_t n, DeviceBuffer &sym_offs, cudaStream_t stream, const char **err) {
    size_t tmp_bytes = 0;
    uint64_t *so = (uint64_t *)sym_offs.ptr;
    SWB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, so, so, (int)(n + 1), stream));
    SWB_CUDA(ctx.scan_tmp.reserve(tmp_bytes + 16));
    SWB_CUDA(cub::DeviceScan::InclusiveSum(ctx.scan_tmp.ptr, tmp_bytes, so, so, (int)(n + 1), stream));
    return SWB_OK;
}

// The tape's symbols as 32-bit words (the wavefront kernel's input); `total_bytes` bounds the symbol count.
static swb_status_t pack_symbols(DeviceCtx &ctx, const uint8_t *data, const void *offs, uint32_t offs32, uint64_t n, uint64_t total_bytes,
                                 DeviceBuffer &sym, const DeviceBuffer &sym_offs, cudaStream_t stream, const char **err) {
    SWB_CUDA(sym.reserve(total_bytes * 4 + 64));
    const unsigned blocks = (unsigned)((n * 32 + 255) / 256);
    utf8_pack_kernel<<<blocks, 256, 0, stream>>>(data, offs, offs32, n, (const uint64_t *)sym_offs.ptr, (uint32_t *)sym.ptr);
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

// Total bytes spanned by the strings of two DEVICE-readable offsets arrays (one small synchronous readback for both).
static swb_status_t device_spans(const void *a_offs, uint32_t a_width, uint64_t n_a, uint64_t *a_bytes, const void *b_offs, uint32_t b_width, uint64_t n_b,
                                 uint64_t *b_bytes, cudaStream_t stream, const char **err) {
    uint64_t ends[4] = {0, 0, 0, 0}; // little-endian: a u32 lands in the low half of the zeroed u64
    SWB_CUDA(cudaMemcpyAsync(&ends[0], a_offs, a_width, cudaMemcpyDeviceToHost, stream));
    SWB_CUDA(cudaMemcpyAsync(&ends[1], (const uint8_t *)a_offs + n_a * a_width, a_width, cudaMemcpyDeviceToHost, stream));
    SWB_CUDA(cudaMemcpyAsync(&ends[2], b_offs, b_width, cudaMemcpyDeviceToHost, stream));
    SWB_CUDA(cudaMemcpyAsync(&ends[3], (const uint8_t *)b_offs + n_b * b_width, b_width, cudaMemcpyDeviceToHost, stream));
    SWB_CUDA(cudaStreamSynchronize(stream));
    if (ends[1] < ends[0] || ends[3] < ends[2])
        return fail(err, SWB_INVALID_ARG, "offsets are not monotonic (last %llu < first %llu)", (unsigned long long)(ends[1] < ends[0] ? ends[1] : ends[3]),
                    (unsigned long long)(ends[1] < ends[0] ? ends[0] : ends[2]));
    *a_bytes = ends[1] - ends[0], *b_bytes = ends[3] - ends[2];
    return SWB_OK;
}

// Codepoint-level batch. n_a / n_b are the string counts of the two tapes, a_bytes / b_bytes their byte spans, or
// SPAN_UNKNOWN when the caller cannot read the offsets (the engine then reads them where it synchronises anyway).
swb_status_t swb::run_device_batch_utf8(swb_engine *eng, int dev_index, PairSpace ps, uint64_t n_a, uint64_t n_b,
                                        uint64_t a_bytes, uint64_t b_bytes, cudaStream_t stream, const char **err) {
    DeviceCtx &ctx = eng->scope->devs[dev_index];
    SWB_CUDA(ctx.flags.reserve(64));
    uint32_t *d_flag = (uint32_t *)ctx.flags.ptr;
    swb_status_t st;
    const uint32_t a32 = ps.offs32 & 1u, b32 = (ps.offs32 >> 1) & 1u;
    // Element-wise pairs: ONE pass renames every pair to a byte alphabet (utf8_rename_kernel; pure-ASCII text goes through
    // it like any other) and the byte kernels do the rest.
    const bool try_dict = ps.n_c != 0 && eng->algorithm != SWB_ALGO_WAVEFRONT;
    if (!try_dict && (a_bytes == SPAN_UNKNOWN || b_bytes == SPAN_UNKNOWN) &&
        (st = device_spans(ps.a_offs, a32 ? 4 : 8, n_a, &a_bytes, ps.b_offs, b32 ? 4 : 8, n_b, &b_bytes, stream, err)) != SWB_OK)
        return st;
    if (ps.n_c == 0 && eng->algorithm != SWB_ALGO_WAVEFRONT) {
        SWB_CUDA(ctx.ren_a.reserve(a_bytes + 64));
        SWB_CUDA(ctx.ren_b.reserve(b_bytes + 64));
        SWB_CUDA(ctx.ren_a_offs.reserve((2 * n_a + 2) * sizeof(uint64_t)));
        SWB_CUDA(ctx.ren_b_offs.reserve((2 * n_b + 2) * sizeof(uint64_t)));
        const unsigned blocks = (unsigned)((ps.n_pairs + RENAME_WARPS - 1) / RENAME_WARPS);
        // tier 0: ASCII keeps its bytes, 127 dictionary slots for the rest; tier 1: 255 slots for everything
        for (int tier = 0; tier < 2; ++tier) {
            SWB_CUDA(cudaMemsetAsync(d_flag + 1, 0, sizeof(uint32_t), stream));
            auto kernel = tier == 0 ? utf8_rename_kernel<true> : utf8_rename_kernel<false>;
            kernel<<<blocks, 32 * RENAME_WARPS, 0, stream>>>(ps.a_data, ps.a_offs, a32, a_bytes, ps.b_data, ps.b_offs, b32, b_bytes, ps.n_pairs,
                                                             (uint8_t *)ctx.ren_a.ptr, (uint8_t *)ctx.ren_b.ptr, (uint64_t *)ctx.ren_a_offs.ptr,
                                                             (uint64_t *)ctx.ren_b_offs.ptr, d_flag + 1);
            g_launches++;
            SWB_CUDA(cudaGetLastError());
            uint32_t h_flags = 0;
            SWB_CUDA(cudaMemcpyAsync(&h_flags, d_flag + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            SWB_CUDA(cudaStreamSynchronize(stream));
            if (h_flags & RENAME_BAD_OFFSETS) return fail(err, SWB_INVALID_ARG, "offsets are not monotonic (or leave the tape, or a string spans 4 GiB or more)");
            if (h_flags == 0) {
                PairSpace ren = ps;
                ren.a_data = (const uint8_t *)ctx.ren_a.ptr, ren.a_offs = (const uint64_t *)ctx.ren_a_offs.ptr;
                ren.b_data = (const uint8_t *)ctx.ren_b.ptr, ren.b_offs = (const uint64_t *)ctx.ren_b_offs.ptr;
                ren.len_shift = 0, ren.offs32 = 0, ren.idx_shift = 1; // two strings per input string: the renamed one and the gap behind it
                return run_device_batch(eng, dev_index, ren, stream, err);
            }
        }
    }
    // Cross products: a survey pass counts symbols, looks for non-ASCII bytes and numbers the batch's multi-byte symbols in
    // one batch-wide dictionary. Pure ASCII: bytes ARE code points, the byte kernels as they are. At most 127 distinct
    // multi-byte symbols: both tapes are translated to byte tapes (count -> scan -> translate) and the byte cross-product
    // kernels do the rest. Anything wider, pairs with too many distinct symbols and the explicit wavefront algorithm:
    // 32-bit symbols (count -> scan -> pack) on the wavefront kernel.
    const bool same_tape = ps.a_data == ps.b_data && ps.a_offs == ps.b_offs && n_a == n_b && a32 == b32;
    const uint64_t *a_so = nullptr, *b_so = nullptr;
    uint32_t h_ascii = 0, h_distinct = 0xFFFFFFFFu;
    const TapeArg ta{ps.a_data, ps.a_offs, a32, n_a}, tb{ps.b_data, ps.b_offs, b32, same_tape ? 0 : n_b};
    if (try_dict) {
        SWB_CUDA(ctx.batch_dict.reserve(sizeof(BatchDict)));
        SWB_CUDA(ctx.sym_a_offs.reserve((n_a + 1) * sizeof(uint64_t)));
        SWB_CUDA(ctx.sym_b_offs.reserve((n_b + 1) * sizeof(uint64_t)));
        BatchDict *dict = (BatchDict *)ctx.batch_dict.ptr;
        SWB_CUDA(cudaMemsetAsync(dict, 0xFF, sizeof(BatchDict), stream));
        utf8_survey_kernel<<<(unsigned)(((ta.n + tb.n) * 32 + 255) / 256), 256, 0, stream>>>(ta, (uint64_t *)ctx.sym_a_offs.ptr, tb, (uint64_t *)ctx.sym_b_offs.ptr, dict);
        g_launches++;
        SWB_CUDA(cudaGetLastError());
        SWB_CUDA(cudaMemcpyAsync(ctx.h_words, dict, offsetof(BatchDict, keys), cudaMemcpyDeviceToHost, stream)); // {ascii, distinct - 1, -, spans}
        SWB_CUDA(cudaStreamSynchronize(stream));
        const uint32_t *head = reinterpret_cast<const uint32_t *>(ctx.h_words);
        h_ascii = head[0], h_distinct = head[1] + 1u;
        if (head[2] == 0u) return fail(err, SWB_INVALID_ARG, "offsets are not monotonic (or leave the tape, or a string spans 4 GiB or more)");
        const uint64_t *span = ctx.h_words + 2;
        if (span[1] < span[0] || (!same_tape && span[3] < span[2])) return fail(err, SWB_INVALID_ARG, "offsets are not monotonic (last offset before the first)");
        if (a_bytes == SPAN_UNKNOWN) a_bytes = span[1] - span[0];
        if (b_bytes == SPAN_UNKNOWN) b_bytes = same_tape ? a_bytes : span[3] - span[2];
    } else {
        SWB_CUDA(cudaMemsetAsync(d_flag, 0x01, sizeof(uint32_t), stream));
        if ((st = count_symbols(ctx, ps.a_data, ps.a_offs, a32, n_a, ctx.sym_a_offs, d_flag, stream, err)) != SWB_OK) return st;
        if (!same_tape && (st = count_symbols(ctx, ps.b_data, ps.b_offs, b32, n_b, ctx.sym_b_offs, d_flag, stream, err)) != SWB_OK) return st;
        SWB_CUDA(cudaMemcpyAsync(&h_ascii, d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        SWB_CUDA(cudaStreamSynchronize(stream));
    }
    a_so = (const uint64_t *)ctx.sym_a_offs.ptr, b_so = same_tape ? a_so : (const uint64_t *)ctx.sym_b_offs.ptr;
    if (h_ascii != 0) return run_device_batch(eng, dev_index, ps, stream, err);
    if (std::max(n_a, n_b) < (1u << 16)) { // a few thousand strings: one block per tape
        utf8_small_scan_kernel<<<same_tape ? 1 : 2, SMALL_SCAN_THREADS, 0, stream>>>((uint64_t *)ctx.sym_a_offs.ptr, n_a + 1, (uint64_t *)ctx.sym_b_offs.ptr, n_b + 1);
        g_launches++;
        SWB_CUDA(cudaGetLastError());
    } else {
        if ((st = scan_symbols(ctx, n_a, ctx.sym_a_offs, stream, err)) != SWB_OK) return st;
        if (!same_tape && (st = scan_symbols(ctx, n_b, ctx.sym_b_offs, stream, err)) != SWB_OK) return st;
    }
    if (try_dict && h_distinct <= BATCH_DICT_CAP) {
        SWB_CUDA(ctx.ren_a.reserve(a_bytes + 64));
        SWB_CUDA(ctx.ren_a_offs.reserve((n_a + 1) * sizeof(uint64_t)));
        if (!same_tape) {
            SWB_CUDA(ctx.ren_b.reserve(b_bytes + 64));
            SWB_CUDA(ctx.ren_b_offs.reserve((n_b + 1) * sizeof(uint64_t)));
        }
        utf8_translate_kernel<<<(unsigned)(((ta.n + tb.n) * 32 + 255) / 256), 256, 0, stream>>>(ta, a_so, (uint8_t *)ctx.ren_a.ptr, (uint64_t *)ctx.ren_a_offs.ptr, tb, b_so,
                                                                                              (uint8_t *)ctx.ren_b.ptr, (uint64_t *)ctx.ren_b_offs.ptr,
                                                                                              (const BatchDict *)ctx.batch_dict.ptr);
        g_launches++;
        SWB_CUDA(cudaGetLastError());
        PairSpace ren = ps;
        ren.a_data = (const uint8_t *)ctx.ren_a.ptr, ren.a_offs = (const uint64_t *)ctx.ren_a_offs.ptr;
        ren.b_data = same_tape ? ren.a_data : (const uint8_t *)ctx.ren_b.ptr;
        ren.b_offs = same_tape ? ren.a_offs : (const uint64_t *)ctx.ren_b_offs.ptr;
        ren.len_shift = 0, ren.offs32 = 0;
        return run_device_batch(eng, dev_index, ren, stream, err);
    }
    if ((st = pack_symbols(ctx, ps.a_data, ps.a_offs, a32, n_a, a_bytes, ctx.sym_a, ctx.sym_a_offs, stream, err)) != SWB_OK) return st;
    if (!same_tape && (st = pack_symbols(ctx, ps.b_data, ps.b_offs, b32, n_b, b_bytes, ctx.sym_b, ctx.sym_b_offs, stream, err)) != SWB_OK) return st;
    PairSpace sym = ps;
    sym.a_data = (const uint8_t *)ctx.sym_a.ptr, sym.a_offs = a_so;
    sym.b_data = same_tape ? sym.a_data : (const uint8_t *)ctx.sym_b.ptr;
    sym.b_offs = b_so;
    sym.len_shift = 2, sym.offs32 = 0;
    return run_device_batch(eng, dev_index, sym, stream, err);
}

// ------------------------------------------------------------------------------------------ device-resident API

static swb_status_t check_engine(swb_engine_t *engine, int device_index, const char **err) {
    swb_status_t st = check_engine_alive(engine, err);
    if (st != SWB_OK) return st;
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
    PairSpace ps = make_pair_space(a_data, a_offsets, b_data, b_offsets, n_pairs, 0, 0, out, 0, 0);
    if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) {
        return run_device_batch_utf8(engine, device_index, ps, n_pairs, n_pairs, SPAN_UNKNOWN, SPAN_UNKNOWN, (cudaStream_t)cuda_stream, err);
    }
    return run_device_batch(engine, device_index, ps, (cudaStream_t)cuda_stream, err);
}

extern "C" swb_status_t swb_compute_pairs_device_view(swb_engine_t *engine, int device_index, const swb_tape_view_t *a, const swb_tape_view_t *b,
                                                      void *out, swb_result_type_t out_type, void *cuda_stream, const char **err) {
    swb_status_t st = check_engine(engine, device_index, err);
    if (st != SWB_OK) return st;
    if (!a || !b) return fail(err, SWB_INVALID_ARG, "NULL view");
    if (a->n != b->n) return fail(err, SWB_INVALID_ARG, "tapes differ in length: %zu vs %zu", a->n, b->n);
    if (a->n == 0) return SWB_OK;
    if (!a->data || !a->offsets || !b->data || !b->offsets || !out) return fail(err, SWB_INVALID_ARG, "NULL tape or output pointer");
    if ((a->offset_bits != 64 && a->offset_bits != 32) || (b->offset_bits != 64 && b->offset_bits != 32)) return fail(err, SWB_INVALID_ARG, "offset_bits must be 64 or 32");
    if (out_type != SWB_RESULT_64 && out_type != SWB_RESULT_32 && out_type != SWB_RESULT_16) return fail(err, SWB_INVALID_ARG, "unknown result type %d", (int)out_type);
    SWB_CUDA(cudaSetDevice(engine->scope->devs[device_index].device));
    const uint32_t offs32 = (a->offset_bits == 32 ? 1u : 0u) | (b->offset_bits == 32 ? 2u : 0u);
    PairSpace ps = make_pair_space(a->data, a->offsets, b->data, b->offsets, a->n, 0, 0, out, (uint32_t)out_type, offs32);
    if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) {
        return run_device_batch_utf8(engine, device_index, ps, a->n, b->n, SPAN_UNKNOWN, SPAN_UNKNOWN, (cudaStream_t)cuda_stream, err);
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
    PairSpace ps = make_pair_space(q_data, q_offsets, c_data, c_offsets, (uint64_t)n_queries * n_candidates, n_candidates, row_stride, out, 0, 0);
    if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) {
        return run_device_batch_utf8(engine, device_index, ps, n_queries, n_candidates, SPAN_UNKNOWN, SPAN_UNKNOWN, (cudaStream_t)cuda_stream, err);
    }
    return run_device_batch(engine, device_index, ps, (cudaStream_t)cuda_stream, err);
}

// Pair-load phase on its own (schedule.cuh): one tile of pairs per CTA, one TMA bulk copy per tape span. (The persistent
// two-stage variant measured slower, profiles/README.md; it is only built with -DSWB_EXPERIMENTS, SWB_LOAD_VARIANT=0.)
static swb_status_t launch_load_phase(swb_engine *engine, DeviceCtx &ctx, const PairSpace &ps, uint32_t max_len, cudaStream_t stream, const char **err) {
    int variant = 1;
    uint32_t tile_override = 0;
#ifdef SWB_EXPERIMENTS
    static const int env_variant = getenv("SWB_LOAD_VARIANT") ? atoi(getenv("SWB_LOAD_VARIANT")) : 1;
    static const uint32_t env_tile = getenv("SWB_LOAD_TILE") ? (uint32_t)atoi(getenv("SWB_LOAD_TILE")) : 0u;
    variant = env_variant, tile_override = env_tile;
#endif
    const uint32_t stages = variant == 0 ? 2u : 1u;
    // pairs per tile: as many as the threads of a block (256) unless the staged spans overflow shared memory; short
    // tiles (long strings) keep several CTAs — several bulk copies — in flight per SM
    const uint32_t budget = (uint32_t)(ctx.max_smem_optin - 1024) / (2u * stages);
    uint32_t tile_pairs = tile_override ? std::min<uint32_t>(tile_override, LOAD_PAIRS) : (uint32_t)LOAD_PAIRS;
    const uint32_t per_cta_target = variant == 0 ? budget : std::min<uint32_t>(budget, 24u * 1024u); // >= 4 CTAs per SM
    while (tile_pairs > 1 && tile_pairs * max_len + 160u > per_cta_target) tile_pairs /= 2;
    const uint32_t cap = (round16(tile_pairs * max_len + 32u) + 127u) & ~127u;
    const size_t smem = 2 * stages * (size_t)cap;
    if (smem > (size_t)ctx.max_smem_optin - 1024) return fail(err, SWB_TOO_LONG, "load-phase kernel supports strings up to %u bytes", budget - 160u);
    const uint64_t tiles = (ps.n_pairs + tile_pairs - 1) / tile_pairs;
    if (variant == 0) {
        SWB_CUDA(allow_smem((const void *)load_phase_kernel, ctx.device, smem));
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
    PairSpace ps = make_pair_space(a_data, a_offsets, b_data, b_offsets, n_pairs, 0, 0, out, 0, 0);
    st = enqueue_schedule_count(ctx, ps, stream, err);
    if (st != SWB_OK) return st;
    if ((st = wait_schedule_count(ctx, err)) != SWB_OK) return st;
    if ((st = refuse_bad_offsets(ctx, err)) != SWB_OK) return st;
    uint32_t max_len = 0;
    for (int c = 0; c < SCHED_MCLASSES; ++c) max_len = std::max(max_len, ctx.h_stats->max_n[c]);
    return launch_load_phase(engine, ctx, ps, max_len, stream, err);
}

// ------------------------------------------------------------------------------------------ typed views

enum MemKind { MEM_HOST = 0, MEM_MANAGED = 1, MEM_DEVICE = 2 };

struct PtrInfo {
    MemKind kind = MEM_HOST;
    int device = -1;
};

static PtrInfo classify(const void *p) {
    PtrInfo info;
    cudaPointerAttributes attr;
    if (!p || cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return info;
    }
    if (attr.type == cudaMemoryTypeManaged) info.kind = MEM_MANAGED;
    else if (attr.type == cudaMemoryTypeDevice) info.kind = MEM_DEVICE, info.device = attr.device;
    return info; // unregistered and page-locked host memory are both staged
}

// One tape as a call sees it: a caller view (any memory kind) or a range of a resident tape.
struct Side {
    const uint8_t *data = nullptr;
    const uint8_t *offs = nullptr; // offsets array as bytes (element 0 = the view's first string)
    uint32_t width = 8;
    uint64_t n = 0;
    PtrInfo data_mem, offs_mem;
    const swb_tape *tape = nullptr;
    uint64_t first = 0; // resident tape: index of the view's first string
    bool host_offsets() const { return tape != nullptr || offs_mem.kind == MEM_HOST; }
    uint64_t off(uint64_t i) const { // only when host_offsets()
        if (tape) return tape->h_offs64[first + i];
        return width == 8 ? reinterpret_cast<const uint64_t *>(offs)[i] : (uint64_t) reinterpret_cast<const uint32_t *>(offs)[i];
    }
    uint32_t bit32() const { return width == 4 ? 1u : 0u; }
};

static swb_status_t side_from_view(const swb_tape_view_t *v, const char *name, Side *out, const char **err) {
    if (!v) return fail(err, SWB_INVALID_ARG, "%s view is NULL", name);
    if (v->offset_bits != 64 && v->offset_bits != 32) return fail(err, SWB_INVALID_ARG, "%s: offset_bits must be 64 or 32, got %d", name, v->offset_bits);
    if (!v->offsets || (!v->data && v->n)) return fail(err, SWB_INVALID_ARG, "%s: NULL tape pointer", name);
    Side s;
    s.data = v->data, s.offs = (const uint8_t *)v->offsets, s.width = v->offset_bits / 8, s.n = v->n;
    s.data_mem = classify(v->data), s.offs_mem = classify(v->offsets);
    if (s.data_mem.kind == MEM_HOST && s.offs_mem.kind != MEM_HOST)
        return fail(err, SWB_INVALID_ARG, "%s: a host byte buffer needs host offsets (they size the copies)", name);
    if (s.host_offsets() && s.n && s.off(s.n) < s.off(0))
        return fail(err, SWB_INVALID_ARG, "%s offsets are not monotonic (last %llu < first %llu)", name, (unsigned long long)s.off(s.n),
                    (unsigned long long)s.off(0));
    *out = s;
    return SWB_OK;
}

static swb_status_t side_from_tape(const swb_tape_t *t, size_t first, size_t n, const char *name, Side *out, const char **err) {
    if (!t) return fail(err, SWB_INVALID_ARG, "%s tape is NULL", name);
    if (!t->scope) return fail(err, SWB_INVALID_ARG, "%s tape: its scope has been destroyed", name);
    if (first > t->n || n > t->n - first) return fail(err, SWB_INVALID_ARG, "%s: strings [%zu, %zu) outside a tape of %llu", name, first, first + n, (unsigned long long)t->n);
    Side s;
    s.tape = t, s.first = first, s.n = n, s.width = t->offset_bits / 8;
    s.data_mem.kind = s.offs_mem.kind = MEM_DEVICE;
    *out = s;
    return SWB_OK;
}

// Device index of the scope entry that must run the whole call because a caller pointer is plain device memory
// (-1: none; -2: conflicting owners or an owner outside the scope).
static int forced_device(const swb_scope *scope, std::initializer_list<PtrInfo> infos) {
    int owner = -1;
    for (const PtrInfo &p : infos) {
        if (p.kind != MEM_DEVICE || p.device < 0) continue;
        if (owner >= 0 && owner != p.device) return -2;
        owner = p.device;
    }
    if (owner < 0) return -1;
    for (size_t i = 0; i < scope->devs.size(); ++i)
        if (scope->devs[i].device == owner) return (int)i;
    return -2;
}

// What one device sees of one side for strings [lo, hi] of the view: pointers the kernels can index with the ORIGINAL
// offsets. Host memory is copied (offsets now, bytes later in chunks); everything else is used in place.
struct SideOnDevice {
    const uint8_t *data_biased = nullptr; // data pointer such that data_biased + offset is the byte
    const uint8_t *offs_at_lo = nullptr;  // offsets array positioned at string lo
    uint8_t *copy_dst = nullptr;          // host data: device address of byte `lo_off` (chunk copies land relative to it)
    uint64_t lo_off = 0, hi_off = 0;      // host-readable offsets only
};

static swb_status_t place_side(DeviceCtx &ctx, int dev_index, const Side &s, uint64_t lo, uint64_t hi, DeviceBuffer &data_buf, DeviceBuffer &offs_buf,
                               SideOnDevice *out, const char **err) {
    SideOnDevice r;
    if (s.host_offsets()) r.lo_off = s.off(lo), r.hi_off = s.off(hi);
    if (s.tape) {
        r.data_biased = (const uint8_t *)s.tape->data[dev_index].ptr + (s.tape->base & 15u) - s.tape->base;
        r.offs_at_lo = (const uint8_t *)s.tape->offs[dev_index].ptr + (s.first + lo) * s.width;
        *out = r;
        return SWB_OK;
    }
    if (s.offs_mem.kind == MEM_HOST) {
        const size_t bytes = (hi - lo + 1) * s.width;
        SWB_CUDA(offs_buf.reserve(bytes));
        SWB_CUDA(cudaMemcpyAsync(offs_buf.ptr, s.offs + lo * s.width, bytes, cudaMemcpyHostToDevice, ctx.copy_stream));
        r.offs_at_lo = (const uint8_t *)offs_buf.ptr;
    } else {
        r.offs_at_lo = s.offs + lo * s.width;
    }
    if (s.data_mem.kind == MEM_HOST) {
        if (r.hi_off < r.lo_off) return fail(err, SWB_INVALID_ARG, "offsets are not monotonic (%llu after %llu)", (unsigned long long)r.hi_off, (unsigned long long)r.lo_off);
        SWB_CUDA(data_buf.reserve(r.hi_off - r.lo_off + 64));
        // keep (device address mod 16) == (tape offset mod 16) so aligned tapes stay aligned on the device
        r.copy_dst = (uint8_t *)data_buf.ptr + (r.lo_off & 15u);
        r.data_biased = r.copy_dst - r.lo_off;
    } else {
        r.data_biased = s.data;
    }
    *out = r;
    return SWB_OK;
}

static swb_status_t drain(swb_scope *scope) {
    for (auto &c : scope->devs) {
        cudaSetDevice(c.device);
        cudaStreamSynchronize(c.copy_stream);
        cudaStreamSynchronize(c.compute_stream);
    }
    return SWB_OK;
}

// Element-wise pairs over two sides of n_pairs strings each.
static swb_status_t host_pairs(swb_engine *engine, const Side &A, const Side &B, uint64_t n_pairs, void *out, uint32_t out_narrow, const char **err) {
    if (n_pairs == 0) return SWB_OK;
    if (!out) return fail(err, SWB_INVALID_ARG, "out is NULL");
    if (n_pairs >= 0xFFFFFFFFull) return fail(err, SWB_INVALID_ARG, "more than 2^32-1 pairs in one call");
    swb_scope *scope = engine->scope;
    const int D = (int)scope->devs.size();
    const size_t elt = result_bytes(out_narrow);
    const PtrInfo out_mem = classify(out);
    const int forced = forced_device(scope, {A.tape ? PtrInfo{} : A.data_mem, A.tape ? PtrInfo{} : A.offs_mem, B.tape ? PtrInfo{} : B.data_mem,
                                             B.tape ? PtrInfo{} : B.offs_mem, out_mem});
    if (forced == -2) return fail(err, SWB_INVALID_ARG, "device pointers must all belong to one device of the scope");
    const bool host_offs = A.host_offsets() && B.host_offsets();
    swb_status_t st;

    // Shard [0, n_pairs) over the devices: by bytes when the offsets are host-readable (bytes track cells for
    // similar-length pairs), by count otherwise (managed / device offsets are never touched from the host).
    std::vector<uint64_t> cut(D + 1, 0);
    if (forced >= 0) {
        for (int d = 0; d <= D; ++d) cut[d] = d > forced ? n_pairs : 0;
    } else if (host_offs) {
        cut[D] = n_pairs;
        const uint64_t a0 = A.off(0), b0 = B.off(0);
        const uint64_t total_bytes = (A.off(n_pairs) - a0) + (B.off(n_pairs) - b0);
        for (int d = 1; d < D; ++d) {
            const uint64_t target = total_bytes / D * d;
            uint64_t lo = cut[d - 1], hi = n_pairs;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) / 2;
                if ((A.off(mid) - a0) + (B.off(mid) - b0) < target) lo = mid + 1;
                else hi = mid;
            }
            cut[d] = lo;
        }
    } else {
        for (int d = 0; d <= D; ++d) cut[d] = n_pairs * d / D;
    }

    struct Shard {
        uint64_t lo = 0, hi = 0;
        PairSpace ps;
        SideOnDevice a, b;
        std::vector<uint64_t> chunk_cut;
        uint64_t a_bytes = 0, b_bytes = 0;
    };
    std::vector<Shard> shards(D);
    const bool staged_out = out_mem.kind == MEM_HOST;

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
        if ((st = place_side(ctx, d, A, sh.lo, sh.hi, ctx.a_data, ctx.a_offs, &sh.a, err)) != SWB_OK) return st;
        if ((st = place_side(ctx, d, B, sh.lo, sh.hi, ctx.b_data, ctx.b_offs, &sh.b, err)) != SWB_OK) return st;
        SWB_CUDA(cudaEventRecord(ctx.ev_offsets, ctx.copy_stream));
        void *d_out = (uint8_t *)out + sh.lo * elt;
        if (staged_out) {
            SWB_CUDA(ctx.out.reserve(n * elt));
            d_out = ctx.out.ptr;
        }
        sh.ps = make_pair_space(sh.a.data_biased, sh.a.offs_at_lo, sh.b.data_biased, sh.b.offs_at_lo, n, 0, 0, d_out, out_narrow, A.bit32() | (B.bit32() << 1));
        SWB_CUDA(cudaStreamWaitEvent(ctx.compute_stream, ctx.ev_offsets, 0));
        if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) {
            if (host_offs) sh.a_bytes = sh.a.hi_off - sh.a.lo_off, sh.b_bytes = sh.b.hi_off - sh.b.lo_off;
        } else if ((st = enqueue_schedule_count(ctx, sh.ps, ctx.compute_stream, err)) != SWB_OK) {
            return st;
        }
        // host bytes in ~16 MiB chunks so compute on chunk i overlaps the copy of chunk i+1 (measured 2/4/8/16/32 MiB:
        // 1.22/1.38/1.45/1.49/1.45 TCUPS e2e at 64x64)
        const bool a_host = sh.a.copy_dst != nullptr, b_host = sh.b.copy_dst != nullptr;
        const uint64_t host_bytes = (a_host ? sh.a.hi_off - sh.a.lo_off : 0) + (b_host ? sh.b.hi_off - sh.b.lo_off : 0);
        const int chunks = (int)std::min<uint64_t>(64, std::max<uint64_t>(1, host_bytes / (16ull << 20)));
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
            if (a_host) {
                const uint64_t c0 = A.off(p0), c1 = A.off(p1);
                if (c0 < sh.a.lo_off || c1 < c0 || c1 > sh.a.hi_off) return drain(scope), fail(err, SWB_INVALID_ARG, "a offsets are not monotonic near string %llu", (unsigned long long)p0);
                if (c1 > c0) SWB_CUDA(cudaMemcpyAsync(sh.a.copy_dst + (c0 - sh.a.lo_off), A.data + c0, c1 - c0, cudaMemcpyHostToDevice, ctx.copy_stream));
            }
            if (b_host) {
                const uint64_t c0 = B.off(p0), c1 = B.off(p1);
                if (c0 < sh.b.lo_off || c1 < c0 || c1 > sh.b.hi_off) return drain(scope), fail(err, SWB_INVALID_ARG, "b offsets are not monotonic near string %llu", (unsigned long long)p0);
                if (c1 > c0) SWB_CUDA(cudaMemcpyAsync(sh.b.copy_dst + (c0 - sh.b.lo_off), B.data + c0, c1 - c0, cudaMemcpyHostToDevice, ctx.copy_stream));
            }
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
        const int chunks = (int)sh.chunk_cut.size() - 1;
        uint8_t *h_out = (uint8_t *)out + sh.lo * elt;
        if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) {
            // decode needs the whole shard resident: no chunk pipelining on this path
            SWB_CUDA(cudaStreamWaitEvent(ctx.compute_stream, ctx.ev_chunks[chunks - 1], 0));
            if (!host_offs) {
                if ((st = device_spans(sh.a.offs_at_lo, A.width, n, &sh.a_bytes, sh.b.offs_at_lo, B.width, n, &sh.b_bytes, ctx.compute_stream, err)) != SWB_OK) return st;
            }
            if ((st = run_device_batch_utf8(engine, d, sh.ps, n, n, sh.a_bytes, sh.b_bytes, ctx.compute_stream, err)) != SWB_OK) return drain(scope), st;
            if (staged_out) SWB_CUDA(cudaMemcpyAsync(h_out, ctx.out.ptr, n * elt, cudaMemcpyDeviceToHost, ctx.compute_stream));
            continue;
        }
        if ((st = wait_schedule_count(ctx, err)) != SWB_OK) return st;
        if ((st = refuse_bad_offsets(ctx, err)) != SWB_OK) return drain(scope), st;
        if (trace_enabled() && d == 0) t_stats = now_ms();
        Plan plan = make_plan(*ctx.h_stats);
        if (plan.identity) {
            for (int c = 0; c < chunks; ++c) {
                const uint64_t p0 = sh.chunk_cut[c] - sh.lo, p1 = sh.chunk_cut[c + 1] - sh.lo;
                SWB_CUDA(cudaStreamWaitEvent(ctx.compute_stream, ctx.ev_chunks[c], 0));
                if ((st = launch_class(engine, ctx, d, sh.ps, p0, p1 - p0, plan.max_m[0], plan.max_n[0], ctx.compute_stream, err, plan.same_shape)) != SWB_OK) return drain(scope), st;
                if (staged_out) SWB_CUDA(cudaMemcpyAsync(h_out + p0 * elt, (uint8_t *)ctx.out.ptr + p0 * elt, (p1 - p0) * elt, cudaMemcpyDeviceToHost, ctx.compute_stream));
            }
        } else {
            SWB_CUDA(cudaStreamWaitEvent(ctx.compute_stream, ctx.ev_chunks[chunks - 1], 0));
            PairSpace ps = sh.ps;
            if ((st = enqueue_scatter(ctx, ps, ctx.compute_stream, err)) != SWB_OK) return st;
            for (int i = 0; i < plan.n_classes; ++i)
                if ((st = launch_class(engine, ctx, d, ps, plan.begin[i], plan.count[i], plan.max_m[i], plan.max_n[i], ctx.compute_stream, err)) != SWB_OK) return drain(scope), st;
            if (staged_out) SWB_CUDA(cudaMemcpyAsync(h_out, ctx.out.ptr, n * elt, cudaMemcpyDeviceToHost, ctx.compute_stream));
        }
    }
    if (trace_enabled()) t_phase2 = now_ms();
    // Phase 3: results are visible to the caller on return.
    for (int d = 0; d < D; ++d) {
        if (shards[d].hi == shards[d].lo) continue;
        SWB_CUDA(cudaSetDevice(scope->devs[d].device));
        SWB_CUDA(cudaStreamSynchronize(scope->devs[d].compute_stream));
    }
    if (trace_enabled())
        fprintf(stderr, "[swb trace] pairs=%llu enqueue=%.3f ms stats_wait=%.3f ms launch=%.3f ms drain=%.3f ms total=%.3f ms\n", (unsigned long long)n_pairs,
                t_phase1 - t_begin, t_stats - t_phase1, t_phase2 - t_stats, now_ms() - t_phase2, now_ms() - t_begin);
    return SWB_OK;
}

// Cross product: query rows sharded over the devices, every device sees all candidates.
static swb_status_t host_cross(swb_engine *engine, const Side &Q, const Side &C, void *out, size_t row_stride, uint32_t out_narrow, const char **err) {
    const uint64_t n_q = Q.n, n_c = C.n;
    if (n_q == 0 || n_c == 0) return SWB_OK;
    if (!out) return fail(err, SWB_INVALID_ARG, "out is NULL");
    if (row_stride < n_c) return fail(err, SWB_INVALID_ARG, "row_stride %zu < n_candidates %llu", row_stride, (unsigned long long)n_c);
    if (n_q * n_c >= 0xFFFFFFFFull) return fail(err, SWB_INVALID_ARG, "more than 2^32-1 pairs in one call");
    swb_scope *scope = engine->scope;
    const int D = (int)scope->devs.size();
    const size_t elt = result_bytes(out_narrow);
    const double t_begin = trace_enabled() ? now_ms() : 0;
    double t_placed = 0, t_launched = 0;
    const PtrInfo out_mem = classify(out);
    const int forced = forced_device(scope, {Q.tape ? PtrInfo{} : Q.data_mem, Q.tape ? PtrInfo{} : Q.offs_mem, C.tape ? PtrInfo{} : C.data_mem,
                                             C.tape ? PtrInfo{} : C.offs_mem, out_mem});
    if (forced == -2) return fail(err, SWB_INVALID_ARG, "device pointers must all belong to one device of the scope");
    const bool staged_out = out_mem.kind == MEM_HOST;
    swb_status_t st;
    bool seen_done[64] = {false};
    if (D > 64) return fail(err, SWB_INVALID_ARG, "scopes of more than 64 devices are not supported");
    for (int d = 0; d < D; ++d) {
        DeviceCtx &ctx = scope->devs[d];
        uint64_t q_lo = n_q * d / D, q_hi = n_q * (d + 1) / D;
        if (forced >= 0) q_lo = d == forced ? 0 : 0, q_hi = d == forced ? n_q : 0;
        if (q_hi == q_lo) continue;
        SWB_CUDA(cudaSetDevice(ctx.device));
        SideOnDevice q, c;
        if ((st = place_side(ctx, d, Q, q_lo, q_hi, ctx.a_data, ctx.a_offs, &q, err)) != SWB_OK) return st;
        if ((st = place_side(ctx, d, C, 0, n_c, ctx.b_data, ctx.b_offs, &c, err)) != SWB_OK) return st;
        if (q.copy_dst && q.hi_off > q.lo_off) SWB_CUDA(cudaMemcpyAsync(q.copy_dst, Q.data + q.lo_off, q.hi_off - q.lo_off, cudaMemcpyHostToDevice, ctx.copy_stream));
        if (c.copy_dst && c.hi_off > c.lo_off) SWB_CUDA(cudaMemcpyAsync(c.copy_dst, C.data + c.lo_off, c.hi_off - c.lo_off, cudaMemcpyHostToDevice, ctx.copy_stream));
        if (q.copy_dst || c.copy_dst || Q.offs_mem.kind == MEM_HOST || C.offs_mem.kind == MEM_HOST) { // something was copied
            SWB_CUDA(cudaEventRecord(ctx.ev_offsets, ctx.copy_stream));
            SWB_CUDA(cudaStreamWaitEvent(ctx.compute_stream, ctx.ev_offsets, 0));
        }
        // Host matrices get a dense device staging matrix and one 2-D copy (the copy engine moves 52 GB/s; kernels
        // writing a mapped host matrix directly were measured at 23.5 GB/s). Managed / device matrices are written in place.
        void *d_out = (uint8_t *)out + q_lo * row_stride * elt;
        uint64_t stride = row_stride;
        if (staged_out) {
            SWB_CUDA(ctx.out.reserve((q_hi - q_lo) * n_c * elt));
            d_out = ctx.out.ptr, stride = n_c;
        }
        PairSpace ps = make_pair_space(q.data_biased, q.offs_at_lo, c.data_biased, c.offs_at_lo, (q_hi - q_lo) * n_c, n_c, stride, d_out, out_narrow,
                                       Q.bit32() | (C.bit32() << 1));
        if (trace_enabled()) t_placed = now_ms();
        if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) {
            uint64_t q_bytes = q.hi_off - q.lo_off, c_bytes = c.hi_off - c.lo_off;
            if (!Q.host_offsets()) q_bytes = SPAN_UNKNOWN;
            if (!C.host_offsets()) c_bytes = SPAN_UNKNOWN;
            if ((st = run_device_batch_utf8(engine, d, ps, q_hi - q_lo, n_c, q_bytes, c_bytes, ctx.compute_stream, err)) != SWB_OK) return drain(scope), st;
        } else if ((st = run_device_batch(engine, d, ps, ctx.compute_stream, err, &seen_done[d])) != SWB_OK) {
            return drain(scope), st;
        }
        if (staged_out)
            SWB_CUDA(cudaMemcpy2DAsync((uint8_t *)out + q_lo * row_stride * elt, row_stride * elt, ctx.out.ptr, n_c * elt, n_c * elt, q_hi - q_lo,
                                       cudaMemcpyDeviceToHost, ctx.compute_stream));
    }
    if (trace_enabled()) t_launched = now_ms();
    for (int d = 0; d < D; ++d) {
        // results written in place by a kernel whose completion report has been read are final and visible: no synchronisation
        if (seen_done[d] && !staged_out && (out_mem.kind == MEM_DEVICE || scope->devs[d].concurrent_managed)) continue;
        SWB_CUDA(cudaSetDevice(scope->devs[d].device));
        SWB_CUDA(cudaStreamSynchronize(scope->devs[d].compute_stream));
    }
    if (trace_enabled())
        fprintf(stderr, "[swb trace] cross %llu x %llu: classify+place=%.1f us run=%.1f us sync=%.1f us total=%.1f us\n", (unsigned long long)n_q,
                (unsigned long long)n_c, (t_placed - t_begin) * 1e3, (t_launched - t_placed) * 1e3, (now_ms() - t_launched) * 1e3, (now_ms() - t_begin) * 1e3);
    return SWB_OK;
}

static swb_status_t check_result_type(swb_result_type_t t, const char **err) {
    if (t != SWB_RESULT_64 && t != SWB_RESULT_32 && t != SWB_RESULT_16) return fail(err, SWB_INVALID_ARG, "unknown result type %d", (int)t);
    return SWB_OK;
}

extern "C" swb_status_t swb_compute_pairs_view(swb_engine_t *engine, const swb_tape_view_t *a, const swb_tape_view_t *b, void *out,
                                               swb_result_type_t out_type, const char **err) {
    swb_status_t st = check_engine_alive(engine, err);
    if (st != SWB_OK) return st;
    if ((st = check_result_type(out_type, err)) != SWB_OK) return st;
    Side A, B;
    if ((st = side_from_view(a, "a", &A, err)) != SWB_OK) return st;
    if ((st = side_from_view(b, "b", &B, err)) != SWB_OK) return st;
    if (A.n != B.n) return fail(err, SWB_INVALID_ARG, "tapes differ in length: %llu vs %llu", (unsigned long long)A.n, (unsigned long long)B.n);
    return host_pairs(engine, A, B, A.n, out, (uint32_t)out_type, err);
}

extern "C" swb_status_t swb_compute_cross_view(swb_engine_t *engine, const swb_tape_view_t *queries, const swb_tape_view_t *candidates,
                                               void *out, size_t row_stride, swb_result_type_t out_type, const char **err) {
    swb_status_t st = check_engine_alive(engine, err);
    if (st != SWB_OK) return st;
    if ((st = check_result_type(out_type, err)) != SWB_OK) return st;
    Side Q, C;
    if ((st = side_from_view(queries, "queries", &Q, err)) != SWB_OK) return st;
    if ((st = side_from_view(candidates, "candidates", &C, err)) != SWB_OK) return st;
    return host_cross(engine, Q, C, out, row_stride, (uint32_t)out_type, err);
}

extern "C" swb_status_t swb_compute_pairs(swb_engine_t *engine, const uint8_t *a_data, const uint64_t *a_offsets,
                                          const uint8_t *b_data, const uint64_t *b_offsets, size_t n_pairs, uint64_t *out,
                                          const char **err) {
    if (!engine) return fail(err, SWB_INVALID_ARG, "engine is NULL");
    if (n_pairs == 0) return SWB_OK;
    if (!a_data || !a_offsets || !b_data || !b_offsets || !out) return fail(err, SWB_INVALID_ARG, "NULL tape or output pointer");
    const swb_tape_view_t a{a_data, a_offsets, n_pairs, 64}, b{b_data, b_offsets, n_pairs, 64};
    return swb_compute_pairs_view(engine, &a, &b, out, SWB_RESULT_64, err);
}

extern "C" swb_status_t swb_compute_cross(swb_engine_t *engine, const uint8_t *q_data, const uint64_t *q_offsets,
                                          size_t n_queries, const uint8_t *c_data, const uint64_t *c_offsets,
                                          size_t n_candidates, uint64_t *out, size_t row_stride, const char **err) {
    if (!engine) return fail(err, SWB_INVALID_ARG, "engine is NULL");
    if (n_queries == 0 || n_candidates == 0) return SWB_OK;
    if (!q_data || !q_offsets || !c_data || !c_offsets || !out) return fail(err, SWB_INVALID_ARG, "NULL tape or output pointer");
    const swb_tape_view_t q{q_data, q_offsets, n_queries, 64}, c{c_data, c_offsets, n_candidates, 64};
    return swb_compute_cross_view(engine, &q, &c, out, row_stride, SWB_RESULT_64, err);
}

extern "C" uint64_t swb_levenshtein_pair(swb_engine_t *engine, const uint8_t *a, size_t a_length, const uint8_t *b,
                                         size_t b_length, const char **err) {
    const uint8_t dummy = 0;
    const uint64_t a_offs[2] = {0, a_length}, b_offs[2] = {0, b_length};
    uint64_t out = UINT64_MAX;
    if (swb_compute_pairs(engine, a ? a : &dummy, a_offs, b ? b : &dummy, b_offs, 1, &out, err) != SWB_OK) return UINT64_MAX;
    return out;
}

// ------------------------------------------------------------------------------------------ resident tapes

extern "C" swb_status_t swb_tape_upload(swb_scope_t *scope, const swb_tape_view_t *view, swb_tape_t **out, const char **err) {
    if (!out) return fail(err, SWB_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!scope) return fail(err, SWB_INVALID_ARG, "scope is NULL");
    Side s;
    swb_status_t st = side_from_view(view, "tape", &s, err);
    if (st != SWB_OK) return st;
    swb_tape *t = new swb_tape();
    t->scope = scope, t->n = s.n, t->offset_bits = view->offset_bits;
    // host copy of the offsets (widened): the cuts of every later call come from here, never from the device
    t->h_offs64.resize(s.n + 1);
    {
        std::vector<uint8_t> raw;
        const uint8_t *src = s.offs;
        if (s.offs_mem.kind != MEM_HOST) {
            raw.resize((s.n + 1) * s.width);
            cudaError_t e = cudaMemcpy(raw.data(), s.offs, raw.size(), cudaMemcpyDefault);
            if (e != cudaSuccess) {
                delete t;
                return fail(err, SWB_CUDA_ERROR, "reading the tape's offsets failed: %s", cudaGetErrorString(e));
            }
            src = raw.data();
        }
        for (uint64_t i = 0; i <= s.n; ++i)
            t->h_offs64[i] = s.width == 8 ? reinterpret_cast<const uint64_t *>(src)[i] : (uint64_t) reinterpret_cast<const uint32_t *>(src)[i];
    }
    for (uint64_t i = 0; i < s.n; ++i)
        if (t->h_offs64[i + 1] < t->h_offs64[i] || t->h_offs64[i + 1] - t->h_offs64[i] >= (1ull << 32)) {
            delete t;
            return fail(err, SWB_INVALID_ARG, "tape offsets are not monotonic at string %llu", (unsigned long long)i);
        }
    t->base = t->h_offs64[0], t->bytes = t->h_offs64[s.n] - t->base;
    const size_t D = scope->devs.size();
    t->data.resize(D), t->offs.resize(D);
    scope->tapes.insert(t);
    for (size_t d = 0; d < D; ++d) {
        cudaError_t e = cudaSetDevice(scope->devs[d].device);
        if (e == cudaSuccess) e = t->data[d].reserve(t->bytes + 64);
        if (e == cudaSuccess) e = t->offs[d].reserve((s.n + 1) * s.width);
        if (e == cudaSuccess && t->bytes) e = cudaMemcpy((uint8_t *)t->data[d].ptr + (t->base & 15u), s.data + t->base, t->bytes, cudaMemcpyDefault);
        if (e == cudaSuccess) e = cudaMemcpy(t->offs[d].ptr, s.offs, (s.n + 1) * s.width, cudaMemcpyDefault);
        if (e != cudaSuccess) {
            swb_tape_destroy(t);
            return fail(err, e == cudaErrorMemoryAllocation ? SWB_BAD_ALLOC : SWB_CUDA_ERROR, "tape upload to device %d failed: %s", scope->devs[d].device,
                        cudaGetErrorString(e));
        }
    }
    *out = t;
    return SWB_OK;
}

extern "C" void swb_tape_destroy(swb_tape_t *tape) {
    if (!tape) return;
    if (tape->scope) {
        for (size_t d = 0; d < tape->data.size(); ++d) {
            cudaSetDevice(tape->scope->devs[d].device);
            tape->data[d].release();
            tape->offs[d].release();
        }
        tape->scope->tapes.erase(tape);
    }
    delete tape;
}

extern "C" size_t swb_tape_len(const swb_tape_t *tape) { return tape ? (size_t)tape->n : 0; }

extern "C" swb_status_t swb_tape_device_view(const swb_tape_t *tape, int device_index, size_t first, size_t n, swb_tape_view_t *out, const char **err) {
    if (!tape || !out) return fail(err, SWB_INVALID_ARG, "NULL argument");
    if (!tape->scope) return fail(err, SWB_INVALID_ARG, "the tape's scope has been destroyed");
    if (device_index < 0 || device_index >= (int)tape->data.size()) return fail(err, SWB_INVALID_ARG, "device_index %d outside the scope", device_index);
    if (first > tape->n || n > tape->n - first) return fail(err, SWB_INVALID_ARG, "strings [%zu, %zu) outside a tape of %llu", first, first + n, (unsigned long long)tape->n);
    out->data = (const uint8_t *)tape->data[device_index].ptr + (tape->base & 15u) - tape->base;
    out->offsets = (const uint8_t *)tape->offs[device_index].ptr + first * (tape->offset_bits / 8);
    out->n = n, out->offset_bits = tape->offset_bits;
    return SWB_OK;
}

extern "C" swb_status_t swb_compute_pairs_tape(swb_engine_t *engine, const swb_tape_t *a, size_t a_first, const swb_tape_t *b, size_t b_first,
                                               size_t n_pairs, void *out, swb_result_type_t out_type, const char **err) {
    swb_status_t st = check_engine_alive(engine, err);
    if (st != SWB_OK) return st;
    if ((st = check_result_type(out_type, err)) != SWB_OK) return st;
    Side A, B;
    if ((st = side_from_tape(a, a_first, n_pairs, "a", &A, err)) != SWB_OK) return st;
    if ((st = side_from_tape(b, b_first, n_pairs, "b", &B, err)) != SWB_OK) return st;
    if (a->scope != engine->scope || b->scope != engine->scope) return fail(err, SWB_INVALID_ARG, "tapes and engine belong to different scopes");
    return host_pairs(engine, A, B, n_pairs, out, (uint32_t)out_type, err);
}

extern "C" swb_status_t swb_compute_cross_tape(swb_engine_t *engine, const swb_tape_t *queries, size_t q_first, size_t n_queries,
                                               const swb_tape_t *candidates, size_t c_first, size_t n_candidates, void *out, size_t row_stride,
                                               swb_result_type_t out_type, const char **err) {
    swb_status_t st = check_engine_alive(engine, err);
    if (st != SWB_OK) return st;
    if ((st = check_result_type(out_type, err)) != SWB_OK) return st;
    Side Q, C;
    if ((st = side_from_tape(queries, q_first, n_queries, "queries", &Q, err)) != SWB_OK) return st;
    if ((st = side_from_tape(candidates, c_first, n_candidates, "candidates", &C, err)) != SWB_OK) return st;
    if (queries->scope != engine->scope || candidates->scope != engine->scope) return fail(err, SWB_INVALID_ARG, "tapes and engine belong to different scopes");
    return host_cross(engine, Q, C, out, row_stride, (uint32_t)out_type, err);
}
