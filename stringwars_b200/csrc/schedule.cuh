// schedule.cuh — length classification / bucketing of a pair batch, and the load-phase-only kernel.
//
// Mixed-length batches (BASELINE.json configs 1 and 3) would make the lanes of a warp run for max(n) columns
// with the widest pattern's word count; the schedule kernel keys every pair by (pattern-length class,
// text-length bin), counts the keys and scatters pair indices into a permutation so that every compute kernel
// launch sees one length class and neighbouring threads see similar text lengths. Uniform batches (configs 2,
// 4, 5) are detected from the same counters and keep the identity order (no scatter, no indirection).
#pragma once
#include "common.cuh"

namespace swb {

constexpr int SCHED_MCLASSES = 9;  // m <= 32, 64, 128, 256, 512, 1024, 2048, 4096, > 4096
constexpr int SCHED_NBINS = 64;    // text length in steps of 32 bytes, last bin open-ended
constexpr int SCHED_KEYS = SCHED_MCLASSES * SCHED_NBINS;

struct ScheduleStats {
    uint32_t counts[SCHED_KEYS];
    uint32_t cursor[SCHED_KEYS]; // scatter cursors (exclusive prefix of counts), written by the host-side scan
    uint32_t max_m[SCHED_MCLASSES];
    uint32_t max_n[SCHED_MCLASSES];
    uint32_t min_m, min_n;
    uint32_t bad;     // strings whose offsets run backwards or span 2^32 bytes or more: the call fails with SWB_INVALID_ARG
    uint32_t extra_max; // longest string of the extra offsets array (the candidates of a cross product)
    uint32_t ticket;  // blocks that have merged their counters (device copy only)
    uint32_t seq;     // host copy only: sequence number of the call whose counters are complete
};

__device__ __forceinline__ int mclass_of(uint32_t m) {
    if (m <= 32u) return 0;
    const int c = 32 - __clz((int)(m - 1u)) - 5; // 33..64 -> 1, 65..128 -> 2, ...
    return c > SCHED_MCLASSES - 1 ? SCHED_MCLASSES - 1 : c;
}
__device__ __forceinline__ int key_of(uint32_t m, uint32_t n) {
    uint32_t bin = (n + 31u) >> 5;
    if (bin > (uint32_t)SCHED_NBINS - 1u) bin = SCHED_NBINS - 1;
    return mclass_of(m) * SCHED_NBINS + (int)bin;
}

#ifdef SWB_SCHEDULE_KERNELS // plain (non-template) kernels: defined in exactly one translation unit, pipeline.cu

// `stats` (device) must be all-zero with min_* = 0xFFFFFFFF on entry; the LAST block to finish publishes the
// merged counters into `published` (mapped, page-locked host memory: no D2H copy, no event — the host polls
// `published->seq`) and restores `stats` to its entry state for the next call.
//
// The kernel is also the offsets validator the ABI promises (SWB_INVALID_ARG for non-monotonic offsets): it reads every
// offset of both tapes anyway, so a string whose end lies before its start (or 4 GiB or more after it) is counted in
// `bad` instead of being keyed, and the host refuses the batch before any compute kernel sees it. `extra_offs`
// (optional) is a second offsets array of extra_n + 1 entries checked the same way — the candidate tape of a cross
// product whose statistics are taken over the queries alone.
__global__ void __launch_bounds__(256) schedule_count_kernel(PairSpace ps, ScheduleStats *stats, ScheduleStats *published, uint32_t seq,
                                                             const void *extra_offs, uint64_t extra_n, uint32_t extra32) {
    __shared__ uint32_t s_counts[SCHED_KEYS];
    __shared__ uint32_t s_max_m[SCHED_MCLASSES], s_max_n[SCHED_MCLASSES];
    __shared__ uint32_t s_min_m, s_min_n, s_bad, s_extra_max;
    __shared__ bool s_last;
    for (int i = threadIdx.x; i < SCHED_KEYS; i += blockDim.x) s_counts[i] = 0;
    if (threadIdx.x < SCHED_MCLASSES) s_max_m[threadIdx.x] = 0, s_max_n[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_min_m = 0xFFFFFFFFu, s_min_n = 0xFFFFFFFFu, s_bad = 0u, s_extra_max = 0u;
    __syncthreads();
    // Threads keep a run (key, count, extrema) in registers and touch shared memory only when the key changes;
    // a warp whose lanes all ended on the same key merges with warp reductions and ONE set of atomics. Uniform
    // batches (every pair the same key) would otherwise serialise a million atomics on a single counter.
    int run_key = -1;
    uint32_t bad = 0, extra_max = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < extra_n && extra_offs; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t o0 = extra32 ? (uint64_t)reinterpret_cast<const uint32_t *>(extra_offs)[i] : reinterpret_cast<const uint64_t *>(extra_offs)[i];
        const uint64_t o1 = extra32 ? (uint64_t)reinterpret_cast<const uint32_t *>(extra_offs)[i + 1] : reinterpret_cast<const uint64_t *>(extra_offs)[i + 1];
        if (o1 < o0 || ((o1 - o0) >> 32) != 0u) ++bad;
        else extra_max = max(extra_max, (uint32_t)(o1 - o0));
    }
    uint32_t run_count = 0, run_max_m = 0, run_max_n = 0, run_min_m = 0xFFFFFFFFu, run_min_n = 0xFFFFFFFFu;
    auto flush = [&](int key, uint32_t count, uint32_t mx_m, uint32_t mx_n, uint32_t mn_m, uint32_t mn_n) {
        if (key < 0) return;
        const int c = key / SCHED_NBINS;
        atomicAdd(&s_counts[key], count);
        atomicMax(&s_max_m[c], mx_m);
        atomicMax(&s_max_n[c], mx_n);
        atomicMin(&s_min_m, mn_m);
        atomicMin(&s_min_n, mn_n);
    };
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < ps.n_pairs; k += (uint64_t)gridDim.x * blockDim.x) {
        PairSpace id = ps;
        id.order = nullptr;
        const PairRequest rq = pair_request(id, k);
        if (rq.a1 < rq.a0 || rq.b1 < rq.b0 || ((rq.a1 - rq.a0) >> 32) != 0u || ((rq.b1 - rq.b0) >> 32) != 0u) {
            ++bad;
            continue;
        }
        const PairView pv = pair_resolve(id, rq);
        const uint32_t m = min(pv.la, pv.lb), n = max(pv.la, pv.lb);
        const int key = key_of(m, n);
        if (key != run_key) {
            flush(run_key, run_count, run_max_m, run_max_n, run_min_m, run_min_n);
            run_key = key, run_count = 0, run_max_m = 0, run_max_n = 0, run_min_m = 0xFFFFFFFFu, run_min_n = 0xFFFFFFFFu;
        }
        run_count += 1;
        run_max_m = max(run_max_m, m), run_max_n = max(run_max_n, n);
        run_min_m = min(run_min_m, m), run_min_n = min(run_min_n, n);
    }
    {
        const int key0 = __shfl_sync(0xffffffffu, run_key, 0);
        if (__all_sync(0xffffffffu, run_key == key0)) {
            const uint32_t count = __reduce_add_sync(0xffffffffu, run_count);
            const uint32_t mx_m = __reduce_max_sync(0xffffffffu, run_max_m), mx_n = __reduce_max_sync(0xffffffffu, run_max_n);
            const uint32_t mn_m = __reduce_min_sync(0xffffffffu, run_min_m), mn_n = __reduce_min_sync(0xffffffffu, run_min_n);
            if ((threadIdx.x & 31) == 0) flush(key0, count, mx_m, mx_n, mn_m, mn_n);
        } else {
            flush(run_key, run_count, run_max_m, run_max_n, run_min_m, run_min_n);
        }
    }
    if (bad) atomicAdd(&s_bad, bad);
    if (extra_max) atomicMax(&s_extra_max, extra_max);
    __syncthreads();
    for (int i = threadIdx.x; i < SCHED_KEYS; i += blockDim.x)
        if (s_counts[i]) atomicAdd(&stats->counts[i], s_counts[i]);
    if (threadIdx.x < SCHED_MCLASSES) {
        if (s_max_m[threadIdx.x]) atomicMax(&stats->max_m[threadIdx.x], s_max_m[threadIdx.x]);
        if (s_max_n[threadIdx.x]) atomicMax(&stats->max_n[threadIdx.x], s_max_n[threadIdx.x]);
    }
    if (threadIdx.x == 0) {
        atomicMin(&stats->min_m, s_min_m);
        atomicMin(&stats->min_n, s_min_n);
        if (s_bad) atomicAdd(&stats->bad, s_bad);
        if (s_extra_max) atomicMax(&stats->extra_max, s_extra_max);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&stats->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    volatile ScheduleStats *vs = stats;
    for (int i = threadIdx.x; i < SCHED_KEYS; i += blockDim.x) {
        published->counts[i] = vs->counts[i];
        stats->counts[i] = 0;
    }
    if (threadIdx.x < SCHED_MCLASSES) {
        published->max_m[threadIdx.x] = vs->max_m[threadIdx.x];
        published->max_n[threadIdx.x] = vs->max_n[threadIdx.x];
        stats->max_m[threadIdx.x] = 0, stats->max_n[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) {
        published->min_m = vs->min_m, published->min_n = vs->min_n;
        published->bad = vs->bad, published->extra_max = vs->extra_max;
        stats->min_m = 0xFFFFFFFFu, stats->min_n = 0xFFFFFFFFu, stats->bad = 0, stats->extra_max = 0, stats->ticket = 0;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile uint32_t *>(&published->seq) = seq;
        __threadfence_system();
    }
}

__global__ void __launch_bounds__(256) schedule_scatter_kernel(PairSpace ps, ScheduleStats *stats, uint32_t *order) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < ps.n_pairs; k += (uint64_t)gridDim.x * blockDim.x) {
        PairSpace id = ps;
        id.order = nullptr;
        const PairView pv = pair_at(id, k);
        const uint32_t m = min(pv.la, pv.lb), n = max(pv.la, pv.lb);
        const uint32_t pos = atomicAdd(&stats->cursor[key_of(m, n)], 1u);
        order[pos] = (uint32_t)k;
    }
}

// Pair-load phase on its own: a block stages the bytes of LOAD_PAIRS consecutive pairs (both tapes) into
// shared memory — one TMA bulk copy per tape span, completion on an mbarrier — then thread p byte-sums pair p
// out of shared memory. Moves exactly the algorithmic bytes of the compute kernels' load phase (+ offsets in,
// 8 B/pair out).
constexpr int LOAD_THREADS = 256;
constexpr int LOAD_PAIRS = 256;

__device__ __forceinline__ uint32_t masked_sum16(const uint8_t *piece, int lo, int hi) {
    // sum of bytes [lo, hi) of the 16-byte shared-memory piece (ragged head / tail of a string)
    const uint4 v = *reinterpret_cast<const uint4 *>(piece);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t acc = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t x = w[q];
        const int b0 = lo - 4 * q, b1 = hi - 4 * q; // keep bytes [b0, b1) of this word
        if (b1 <= 0 || b0 >= 4) continue;
        if (b0 > 0) x &= 0xFFFFFFFFu << (8 * b0);
        if (b1 < 4) x &= 0xFFFFFFFFu >> (8 * (4 - b1));
        acc = __dp4a(x, 0x01010101u, acc);
    }
    return acc;
}

__device__ __forceinline__ uint32_t whole_sum16(const uint8_t *piece) {
    const uint4 v = *reinterpret_cast<const uint4 *>(piece);
    uint32_t acc = __dp4a(v.x, 0x01010101u, 0u);
    acc = __dp4a(v.y, 0x01010101u, acc);
    acc = __dp4a(v.z, 0x01010101u, acc);
    return __dp4a(v.w, 0x01010101u, acc);
}

// Byte sum of the string occupying bytes [s, e) of a staged span: only a ragged head or tail piece pays for
// masks, every whole piece is one LDS.128 + four DP4A (the first version masked every piece and made this
// measurement kernel instruction-bound: issue 81 %, profiles/r1_c_load_phase_ascii64.txt).
__device__ __forceinline__ uint32_t span_sum(const uint8_t *span, uint32_t s, uint32_t e) {
    if (e <= s) return 0u;
    uint32_t acc = 0;
    uint32_t q = s >> 4;
    const uint32_t q_end = e >> 4; // pieces [q, q_end) lie wholly inside once the head is handled
    if (s & 15u) {
        acc = masked_sum16(span + 16u * q, (int)(s & 15u), (int)min(16u, e - 16u * q));
        ++q;
    }
    for (; q < q_end; ++q) acc += whole_sum16(span + 16u * q);
    if ((e & 15u) && q == q_end) acc += masked_sum16(span + 16u * q, 0, (int)(e & 15u));
    return acc;
}

// 1-D bulk async copy global -> shared (TMA, `cp.async.bulk`, SASS UBLKCP) completing on an mbarrier.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}

struct LoadTileMeta {
    uint32_t mis_a, mis_b; // misalignment of the tile's first byte inside its first 16-byte piece
    uint64_t a_lo, b_lo;   // tape offsets of the tile's first strings
};

// Persistent CTAs, two shared-memory stages: while the threads byte-sum tile i out of stage i&1, the TMA engine
// is already filling stage (i+1)&1 with tile i+gridDim.x. Per-pair offsets are fetched before waiting on the
// mbarrier so their latency hides behind the bulk copy.
__global__ void __launch_bounds__(LOAD_THREADS) load_phase_kernel(PairSpace ps, uint32_t cap_a, uint32_t cap_b, uint32_t tile_pairs) {
    extern __shared__ __align__(128) uint8_t lsm[];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ LoadTileMeta meta[2];
    const uint32_t stage_bytes = cap_a + cap_b;
    const uint64_t tiles = (ps.n_pairs + tile_pairs - 1) / tile_pairs; // tile_pairs <= LOAD_THREADS pairs per tile
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto issue = [&](uint64_t tile, int stage) { // thread 0 only
        const uint64_t p0 = tile * tile_pairs, p1 = min(p0 + (uint64_t)tile_pairs, ps.n_pairs);
        const uint64_t a_lo = a_off(ps, p0), a_hi = a_off(ps, p1), b_lo = b_off(ps, p0), b_hi = b_off(ps, p1);
        const uint8_t *ga = ps.a_data + a_lo, *gb = ps.b_data + b_lo;
        const uint32_t mis_a = (uint32_t)(reinterpret_cast<uintptr_t>(ga) & 15u), mis_b = (uint32_t)(reinterpret_cast<uintptr_t>(gb) & 15u);
        // whole 16-byte pieces covering each span: exactly what the compute kernels' staging loads touch
        const uint32_t bytes_a = a_hi > a_lo ? ((mis_a + (uint32_t)(a_hi - a_lo) + 15u) & ~15u) : 0u;
        const uint32_t bytes_b = b_hi > b_lo ? ((mis_b + (uint32_t)(b_hi - b_lo) + 15u) & ~15u) : 0u;
        meta[stage] = LoadTileMeta{mis_a, mis_b, a_lo, b_lo};
        uint8_t *s_a = lsm + (size_t)stage * stage_bytes, *s_b = s_a + cap_a;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[stage])), "r"(bytes_a + bytes_b) : "memory");
        if (bytes_a) bulk_g2s(s_a, ga - mis_a, bytes_a, &mbar[stage]);
        if (bytes_b) bulk_g2s(s_b, gb - mis_b, bytes_b, &mbar[stage]);
    };
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < tiles) issue(blockIdx.x, 0);
    __syncthreads();
    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int stage = it & 1;
        if (threadIdx.x == 0 && tile + gridDim.x < tiles) issue(tile + gridDim.x, stage ^ 1);
        const uint64_t p0 = tile * tile_pairs, p1 = min(p0 + (uint64_t)tile_pairs, ps.n_pairs);
        const uint64_t p = p0 + threadIdx.x;
        uint64_t oa0 = 0, oa1 = 0, ob0 = 0, ob1 = 0;
        if (p < p1) oa0 = a_off(ps, p), oa1 = a_off(ps, p + 1), ob0 = b_off(ps, p), ob1 = b_off(ps, p + 1);
        {
            uint32_t done = 0;
            const uint32_t parity = (it >> 1) & 1u;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_u32(&mbar[stage])), "r"(parity) : "memory");
        }
        if (p < p1) {
            const LoadTileMeta mt = meta[stage];
            const uint8_t *s_a = lsm + (size_t)stage * stage_bytes, *s_b = s_a + cap_a;
            uint32_t acc = 0;
            acc += span_sum(s_a, mt.mis_a + (uint32_t)(oa0 - mt.a_lo), mt.mis_a + (uint32_t)(oa1 - mt.a_lo));
            acc += span_sum(s_b, mt.mis_b + (uint32_t)(ob0 - mt.b_lo), mt.mis_b + (uint32_t)(ob1 - mt.b_lo));
            ps.out[p] = acc;
        }
        __syncthreads(); // stage `stage` may be refilled (and its meta rewritten) from the next iteration on
    }
}

// Variant with one tile per CTA (no persistence, single stage): latency is hidden by CTA-level parallelism
// instead of a software pipeline. Selected with SWB_LOAD_VARIANT=1 for A/B measurements.
__global__ void __launch_bounds__(LOAD_THREADS) load_phase_simple_kernel(PairSpace ps, uint32_t cap_a, uint32_t cap_b, uint32_t tile_pairs) {
    extern __shared__ __align__(128) uint8_t lsm[];
    __shared__ __align__(8) uint64_t mbar;
    uint8_t *s_a = lsm, *s_b = lsm + cap_a;
    const uint64_t p0 = (uint64_t)blockIdx.x * tile_pairs;
    const uint64_t p1 = min(p0 + (uint64_t)tile_pairs, ps.n_pairs);
    const uint64_t a_lo = a_off(ps, p0), a_hi = a_off(ps, p1), b_lo = b_off(ps, p0), b_hi = b_off(ps, p1);
    const uint8_t *ga = ps.a_data + a_lo, *gb = ps.b_data + b_lo;
    const uint32_t mis_a = (uint32_t)(reinterpret_cast<uintptr_t>(ga) & 15u), mis_b = (uint32_t)(reinterpret_cast<uintptr_t>(gb) & 15u);
    const uint32_t bytes_a = a_hi > a_lo ? ((mis_a + (uint32_t)(a_hi - a_lo) + 15u) & ~15u) : 0u;
    const uint32_t bytes_b = b_hi > b_lo ? ((mis_b + (uint32_t)(b_hi - b_lo) + 15u) & ~15u) : 0u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes_a + bytes_b) : "memory");
        if (bytes_a) bulk_g2s(s_a, ga - mis_a, bytes_a, &mbar);
        if (bytes_b) bulk_g2s(s_b, gb - mis_b, bytes_b, &mbar);
    }
    // A tile of fewer pairs than threads (long strings) gives each pair TPP = LOAD_THREADS / tile_pairs threads: every
    // thread byte-sums its slice of the pair's two spans and the group adds up with shuffles, so the time a block spends
    // reading shared memory does not grow with the string length (256-byte strings: 4 threads per pair).
    const uint32_t tpp = max(1u, (uint32_t)LOAD_THREADS / tile_pairs) > 32u ? 32u : max(1u, (uint32_t)LOAD_THREADS / tile_pairs);
    const uint32_t sub = threadIdx.x % tpp;
    const uint64_t p = p0 + threadIdx.x / tpp;
    const bool mine = threadIdx.x / tpp < tile_pairs && p < p1;
    uint64_t oa0 = 0, oa1 = 0, ob0 = 0, ob1 = 0;
    if (mine) oa0 = a_off(ps, p), oa1 = a_off(ps, p + 1), ob0 = b_off(ps, p), ob1 = b_off(ps, p + 1);
    __syncthreads(); // mbarrier init visible to every thread before anyone waits on it
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    }
    uint32_t acc = 0;
    if (mine) {
        const uint32_t sa = mis_a + (uint32_t)(oa0 - a_lo), la = (uint32_t)(oa1 - oa0), sb = mis_b + (uint32_t)(ob0 - b_lo), lb = (uint32_t)(ob1 - ob0);
        acc += span_sum(s_a, sa + (uint32_t)((uint64_t)la * sub / tpp), sa + (uint32_t)((uint64_t)la * (sub + 1u) / tpp));
        acc += span_sum(s_b, sb + (uint32_t)((uint64_t)lb * sub / tpp), sb + (uint32_t)((uint64_t)lb * (sub + 1u) / tpp));
    }
    for (uint32_t d = tpp >> 1; d >= 1u; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (mine && sub == 0u) ps.out[p] = acc;
}

#endif // SWB_SCHEDULE_KERNELS

} // namespace swb
