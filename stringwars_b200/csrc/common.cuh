// common.cuh — shared device-side helpers for the stringwars_b200 kernels (sm_100a).
//
// Data layout in HBM: a "tape" is one contiguous byte buffer plus N+1 u64 offsets, the
// Arrow LargeBinary layout of `BytesTapeView<u64>` (StringWars similarities/bench.rs:62, 292-306). Kernels
// address pairs through a PairSpace so that element-wise pairs (BASELINE.json configs) and the harness's
// cross-product shape (bench.rs:123-148) share every kernel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace swb {

struct PairSpace {
    const uint8_t *a_data;
    const uint64_t *a_offs;
    const uint8_t *b_data;
    const uint64_t *b_offs;
    uint64_t n_pairs;      // total pairs (n_q * n_c in cross mode)
    uint64_t n_c;          // 0: element-wise pairs; otherwise candidates per query
    uint64_t stride;       // cross mode: output row stride (elements)
    const uint32_t *order; // optional permutation (length-bucketed schedule); nullptr = identity
    uint64_t *out;         // u64 distances or i64 scores (element width: out_narrow)
    uint32_t len_shift;    // log2(bytes per symbol): 0 for byte tapes, 2 for decoded UTF-8 (u32 symbols)
    uint32_t out_narrow;   // 0: 64-bit results (the ABI's usize / isize); 1: 32-bit; 2: 16-bit (device staging that the host widens)
    uint32_t offs32;       // bit 0: a_offs points at u32 offsets, bit 1: b_offs does (`BytesTapeView<u32>`, AnyBytesTape::View32)
    uint32_t idx_shift;    // string i of a pair space is string (i << idx_shift) of its tape (1: the renamed code-point tapes, whose odd strings are the gaps)
};

// Offset i of tape A / B, whatever the offset width.
__device__ __forceinline__ uint64_t a_off(const PairSpace &ps, uint64_t i) {
    return (ps.offs32 & 1u) ? (uint64_t)reinterpret_cast<const uint32_t *>(ps.a_offs)[i] : ps.a_offs[i];
}
__device__ __forceinline__ uint64_t b_off(const PairSpace &ps, uint64_t i) {
    return (ps.offs32 & 2u) ? (uint64_t)reinterpret_cast<const uint32_t *>(ps.b_offs)[i] : ps.b_offs[i];
}

// One result (distance or signed score) at element index `idx` of the output, in the width the host asked for.
__device__ __forceinline__ void store_result(const PairSpace &ps, uint64_t idx, int64_t v) {
    if (ps.out_narrow == 0u) ps.out[idx] = (uint64_t)v;
    else if (ps.out_narrow == 1u) reinterpret_cast<uint32_t *>(ps.out)[idx] = (uint32_t)v;
    else reinterpret_cast<uint16_t *>(ps.out)[idx] = (uint16_t)v;
}

// Self-validating speculative launches. A kernel launched on the bet that a batch looks like the previous one checks every
// pair it is given — offsets that run backwards or span 4 GiB, a pattern that does not fit the kernel, a string longer than
// anything the host has seen — skips what fails ("miss") and keeps the batch's length extrema; the last block to finish
// publishes {misses, extrema} into mapped host memory and re-arms the device copy. The host polls one word: no statistics
// kernel, no second launch, no copy on the critical path of a repeated call.
struct SpecReport {
    uint32_t misses, max_m, max_n, min_m, min_n, ticket; // device; entry state: 0, 0, 0, ~0, ~0, 0
};
struct SpecPublished {
    uint32_t misses, max_m, max_n, min_m, min_n, seq; // mapped host memory; seq is written last
};
struct SpecArgs {
    SpecReport *report = nullptr; // nullptr: not a speculative launch (nothing is counted or published)
    SpecPublished *published = nullptr;
    uint32_t seq = 0;
};

struct PairView {
    const uint8_t *a, *b; // first byte of each string
    uint32_t la, lb;      // lengths in SYMBOLS
    uint64_t out_index;
};

// pair_at in two halves, for kernels that request a pair's offsets long before they need them: pair_request issues
// the four offset loads and does no arithmetic on their results, pair_resolve turns them into a PairView.
struct PairRequest {
    uint64_t a0, a1, b0, b1; // raw offsets (in flight until first use)
    uint64_t out_index;
};

__device__ __forceinline__ PairRequest pair_request(const PairSpace &ps, uint64_t slot) {
    uint64_t k = ps.order ? (uint64_t)ps.order[slot] : slot;
    uint64_t ia = k, ib = k, oi = k;
    if (ps.n_c) {
        ia = k / ps.n_c;
        ib = k - ia * ps.n_c;
        oi = ia * ps.stride + ib;
    }
    ia <<= ps.idx_shift, ib <<= ps.idx_shift;
    PairRequest r;
    if (ps.offs32 & 1u) {
        const uint32_t *a32 = reinterpret_cast<const uint32_t *>(ps.a_offs);
        r.a0 = a32[ia], r.a1 = a32[ia + 1];
    } else {
        r.a0 = ps.a_offs[ia], r.a1 = ps.a_offs[ia + 1];
    }
    if (ps.offs32 & 2u) {
        const uint32_t *b32 = reinterpret_cast<const uint32_t *>(ps.b_offs);
        r.b0 = b32[ib], r.b1 = b32[ib + 1];
    } else {
        r.b0 = ps.b_offs[ib], r.b1 = ps.b_offs[ib + 1];
    }
    r.out_index = oi;
    return r;
}

__device__ __forceinline__ PairView pair_resolve(const PairSpace &ps, const PairRequest &r) {
    PairView v;
    v.a = ps.a_data + r.a0;
    v.b = ps.b_data + r.b0;
    // a string whose offsets run backwards reads as empty: the schedule kernel counts it and the host fails the call with
    // SWB_INVALID_ARG, but a speculative kernel launched beside that check must not chase the wild length
    v.la = r.a1 < r.a0 ? 0u : (uint32_t)((r.a1 - r.a0) >> ps.len_shift);
    v.lb = r.b1 < r.b0 ? 0u : (uint32_t)((r.b1 - r.b0) >> ps.len_shift);
    v.out_index = r.out_index;
    return v;
}

__device__ __forceinline__ PairView pair_at(const PairSpace &ps, uint64_t slot) { return pair_resolve(ps, pair_request(ps, slot)); }

// Both strings of a request are well-formed: neither runs backwards nor spans 4 GiB or more.
__device__ __forceinline__ bool request_ok(const PairRequest &r) {
    return r.a1 >= r.a0 && r.b1 >= r.b0 && ((r.a1 - r.a0) >> 32) == 0u && ((r.b1 - r.b0) >> 32) == 0u;
}

// Tail of a speculative kernel: every thread of the block calls it once. `miss`: this thread skipped its pair; m / n: the
// pair's shorter / longer length (ignored when `counted` is false: threads beyond the batch, skipped pairs).
__device__ __forceinline__ void spec_finish(const SpecArgs &sa, bool miss, bool counted, uint32_t m, uint32_t n) {
    if (!sa.report) return;
    const uint32_t all = 0xffffffffu;
    const uint32_t hi_m = __reduce_max_sync(all, counted ? m : 0u), hi_n = __reduce_max_sync(all, counted ? n : 0u);
    const uint32_t lo_m = __reduce_min_sync(all, counted ? m : all), lo_n = __reduce_min_sync(all, counted ? n : all);
    const uint32_t any_miss = __ballot_sync(all, miss);
    if ((threadIdx.x & 31u) == 0u) { // read first: after the first blocks of a uniform batch nothing is left to update
        volatile SpecReport *r = sa.report;
        if (any_miss) atomicAdd(&sa.report->misses, (uint32_t)__popc(any_miss));
        if (hi_m > r->max_m) atomicMax(&sa.report->max_m, hi_m);
        if (hi_n > r->max_n) atomicMax(&sa.report->max_n, hi_n);
        if (lo_m < r->min_m) atomicMin(&sa.report->min_m, lo_m);
        if (lo_n < r->min_n) atomicMin(&sa.report->min_n, lo_n);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&sa.report->ticket, 1u) == gridDim.x * gridDim.y - 1u) {
            __threadfence();
            volatile SpecReport *r = sa.report;
            sa.published->misses = r->misses, sa.published->max_m = r->max_m, sa.published->max_n = r->max_n;
            sa.published->min_m = r->min_m, sa.published->min_n = r->min_n;
            r->misses = 0u, r->max_m = 0u, r->max_n = 0u, r->min_m = all, r->min_n = all, r->ticket = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t *>(&sa.published->seq) = sa.seq;
            __threadfence_system();
        }
    }
}

// Read-only 16-byte / 4-byte / 1-byte global loads (LDG through the non-coherent path).
__device__ __forceinline__ uint4 ldg16(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ uint32_t ldg4(const void *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }
__device__ __forceinline__ uint32_t ldg1(const uint8_t *p) { return __ldg(p); }

// Byte b (0..15, run-time) of a 16-byte piece held in registers.
__device__ __forceinline__ uint32_t piece_byte(const uint4 &p, uint32_t b) {
    const uint32_t w = (b & 8u) ? ((b & 4u) ? p.w : p.z) : ((b & 4u) ? p.y : p.x);
    return (w >> (8u * (b & 3u))) & 0xFFu;
}

// The 16 bytes that start `mis` bytes (0..15, run-time) into the 32-byte window {lo, hi}: two rounds of word
// selects (by 2 words, by 1 word) and four funnel shifts.
__device__ __forceinline__ uint4 realign16(const uint4 &lo, const uint4 &hi, uint32_t mis) {
    const bool by2 = (mis & 8u) != 0u, by1 = (mis & 4u) != 0u;
    const uint32_t t0 = by2 ? lo.z : lo.x, t1 = by2 ? lo.w : lo.y, t2 = by2 ? hi.x : lo.z, t3 = by2 ? hi.y : lo.w, t4 = by2 ? hi.z : hi.x,
                   t5 = by2 ? hi.w : hi.y;
    const uint32_t u0 = by1 ? t1 : t0, u1 = by1 ? t2 : t1, u2 = by1 ? t3 : t2, u3 = by1 ? t4 : t3, u4 = by1 ? t5 : t4;
    const uint32_t sh = (mis & 3u) * 8u;
    return make_uint4(__funnelshift_r(u0, u1, sh), __funnelshift_r(u1, u2, sh), __funnelshift_r(u2, u3, sh), __funnelshift_r(u3, u4, sh));
}

// A string in global memory as a stream of 16-byte register pieces, whatever its alignment: piece i holds bytes
// [16 i, 16 i + 16) of the window that starts `lead` bytes BEFORE the first string byte (lead < 16 is the caller's
// choice of phase: 0 streams the string itself, (position of byte 0) % 16 streams pieces aligned to some other grid).
// Only aligned 16-byte loads are issued, and only for aligned pieces that hold at least one string byte.
struct PieceStream {
    const uint8_t *grid; // aligned address of the aligned piece that holds window[0]
    uint32_t mis;        // window[0] is byte `mis` of that piece
    uint32_t lo, hi;     // the string occupies bytes [lo, hi) counted from `grid`
    __device__ __forceinline__ void open(const uint8_t *s, uint32_t len, uint32_t lead) {
        const uint8_t *window = s - lead;
        mis = (uint32_t)(reinterpret_cast<uintptr_t>(window) & 15u);
        grid = window - mis;
        lo = mis + lead, hi = lo + len;
    }
    // aligned piece number a of the grid (zeros if it holds no string byte: never loaded)
    __device__ __forceinline__ uint4 aligned(uint32_t a) const {
        return (16u * a + 16u > lo && 16u * a < hi) ? ldg16(grid + 16u * a) : make_uint4(0, 0, 0, 0);
    }
    // piece i of the window out of its two aligned pieces (i and i + 1)
    __device__ __forceinline__ uint4 combine(const uint4 &a_i, const uint4 &a_i1) const { return mis ? realign16(a_i, a_i1, mis) : a_i; }
    __device__ __forceinline__ uint4 piece(uint32_t i) const {
        const uint4 lo_piece = aligned(i);
        return mis ? realign16(lo_piece, aligned(i + 1), mis) : lo_piece;
    }
};

// Byte k (0..3, compile-time) of a 32-bit word.
template <int KB>
__device__ __forceinline__ uint32_t byte_of(uint32_t w) {
    return __byte_perm(w, 0, 0x4440 + KB);
}

// Visit every byte of a string that lives in GLOBAL memory, in order: f(index, byte).
// Aligned 16-byte pieces (one LDG.128, next piece prefetched) with byte loads for the ragged head and tail.
template <typename F>
__device__ __forceinline__ void for_each_byte_global(const uint8_t *s, uint32_t len, F &&f) {
    uint32_t i = 0;
    const uint32_t head = min(len, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(s) & 15u)) & 15u));
    for (; i < head; ++i) f(i, ldg1(s + i));
    if (i + 16 <= len) {
        uint4 nxt = ldg16(s + i);
        while (i + 16 <= len) {
            uint4 cur = nxt;
            if (i + 32 <= len) nxt = ldg16(s + i + 16);
            const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                f(i + 4 * q + 0, byte_of<0>(w[q]));
                f(i + 4 * q + 1, byte_of<1>(w[q]));
                f(i + 4 * q + 2, byte_of<2>(w[q]));
                f(i + 4 * q + 3, byte_of<3>(w[q]));
            }
            i += 16;
        }
    }
    for (; i < len; ++i) f(i, ldg1(s + i));
}

// Copy `len` bytes from global `src` into shared `dst` using the lanes of one group (lane_in_group of
// group_size) with aligned 16-byte loads. `dst` must be 16-byte aligned; the copy preserves the source's
// misalignment, i.e. byte i of the string lands at dst[(src & 15) + i]. Returns that misalignment.
// Over-reads up to 15 bytes on both sides of the string inside its own 16-byte pieces (always inside the
// allocation granule of cudaMalloc, never crossing a 16-byte boundary that holds no string byte).
__device__ __forceinline__ uint32_t stage_string(uint8_t *dst, const uint8_t *src, uint32_t len, int lane_in_group,
                                                 int group_size) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint8_t *base = src - mis;
    const uint32_t pieces = len ? (mis + len + 15u) >> 4 : 0u;
    for (uint32_t p = lane_in_group; p < pieces; p += group_size)
        *reinterpret_cast<uint4 *>(dst + 16u * p) = ldg16(base + 16u * p);
    return mis;
}

} // namespace swb
