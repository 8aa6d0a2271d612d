// lev_bitpar.cuh — unit-cost Levenshtein by the Myers (1999) / Hyyrö (2003) bit-vector recurrence, sm_100a.
//
// Why this exists next to the wavefront kernel: with unit costs the vertical/horizontal DP deltas are in
// {-1, 0, +1}, so 32 DP rows fit two 32-bit masks and one column step costs ~12 LOP3/IADD/SHF per 32 cells
// instead of ~4-5 integer instructions per cell. Integer ALU issue is the binding roofline of this path
// (DESIGN.md §5), so this is the kernel that moves the HBM-roofline fraction.
//
// Mapping: one THREAD per pair, K 32-bit words of vertical state (pattern = the shorter string, <= 32*K
// symbols). The match mask Eq(c) is never stored for 256 symbols (2 KB/word would kill occupancy): it is
// rebuilt per column from two 16-entry nibble tables, Eq(c) = HI[c >> 4] & LO[c & 15], because a pattern
// byte equals c iff both nibbles do. The tables live in shared memory, laid out [table][nibble][thread][K]
// so a warp's lookups are one conflict-free LDS.32/64/128 each. Bytes stream straight from HBM with
// aligned 16-byte read-only loads, one piece prefetched (thread-private strings: nothing to share, so no
// smem staging here — the group-per-pair wavefront kernel is the one that stages).
//
// Parity: bit-exact with the CPU oracle (both of its algorithm families), the restatement of
// `rapidfuzz::distance::levenshtein::distance` as called at StringWars similarities/bench.rs:416-419.
#pragma once
#include "common.cuh"

namespace swb {

constexpr int BP_THREADS = 128;

template <int K>
struct WordVec;
template <>
struct WordVec<1> {
    using type = uint32_t;
};
template <>
struct WordVec<2> {
    using type = uint2;
};
template <>
struct WordVec<4> {
    using type = uint4;
};

// Load the K table words of entry `e` for this thread.
template <int K>
__device__ __forceinline__ void load_entry(const uint32_t *tbl, int e, int tid, uint32_t (&w)[K]) {
    const uint32_t *p = tbl + ((size_t)e * BP_THREADS + tid) * K;
    if constexpr (K == 1) {
        w[0] = p[0];
    } else if constexpr (K == 2) {
        uint2 v = *reinterpret_cast<const uint2 *>(p);
        w[0] = v.x, w[1] = v.y;
    } else {
#pragma unroll
        for (int q = 0; q < K / 4; ++q) {
            uint4 v = *reinterpret_cast<const uint4 *>(p + 4 * q);
            w[4 * q + 0] = v.x, w[4 * q + 1] = v.y, w[4 * q + 2] = v.z, w[4 * q + 3] = v.w;
        }
    }
}

template <int K>
__global__ void __launch_bounds__(BP_THREADS) lev_bitpar_kernel(PairSpace ps, uint64_t begin, uint64_t count) {
    extern __shared__ __align__(16) uint32_t tbl[]; // [32 entries][BP_THREADS][K]
    const int tid = threadIdx.x;
    const uint64_t local = (uint64_t)blockIdx.x * BP_THREADS + tid;

    // Zero this thread's 32 x K table words (coalesced across the warp by construction of the layout).
#pragma unroll 4
    for (int e = 0; e < 32; ++e) {
        uint32_t *p = tbl + ((size_t)e * BP_THREADS + tid) * K;
#pragma unroll
        for (int w = 0; w < K; ++w) p[w] = 0u;
    }
    if (local >= count) return;

    const PairView pv = pair_at(ps, begin + local);
    // pattern = shorter string (distance is symmetric), text = longer one
    const bool swap = pv.la > pv.lb;
    const uint8_t *pat = swap ? pv.b : pv.a;
    const uint8_t *txt = swap ? pv.a : pv.b;
    const uint32_t m = swap ? pv.lb : pv.la;
    const uint32_t n = swap ? pv.la : pv.lb;
    if (m == 0) {
        ps.out[pv.out_index] = n;
        return;
    }

    // Build the nibble tables: bit i of HI[h][w] (LO[l][w]) is set iff pattern byte 32*w+i has high (low) nibble h (l).
    for_each_byte_global(pat, m, [&](uint32_t i, uint32_t c) {
        const uint32_t bit = 1u << (i & 31u);
        uint32_t *hi = tbl + ((size_t)(c >> 4) * BP_THREADS + tid) * K + (i >> 5);
        uint32_t *lo = tbl + ((size_t)(16u + (c & 15u)) * BP_THREADS + tid) * K + (i >> 5);
        *hi |= bit;
        *lo |= bit;
    });

    uint32_t Pv[K], Mv[K];
#pragma unroll
    for (int w = 0; w < K; ++w) Pv[w] = 0xFFFFFFFFu, Mv[w] = 0u;
    uint32_t score = m;
    const uint32_t last_word = (m - 1u) >> 5, last_bit = (m - 1u) & 31u;

    for_each_byte_global(txt, n, [&](uint32_t, uint32_t c) {
        uint32_t hi[K], lo[K];
        load_entry<K>(tbl, (int)(c >> 4), tid, hi);
        load_entry<K>(tbl, (int)(16u + (c & 15u)), tid, lo);
        // Horizontal delta entering word 0 is +1 (row 0 of the DP matrix is 0,1,2,...).
        uint32_t ph_prev = 0x80000000u, mh_prev = 0u;
#pragma unroll
        for (int w = 0; w < K; ++w) {
            const uint32_t eq = hi[w] & lo[w];
            const uint32_t xv = eq | Mv[w];
            const uint32_t eqc = eq | (mh_prev >> 31);
            const uint32_t xh = (((eqc & Pv[w]) + Pv[w]) ^ Pv[w]) | eqc;
            const uint32_t ph = Mv[w] | ~(xh | Pv[w]);
            const uint32_t mh = Pv[w] & xh;
            if (w == (int)last_word || K == 1) score += ((ph >> last_bit) & 1u) - ((mh >> last_bit) & 1u);
            const uint32_t phs = __funnelshift_l(ph_prev, ph, 1);
            const uint32_t mhs = __funnelshift_l(mh_prev, mh, 1);
            Pv[w] = mhs | ~(xv | phs);
            Mv[w] = phs & xv;
            ph_prev = ph, mh_prev = mh;
        }
    });
    ps.out[pv.out_index] = score;
}

} // namespace swb
