// lev_bitpar_coop.cuh — bit-vector Levenshtein for LONG patterns: a group of G lanes per pair, sm_100a.
//
// Same recurrence and nibble tables as lev_bitpar.cuh, but the pattern's bit-vector (32*G*K rows) is split over
// the G lanes of a sub-warp group, K 32-bit words per lane. A block of 32*K rows only needs the horizontal delta
// (+1 / 0 / -1, two bits) leaving the block above it, so the lanes run as an anti-diagonal wavefront over
// (row block, column): at step t lane l advances column t - l and hands its two delta bits to lane l + 1 with
// one `__shfl_up_sync`. Both strings are staged into shared memory by the group with coalesced 16-byte loads;
// each lane builds the table slice of its own rows. The pattern is top-aligned (see lev_bitpar.cuh), so the
// distance accumulates in the LAST lane only, and lanes that hold nothing but free rows stay idle (the first
// lane holding pattern rows takes the constant +1 delta of DP row 0).
//
// Capacity 32*G*K rows: <8,2> 512, <16,2> 1024, <32,2> 2048, <32,4> 4096. Longer patterns go to the wavefront
// kernel's row strips. Parity: bit-exact with the CPU oracle on every case of tests/test_gpu_parity.py.
#pragma once
#include "common.cuh"

namespace swb {

constexpr int BPC_THREADS = 128;

template <int G, int K>
__global__ void __launch_bounds__(BPC_THREADS)
    lev_bitpar_coop_kernel(PairSpace ps, uint64_t begin, uint64_t count, uint32_t slot_rows, uint32_t slot_cols) {
    constexpr int T = BPC_THREADS, PAIRS_PER_BLOCK = T / G;
    constexpr uint32_t ESTRIDE = T * K * 4; // bytes between consecutive table entries
    constexpr uint32_t ROWS_PER_LANE = 32u * K, CAPACITY = ROWS_PER_LANE * G;
    extern __shared__ __align__(16) unsigned char smem_bytes[]; // [32 entries][T][K] u32 | per pair: rows slot, cols slot
    const int tid = threadIdx.x, group = tid / G, lane = tid % G;
    unsigned char *const tbl_hi = smem_bytes + tid * (K * 4), *const tbl_lo = tbl_hi + 16 * ESTRIDE;
    unsigned char *const strings = smem_bytes + 32 * ESTRIDE + (size_t)group * (slot_rows + slot_cols);

#pragma unroll 8
    for (int e = 0; e < 32; ++e) {
#pragma unroll
        for (int w = 0; w < K; ++w) *reinterpret_cast<uint32_t *>(tbl_hi + e * ESTRIDE + 4 * w) = 0u;
    }

    const uint64_t local = (uint64_t)blockIdx.x * PAIRS_PER_BLOCK + group;
    const bool have = local < count;
    uint32_t m = 0, n = 0, mis_r = 0, mis_c = 0;
    uint64_t out_index = 0;
    if (have) {
        const PairView pv = pair_at(ps, begin + local);
        const bool swap = pv.la > pv.lb; // pattern = shorter string
        m = swap ? pv.lb : pv.la;
        n = swap ? pv.la : pv.lb;
        out_index = pv.out_index;
        mis_r = stage_string(strings, swap ? pv.b : pv.a, m, lane, G);
        mis_c = stage_string(strings + slot_rows, swap ? pv.a : pv.b, n, lane, G);
    }
    __syncwarp();
    const unsigned char *rows_s = strings + mis_r, *cols_s = strings + slot_rows + mis_c;

    const uint32_t s = CAPACITY - m;                  // pattern symbol i lives at global bit position s + i
    const uint32_t base = ROWS_PER_LANE * (uint32_t)lane; // first global bit position of this lane
    const bool lane_live = have && m > 0 && base + ROWS_PER_LANE > s; // holds at least one pattern row
    const bool lane_first = base <= s;                // the lowest live lane: DP row 0 enters here (delta +1)

    // ---- this lane's slice of the nibble tables
    if (lane_live) {
        const uint32_t i_lo = base > s ? base - s : 0u, i_hi = min(m, base + ROWS_PER_LANE - s);
        for (uint32_t i = i_lo; i < i_hi; ++i) {
            const uint32_t c = rows_s[i], q = s + i - base; // local bit position 0 .. 32K-1
            const uint32_t bit = 1u << (q & 31u), word_off = (q >> 5) * 4u;
            *reinterpret_cast<uint32_t *>(tbl_hi + (c >> 4) * ESTRIDE + word_off) |= bit;
            *reinterpret_cast<uint32_t *>(tbl_lo + (c & 15u) * ESTRIDE + word_off) |= bit;
        }
    }

    uint32_t Pv[K], Mv[K];
#pragma unroll
    for (int w = 0; w < K; ++w) {
        const uint32_t lo_bit = base + 32u * w; // rows below the pattern: vertical delta 0; pattern rows: +1
        Pv[w] = s >= lo_bit + 32u ? 0u : (s <= lo_bit ? 0xFFFFFFFFu : 0xFFFFFFFFu << (s - lo_bit));
        Mv[w] = 0u;
    }
    uint32_t score = m, send = 0;
    const uint32_t n_live = lane_live ? n : 0u;

    uint32_t n_max = n;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, d));
    const uint32_t steps = n_max + G - 1;
    for (uint32_t t = 0; t < steps; ++t) {
        const uint32_t recv = __shfl_up_sync(0xffffffffu, send, 1, G);
        const uint32_t j = t - (uint32_t)lane;
        if (j < n_live) {
            const uint32_t c = cols_s[j];
            uint32_t hi[K], lo[K];
            const unsigned char *ph_ = tbl_hi + (c >> 4) * ESTRIDE, *pl_ = tbl_lo + (c & 15u) * ESTRIDE;
            if constexpr (K == 2) {
                const uint2 vh = *reinterpret_cast<const uint2 *>(ph_), vl = *reinterpret_cast<const uint2 *>(pl_);
                hi[0] = vh.x, hi[1] = vh.y, lo[0] = vl.x, lo[1] = vl.y;
            } else if constexpr (K == 4) {
                const uint4 vh = *reinterpret_cast<const uint4 *>(ph_), vl = *reinterpret_cast<const uint4 *>(pl_);
                hi[0] = vh.x, hi[1] = vh.y, hi[2] = vh.z, hi[3] = vh.w, lo[0] = vl.x, lo[1] = vl.y, lo[2] = vl.z, lo[3] = vl.w;
            } else {
#pragma unroll
                for (int w = 0; w < K; ++w) {
                    hi[w] = *reinterpret_cast<const uint32_t *>(ph_ + 4 * w);
                    lo[w] = *reinterpret_cast<const uint32_t *>(pl_ + 4 * w);
                }
            }
            uint32_t hp_in = lane_first ? 1u : (recv & 1u), hn_in = lane_first ? 0u : (recv >> 1);
#pragma unroll
            for (int w = 0; w < K; ++w) {
                const uint32_t eqc = (hi[w] & lo[w]) | hn_in;
                const uint32_t xv = (hi[w] & lo[w]) | Mv[w];
                const uint32_t xh = (((eqc & Pv[w]) + Pv[w]) ^ Pv[w]) | eqc;
                const uint32_t ph = Mv[w] | ~(xh | Pv[w]);
                const uint32_t mh = Pv[w] & xh;
                const uint32_t phs = ph * 2u + hp_in;
                const uint32_t mhs = mh * 2u + hn_in;
                Pv[w] = mhs | ~(xv | phs);
                Mv[w] = phs & xv;
                hp_in = ph >> 31, hn_in = mh >> 31;
            }
            send = hp_in | (hn_in << 1);
            if (lane == G - 1) score += hp_in - hn_in; // the last pattern row is bit 31 of the last lane's last word
        }
    }
    if (have && lane == G - 1) ps.out[out_index] = m == 0 ? n : score;
}

} // namespace swb
