// lev_bitpar_coop.cuh — bit-vector Levenshtein for LONG patterns: a group of G lanes per pair, sm_100a.
//
// Same recurrence and nibble tables as lev_bitpar.cuh, but the pattern's bit-vector is split over the G lanes of
// a sub-warp group, K 32-bit words per lane (capacity CAP = 32*G*K rows per strip). A block of 32*K rows only
// needs the horizontal delta (+1 / 0 / -1, two bits) leaving the block above it, so the lanes run as an
// anti-diagonal wavefront over (row block, column): at step t lane l advances column t - l and hands its two
// delta bits to lane l + 1 with one `__shfl_up_sync`. The group stages the strings into shared memory with
// coalesced 16-byte loads; each lane builds the table slice of its own rows.
//
// No length limits:
//  * the TEXT is streamed through shared memory in chunks of `chunk_cols` bytes (the per-lane state simply
//    carries over; each chunk costs G-1 fill/drain steps);
//  * a PATTERN longer than CAP is processed in row strips of CAP rows; the two delta bits leaving a strip's last
//    lane are parked per column in a global scratch row (1 byte per column) and fed to lane 0 of the next strip.
// The pattern is top-aligned over all strips (see lev_bitpar.cuh); lanes of strip 0 that hold nothing but free rows
// stay idle (the first lane holding pattern rows takes the constant +1 delta of DP row 0). The distance is
// n + sum of popc(Pv) - popc(Mv) over all lanes and strips after the last column: nothing is tracked per step.
//
// <G,K> = <8,2> 512, <16,2> 1024, <32,2> 2048, <32,4> 4096 rows per strip. Parity: bit-exact with the CPU oracle.
#pragma once
#include "lev_bitpar.cuh"

namespace swb {

constexpr int BPC_THREADS = 128;

// STRIPS = false compiles the row-strip machinery out (every pattern of the launch fits CAP rows).
template <int G, int K, bool STRIPS>
__global__ void __launch_bounds__(BPC_THREADS)
    lev_bitpar_coop_kernel(PairSpace ps, uint64_t begin, uint64_t count, uint32_t chunk_cols, uint8_t *hscratch, uint64_t scratch_stride) {
    constexpr int T = BPC_THREADS, PAIRS_PER_BLOCK = T / G;
    using NT = NibbleTables<K, T>;
    constexpr uint32_t ROWS_PER_LANE = 32u * K, CAP = ROWS_PER_LANE * G;
    constexpr uint32_t SLOT_ROWS = CAP + 32u;
    extern __shared__ __align__(16) unsigned char smem_bytes[]; // NibbleTables<K, T> | per pair: rows slot, cols slot
    const uint32_t slot_cols = chunk_cols + 32u;
    const int tid = threadIdx.x, group = tid / G, lane = tid % G;
    const uint32_t base_hi = NT::base((uint32_t)__cvta_generic_to_shared(smem_bytes), tid), base_lo = base_hi + NT::TABLE;
    const uint32_t opaque0 = (uint32_t)(count >> 63); // run-time 0: see mad_lo in lev_bitpar.cuh
    const uint32_t one = 1u + opaque0, two = 2u + opaque0;
    unsigned char *const rows_slot = smem_bytes + NT::BYTES + (size_t)group * (SLOT_ROWS + slot_cols);
    unsigned char *const cols_slot = rows_slot + SLOT_ROWS;

    const uint64_t local = (uint64_t)blockIdx.x * PAIRS_PER_BLOCK + group;
    const bool have = local < count;
    uint32_t m = 0, n = 0;
    const uint8_t *pat = nullptr, *txt = nullptr;
    uint64_t out_index = 0;
    if (have) {
        const PairView pv = pair_at(ps, begin + local);
        const bool swap = pv.la > pv.lb; // pattern = shorter string
        m = swap ? pv.lb : pv.la, n = swap ? pv.la : pv.lb;
        pat = swap ? pv.b : pv.a, txt = swap ? pv.a : pv.b;
        out_index = pv.out_index;
    }
    uint8_t *const hs = (STRIPS && hscratch) ? hscratch + local * scratch_stride : nullptr;
    const uint32_t strips = STRIPS ? (m + CAP - 1u) / CAP : (m ? 1u : 0u); // 0 for an empty pattern
    const uint32_t s = strips * CAP - m;          // pattern symbol i lives at global bit position s + i (s < CAP)

    uint32_t n_max = n, strips_max = strips;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, d));
        strips_max = max(strips_max, __shfl_xor_sync(0xffffffffu, strips_max, d));
    }
    // D[m][n] = n + (sum of the vertical deltas of the last column) = n + sum over strips and lanes of
    // popc(Pv) - popc(Mv) after the last column; nothing is tracked per step.
    int delta_sum = 0;

    for (uint32_t r = 0; r < strips_max; ++r) {
        const bool strip_live = have && r < strips;
        // ---- stage the pattern bytes of this strip, then build this lane's slice of the nibble tables
        const uint32_t strip_lo = r * CAP;                                      // first global bit position of the strip
        const uint32_t i_begin = strip_lo > s ? strip_lo - s : 0u;               // first pattern index in the strip
        const uint32_t i_end = strip_live ? min(m, strip_lo + CAP - s) : i_begin; // one past the last
        __syncwarp();
        const uint32_t mis_r = strip_live ? stage_string(rows_slot, pat + i_begin, i_end - i_begin, lane, G) : 0u;
        zero_tables<K, T>(base_hi);
        __syncwarp();
        const unsigned char *rows_s = rows_slot + mis_r;
        const uint32_t base = strip_lo + ROWS_PER_LANE * (uint32_t)lane; // first global bit position of this lane
        const bool lane_live = strip_live && base + ROWS_PER_LANE > s;  // holds at least one pattern row
        const bool lane_first = r == 0 && base <= s;                    // lowest live lane: DP row 0 enters here (+1)
        if (lane_live) {
            const uint32_t i_lo = max(i_begin, base > s ? base - s : 0u), i_hi = min(i_end, base + ROWS_PER_LANE - s);
            for (uint32_t i = i_lo; i < i_hi; ++i) {
                add_symbol<K, T>(base_hi, s + i - base, rows_s[i - i_begin]); // local bit position 0 .. 32K-1
            }
        }
        uint32_t Pv[K], Mv[K];
        first_column<K>(s, Pv, Mv, base);
        const bool from_scratch = STRIPS && r > 0 && lane == 0;                          // delta bits come from the previous strip
        const bool to_scratch = STRIPS && strip_live && r + 1 < strips && lane == G - 1; // ... or go to the next one
        uint32_t send = 0;

        // ---- the text, one shared-memory chunk at a time
        for (uint32_t c0 = 0; c0 < n_max; c0 += chunk_cols) {
            const uint32_t len_c = (lane_live && c0 < n) ? min(n - c0, chunk_cols) : 0u;
            const uint32_t len_stage = (strip_live && c0 < n) ? min(n - c0, chunk_cols) : 0u;
            __syncwarp(); // every lane is done with the previous chunk
            const uint32_t mis_c = len_stage ? stage_string(cols_slot, txt + c0, len_stage, lane, G) : 0u;
            __syncwarp();
            const unsigned char *cols_s = cols_slot + mis_c;
            uint32_t next_in = (STRIPS && from_scratch && len_c) ? hs[c0] : 0u; // prefetched delta bits of the next column
            const uint32_t steps = min(n_max - c0, chunk_cols) + G - 1;
            for (uint32_t t = 0; t < steps; ++t) {
                const uint32_t recv = __shfl_up_sync(0xffffffffu, send, 1, G);
                const uint32_t jj = t - (uint32_t)lane; // column inside the chunk
                if (jj < len_c) {
                    const uint32_t c = cols_s[jj];
                    uint32_t hi[K], lo[K];
                    load_words<K, T>(base_hi + (c >> 4) * NT::ROW, hi);
                    load_words<K, T>(base_lo + (c & 15u) * NT::ROW, lo);
                    uint32_t in = recv;
                    if (STRIPS && from_scratch) {
                        in = next_in;
                        if (jj + 1 < len_c) next_in = hs[c0 + jj + 1];
                    }
                    const uint32_t hp_in = lane_first ? 1u : (in & 1u), hn_in = lane_first ? 0u : (in >> 1);
                    uint32_t ph_top, mh_top;
                    column_step_io<K, true>(hi, lo, Pv, Mv, one, two, hp_in, hn_in, ph_top, mh_top);
                    const uint32_t hp_out = ph_top >> 31, hn_out = mh_top >> 31;
                    send = hp_out | (hn_out << 1);
                    if (STRIPS && to_scratch) hs[c0 + jj] = (uint8_t)send;
                }
            }
        }
        if (lane_live) { // this lane's rows of this strip have seen the whole text
#pragma unroll
            for (int w = 0; w < K; ++w) delta_sum += __popc(Pv[w]) - __popc(Mv[w]);
        }
    }
#pragma unroll
    for (int d = G / 2; d >= 1; d >>= 1) delta_sum += __shfl_xor_sync(0xffffffffu, delta_sum, d, G);
    if (have && lane == G - 1) ps.out[out_index] = (uint64_t)((int64_t)n + delta_sum);
}

} // namespace swb
