// lev_bitpar_coop.cuh — bit-vector Levenshtein for LONG patterns: a group of G lanes per pair, sm_100a.
//
// Same recurrence and nibble tables as lev_bitpar.cuh, but the pattern's bit-vector is split over the G lanes of
// a sub-warp group, K 32-bit words per lane (capacity CAP = 32*G*K rows per strip). A block of 32*K rows only
// needs the horizontal delta (+1 / 0 / -1, two bits) leaving the block above it, so the lanes run as an
// anti-diagonal wavefront over (row block, column group): at step t lane l advances the four columns of group t - l
// and hands their delta bits to lane l + 1 with one `__shfl_up_sync`. The group stages the strings into shared memory with
// coalesced 16-byte loads; each lane builds the table slice of its own rows.
//
// No length limits:
//  * the TEXT is streamed through shared memory in chunks of `chunk_cols` bytes (the per-lane state simply
//    carries over; each chunk costs G-1 fill/drain steps);
//  * a PATTERN longer than CAP is processed in row strips of CAP rows; the two delta bits leaving a strip's last
//    lane are parked per column in a global scratch row (1 byte per four columns) and fed to lane 0 of the next strip.
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

        // ---- the text, one shared-memory chunk at a time, FOUR columns per wavefront step: lane l computes columns
        // 4(t - l) .. 4(t - l) + 3 of its row block at step t and hands the four pairs of delta bits leaving the block to
        // lane l + 1 in one shuffle (the shuffle, the loop control and the text fetch are paid once per four columns).
        // Delta bits travel most-significant-first in two nibbles (+1 bits in 3..0, -1 bits in 7..4) so that the
        // producer appends a bit with one funnel shift and the consumer takes one off the top with one shift.
        for (uint32_t c0 = 0; c0 < n_max; c0 += chunk_cols) {
            const uint32_t len_c = (lane_live && c0 < n) ? min(n - c0, chunk_cols) : 0u;
            const uint32_t len_stage = (strip_live && c0 < n) ? min(n - c0, chunk_cols) : 0u;
            __syncwarp(); // every lane is done with the previous chunk
            if (len_stage) { // text bytes c0 .. c0 + len_stage - 1 land at cols_slot[0 ..] whatever their alignment in memory
                PieceStream ts;
                ts.open(txt + c0, len_stage, 0u);
                for (uint32_t p = lane; 16u * p < len_stage; p += G) *reinterpret_cast<uint4 *>(cols_slot + 16u * p) = ts.piece(p);
            }
            __syncwarp();
            const uint32_t *cols32 = reinterpret_cast<const uint32_t *>(cols_slot);
            const uint32_t my_groups = (len_c + 3u) >> 2;
            const uint32_t steps = ((min(n_max - c0, chunk_cols) + 3u) >> 2) + G - 1;
            uint8_t *const hs_c = STRIPS ? hs + (c0 >> 2) : nullptr; // one scratch byte per four columns
            uint32_t word_next = 0u, in_next = 0u;
            {
                const uint32_t gg = 0u - (uint32_t)lane;
                if (gg < my_groups) word_next = cols32[gg];
                if (STRIPS && from_scratch && my_groups) in_next = hs_c[0];
            }
            for (uint32_t t = 0; t < steps; ++t) {
                const uint32_t recv = __shfl_up_sync(0xffffffffu, send, 1, G);
                const uint32_t gg = t - (uint32_t)lane; // column group inside the chunk
                const uint32_t word = word_next;
                if (gg + 1u < my_groups) word_next = cols32[gg + 1u];
                if (gg < my_groups) {
                    const NibbleWord nw = split_nibbles(word);
                    uint32_t hi[4][K], lo[4][K];
                    load_words<K, T>(row_of<0>(nw.hi, base_hi), hi[0]);
                    load_words<K, T>(row_of<0>(nw.lo, base_lo), lo[0]);
                    load_words<K, T>(row_of<1>(nw.hi, base_hi), hi[1]);
                    load_words<K, T>(row_of<1>(nw.lo, base_lo), lo[1]);
                    load_words<K, T>(row_of<2>(nw.hi, base_hi), hi[2]);
                    load_words<K, T>(row_of<2>(nw.lo, base_lo), lo[2]);
                    load_words<K, T>(row_of<3>(nw.hi, base_hi), hi[3]);
                    load_words<K, T>(row_of<3>(nw.lo, base_lo), lo[3]);
                    uint32_t in = recv;
                    if (STRIPS && from_scratch) {
                        in = in_next;
                        if (gg + 1u < my_groups) in_next = hs_c[gg + 1u];
                    }
                    if (lane_first) in = 0x0Fu; // DP row 0 enters here: +1 in every column
                    uint32_t in_p = in << 28, in_n = (in >> 4) << 28; // next column's bit on top
                    uint32_t out_p = 0u, out_n = 0u;
                    const uint32_t cols_here = min(4u, len_c - 4u * gg);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k == 0 || (uint32_t)k < cols_here) {
                            uint32_t ph_top, mh_top;
                            column_step_io<K, true>(hi[k], lo[k], Pv, Mv, one, two, in_p >> 31, in_n >> 31, ph_top, mh_top);
                            out_p = __funnelshift_l(ph_top, out_p, 1), out_n = __funnelshift_l(mh_top, out_n, 1);
                        } else { // past the end of the text: keep the bit order
                            out_p <<= 1, out_n <<= 1;
                        }
                        in_p = mad_lo(in_p, two, 0u), in_n = mad_lo(in_n, two, 0u);
                    }
                    send = out_p | (out_n << 4);
                    if (STRIPS && to_scratch) hs_c[gg] = (uint8_t)send;
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
    if (have && lane == G - 1) store_result(ps, out_index, (int64_t)n + delta_sum);
}

} // namespace swb
