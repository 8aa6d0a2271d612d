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
#include "common.cuh"

namespace swb {

constexpr int BPC_THREADS = 128;

// STRIPS = false compiles the row-strip machinery out (every pattern of the launch fits CAP rows).
template <int G, int K, bool STRIPS>
__global__ void __launch_bounds__(BPC_THREADS)
    lev_bitpar_coop_kernel(PairSpace ps, uint64_t begin, uint64_t count, uint32_t chunk_cols, uint8_t *hscratch, uint64_t scratch_stride) {
    constexpr int T = BPC_THREADS, PAIRS_PER_BLOCK = T / G;
    constexpr uint32_t ESTRIDE = T * K * 4; // bytes between consecutive table entries
    constexpr uint32_t ROWS_PER_LANE = 32u * K, CAP = ROWS_PER_LANE * G;
    constexpr uint32_t SLOT_ROWS = CAP + 32u;
    extern __shared__ __align__(16) unsigned char smem_bytes[]; // [32 entries][T][K] u32 | per pair: rows slot, cols slot
    const uint32_t slot_cols = chunk_cols + 32u;
    const int tid = threadIdx.x, group = tid / G, lane = tid % G;
    unsigned char *const tbl_hi = smem_bytes + tid * (K * 4), *const tbl_lo = tbl_hi + 16 * ESTRIDE;
    unsigned char *const rows_slot = smem_bytes + 32 * ESTRIDE + (size_t)group * (SLOT_ROWS + slot_cols);
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
#pragma unroll 8
        for (int e = 0; e < 32; ++e) {
#pragma unroll
            for (int w = 0; w < K; ++w) *reinterpret_cast<uint32_t *>(tbl_hi + e * ESTRIDE + 4 * w) = 0u;
        }
        __syncwarp();
        const unsigned char *rows_s = rows_slot + mis_r;
        const uint32_t base = strip_lo + ROWS_PER_LANE * (uint32_t)lane; // first global bit position of this lane
        const bool lane_live = strip_live && base + ROWS_PER_LANE > s;  // holds at least one pattern row
        const bool lane_first = r == 0 && base <= s;                    // lowest live lane: DP row 0 enters here (+1)
        if (lane_live) {
            const uint32_t i_lo = max(i_begin, base > s ? base - s : 0u), i_hi = min(i_end, base + ROWS_PER_LANE - s);
            for (uint32_t i = i_lo; i < i_hi; ++i) {
                const uint32_t c = rows_s[i - i_begin], q = s + i - base; // local bit position 0 .. 32K-1
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
                    uint32_t in = recv;
                    if (STRIPS && from_scratch) {
                        in = next_in;
                        if (jj + 1 < len_c) next_in = hs[c0 + jj + 1];
                    }
                    uint32_t hp_in = lane_first ? 1u : (in & 1u), hn_in = lane_first ? 0u : (in >> 1);
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
