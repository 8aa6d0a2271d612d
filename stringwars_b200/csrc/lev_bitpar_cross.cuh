// lev_bitpar_cross.cuh — cross-product bit-vector Levenshtein: one thread per QUERY, candidates swept, sm_100a.
//
// This is the shape the StringWars harness drives its batched engines with: `compute_into(queries, candidates,
// &mut matrix)` fills a side x side matrix (similarities/bench.rs:123-148, 599-603), so every query meets every
// candidate. The element-wise kernels rebuild the pattern's nibble tables for every pair; here a thread owns one
// query, builds its tables ONCE, and sweeps a range of candidates. A candidate is the same for the whole block,
// so it is staged into shared memory once (coalesced 16-byte loads) and every thread reads it with broadcast
// LDS — no per-thread global traffic in the inner loop. Per-pair fixed costs (table zeroing and build, offset
// loads, index division) are amortised over `cands_per_block` candidates, which is what short strings
// (the harness's default "words" dataset, ~5 bytes) are dominated by in the element-wise kernels.
//
// The query is always the pattern (rows), so queries must fit the K words: the host picks K from the longest
// query and falls back to the element-wise kernels when queries exceed 256 bytes.
#pragma once
#include "lev_bitpar.cuh"

namespace swb {

// Threads (= queries) per block: a candidate is staged once per block, so wider blocks than the pair kernel's.
template <int K>
struct CrossConfig {
    static constexpr int THREADS = K <= 4 ? 128 : 64;
};

constexpr int CROSS_TILE = 16; // candidates per result tile: 16 x u64 = one 128-byte segment per query row

template <int K>
__global__ void __launch_bounds__(CrossConfig<K>::THREADS)
    lev_bitpar_cross_kernel(PairSpace ps, uint32_t n_queries, uint32_t cands_per_block, uint32_t chunk, uint32_t n_cap, SpecArgs sa) {
    constexpr int T = CrossConfig<K>::THREADS;
    using NT = NibbleTables<K, T>;
    extern __shared__ __align__(16) unsigned char xsm[]; // NibbleTables<K, T>::BYTES | result tile | candidate chunk (chunk + 32 B)
    // Results leave the block as whole 128-byte row segments: a thread parks the scores of CROSS_TILE consecutive
    // candidates in a shared tile, then the block writes 16 consecutive u64 per query row (a thread's own stores
    // would be one 8-byte piece per row of the matrix).
    constexpr uint32_t TILE_PITCH = CROSS_TILE + 1; // odd pitch: a warp's 32 parked scores fall into 32 banks
    uint32_t *const tile = reinterpret_cast<uint32_t *>(xsm + NT::BYTES); // [T][TILE_PITCH] scores
    unsigned char *const cand_slot = xsm + NT::BYTES + ((T * TILE_PITCH * 4 + 15) & ~15);
    const int tid = threadIdx.x;
    const uint32_t q = blockIdx.x * T + tid;
    const uint32_t n_c = (uint32_t)ps.n_c;
    const uint32_t c_lo = blockIdx.y * cands_per_block, c_hi = min(n_c, c_lo + cands_per_block);
    const uint32_t opaque0 = (uint32_t)(ps.n_pairs >> 63); // run-time 0: see mad_lo in lev_bitpar.cuh
    const uint32_t one = 1u + opaque0, two = 2u + opaque0;
    const uint32_t base_hi = NT::base((uint32_t)__cvta_generic_to_shared(xsm), tid), base_lo = base_hi + NT::TABLE;

    zero_tables<K, T>(base_hi);
    const bool have = q < n_queries;
    uint32_t m = 0;
    const uint8_t *pat = nullptr;
    if (have) {
        const uint64_t o0 = a_off(ps, q), o1 = a_off(ps, q + 1);
        pat = ps.a_data + o0;
        m = o1 < o0 ? 0xFFFFFFFFu : (uint32_t)min(o1 - o0, (uint64_t)0xFFFFFFFFu); // backwards offsets: "does not fit", skipped
    }
    const bool fits = m <= 32u * K; // guaranteed by the host's choice of K; oversize queries are skipped, not corrupted
    bool miss = have && !fits;      // speculative launches report what they skipped (spec_finish)
    uint32_t longest_cand = 0;
    const uint32_t s = 32u * K - (fits ? m : 0u);
    if (have && fits) for_each_byte_global(pat, m, [&](uint32_t i, uint32_t c) { add_symbol<K, T>(base_hi, s + i, c); });
    const bool active = have && fits && m > 0;

    for (uint32_t c = c_lo; c < c_hi; ++c) {
        const uint64_t b0 = b_off(ps, c), b1 = b_off(ps, c + 1);
        const uint8_t *txt = ps.b_data + b0;
        // a speculative launch runs beside the offsets check: it never reads past a length the host has seen (n_cap); what it
        // writes for such a candidate is recomputed, the host re-runs any batch whose statistics exceed the cap
        const uint32_t n = b1 < b0 ? 0u : (uint32_t)min(b1 - b0, (uint64_t)n_cap);
        miss |= b1 < b0 || b1 - b0 > (uint64_t)n_cap;
        longest_cand = max(longest_cand, n);
        uint32_t Pv[K], Mv[K];
        first_column<K>(s, Pv, Mv);
        auto step = [&](uint32_t row_h, uint32_t row_l) {
            uint32_t hi[K], lo[K];
            load_words<K, T>(row_h, hi);
            load_words<K, T>(row_l, lo);
            column_step<K>(hi, lo, Pv, Mv, one, two);
        };
        auto step1 = [&](uint32_t ch) { step(base_hi + (ch >> 4) * NT::ROW, base_lo + (ch & 15u) * NT::ROW); };
        for (uint32_t c0 = 0; c0 < n; c0 += chunk) {
            const uint32_t len = min(n - c0, chunk);
            __syncthreads(); // the previous chunk / candidate has been consumed by every thread
            const uint32_t mis = stage_string(cand_slot, txt + c0, len, tid, T);
            __syncthreads();
            if (active) {
                const unsigned char *cs = cand_slot + mis;
                uint32_t j = 0;
                const uint32_t head = min(len, (4u - (mis & 3u)) & 3u); // up to the next 4-byte boundary of the slot
                for (; j < head; ++j) step1(cs[j]);
                for (; j + 4 <= len; j += 4) {
                    const NibbleWord nw = split_nibbles(*reinterpret_cast<const uint32_t *>(cs + j)); // broadcast: same address in every lane
                    step(row_of<0>(nw.hi, base_hi), row_of<0>(nw.lo, base_lo));
                    step(row_of<1>(nw.hi, base_hi), row_of<1>(nw.lo, base_lo));
                    step(row_of<2>(nw.hi, base_hi), row_of<2>(nw.lo, base_lo));
                    step(row_of<3>(nw.hi, base_hi), row_of<3>(nw.lo, base_lo));
                }
                for (; j < len; ++j) step1(cs[j]);
            }
        }
        uint32_t score = n; // D[m][n] = n + popc(Pv) - popc(Mv) of the final column (free rows carry delta 0)
#pragma unroll
        for (int w = 0; w < K; ++w) score += __popc(Pv[w]) - __popc(Mv[w]);
        // park the score; flush the tile every CROSS_TILE candidates and after the last one
        const uint32_t slot = (c - c_lo) % CROSS_TILE;
        tile[tid * TILE_PITCH + slot] = (have && fits) ? (m == 0 ? n : score) : 0xFFFFFFFFu; // 0xFFFFFFFF: no result for this row
        if (slot == CROSS_TILE - 1 || c + 1 == c_hi) {
            __syncthreads();
            const uint32_t c_first = c - slot, width = slot + 1;
            const uint32_t q_block = blockIdx.x * T;
            // lane -> (row, column): 16 lanes cover one row's segment, a warp covers two rows per turn
            for (uint32_t e = tid; e < (uint32_t)T * CROSS_TILE; e += T) {
                const uint32_t row = e / CROSS_TILE, col = e % CROSS_TILE;
                const uint32_t qq = q_block + row;
                const uint32_t v = tile[row * TILE_PITCH + col];
                // rows of queries that do not fit this kernel's K are left untouched (the host never sends them here)
                if (col < width && qq < n_queries && v != 0xFFFFFFFFu) store_result(ps, (uint64_t)qq * ps.stride + c_first + col, v);
            }
            __syncthreads();
        }
    }
    // statistics in the terms the host keeps them: "m" = the query (pattern), "n" = the candidates this block saw
    spec_finish(sa, miss, have && fits && c_hi > c_lo, m, longest_cand);
}

} // namespace swb
