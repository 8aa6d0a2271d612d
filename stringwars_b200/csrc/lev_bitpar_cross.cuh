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

template <int K>
__global__ void __launch_bounds__(BitparConfig<K>::THREADS)
    lev_bitpar_cross_kernel(PairSpace ps, uint32_t n_queries, uint32_t cands_per_block, uint32_t chunk) {
    constexpr int T = BitparConfig<K>::THREADS;
    constexpr uint32_t ESTRIDE = T * K * 4;
    extern __shared__ __align__(16) unsigned char xsm[]; // [32 entries][T][K] u32 | candidate chunk (chunk + 32 B)
    unsigned char *const tbl_hi = xsm, *const tbl_lo = xsm + 16 * ESTRIDE;
    unsigned char *const cand_slot = xsm + 32 * ESTRIDE;
    const int tid = threadIdx.x;
    const uint32_t q = blockIdx.x * T + tid;
    const uint32_t n_c = (uint32_t)ps.n_c;
    const uint32_t c_lo = blockIdx.y * cands_per_block, c_hi = min(n_c, c_lo + cands_per_block);

#pragma unroll 8
    for (int e = 0; e < 32; ++e) {
#pragma unroll
        for (int w = 0; w < K; ++w) *reinterpret_cast<uint32_t *>(xsm + e * ESTRIDE + word_offset<K, T>(tid, w)) = 0u;
    }
    const bool have = q < n_queries;
    uint32_t m = 0;
    const uint8_t *pat = nullptr;
    if (have) {
        const uint64_t o0 = ps.a_offs[q], o1 = ps.a_offs[q + 1];
        pat = ps.a_data + o0;
        m = (uint32_t)(o1 - o0);
    }
    const bool fits = m <= 32u * K; // guaranteed by the host's choice of K; oversize queries are skipped, not corrupted
    const uint32_t s = 32u * K - (fits ? m : 0u);
    if (have && fits)
        for_each_byte_global(pat, m, [&](uint32_t i, uint32_t c) {
            const uint32_t pos = s + i, bit = 1u << (pos & 31u), word_off = word_offset<K, T>(tid, pos >> 5);
            *reinterpret_cast<uint32_t *>(tbl_hi + (c >> 4) * ESTRIDE + word_off) |= bit;
            *reinterpret_cast<uint32_t *>(tbl_lo + (c & 15u) * ESTRIDE + word_off) |= bit;
        });
    uint32_t Pv0[K];
#pragma unroll
    for (int w = 0; w < K; ++w) {
        const uint32_t lo_bit = 32u * w;
        Pv0[w] = s >= lo_bit + 32u ? 0u : (s <= lo_bit ? 0xFFFFFFFFu : 0xFFFFFFFFu << (s - lo_bit));
    }
    const bool active = have && fits && m > 0;

    for (uint32_t c = c_lo; c < c_hi; ++c) {
        const uint64_t b0 = ps.b_offs[c], b1 = ps.b_offs[c + 1];
        const uint8_t *txt = ps.b_data + b0;
        const uint32_t n = (uint32_t)(b1 - b0);
        uint32_t Pv[K], Mv[K];
#pragma unroll
        for (int w = 0; w < K; ++w) Pv[w] = Pv0[w], Mv[w] = 0u;
        auto step = [&](uint32_t h, uint32_t l) {
            uint32_t hi[K], lo[K];
            load_words<K, T>(tbl_hi + h * ESTRIDE, tid, hi);
            load_words<K, T>(tbl_lo + l * ESTRIDE, tid, lo);
            uint32_t hp_in = 1u, hn_in = 0u;
#pragma unroll
            for (int w = 0; w < K; ++w) {
                const uint32_t eqc = w == 0 ? (hi[w] & lo[w]) : ((hi[w] & lo[w]) | hn_in);
                const uint32_t xv = (hi[w] & lo[w]) | Mv[w];
                const uint32_t xh = (((eqc & Pv[w]) + Pv[w]) ^ Pv[w]) | eqc;
                const uint32_t ph = Mv[w] | ~(xh | Pv[w]);
                const uint32_t mh = Pv[w] & xh;
                const uint32_t phs = ph * 2u + hp_in;
                const uint32_t mhs = mh * 2u + hn_in;
                Pv[w] = mhs | ~(xv | phs);
                Mv[w] = phs & xv;
                if (w + 1 < K) hp_in = ph >> 31, hn_in = mh >> 31;
            }
        };
        for (uint32_t c0 = 0; c0 < n; c0 += chunk) {
            const uint32_t len = min(n - c0, chunk);
            __syncthreads(); // the previous chunk / candidate has been consumed by every thread
            const uint32_t mis = stage_string(cand_slot, txt + c0, len, tid, T);
            __syncthreads();
            if (active) {
                const unsigned char *cs = cand_slot + mis;
                uint32_t j = 0;
                const uint32_t head = min(len, (4u - (mis & 3u)) & 3u); // up to the next 4-byte boundary of the slot
                for (; j < head; ++j) {
                    const uint32_t ch = cs[j];
                    step(ch >> 4, ch & 15u);
                }
                for (; j + 4 <= len; j += 4) {
                    const uint32_t word = *reinterpret_cast<const uint32_t *>(cs + j); // broadcast: same address in every lane
                    const uint32_t wh = (word >> 4) & 0x0F0F0F0Fu, wl = word & 0x0F0F0F0Fu;
                    step(wh & 0xFFu, wl & 0xFFu);
                    step((wh >> 8) & 0xFFu, (wl >> 8) & 0xFFu);
                    step((wh >> 16) & 0xFFu, (wl >> 16) & 0xFFu);
                    step(wh >> 24, wl >> 24);
                }
                for (; j < len; ++j) {
                    const uint32_t ch = cs[j];
                    step(ch >> 4, ch & 15u);
                }
            }
        }
        uint32_t score = n; // D[m][n] = n + popc(Pv) - popc(Mv) of the final column (free rows carry delta 0)
#pragma unroll
        for (int w = 0; w < K; ++w) score += __popc(Pv[w]) - __popc(Mv[w]);
        if (have && fits) ps.out[(uint64_t)q * ps.stride + c] = m == 0 ? n : score;
    }
}

} // namespace swb
