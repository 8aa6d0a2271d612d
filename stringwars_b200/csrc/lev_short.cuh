// lev_short.cuh — unit-cost Levenshtein for SHORT patterns (<= 24 bytes: BASELINE.json config 1, "words"), sm_100a.
//
// The table kernels (lev_bitpar.cuh) pay a fixed cost per pair that short strings cannot amortise: 32 table rows
// zeroed, two shared-memory read-modify-writes per pattern symbol in dependent chains, two table lookups per column.
// A word pair has at most 16 x 16 cells, so that fixed cost was everything (15 G pairs/s = 4 % of the HBM roofline,
// VERDICT r1 item 5). Here nothing touches shared memory:
//   * one thread per pair, the pattern (shorter string) lives in W <= 6 registers of four bytes each;
//   * the match mask of a text symbol c is derived in registers, four pattern bytes per step (SWAR):
//         t = word ^ splat(c);  flags = ~(((t & 0x7F7F7F7F) + 0x7F7F7F7F) | t) & 0x80808080   (0x80 where byte == c, exact)
//     and two IDP.4A with weights (1,2,4,8) / (16,32,64,128) pack eight flags into one byte, shifted left by 7;
//   * the bit-vector therefore sits at bit 7: the seven rows below it are "free" rows (vertical delta 0, never
//     matching) that replicate DP row 0 exactly as the top-aligned table kernels' free rows do, so column_step is
//     shared unchanged and no shift is spent per column;
//   * garbage rows above the pattern (bytes past its end are whatever followed it in the tape) cannot influence the rows
//     below them (carries only travel upwards) and are masked out of the final popcounts.
// Per column: 3 W + 7 ALU-pipe instructions, 2 W + 6 on the FMA pipe, no LSU work, no dependent memory latency.
// Pair-load phase: offsets and string pieces are read in input order (no permutation), so a warp's 32 pairs read
// contiguous spans of both tapes.
//
// Parity: bit-exact with the CPU oracle, i.e. `rapidfuzz::distance::levenshtein::distance` at
// StringWars similarities/bench.rs:416-419.
#pragma once
#include "lev_bitpar.cuh"

namespace swb {

constexpr int SHORT_THREADS = 128;
constexpr uint32_t SHORT_MAX_M = 24; // 6 words x 4 bytes at bits 7..30

// (c, c, c, c) for byte KB of `word`: IDP.4A picks the byte, IMAD splats it — both on the FMA pipe.
template <int KB>
__device__ __forceinline__ uint32_t splat_byte(uint32_t word, uint32_t ones_mul) {
    return dp4a(word, 1u << (8 * KB), 0u) * ones_mul;
}

// 0x80 in every byte of `word` that equals the splatted symbol (exact for all 256 byte values).
__device__ __forceinline__ uint32_t eq_flags(uint32_t word, uint32_t splat, uint32_t one) {
    const uint32_t t = word ^ splat;
    const uint32_t s = mad_lo(t & 0x7F7F7F7Fu, one, 0x7F7F7F7Fu); // bit 7 of a byte: its low seven bits are non-zero
    return lop3<0x01>(s, t, 0x7F7F7F7Fu);                         // ~(s | t | 0x7F..): bit 7 where the whole byte is zero
}

// Match mask of one text symbol against W pattern words, positioned at bit 7: bit 7 + i is set iff pattern byte i == c.
// Run-time copies of the constants the FMA-pipe instructions multiply by (an immediate power of two would be rewritten
// into ALU-pipe shifts by ptxas, see mad_lo in lev_bitpar.cuh).
struct ShortConsts {
    uint32_t one, two, ones_mul, k256, k65536;
};

template <int W>
__device__ __forceinline__ uint32_t eq_mask7(const uint32_t (&pat)[W], uint32_t splat, const ShortConsts &k) {
    const uint32_t one = k.one;
    uint32_t acc[(W + 1) / 2];
#pragma unroll
    for (int w = 0; w < W; w += 2) {
        uint32_t a = dp4a(eq_flags(pat[w], splat, one), 0x08040201u, 0u);
        if (w + 1 < W) a = dp4a(eq_flags(pat[w + 1], splat, one), 0x80402010u, a);
        acc[w / 2] = a; // (eight match bits) << 7
    }
    uint32_t eq = acc[0];
    if constexpr (W > 2) eq = mad_lo(acc[1], k.k256, eq);
    if constexpr (W > 4) eq = mad_lo(acc[2], k.k65536, eq);
    return eq;
}

template <int W>
__device__ __forceinline__ void short_columns(const uint32_t (&pat)[W], const uint4 &piece, uint32_t cols, uint32_t &Pv, uint32_t &Mv,
                                              const ShortConsts &k) {
    const uint32_t one = k.one, two = k.two, ones_mul = k.ones_mul;
    // `cols` <= 16 symbols of one text piece; warp-uniform trip count is the caller's business (lanes past their text idle)
    auto col = [&](uint32_t splat) {
        const uint32_t eq = eq_mask7<W>(pat, splat, k);
        uint32_t hi[1] = {eq}, lo[1] = {0xFFFFFFFFu}, P[1] = {Pv}, M[1] = {Mv};
        column_step<1>(hi, lo, P, M, one, two);
        Pv = P[0], Mv = M[0];
    };
    const uint32_t words[4] = {piece.x, piece.y, piece.z, piece.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (4u * q >= cols) break;
        const uint32_t w = words[q];
        col(splat_byte<0>(w, ones_mul));
        if (4u * q + 1u < cols) col(splat_byte<1>(w, ones_mul));
        if (4u * q + 2u < cols) col(splat_byte<2>(w, ones_mul));
        if (4u * q + 3u < cols) col(splat_byte<3>(w, ones_mul));
    }
}

// W = pattern words of the launch: 2 (<= 8 bytes), 4 (<= 16), 6 (<= 24).
// LONG_PAT picks which string of a pair is the pattern. false: the SHORTER one (any text length; W covers the class's
// longest shorter string). true: the LONGER one, for batches whose every string fits the registers (W covers the longest
// string of the batch — config 1's words): the columns are then the shorter string's symbols, and a column costs the same
// whatever the pattern length, so a pair of lengths (la, lb) costs min(la, lb) column steps instead of max(la, lb).
// One pair; returns true when the pair was SKIPPED (malformed offsets, a pattern that does not fit W words, a string longer
// than n_cap). m_out / n_out: the pair's shorter / longer length when it was not skipped.
template <int W, bool LONG_PAT>
__device__ __forceinline__ bool lev_short_pair(const PairSpace &ps, uint64_t slot, uint32_t n_cap, uint32_t opaque0, uint32_t &m_out, uint32_t &n_out) {
    const ShortConsts k{1u + opaque0, 2u + opaque0, 0x01010101u + opaque0, 256u + opaque0, 65536u + opaque0};
    const PairRequest rq = pair_request(ps, slot);
    if (!request_ok(rq)) return true;
    const PairView pv = pair_resolve(ps, rq);
    const bool swap = LONG_PAT ? pv.la < pv.lb : pv.la > pv.lb;
    const uint8_t *pat_p = swap ? pv.b : pv.a;
    const uint8_t *txt_p = swap ? pv.a : pv.b;
    const uint32_t m = swap ? pv.lb : pv.la;
    const uint32_t n = swap ? pv.la : pv.lb;
    m_out = min(m, n), n_out = max(m, n);
    if (max(m, n) > n_cap) return true; // a length the host has not seen yet (speculative launch)
    if (min(m, n) == 0) {
        store_result(ps, pv.out_index, max(m, n));
        return false;
    }
    if (m > 4u * W) return true; // not this kernel's class

    // pattern: up to two 16-byte pieces, re-aligned in registers to start at its first byte
    PieceStream pat_stream, txt_stream;
    pat_stream.open(pat_p, m, 0u);
    txt_stream.open(txt_p, n, 0u);
    const uint4 p0 = pat_stream.piece(0);
    uint4 p1 = make_uint4(0, 0, 0, 0);
    if (W > 4 && m > 16u) p1 = pat_stream.piece(1);
    uint4 t_cur = txt_stream.piece(0);
    uint32_t pat[W];
    {
        const uint32_t all[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int w = 0; w < W; ++w) pat[w] = all[w];
    }
    // rows: bits 7 .. 7 + m - 1 carry vertical delta +1; the seven free rows below carry 0
    uint32_t Pv = 0xFFFFFF80u, Mv = 0u;
    const uint32_t pieces = (n + 15u) >> 4;
    for (uint32_t i = 0; i < pieces; ++i) {
        const uint4 cur = t_cur;
        if (i + 1u < pieces) t_cur = txt_stream.piece(i + 1u);
        short_columns<W>(pat, cur, min(16u, n - 16u * i), Pv, Mv, k);
    }
    const uint32_t rows = ((1u << m) - 1u) << 7; // the pattern's rows (m <= 24)
    store_result(ps, pv.out_index, (int64_t)n + __popc(Pv & rows) - __popc(Mv & rows));
    return false;
}

template <int W, bool LONG_PAT>
__global__ void __launch_bounds__(SHORT_THREADS) lev_short_kernel(PairSpace ps, uint64_t begin, uint64_t count, uint32_t n_cap, SpecArgs sa) {
    const uint64_t local = (uint64_t)blockIdx.x * SHORT_THREADS + threadIdx.x;
    const uint32_t opaque0 = (uint32_t)(count >> 63); // run-time 0: see mad_lo in lev_bitpar.cuh
    bool miss = false;
    uint32_t m = 0, n = 0;
    if (local < count) miss = lev_short_pair<W, LONG_PAT>(ps, begin + local, n_cap, opaque0, m, n);
    spec_finish(sa, miss, local < count && !miss, m, n);
}

} // namespace swb
