// lev_bitpar_ws.cuh — the thread-per-pair bit-vector kernel, warp-specialised (sm_100a).
//
// lev_bitpar_kernel runs three phases per pair on the same thread: zero the nibble tables, build them (both bound by the
// shared-memory pipe), then the column loop (bound by the ALU pipe). With seven warps per scheduler the phases of
// different warps overlap only statistically: ncu shows the ALU pipe 70 % busy while the column loop alone saturates it
// with two warps per scheduler (scripts/ubench/recur.cu: 43 cycles per column per scheduler at 2..8 warps).
// Here the phases belong to different warps. A block is four COMPUTE warps and four BUILDER warps; builder warp w fills the
// tables of the next 32 pairs of compute warp w in one of two shared-memory buffers while that compute warp runs the
// column loop out of the other. The hand-over is a pair of mbarriers per (warp, buffer): `full` (tables ready) and `empty`
// (tables consumed). Compute warps therefore issue nothing but the column loop — table lookups and the recurrence — and
// the LSU-bound build overlaps it by construction instead of by chance.
//
// Same tables, same build (build_pieces / add_symbol), same column_step, same results as lev_bitpar_kernel; pairs that do
// not fit the K words (or whose text is longer than n_cap) are skipped by both roles, so the kernel can be launched
// speculatively exactly like lev_bitpar_kernel.
#pragma once
#include "lev_bitpar.cuh"

namespace swb {

constexpr int WS_COMPUTE_WARPS = 4;
constexpr int WS_COMPUTE_THREADS = 32 * WS_COMPUTE_WARPS; // pairs per group
constexpr int WS_THREADS = 2 * WS_COMPUTE_THREADS;

template <int K>
struct WsConfig {
    static constexpr size_t TABLE_BYTES = (size_t)32 * WS_COMPUTE_THREADS * K * sizeof(uint32_t);
    static constexpr size_t SMEM = 2 * TABLE_BYTES;
};

__device__ __forceinline__ uint32_t ws_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_smem(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(ws_smem(bar)), "r"(parity)
                     : "memory");
}

// What both roles derive from a pair's offsets: pattern = shorter string, top-aligned at bit s = 32K - m.
struct WsPair {
    const uint8_t *pat, *txt;
    uint32_t m, n;
    uint64_t out_index;
    bool run; // the pair fits this kernel and has a non-empty pattern: tables are built, columns are run
};
template <int K>
__device__ __forceinline__ WsPair ws_pair(const PairSpace &ps, const PairRequest &rq, uint32_t n_cap) {
    const PairView pv = pair_resolve(ps, rq);
    const bool swap = pv.la > pv.lb;
    WsPair p;
    p.pat = swap ? pv.b : pv.a, p.txt = swap ? pv.a : pv.b;
    p.m = swap ? pv.lb : pv.la, p.n = swap ? pv.la : pv.lb;
    p.out_index = pv.out_index;
    p.run = p.m != 0u && p.m <= 32u * K && p.n <= n_cap;
    return p;
}

// Builder: zero and fill this slot's tables (the build of lev_bitpar_kernel, unchanged).
template <int K>
__device__ __forceinline__ void ws_build(const WsPair &p, uint32_t base_hi, int ctid, uint32_t one) {
    constexpr int T = WS_COMPUTE_THREADS;
    using NT = NibbleTables<K, T>;
    const uint32_t m = p.m, s = 32u * K - m;
    PieceStream pat_stream;
    pat_stream.open(p.pat, m, s & 15u);
    const uint32_t first_piece = s >> 4, p_begin = (s + 15u) & ~15u;
    constexpr uint32_t NQ = (K + 1) / 2;
    constexpr bool ROLLED = K >= 4;
    constexpr int QP = K == 1 ? 2 : 4;
    const uint32_t g = NT::SHARERS == 1 ? 0u : ((uint32_t)ctid / NT::TPR) % NT::SHARERS;
    auto piece_pos = [&](uint32_t q, int pp) -> uint32_t { return 32u * ((2u * q + g + (uint32_t)(pp >> 1)) % K) + 16u * (uint32_t)(pp & 1); };
    auto load_quad = [&](uint32_t q, uint4(&pc)[QP]) {
#pragma unroll
        for (int pp = 0; pp < QP; ++pp) {
            const uint32_t pos = piece_pos(q, pp);
            if (pos >= p_begin) pc[pp] = pat_stream.piece((pos >> 4) - first_piece);
        }
    };
    auto build_quad = [&](uint32_t q, const uint4(&pc)[QP]) {
        uint32_t hb[QP], exists[QP];
        bool any = false;
#pragma unroll
        for (int pp = 0; pp < QP; ++pp) {
            const uint32_t pos = piece_pos(q, pp);
            hb[pp] = base_hi + NT::word(pos >> 5) + ((pos >> 3) & 2u);
            exists[pp] = pos >= p_begin ? one : 0u;
            any = any || pos >= p_begin;
        }
        if (__any_sync(__activemask(), any)) build_pieces<QP, ROLLED>(hb, pc, NT::TABLE, exists);
    };
    uint4 quad_a[QP], quad_b[QP];
#pragma unroll
    for (int pp = 0; pp < QP; ++pp) quad_a[pp] = quad_b[pp] = make_uint4(0, 0, 0, 0);
    uint4 pat_first = make_uint4(0, 0, 0, 0);
    if (s & 15u) pat_first = pat_stream.piece(0);
    load_quad(0, quad_a);
    zero_tables<K, T>(base_hi);
    for (uint32_t pos = s; pos < min(p_begin, 32u * K); ++pos) add_symbol<K, T>(base_hi, pos, piece_byte(pat_first, pos & 15u));
#pragma unroll 1
    for (uint32_t q = 0; q < NQ; ++q) {
        if (q + 1 < NQ) load_quad(q + 1, quad_b);
        build_quad(q, quad_a);
#pragma unroll
        for (int pp = 0; pp < QP; ++pp) quad_a[pp] = quad_b[pp];
    }
}

// The text's first two aligned pieces, requested before the tables are waited for.
struct WsText {
    PieceStream stream;
    uint4 a0, a1;
};
__device__ __forceinline__ WsText ws_open_text(const WsPair &p) {
    WsText t;
    t.stream.open(p.txt, p.n, 0u);
    t.a0 = t.stream.aligned(0), t.a1 = t.stream.aligned(1);
    return t;
}

// Compute: the column loop of lev_bitpar_kernel over tables somebody else has built. Returns the distance.
template <int K>
__device__ __forceinline__ uint32_t ws_columns(const WsPair &p, WsText &t, uint32_t base_hi, uint32_t one, uint32_t two) {
    constexpr int T = WS_COMPUTE_THREADS;
    using NT = NibbleTables<K, T>;
    const uint32_t base_lo = base_hi + NT::TABLE;
    const uint32_t n = p.n, s = 32u * K - p.m;
    const uint32_t txt_aligned = (t.stream.hi + 15u) >> 4;
    const uint8_t *txt_next = t.stream.grid + 32;
    uint32_t Pv[K], Mv[K];
    first_column<K>(s, Pv, Mv);
    auto step = [&](uint32_t row_h, uint32_t row_l) {
        uint32_t hi[K], lo[K];
        load_words<K, T>(row_h, hi);
        load_words<K, T>(row_l, lo);
        column_step<K>(hi, lo, Pv, Mv, one, two);
    };
    auto step4 = [&](uint32_t word) {
        const NibbleWord nw = split_nibbles(word);
        step(row_of<0>(nw.hi, base_hi), row_of<0>(nw.lo, base_lo));
        step(row_of<1>(nw.hi, base_hi), row_of<1>(nw.lo, base_lo));
        step(row_of<2>(nw.hi, base_hi), row_of<2>(nw.lo, base_lo));
        step(row_of<3>(nw.hi, base_hi), row_of<3>(nw.lo, base_lo));
    };
    auto step1 = [&](uint32_t c) { step(base_hi + (c >> 4) * NT::ROW, base_lo + (c & 15u) * NT::ROW); };
    const uint32_t full = n >> 4, rem = n & 15u;
#pragma unroll 1
    for (uint32_t i = 0; i < full; ++i) {
        const uint4 cur = t.stream.combine(t.a0, t.a1);
        t.a0 = t.a1;
        if (i + 2u < txt_aligned) t.a1 = ldg16(txt_next);
        txt_next += 16;
        if constexpr (K >= 4) {
#pragma unroll 1
            for (uint32_t w = 0; w < 4u; ++w) step4((w & 2u) ? ((w & 1u) ? cur.w : cur.z) : ((w & 1u) ? cur.y : cur.x));
        } else {
            step4(cur.x);
            step4(cur.y);
            step4(cur.z);
            step4(cur.w);
        }
    }
    if (rem) {
        const uint4 cur = t.stream.combine(t.a0, t.a1);
#pragma unroll 1
        for (uint32_t w = 0; 4u * w < rem; ++w) {
            uint32_t word = (w & 2u) ? ((w & 1u) ? cur.w : cur.z) : ((w & 1u) ? cur.y : cur.x);
            if (4u * w + 4u <= rem) {
                step4(word);
            } else {
                for (uint32_t k = 4u * w; k < rem; ++k, word >>= 8) step1(word & 0xFFu);
            }
        }
    }
    uint32_t score = n;
#pragma unroll
    for (int w = 0; w < K; ++w) score += __popc(Pv[w]) - __popc(Mv[w]);
    return score;
}

// groups = ceil(count / 128); a block takes groups blockIdx.x, blockIdx.x + gridDim.x, ...
template <int K>
__global__ void __launch_bounds__(WS_THREADS, 3) lev_bitpar_ws_kernel(PairSpace ps, uint64_t begin, uint64_t count, uint32_t n_cap, uint32_t groups) {
    using NT = NibbleTables<K, WS_COMPUTE_THREADS>;
    extern __shared__ __align__(16) unsigned char ws_tbl[]; // two table buffers of WsConfig<K>::TABLE_BYTES
    __shared__ __align__(8) uint64_t bar_full[WS_COMPUTE_WARPS][2], bar_empty[WS_COMPUTE_WARPS][2];
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool builder = warp >= WS_COMPUTE_WARPS;
    const int w = warp & (WS_COMPUTE_WARPS - 1), ctid = tid & (WS_COMPUTE_THREADS - 1);
    if (tid < 2 * WS_COMPUTE_WARPS) {
        mbar_init(&bar_full[tid >> 1][tid & 1], 32);
        mbar_init(&bar_empty[tid >> 1][tid & 1], 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t opaque0 = (uint32_t)(count >> 63); // run-time 0: see mad_lo in lev_bitpar.cuh
    const uint32_t one = 1u + opaque0, two = 2u + opaque0;
    const uint32_t smem0 = ws_smem(ws_tbl);
    const uint32_t base[2] = {NT::base(smem0, ctid), NT::base(smem0 + (uint32_t)WsConfig<K>::TABLE_BYTES, ctid)};
    // The offsets of a group are requested one iteration ahead of their first use: with three or four compute warps per
    // scheduler nothing else would hide that load (and the string loads that depend on it) between two pairs.
    uint32_t it = 0;
    PairRequest rq_next;
    rq_next.a0 = rq_next.a1 = rq_next.b0 = rq_next.b1 = rq_next.out_index = 0;
    if ((uint64_t)blockIdx.x * WS_COMPUTE_THREADS + ctid < count) rq_next = pair_request(ps, begin + (uint64_t)blockIdx.x * WS_COMPUTE_THREADS + ctid);
    for (uint32_t g = blockIdx.x; g < groups; g += gridDim.x, ++it) {
        const uint32_t b = it & 1u, use = it >> 1; // this is the use-th time buffer b is filled / drained
        const uint64_t local = (uint64_t)g * WS_COMPUTE_THREADS + ctid;
        const bool have = local < count;
        const PairRequest rq = rq_next;
        const uint64_t local_next = local + (uint64_t)gridDim.x * WS_COMPUTE_THREADS;
        if (g + gridDim.x < groups && local_next < count) rq_next = pair_request(ps, begin + local_next);
        WsPair p;
        p.run = false, p.m = 1u, p.n = 0u, p.out_index = 0;
        if (have) p = ws_pair<K>(ps, rq, n_cap);
        if (builder) {
            if (use > 0) mbar_wait(&bar_empty[w][b], (use - 1u) & 1u); // the compute warp is done with the previous tables
            if (p.run) ws_build<K>(p, base[b], ctid, one);
            mbar_arrive(&bar_full[w][b]);
        } else {
            WsText text;
            if (p.run) text = ws_open_text(p); // in flight while the tables are waited for
            mbar_wait(&bar_full[w][b], use & 1u);
            if (p.run) store_result(ps, p.out_index, ws_columns<K>(p, text, base[b], one, two));
            else if (have && p.m == 0u) store_result(ps, p.out_index, p.n);
            mbar_arrive(&bar_empty[w][b]);
        }
    }
}

} // namespace swb
