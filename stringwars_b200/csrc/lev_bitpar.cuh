// lev_bitpar.cuh — unit-cost Levenshtein by the Myers (1999) / Hyyrö (2003) bit-vector recurrence, sm_100a.
//
// Why this exists next to the wavefront kernel: with unit costs the vertical/horizontal DP deltas are in
// {-1, 0, +1}, so 32 DP rows fit two 32-bit masks and one column step costs ~9 integer ops per 32 cells
// instead of ~4-5 integer instructions per cell. Integer ALU issue is the binding roofline of this path
// (DESIGN.md §4), so this is the kernel that moves the HBM-roofline fraction.
//
// Mapping: one THREAD per pair, K 32-bit words of vertical state (pattern = the shorter string, <= 32*K
// symbols). The match mask Eq(c) is never stored for 256 symbols (2 KB/word would kill occupancy): it is
// rebuilt per column from two 16-entry nibble tables, Eq(c) = HI[c >> 4] & LO[c & 15], because a pattern
// byte equals c iff both nibbles do. The tables live in shared memory (NibbleTables below). Bytes stream
// straight from HBM with aligned 16-byte read-only loads, one piece prefetched (thread-private strings: nothing
// to share, so no smem staging here — the group-per-pair wavefront kernel is the one that stages).
//
// Instruction-level design. The kernel is bound by the ALU pipe (LOP3 / SHF / IADD3: one warp instruction per two
// cycles per scheduler); measured on a B200 (scripts/ubench/): the FMA pipe (IMAD, IDP.4A) issues beside it at
// the same rate, IMAD.WIDE / IMAD.HI run at a quarter rate and stall the ALU stream, so they are avoided.
//  * The pattern is TOP-aligned: symbol i sits at bit s+i, s = 32K-m. The s rows below it are "free" rows
//    (vertical delta 0, never matching) that replicate DP row 0, so the result is unchanged and no bit index
//    depends on m at run time.
//  * The distance is not tracked per column at all: D[m][n] = n + popc(Pv) - popc(Mv) of the final column.
//  * Hyyrö's single-mask form, D0 = (((Eq & Pv) + Pv) ^ Pv) | Eq | Mv, over the whole 32K-bit column: 7 LOP3 per
//    word (Eq = HI & LO only ever appears inside 3-input LOP3s), one carry chain for the add, one funnel shift per
//    shifted word above word 0 (column_step).
//  * A table address is ONE instruction: the four symbols of a text word are nibble-split together
//    (w & 0x0F0F0F0F, (w >> 4) & 0x0F0F0F0F) and IDP.4A with a one-hot byte selector of 128 computes
//    base + nibble * 128 — which is why a table entry row is 128 bytes.
//  * The table build is 16-bit read-modify-writes in up to eight independent chains, lanes rotated over the words so
//    that no two lanes of a warp meet in a bank (build_pieces).
//
// Parity: bit-exact with the CPU oracle (both of its algorithm families), the restatement of
// `rapidfuzz::distance::levenshtein::distance` as called at StringWars similarities/bench.rs:416-419.
#pragma once
#include "common.cuh"
#include <type_traits>

namespace swb {

// One LOP3 with an explicit truth table (a = 0xF0, b = 0xCC, c = 0xAA). Written as PTX so the front end does
// not "share" the sub-expression HI & LO between two 3-input ops (its CSE turns 2 LOP3 into 3).
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}
// x * mul + add as one IMAD (FMA pipe). Callers pass `mul` as a run-time value the assembler cannot see through:
// with an immediate power of two ptxas rewrites the multiply into ALU-pipe LEA / SHF, the pipe this kernel saturates.
__device__ __forceinline__ uint32_t mad_lo(uint32_t x, uint32_t mul, uint32_t add) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(mul), "r"(add));
    return r;
}
// sum of byte_i(a) * byte_i(b) + c (IDP.4A, FMA pipe, full rate).
__device__ __forceinline__ uint32_t dp4a(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// sum = a + b over K 32-bit words (little-endian), one carry chain.
template <int K>
__device__ __forceinline__ void add_wide(const uint32_t (&a)[K], const uint32_t (&b)[K], uint32_t (&s)[K], uint32_t one) {
    if constexpr (K == 1) {
        s[0] = mad_lo(a[0], one, b[0]);
    } else if constexpr (K == 2) {
        asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(s[0]), "=r"(s[1]) : "r"(a[0]), "r"(a[1]), "r"(b[0]), "r"(b[1]));
    } else if constexpr (K == 4) {
        asm("add.cc.u32 %0, %4, %8;\n\taddc.cc.u32 %1, %5, %9;\n\taddc.cc.u32 %2, %6, %10;\n\taddc.u32 %3, %7, %11;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    } else {
        static_assert(K == 8, "add_wide: K must be 1, 2, 4 or 8");
        asm("add.cc.u32 %0, %8, %16;\n\taddc.cc.u32 %1, %9, %17;\n\taddc.cc.u32 %2, %10, %18;\n\taddc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\taddc.cc.u32 %5, %13, %21;\n\taddc.cc.u32 %6, %14, %22;\n\taddc.u32 %7, %15, %23;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]),
              "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    }
}

#ifndef SWB_BITPAR_THREADS_SMALL
#define SWB_BITPAR_THREADS_SMALL 64 // threads per block for K <= 4: measured 64 > 128 > 256 (+2 % / -2 %) at 64x64: finer blocks mix the phases
#endif
#ifndef SWB_BITPAR_THREADS_K8
#define SWB_BITPAR_THREADS_K8 64
#endif
template <int K>
struct BitparConfig {
    static constexpr int THREADS = K <= 4 ? SWB_BITPAR_THREADS_SMALL : SWB_BITPAR_THREADS_K8; // 128 * K bytes of tables per thread; K = 8: 32 = 64 > 128
    static constexpr int PAIRS_PER_THREAD = 1;
    static constexpr size_t SMEM = (size_t)32 * THREADS * K * sizeof(uint32_t);
};

// Shared-memory layout of the nibble tables of a block of T threads with K words each.
// A thread's K words are split into PARTS parts of at most 4 words (one LDS.32/64/128 each); a ROW is the 128
// bytes that TPR consecutive threads' slots of one part make, and an entry of a table is one row:
//     [group of TPR threads][part][table HI|LO][16 entries][TPR threads][SB bytes]
// so (a) a warp's lookup is conflict-free (each quarter/half/full warp reads one contiguous row per wavefront),
// and (b) the address of entry n is base + n * 128: one IDP.4A from a packed nibble word.
template <int K, int T>
struct NibbleTables {
    static constexpr int PARTS = K <= 4 ? 1 : K / 4;
    static constexpr int SB = (K < 4 ? K : 4) * 4; // slot bytes per thread and part
    static constexpr int TPR = 128 / SB;           // threads per row
    static constexpr uint32_t ROW = 128, TABLE = 16 * ROW;
    static constexpr uint32_t PART_STRIDE = 2 * TABLE;
    static constexpr uint32_t BYTES = (uint32_t)T * K * 4 * 32;
    // lanes of a warp whose slots share a bank (they differ only in the group): the build rotates them over words
    static constexpr uint32_t SHARERS = 32 / TPR;
    // shared address of this thread's part-0 slot in entry 0 of the HI table; LO is + TABLE, part p is + p * PART_STRIDE
    __device__ static __forceinline__ uint32_t base(uint32_t smem, int tid) {
        return smem + (uint32_t)(tid / TPR) * (PARTS * PART_STRIDE) + (uint32_t)(tid % TPR) * SB;
    }
    // byte offset of word w inside the thread's slots (relative to base)
    __device__ static __forceinline__ uint32_t word(uint32_t w) { return (w >> 2) * PART_STRIDE + (w & 3u) * 4u; }
};

// K table words of THIS thread from the entry row at 32-bit shared-space address `addr` (base + n * 128).
template <int K, int T>
__device__ __forceinline__ void load_words(uint32_t addr, uint32_t (&w)[K]) {
    if constexpr (K == 1) {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[0]) : "r"(addr));
    } else if constexpr (K == 2) {
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(addr));
    } else {
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
        if constexpr (K == 8)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];"
                         : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                         : "r"(addr), "n"(NibbleTables<K, T>::PART_STRIDE));
    }
}

// Zero this thread's 32 x K table words.
template <int K, int T>
__device__ __forceinline__ void zero_tables(uint32_t base) {
    using NT = NibbleTables<K, T>;
#pragma unroll
    for (int p = 0; p < NT::PARTS; ++p)
#pragma unroll 8
        for (int e = 0; e < 32; ++e) { // 16 HI rows, then 16 LO rows
            const uint32_t addr = base + p * NT::PART_STRIDE + e * NT::ROW;
            if constexpr (K == 1) asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(0u) : "memory");
            else if constexpr (K == 2) asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(addr), "r"(0u) : "memory");
            else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
        }
}

// One DP column: Hyyrö's single-mask form of the recurrence over the whole 32K-bit column at once.
// D0 = (((Eq & Pv) + Pv) ^ Pv) | Eq | Mv marks the cells whose diagonal delta is 0 and replaces both of Myers' Xv and
// Xh, so Eq = HI & LO is never materialised (it only appears inside two 3-input LOP3s): 7 logic ops per word. The
// add is ONE 32K-bit add (a carry chain, 1 op per word) instead of Myers' per-word blocks, so no word waits for the
// deltas of the word above it and no block carry has to be folded into Eq. Shifted deltas: word 0 takes the
// horizontal delta entering the block as x * 2 + carry (IMAD, FMA pipe); the words above take the bit leaving the
// word below through one funnel shift (SHF.L.W). `one`, `two`: run-time 1 and 2 (see mad_lo).
//
// CARRY = false: the block starts at DP row 0, whose horizontal delta is always +1 (thread-per-pair kernels).
// CARRY = true: the block sits below another one (lane-group kernel): hp_in / hn_in are the +1 / -1 horizontal delta
// bits leaving the block above; a -1 acts as a match in the block's first row (Myers' block rule, one more LOP3).
// ph_top / mh_top return the top word's unshifted deltas: their bit 31 is what leaves this block.
template <int K, bool CARRY>
__device__ __forceinline__ void column_step_io(const uint32_t (&hi)[K], const uint32_t (&lo)[K], uint32_t (&Pv)[K], uint32_t (&Mv)[K],
                                               uint32_t one, uint32_t two, uint32_t hp_in, uint32_t hn_in, uint32_t &ph_top,
                                               uint32_t &mh_top) {
    uint32_t t[K], x[K], sum[K];
#pragma unroll
    for (int w = 0; w < K; ++w) {
        if (CARRY && w == 0) {
            const uint32_t eqc = lop3<0xEA>(hi[w], lo[w], hn_in); // Eq | (-1 entering row 0)
            t[w] = eqc & Pv[w];
            x[w] = eqc | Mv[w];
        } else {
            t[w] = lop3<0x80>(hi[w], lo[w], Pv[w]); // Eq & Pv
            x[w] = lop3<0xEA>(hi[w], lo[w], Mv[w]); // Eq | Mv
        }
    }
    add_wide<K>(t, Pv, sum, one);
    uint32_t ph_below = 0u, mh_below = 0u;
#pragma unroll
    for (int w = 0; w < K; ++w) {
        const uint32_t d0 = lop3<0xBE>(sum[w], Pv[w], x[w]); // ((t + Pv) ^ Pv) | x
        const uint32_t ph = lop3<0xF1>(Mv[w], d0, Pv[w]);    // Mv | ~(D0 | Pv)
        const uint32_t mh = Pv[w] & d0;
        const uint32_t phs = w == 0 ? mad_lo(ph, two, CARRY ? hp_in : one) : __funnelshift_l(ph_below, ph, 1);
        const uint32_t mhs = w == 0 ? mad_lo(mh, two, CARRY ? hn_in : 0u) : __funnelshift_l(mh_below, mh, 1);
        ph_below = ph, mh_below = mh;
        Pv[w] = lop3<0xF1>(mhs, d0, phs); // Mh' | ~(D0 | Ph')
        Mv[w] = phs & d0;
    }
    ph_top = ph_below, mh_top = mh_below;
}
template <int K>
__device__ __forceinline__ void column_step(const uint32_t (&hi)[K], const uint32_t (&lo)[K], uint32_t (&Pv)[K], uint32_t (&Mv)[K],
                                            uint32_t one, uint32_t two) {
    uint32_t ph_top, mh_top;
    column_step_io<K, false>(hi, lo, Pv, Mv, one, two, 0u, 0u, ph_top, mh_top);
}

// Column 0 of the DP matrix for a top-aligned pattern whose first symbol sits at global bit position s, for the K
// words that start at global bit position `first`: rows below the pattern (bit < s) carry vertical delta 0, pattern
// rows +1.
template <int K>
__device__ __forceinline__ void first_column(uint32_t s, uint32_t (&Pv)[K], uint32_t (&Mv)[K], uint32_t first = 0u) {
#pragma unroll
    for (int w = 0; w < K; ++w) {
        const uint32_t lo_bit = first + 32u * w;
        Pv[w] = s >= lo_bit + 32u ? 0u : (s <= lo_bit ? 0xFFFFFFFFu : 0xFFFFFFFFu << (s - lo_bit));
        Mv[w] = 0u;
    }
}

// The four symbols of a 32-bit word, nibble-split: byte k of `hi` / `lo` is the high / low nibble of symbol k.
struct NibbleWord {
    uint32_t hi, lo;
};
__device__ __forceinline__ NibbleWord split_nibbles(uint32_t word) { return {(word >> 4) & 0x0F0F0F0Fu, word & 0x0F0F0F0Fu}; }
// Address of the entry row selected by byte KB of a packed nibble word: base + nibble * 128 (one IDP.4A).
template <int KB>
__device__ __forceinline__ uint32_t row_of(uint32_t nibbles, uint32_t base) {
    return dp4a(nibbles, 128u << (8 * KB), base);
}

// Table build for aligned 16-symbol pieces. A piece fills one 16-bit half of one table word; per symbol and table it
// is one 16-bit read-modify-write of shared memory: IDP.4A (entry address), LDS.U16, add the position bit (each
// position bit is set exactly once per table, so OR == ADD, and the add is an IMAD: FMA pipe), STS.U16. Pieces of different words or different
// halves can never touch the same address, so NP pieces are 2*NP independent update chains: each symbol step loads
// all 2*NP half-words first and stores them afterwards (four pieces = the whole pattern at K = 2 = eight chains,
// 16 dependent steps instead of 64).
// hi_base[p]: shared address of the 16-bit half piece p fills, in entry 0 of the HI table (LO is + table_bytes).
// one[p]: run-time 1 for a piece that exists, 0 for one that lies below the pattern (its zero symbols then add nothing:
// lanes with shorter patterns stay in step with their warp instead of branching).
// ROLLED keeps the loop over the four 32-bit words of a piece as a loop (4x less code; the K >= 4 kernels are
// instruction-cache bound otherwise: 1.5 warps per scheduler walking 70 KB of code).
template <int NP, bool ROLLED>
__device__ __forceinline__ void build_pieces(const uint32_t (&hi_base)[NP], const uint4 (&piece)[NP], uint32_t table_bytes,
                                             const uint32_t (&one)[NP]) {
    auto word_steps = [&](auto q_static, uint32_t q) {
        NibbleWord nw[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) nw[p] = split_nibbles(q == 0 ? piece[p].x : (q == 1 ? piece[p].y : (q == 2 ? piece[p].z : piece[p].w)));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t ah[NP], al[NP], vh[NP], vl[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint32_t lo_base = hi_base[p] + table_bytes;
                if (k == 0) ah[p] = row_of<0>(nw[p].hi, hi_base[p]), al[p] = row_of<0>(nw[p].lo, lo_base);
                if (k == 1) ah[p] = row_of<1>(nw[p].hi, hi_base[p]), al[p] = row_of<1>(nw[p].lo, lo_base);
                if (k == 2) ah[p] = row_of<2>(nw[p].hi, hi_base[p]), al[p] = row_of<2>(nw[p].lo, lo_base);
                if (k == 3) ah[p] = row_of<3>(nw[p].hi, hi_base[p]), al[p] = row_of<3>(nw[p].lo, lo_base);
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) vh[p] = lds16(ah[p]), vl[p] = lds16(al[p]);
            // position bit 4q + k of the half: a compile-time constant when unrolled, (16^q) << k when rolled
            const uint32_t bit = decltype(q_static)::value >= 0 ? (1u << (4 * decltype(q_static)::value + k)) : ((1u << k) << (4u * q));
#pragma unroll
            for (int p = 0; p < NP; ++p) sts16(ah[p], mad_lo(one[p], bit, vh[p])), sts16(al[p], mad_lo(one[p], bit, vl[p]));
        }
    };
    if constexpr (ROLLED) {
#pragma unroll 1
        for (uint32_t q = 0; q < 4; ++q) word_steps(std::integral_constant<int, -1>{}, q);
    } else {
        word_steps(std::integral_constant<int, 0>{}, 0u);
        word_steps(std::integral_constant<int, 1>{}, 1u);
        word_steps(std::integral_constant<int, 2>{}, 2u);
        word_steps(std::integral_constant<int, 3>{}, 3u);
    }
}

// One symbol into the tables (ragged heads and tails, unaligned patterns): bit `pos` of HI[c >> 4] and LO[c & 15].
template <int K, int T>
__device__ __forceinline__ void add_symbol(uint32_t base, uint32_t pos, uint32_t c) {
    using NT = NibbleTables<K, T>;
    const uint32_t bit = 1u << (pos & 31u);
    const uint32_t ah = base + NT::word(pos >> 5) + (c >> 4) * NT::ROW, al = base + NT::TABLE + NT::word(pos >> 5) + (c & 15u) * NT::ROW;
    const uint32_t vh = lds32(ah);
    sts32(ah, vh | bit);
    const uint32_t vl = lds32(al);
    sts32(al, vl | bit);
}

// A thread computes any pair whose pattern fits its K words and SKIPS (writes nothing for) a longer one: that is
// what lets the host launch this kernel speculatively, before the batch's length statistics are known — the
// statistics then tell the host whether anything was skipped (engine.cu, run_device_batch).
// One pair; returns true when the pair was SKIPPED (malformed offsets, a pattern beyond 32 K symbols, a text longer than
// n_cap). m_out / n_out: the pair's shorter / longer length when it was not skipped.
template <int K>
__device__ __forceinline__ bool lev_bitpar_pair(const PairSpace &ps, uint64_t begin, uint64_t local, uint32_t n_cap, uint32_t opaque0, uint32_t &m_out,
                                                uint32_t &n_out) {
    constexpr int T = BitparConfig<K>::THREADS;
    using NT = NibbleTables<K, T>;
    extern __shared__ __align__(16) unsigned char tbl_bytes[]; // NibbleTables<K, T>::BYTES
    const int tid = threadIdx.x;
    // 1 and 2 as values ptxas cannot fold: see mad_lo
    const uint32_t one = 1u + opaque0, two = 2u + opaque0;
    const uint32_t base_hi = NT::base((uint32_t)__cvta_generic_to_shared(tbl_bytes), tid), base_lo = base_hi + NT::TABLE;

    const PairRequest pr = pair_request(ps, begin + local);
#ifdef SWB_BITPAR_ZERO_FIRST
    zero_tables<K, T>(base_hi); // offsets in flight while the tables are zeroed
#endif
    if (!request_ok(pr)) return true;
    const PairView pv = pair_resolve(ps, pr);
    // pattern = shorter string (distance is symmetric), text = longer one
    const bool swap = pv.la > pv.lb;
    const uint8_t *pat = swap ? pv.b : pv.a;
    const uint8_t *txt = swap ? pv.a : pv.b;
    const uint32_t m = swap ? pv.lb : pv.la;
    const uint32_t n = swap ? pv.la : pv.lb;
    m_out = m, n_out = n;
    // a text longer than anything the host has seen (a speculative launch validates as it goes: it must not chase an
    // unchecked length), or not this kernel's class (speculative launch on a batch that changed shape)
    if (n > n_cap) return true;
    if (m == 0) {
        store_result(ps, pv.out_index, n);
        return false;
    }
    if (m > 32u * K) return true;
    const uint32_t s = 32u * K - m; // pattern symbol i lives at bit position s + i

    // ---- requests that can leave before the tables are zeroed
    // Pattern: streamed as 16-symbol pieces aligned to the TABLE positions (piece i = positions [16 i, 16 i + 16)),
    // whatever the alignment of the bytes in memory (PieceStream): 16 symbols share one table word and one 16-bit
    // half of it, and because the pattern is top-aligned every piece from position p_begin up to 32K is full; the
    // at most 15 symbols before p_begin go in one by one. The table is built a QUAD at a time: the four pieces of
    // words a and a+1 (64 pattern bytes). Lanes whose slots share a bank (NT::SHARERS of them per warp) start at
    // different words: lane group g works on words (2q + g) % K and +1 while group 0 works on words 2q and 2q+1, so
    // in a batch of equal lengths every access of the build is conflict-free.
    PieceStream pat_stream;
    pat_stream.open(pat, m, s & 15u);
    const uint32_t first_piece = s >> 4;         // table piece that holds position s
    const uint32_t p_begin = (s + 15u) & ~15u;   // first position of the first full piece
    constexpr uint32_t NQ = (K + 1) / 2; // quads
    constexpr bool ROLLED = K >= 4;       // trade a few selects for 4x less code where the instruction cache binds
#ifndef SWB_ROLLED_MAIN_MIN_K
#define SWB_ROLLED_MAIN_MIN_K 4
#endif
    constexpr bool ROLLED_MAIN = K >= SWB_ROLLED_MAIN_MIN_K;
    constexpr int QP = K == 1 ? 2 : 4;   // pieces per quad
    const uint32_t g = NT::SHARERS == 1 ? 0u : ((uint32_t)tid / NT::TPR) % NT::SHARERS;
    // piece p of quad q: word (2q + g + p/2) % K, half p % 2
    auto piece_pos = [&](uint32_t q, int p) -> uint32_t { return 32u * ((2u * q + g + (uint32_t)(p >> 1)) % K) + 16u * (uint32_t)(p & 1); };
    auto load_quad = [&](uint32_t q, uint4 (&pc)[QP]) {
#pragma unroll
        for (int p = 0; p < QP; ++p) {
            const uint32_t pos = piece_pos(q, p);
            if (pos >= p_begin) pc[p] = pat_stream.piece((pos >> 4) - first_piece);
        }
    };
    auto build_quad = [&](uint32_t q, const uint4 (&pc)[QP]) {
        uint32_t hb[QP], exists[QP];
        bool any = false;
#pragma unroll
        for (int p = 0; p < QP; ++p) {
            const uint32_t pos = piece_pos(q, p);
            hb[p] = base_hi + NT::word(pos >> 5) + ((pos >> 3) & 2u);
            exists[p] = pos >= p_begin ? one : 0u;
            any = any || pos >= p_begin;
        }
        // skipped only if no lane of the warp has a piece in this quad (warp-uniform: no divergence in ragged batches)
        if (__any_sync(__activemask(), any)) build_pieces<QP, ROLLED>(hb, pc, NT::TABLE, exists);
    };
    uint4 quad_a[QP], quad_b[QP];
#pragma unroll
    for (int p = 0; p < QP; ++p) quad_a[p] = quad_b[p] = make_uint4(0, 0, 0, 0);
    uint4 pat_first = make_uint4(0, 0, 0, 0);
    if (s & 15u) pat_first = pat_stream.piece(0); // positions [16 first_piece, p_begin): its bytes s % 16 .. 15 are symbols
    load_quad(0, quad_a);
    // Text: streamed from its first byte, two aligned pieces ahead of the columns.
    PieceStream txt_stream;
    txt_stream.open(txt, n, 0u);
    const uint32_t txt_aligned = (txt_stream.hi + 15u) >> 4; // aligned pieces that hold text bytes
    uint4 txt_a0 = txt_stream.aligned(0), txt_a1 = txt_stream.aligned(1);
    const uint8_t *txt_next = txt_stream.grid + 32;          // the aligned piece to request next

#ifndef SWB_BITPAR_ZERO_FIRST
    zero_tables<K, T>(base_hi);
#endif

    // ---- build the nibble tables: bit p of HI[h] (LO[l]) is set iff the symbol at position p has high (low) nibble h (l)
    for (uint32_t pos = s; pos < min(p_begin, 32u * K); ++pos) add_symbol<K, T>(base_hi, pos, piece_byte(pat_first, pos & 15u));
    // the next quad's pieces are in flight while the current quad is built
#pragma unroll 1
    for (uint32_t q = 0; q < NQ; ++q) {
        if (q + 1 < NQ) load_quad(q + 1, quad_b);
        build_quad(q, quad_a);
#pragma unroll
        for (int p = 0; p < QP; ++p) quad_a[p] = quad_b[p];
    }

    // ---- column loop. No score is tracked per column: D[m][n] = D[0][n] + (sum of the vertical deltas of column n)
    //                                                            = n + popc(Pv) - popc(Mv) at the end.
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
        const uint4 cur = txt_stream.combine(txt_a0, txt_a1);
        txt_a0 = txt_a1;
        if (i + 2u < txt_aligned) txt_a1 = ldg16(txt_next); // whole or partial; nothing is requested past the text
        txt_next += 16;
        if constexpr (ROLLED_MAIN) {
#pragma unroll 1
            for (uint32_t w = 0; w < 4u; ++w) step4((w & 2u) ? ((w & 1u) ? cur.w : cur.z) : ((w & 1u) ? cur.y : cur.x));
        } else {
            step4(cur.x);
            step4(cur.y);
            step4(cur.z);
            step4(cur.w);
        }
    }
    if (rem) { // the last, partial piece: whole words, then single symbols
        const uint4 cur = txt_stream.combine(txt_a0, txt_a1);
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
    store_result(ps, pv.out_index, score);
    return false;
}

template <int K>
__global__ void __launch_bounds__(BitparConfig<K>::THREADS) lev_bitpar_kernel(PairSpace ps, uint64_t begin, uint64_t count, uint32_t n_cap, SpecArgs sa) {
    const uint64_t local = (uint64_t)blockIdx.x * BitparConfig<K>::THREADS + threadIdx.x;
    const uint32_t opaque0 = (uint32_t)(count >> 63); // run-time 0 (count < 2^63 always)
    bool miss = false;
    uint32_t m = 0, n = 0;
    if (local < count) miss = lev_bitpar_pair<K>(ps, begin, local, n_cap, opaque0, m, n);
    spec_finish(sa, miss, local < count && !miss, m, n);
}

} // namespace swb
