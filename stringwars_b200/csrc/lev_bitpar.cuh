// lev_bitpar.cuh — unit-cost Levenshtein by the Myers (1999) / Hyyrö (2003) bit-vector recurrence, sm_100a.
//
// Why this exists next to the wavefront kernel: with unit costs the vertical/horizontal DP deltas are in
// {-1, 0, +1}, so 32 DP rows fit two 32-bit masks and one column step costs ~10 LOP3/SHF per 32 cells
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
// Instruction-level design (the kernel is ALU-pipe bound, ncu: profiles/r1_*):
//  * The pattern is TOP-aligned: symbol i sits at bit s+i, s = 32K-m. The s rows below it are "free" rows
//    (vertical delta 0, never matching) that replicate DP row 0, so the result is unchanged and no bit index
//    depends on m at run time.
//  * The distance is not tracked per column at all: D[m][n] = n + popc(Pv) - popc(Mv) of the final column.
//  * `x << 1 | carry` is written as x*2+carry and the adds as mul-adds: they issue on the FMA pipe (IMAD),
//    which is otherwise idle, instead of the saturated ALU pipe.
//  * Symbols are nibble-split four at a time ((w >> 4) & 0x0F0F0F0F, w & 0x0F0F0F0F); a table address is one
//    byte extract plus one IMAD.
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
// x * mul + carry_in as one 64-bit multiply-add (IMAD.WIDE.U32, FMA pipe): with mul = 2 the low half is
// (x << 1) | carry_in and the high half is the bit that left the word = the carry into the next word. `mul` is
// a run-time value the assembler cannot see through, otherwise ptxas rewrites the multiply into ALU-pipe
// LEA / LEA.HI.X / SHF, and the ALU pipe is the one this kernel saturates.
__device__ __forceinline__ void mad_wide(uint32_t x, uint32_t mul, uint32_t carry_in, uint32_t &lo, uint32_t &hi) {
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(x), "r"(mul), "l"((uint64_t)carry_in));
    lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
}
__device__ __forceinline__ void mul_wide(uint32_t x, uint32_t mul, uint32_t &lo, uint32_t &hi) {
    uint64_t r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x), "r"(mul));
    lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t x, uint32_t mul, uint32_t add) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(mul), "r"(add));
    return r;
}

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

template <int K>
struct BitparConfig {
    static constexpr int THREADS = K <= 4 ? 128 : 64; // 128 * K bytes of tables per thread
};

// Byte offset of word w of THIS thread inside a table entry. K <= 4: the thread's K words are contiguous
// ([thread][K]); K = 8: two 4-word parts, each part [thread][4], so that every LDS.128 of a warp covers 512
// contiguous bytes (a [thread][8] layout would put 32-byte strides under 16-byte loads: 2-way bank conflicts).
template <int K, int T>
__device__ __forceinline__ uint32_t word_offset(int tid, uint32_t w) {
    if constexpr (K <= 4) return (uint32_t)tid * (K * 4) + w * 4u;
    else return (w >> 2) * (T * 16u) + (uint32_t)tid * 16u + (w & 3u) * 4u;
}

// K table words of THIS thread from the entry at 32-bit shared-space address `addr` (entry base + the thread's
// word_offset): explicit ld.shared so the address is one 32-bit register plus an immediate.
template <int K, int T>
__device__ __forceinline__ void load_words(uint32_t addr, uint32_t (&w)[K]) {
    if constexpr (K == 1) {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[0]) : "r"(addr));
    } else if constexpr (K == 2) {
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(addr));
    } else {
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
        if constexpr (K == 8)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(addr), "n"(T * 16));
    }
}
// Same through a generic pointer to the entry (lev_bitpar_cross.cuh).
template <int K, int T>
__device__ __forceinline__ void load_words(const unsigned char *entry, int tid, uint32_t (&w)[K]) {
    load_words<K, T>((uint32_t)__cvta_generic_to_shared(entry) + word_offset<K, T>(tid, 0), w);
}

// Table build for aligned 16-symbol pieces. A piece fills one 16-bit half of one table word; per symbol and table it
// is one read-modify-write of shared memory: AND (isolate the nibble), IMAD (scale it by the entry stride and add
// the table base), LDS, add the position bit (each position bit is set exactly once per table, so OR == ADD and
// the add can issue on either pipe), STS. `BuildMul` carries the run-time multipliers (see the column loop: with
// immediates ptxas turns these IMADs into ALU-pipe LEAs).
struct BuildMul {
    uint32_t mul0, mul4, mul8, two; // 2^E, 2^(E-4), 2^(E-8) for nibbles at bit 0, 4, 8; and 2
};
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// Shared addresses of the entries selected by the 4 symbols of one 32-bit pattern word: hi[k] in the HI table
// (based at base_hi), lo[k] in the LO table (based at base_lo).
__device__ __forceinline__ void symbol_slots(uint32_t word, uint32_t base_hi, uint32_t base_lo, const BuildMul &bm, uint32_t (&hi)[4],
                                             uint32_t (&lo)[4]) {
    const uint32_t w12 = __umulhi(word, 1u << 20), w24 = __umulhi(word, 1u << 8); // word >> 12, word >> 24 (IMAD.HI)
    hi[0] = mad_lo(word & 0xF0u, bm.mul4, base_hi), lo[0] = mad_lo(word & 0xFu, bm.mul0, base_lo);
    hi[1] = mad_lo(w12 & 0xFu, bm.mul0, base_hi), lo[1] = mad_lo(word & 0xF00u, bm.mul8, base_lo);
    hi[2] = mad_lo(w12 & 0xF00u, bm.mul8, base_hi), lo[2] = mad_lo(w12 & 0xF0u, bm.mul4, base_lo);
    hi[3] = mad_lo(w24 & 0xF0u, bm.mul4, base_hi), lo[3] = mad_lo(w24 & 0xFu, bm.mul0, base_lo);
}

// Two pieces that live in DIFFERENT table words: four read-modify-write chains that can never touch the same
// address (two tables x two words), so each symbol step loads all four words first and stores them afterwards.
// base_*: shared address of this thread's word in entry 0 of the HI / LO table, for piece a and piece b.
__device__ __forceinline__ void build_piece_pair(uint32_t hi_a, uint32_t lo_a, uint32_t hi_b, uint32_t lo_b, const uint4 piece_a,
                                                 const uint4 piece_b, uint32_t bit, const BuildMul &bm) {
    const uint32_t wa[4] = {piece_a.x, piece_a.y, piece_a.z, piece_a.w}, wb[4] = {piece_b.x, piece_b.y, piece_b.z, piece_b.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t ah[4], al[4], bh[4], bl[4];
        symbol_slots(wa[q], hi_a, lo_a, bm, ah, al);
        symbol_slots(wb[q], hi_b, lo_b, bm, bh, bl);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t v0 = lds32(ah[k]), v1 = lds32(al[k]), v2 = lds32(bh[k]), v3 = lds32(bl[k]);
            sts32(ah[k], v0 + bit), sts32(al[k], v1 + bit), sts32(bh[k], v2 + bit), sts32(bl[k], v3 + bit);
            bit = mad_lo(bit, bm.two, 0u);
        }
    }
}

__device__ __forceinline__ void build_piece(uint32_t hi_a, uint32_t lo_a, const uint4 piece_a, uint32_t bit, const BuildMul &bm) {
    const uint32_t wa[4] = {piece_a.x, piece_a.y, piece_a.z, piece_a.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t ah[4], al[4];
        symbol_slots(wa[q], hi_a, lo_a, bm, ah, al);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t v0 = lds32(ah[k]), v1 = lds32(al[k]);
            sts32(ah[k], v0 + bit), sts32(al[k], v1 + bit);
            bit = mad_lo(bit, bm.two, 0u);
        }
    }
}

// A thread computes any pair whose pattern fits its K words and SKIPS (writes nothing for) a longer one: that is
// what lets the host launch this kernel speculatively, before the batch's length statistics are known — the
// statistics then tell the host whether anything was skipped (engine.cu, run_device_batch).
template <int K>
__global__ void __launch_bounds__(BitparConfig<K>::THREADS) lev_bitpar_kernel(PairSpace ps, uint64_t begin, uint64_t count) {
    constexpr int T = BitparConfig<K>::THREADS;
    constexpr uint32_t ESTRIDE = T * K * 4; // bytes between consecutive table entries
    extern __shared__ __align__(16) unsigned char tbl_bytes[]; // [32 entries][T][K] u32
    const int tid = threadIdx.x;
    const uint64_t local = (uint64_t)blockIdx.x * T + tid;
    unsigned char *const tbl_hi = tbl_bytes, *const tbl_lo = tbl_bytes + 16 * ESTRIDE;
    // 1, 2 and the nibble scales as values ptxas cannot fold (count < 2^63 always): with immediates it rewrites the
    // multiply-adds below into ALU-pipe LEA / SHF, and the ALU pipe is the one this kernel saturates.
    const uint32_t opaque0 = (uint32_t)(count >> 63);
    const uint32_t one = 1u + opaque0, two = 2u + opaque0;
    constexpr int E = K == 1 ? 9 : (K == 2 ? 10 : 11);
    static_assert(ESTRIDE == (1u << E), "nibble offsets assume a power-of-two entry stride");
    const BuildMul bm = {(1u << E) + opaque0, (1u << (E - 4)) + opaque0, (1u << (E - 8)) + opaque0, two};
    const uint32_t tbl_s = (uint32_t)__cvta_generic_to_shared(tbl_bytes);
    const uint32_t base_hi = tbl_s + word_offset<K, T>(tid, 0), base_lo = base_hi + 16u * ESTRIDE;

    // The pair's offsets are requested first; the table zeroing below runs while they are in flight.
    const bool live = local < count;
    PairView pv = {nullptr, nullptr, 0u, 0u, 0u};
    if (live) pv = pair_at(ps, begin + local);

    // Zero this thread's 32 x K table words.
#pragma unroll 8
    for (int e = 0; e < 32; ++e) {
#pragma unroll
        for (int w = 0; w < K; ++w) *reinterpret_cast<uint32_t *>(tbl_bytes + e * ESTRIDE + word_offset<K, T>(tid, w)) = 0u;
    }
    if (!live) return;

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
    if (m > 32u * K) return; // not this kernel's class (speculative launch on a batch that changed shape)
    const uint32_t s = 32u * K - m; // pattern symbol i lives at bit position s + i

    // ---- build the nibble tables: bit p of HI[h] (LO[l]) is set iff the symbol at position p has high (low) nibble h (l)
    auto add_symbol = [&](uint32_t pos, uint32_t c) {
        const uint32_t bit = 1u << (pos & 31u);
        const uint32_t word_off = word_offset<K, T>(tid, pos >> 5);
        uint32_t *hi = reinterpret_cast<uint32_t *>(tbl_hi + (c >> 4) * ESTRIDE + word_off);
        uint32_t *lo = reinterpret_cast<uint32_t *>(tbl_lo + (c & 15u) * ESTRIDE + word_off);
        *hi |= bit;
        *lo |= bit;
    };
    {
        uint32_t i = 0;
        const uint32_t head = min(m, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(pat) & 15u)) & 15u));
        for (; i < head; ++i) add_symbol(s + i, ldg1(pat + i));
        if (((s + i) & 15u) == 0u) {
            // Aligned pieces: 16 symbols share one table word and one 16-bit half of it. The pattern is top-aligned,
            // so every piece from position p_begin up to 32K is full. Pieces are built two at a time, one from word a
            // and one from word a+1 (four independent update chains), and the lanes of a warp start at DIFFERENT
            // words: with [thread][K] slots a warp's 32-bit accesses to one word index hit only 32/K' banks (K' =
            // min(K, 4) lanes share a bank), so lane group g = (lane / (32/K')) % K' works on word (2i + g) % K while
            // group 0 works on word 2i. In a batch of equal lengths every table access is then conflict-free.
            const uint32_t p_begin = s + i;
            constexpr uint32_t KP = K < 4 ? K : 4;
            const uint32_t g = KP == 1 ? 0u : ((uint32_t)tid / (32u / KP)) % KP;
#pragma unroll 1
            for (uint32_t it = 0; it < (K + 1) / 2; ++it) {
                const uint32_t wa = (2u * it + g) % K, wb = (wa + 1u) % K;
                const uint32_t off_a = word_offset<K, T>(0, wa), off_b = word_offset<K, T>(0, wb);
#pragma unroll 1
                for (uint32_t half = 0; half < 32u; half += 16u) {
                    const uint32_t pos_a = 32u * wa + half, pos_b = 32u * wb + half;
                    const bool act_a = pos_a >= p_begin, act_b = K > 1 && pos_b >= p_begin;
                    const uint32_t bit0 = 1u << half;
                    if (act_a && act_b) {
                        const uint4 piece_a = ldg16(pat + (pos_a - s)), piece_b = ldg16(pat + (pos_b - s));
                        build_piece_pair(base_hi + off_a, base_lo + off_a, base_hi + off_b, base_lo + off_b, piece_a, piece_b, bit0, bm);
                    } else if (act_a) {
                        build_piece(base_hi + off_a, base_lo + off_a, ldg16(pat + (pos_a - s)), bit0, bm);
                    } else if (act_b) {
                        build_piece(base_hi + off_b, base_lo + off_b, ldg16(pat + (pos_b - s)), bit0, bm);
                    }
                }
            }
            i = m;
        } else {
            for (; i + 16 <= m; i += 16) {
                const uint4 piece = ldg16(pat + i);
                const uint32_t w[4] = {piece.x, piece.y, piece.z, piece.w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int k = 0; k < 4; ++k) add_symbol(s + i + 4 * q + k, (w[q] >> (8 * k)) & 0xFFu);
            }
        }
        for (; i < m; ++i) add_symbol(s + i, ldg1(pat + i));
    }

    // ---- column loop
    uint32_t Pv[K], Mv[K];
#pragma unroll
    for (int w = 0; w < K; ++w) {
        // rows below the pattern (bit < s): vertical delta 0; pattern rows: +1 (column 0 of the DP matrix)
        const uint32_t lo_bit = 32u * w;
        Pv[w] = s >= lo_bit + 32u ? 0u : (s <= lo_bit ? 0xFFFFFFFFu : 0xFFFFFFFFu << (s - lo_bit));
        Mv[w] = 0u;
    }
    // No score is tracked per column: D[m][n] = D[0][n] + (sum of the vertical deltas of column n)
    //                                          = n + popc(Pv) - popc(Mv) at the end (free rows carry delta 0).
    // Hyyro's single-mask form of the recurrence over the whole 32K-bit column at once:
    // D0 = (((Eq & Pv) + Pv) ^ Pv) | Eq | Mv marks the cells whose diagonal delta is 0 and replaces both of Myers'
    // Xv and Xh, so Eq = HI & LO is never materialised (it only appears inside two 3-input LOP3s): 7 logic ops per
    // word. The add is ONE 32K-bit add (a carry chain, 1 op per word) instead of Myers' per-word blocks, so no
    // word waits for the deltas of the word above it and no block carry has to be folded into Eq.
    auto step = [&](uint32_t addr_h, uint32_t addr_l) {
        uint32_t hi[K], lo[K], t[K], x[K], sum[K];
        load_words<K, T>(addr_h, hi);
        load_words<K, T>(addr_l, lo);
#pragma unroll
        for (int w = 0; w < K; ++w) {
            t[w] = lop3<0x80>(hi[w], lo[w], Pv[w]); // Eq & Pv
            x[w] = lop3<0xEA>(hi[w], lo[w], Mv[w]); // Eq | Mv
        }
        add_wide<K>(t, Pv, sum, one);
        uint32_t ph_below = 0u, mh_below = 0u;
#pragma unroll
        for (int w = 0; w < K; ++w) {
            const uint32_t d0 = lop3<0xBE>(sum[w], Pv[w], x[w]); // ((t + Pv) ^ Pv) | x
            const uint32_t ph = lop3<0xF1>(Mv[w], d0, Pv[w]);    // Mv | ~(D0 | Pv)
            const uint32_t mh = Pv[w] & d0;
            // shifted deltas: word 0 takes the horizontal delta of DP row 0 (+1) as x * 2 + 1 / x * 2 (IMAD, FMA pipe);
            // the words above take the bit leaving the word below through one funnel shift (SHF.L.W)
            const uint32_t phs = w == 0 ? mad_lo(ph, two, one) : __funnelshift_l(ph_below, ph, 1);
            const uint32_t mhs = w == 0 ? mad_lo(mh, two, 0u) : __funnelshift_l(mh_below, mh, 1);
            ph_below = ph, mh_below = mh;
            Pv[w] = lop3<0xF1>(mhs, d0, phs); // Mh' | ~(D0 | Ph')
            Mv[w] = phs & d0;
        }
    };
    // Shared address of this thread's words in the table entry selected by the nibble at bit P (P = 0, 4, 8) of
    // `word`: base + ((word >> P) & 15) * ESTRIDE = one AND (the only ALU-pipe op) and one IMAD that scales the
    // isolated nibble by 2^(E-P) and adds the base. The multipliers are run-time values so that ptxas keeps the
    // IMAD (with an immediate it emits LEA, ALU pipe). Nibbles above bit 11 are first brought down by two
    // multiply-high shifts per 32-bit word (word >> 12, word >> 24: IMAD.HI, FMA pipe).
    const uint32_t mul0 = bm.mul0, mul4 = bm.mul4, mul8 = bm.mul8;
    auto step4 = [&](uint32_t word) {
        const uint32_t w12 = __umulhi(word, 1u << 20), w24 = __umulhi(word, 1u << 8); // word >> 12, word >> 24
        step(mad_lo(word & 0xF0u, mul4, base_hi), mad_lo(word & 0xFu, mul0, base_lo));
        step(mad_lo(w12 & 0xFu, mul0, base_hi), mad_lo(word & 0xF00u, mul8, base_lo));
        step(mad_lo(w12 & 0xF00u, mul8, base_hi), mad_lo(w12 & 0xF0u, mul4, base_lo));
        step(mad_lo(w24 & 0xF0u, mul4, base_hi), mad_lo(w24 & 0xFu, mul0, base_lo));
    };
    auto step1 = [&](uint32_t c) { step(base_hi + (c >> 4) * ESTRIDE, base_lo + (c & 15u) * ESTRIDE); };
    {
        uint32_t j = 0;
        const uint32_t head = min(n, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(txt) & 15u)) & 15u));
        for (; j < head; ++j) {
            const uint32_t c = ldg1(txt + j);
            step1(c);
        }
        if (j + 16 <= n) {
            uint4 nxt = ldg16(txt + j);
            while (j + 16 <= n) {
                const uint4 cur = nxt;
                if (j + 32 <= n) nxt = ldg16(txt + j + 16);
                step4(cur.x);
                step4(cur.y);
                step4(cur.z);
                step4(cur.w);
                j += 16;
            }
        }
        for (; j < n; ++j) {
            const uint32_t c = ldg1(txt + j);
            step1(c);
        }
    }
    uint32_t score = n;
#pragma unroll
    for (int w = 0; w < K; ++w) score += __popc(Pv[w]) - __popc(Mv[w]);
    ps.out[pv.out_index] = score;
}

} // namespace swb
