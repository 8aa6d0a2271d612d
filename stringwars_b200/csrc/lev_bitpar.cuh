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

namespace swb {

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

// K table words of the entry that starts at `entry` (shared memory) for thread `tid`.
template <int K, int T>
__device__ __forceinline__ void load_words(const unsigned char *entry, int tid, uint32_t (&w)[K]) {
    if constexpr (K == 1) {
        w[0] = *reinterpret_cast<const uint32_t *>(entry + tid * 4);
    } else if constexpr (K == 2) {
        const uint2 v = *reinterpret_cast<const uint2 *>(entry + tid * 8);
        w[0] = v.x, w[1] = v.y;
    } else {
#pragma unroll
        for (int q = 0; q < K / 4; ++q) {
            const uint4 v = *reinterpret_cast<const uint4 *>(entry + (K == 4 ? 0 : q * T * 16) + tid * 16);
            w[4 * q + 0] = v.x, w[4 * q + 1] = v.y, w[4 * q + 2] = v.z, w[4 * q + 3] = v.w;
        }
    }
}

// Table build for TWO aligned 16-symbol pieces that live in DIFFERENT table words (and, per piece, the HI and the
// LO table): four read-modify-write chains through shared memory that can never touch the same address, so each
// symbol step loads all four words first and stores them afterwards (the compiler cannot prove the independence
// itself and otherwise serialises LDS -> OR -> STS four times; a single chain was 49 % of this kernel's stall
// samples, profiles/r1_d_*).
template <uint32_t ESTRIDE>
__device__ __forceinline__ void build_piece_pair(unsigned char *__restrict__ hi_a, unsigned char *__restrict__ lo_a,
                                                 unsigned char *__restrict__ hi_b, unsigned char *__restrict__ lo_b, const uint4 piece_a,
                                                 const uint4 piece_b, const uint32_t bit0) {
    const uint32_t wa[4] = {piece_a.x, piece_a.y, piece_a.z, piece_a.w}, wb[4] = {piece_b.x, piece_b.y, piece_b.z, piece_b.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t ah = (wa[q] >> 4) & 0x0F0F0F0Fu, al = wa[q] & 0x0F0F0F0Fu;
        const uint32_t bh = (wb[q] >> 4) & 0x0F0F0F0Fu, bl = wb[q] & 0x0F0F0F0Fu;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t bit = bit0 * (1u << (4 * q + k));
            uint32_t *const p0 = reinterpret_cast<uint32_t *>(hi_a + ((ah >> (8 * k)) & 0xFFu) * ESTRIDE);
            uint32_t *const p1 = reinterpret_cast<uint32_t *>(lo_a + ((al >> (8 * k)) & 0xFFu) * ESTRIDE);
            uint32_t *const p2 = reinterpret_cast<uint32_t *>(hi_b + ((bh >> (8 * k)) & 0xFFu) * ESTRIDE);
            uint32_t *const p3 = reinterpret_cast<uint32_t *>(lo_b + ((bl >> (8 * k)) & 0xFFu) * ESTRIDE);
            // the four addresses are distinct by construction (two tables x two words): load all, then store all
            const uint32_t v0 = *p0, v1 = *p1, v2 = *p2, v3 = *p3;
            *p0 = v0 | bit, *p1 = v1 | bit, *p2 = v2 | bit, *p3 = v3 | bit;
        }
    }
}

template <uint32_t ESTRIDE>
__device__ __forceinline__ void build_piece(unsigned char *__restrict__ hi_a, unsigned char *__restrict__ lo_a, const uint4 piece_a,
                                            const uint32_t bit0) {
    const uint32_t wa[4] = {piece_a.x, piece_a.y, piece_a.z, piece_a.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t ah = (wa[q] >> 4) & 0x0F0F0F0Fu, al = wa[q] & 0x0F0F0F0Fu;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t bit = bit0 * (1u << (4 * q + k));
            uint32_t *const p0 = reinterpret_cast<uint32_t *>(hi_a + ((ah >> (8 * k)) & 0xFFu) * ESTRIDE);
            uint32_t *const p1 = reinterpret_cast<uint32_t *>(lo_a + ((al >> (8 * k)) & 0xFFu) * ESTRIDE);
            const uint32_t v0 = *p0, v1 = *p1;
            *p0 = v0 | bit, *p1 = v1 | bit;
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

    // Zero this thread's 32 x K table words.
#pragma unroll 8
    for (int e = 0; e < 32; ++e) {
#pragma unroll
        for (int w = 0; w < K; ++w) *reinterpret_cast<uint32_t *>(tbl_bytes + e * ESTRIDE + word_offset<K, T>(tid, w)) = 0u;
    }
    if (local >= count) return;

    const PairView pv = pair_at(ps, begin + local);
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
            // so the pieces from position s+i up to 32K are all full. Pieces 32 positions apart sit in different
            // words: they are built in interleaved pairs (four independent shared-memory update chains).
            const uint32_t p_begin = s + i, p_end = 32u * K;
#pragma unroll 1
            for (uint32_t half = 0; half < 32u; half += 16u) {
                uint32_t pos = (p_begin & ~31u) + half;
                if (pos < p_begin) pos += 32u;
                const uint32_t bit0 = 1u << half;
#pragma unroll 1
                for (; pos + 32u < p_end; pos += 64u) {
                    const uint4 piece_a = ldg16(pat + (pos - s)), piece_b = ldg16(pat + (pos + 32u - s));
                    const uint32_t off_a = word_offset<K, T>(tid, pos >> 5), off_b = word_offset<K, T>(tid, (pos >> 5) + 1u);
                    build_piece_pair<ESTRIDE>(tbl_hi + off_a, tbl_lo + off_a, tbl_hi + off_b, tbl_lo + off_b, piece_a, piece_b, bit0);
                }
                if (pos < p_end) {
                    const uint32_t off_a = word_offset<K, T>(tid, pos >> 5);
                    build_piece<ESTRIDE>(tbl_hi + off_a, tbl_lo + off_a, ldg16(pat + (pos - s)), bit0);
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
    auto step = [&](uint32_t h, uint32_t l) {
        uint32_t hi[K], lo[K];
        load_words<K, T>(tbl_hi + h * ESTRIDE, tid, hi);
        load_words<K, T>(tbl_lo + l * ESTRIDE, tid, lo);
        uint32_t hp_in = 1u, hn_in = 0u; // horizontal delta entering word 0: +1 (DP row 0 is 0,1,2,...)
#pragma unroll
        for (int w = 0; w < K; ++w) {
            const uint32_t eqc = w == 0 ? (hi[w] & lo[w]) : ((hi[w] & lo[w]) | hn_in);
            const uint32_t xv = (hi[w] & lo[w]) | Mv[w];
            const uint32_t xh = (((eqc & Pv[w]) + Pv[w]) ^ Pv[w]) | eqc;
            const uint32_t ph = Mv[w] | ~(xh | Pv[w]);
            const uint32_t mh = Pv[w] & xh;
            const uint32_t phs = ph * 2u + hp_in; // (ph << 1) | carry, on the FMA pipe
            const uint32_t mhs = mh * 2u + hn_in;
            Pv[w] = mhs | ~(xv | phs);
            Mv[w] = phs & xv;
            if (w + 1 < K) hp_in = ph >> 31, hn_in = mh >> 31; // carries into the next word; none leave the last one
        }
    };
    auto step4 = [&](uint32_t word) {
        const uint32_t wh = (word >> 4) & 0x0F0F0F0Fu, wl = word & 0x0F0F0F0Fu;
        step(wh & 0xFFu, wl & 0xFFu);
        step((wh >> 8) & 0xFFu, (wl >> 8) & 0xFFu);
        step((wh >> 16) & 0xFFu, (wl >> 16) & 0xFFu);
        step(wh >> 24, wl >> 24);
    };
    {
        uint32_t j = 0;
        const uint32_t head = min(n, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(txt) & 15u)) & 15u));
        for (; j < head; ++j) {
            const uint32_t c = ldg1(txt + j);
            step(c >> 4, c & 15u);
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
            step(c >> 4, c & 15u);
        }
    }
    uint32_t score = n;
#pragma unroll
    for (int w = 0; w < K; ++w) score += __popc(Pv[w]) - __popc(Mv[w]);
    ps.out[pv.out_index] = score;
}

} // namespace swb
