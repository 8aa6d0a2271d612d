// utf8.cuh — device-side split of UTF-8 tapes into 32-bit symbols for the codepoint-level engine.
//
// `LevenshteinDistancesUtf8` (StringWars similarities/bench.rs:386-399) and the rapidfuzz baseline over
// `str::chars()` (bench.rs:437) compare Unicode scalar values. Only EQUALITY of symbols matters for an edit
// distance, so the device does not decode to scalar values: a symbol is the big-endian packing of a lead byte and
// up to three continuation bytes, which is a bijection with scalar values on valid UTF-8. A symbol starts at
// every byte that is not 10xxxxxx; stray leading continuation bytes are ignored (same rule as the CPU oracle).
#pragma once
#include "common.cuh"
#include <type_traits>

namespace swb {

__device__ __forceinline__ uint64_t tape_offset(const void *offs, uint32_t offs32, uint64_t i) {
    return offs32 ? (uint64_t) reinterpret_cast<const uint32_t *>(offs)[i] : reinterpret_cast<const uint64_t *>(offs)[i];
}

// Lead-byte bits of one 32-bit word: bit k is set iff byte k is NOT a continuation byte (10xxxxxx).
__device__ __forceinline__ uint32_t lead_bits4(uint32_t w) {
    const uint32_t cont = (w >> 7) & ~(w >> 6) & 0x01010101u; // bit 0 of byte k: byte k is 10xxxxxx
    const uint32_t lead = cont ^ 0x01010101u;
    return ((lead * 0x01020408u) >> 24) & 0xFu; // gather bits 0, 8, 16, 24 into a nibble
}
// 16 lead bits of a 16-byte piece (bit j: byte j starts a symbol).
__device__ __forceinline__ uint32_t lead_bits16(const uint4 &p) {
    return lead_bits4(p.x) | (lead_bits4(p.y) << 4) | (lead_bits4(p.z) << 8) | (lead_bits4(p.w) << 12);
}

// One warp per string: bytes_out[i + 1] = 4 * (number of symbols of string i); bytes_out[0] = 0.
// `ascii_flag` is cleared (atomicAnd) when any byte >= 0x80 is seen. Lanes read re-aligned 16-byte pieces (PieceStream).
__global__ void __launch_bounds__(256) utf8_count_kernel(const uint8_t *data, const void *offs, uint32_t offs32, uint64_t n_strings,
                                                         uint64_t *bytes_out, uint32_t *ascii_flag) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) bytes_out[0] = 0;
    if (warp >= n_strings) return;
    const uint64_t o0 = tape_offset(offs, offs32, warp), o1 = tape_offset(offs, offs32, warp + 1);
    const uint32_t len = o1 < o0 ? 0u : (uint32_t)(o1 - o0);
    PieceStream st;
    st.open(data + o0, len, 0u);
    uint32_t count = 0, high = 0;
    for (uint32_t i = lane; 16u * i < len; i += 32) {
        const uint4 p = st.piece(i);
        const uint32_t valid = len - 16u * i >= 16u ? 0xFFFFu : ((1u << (len - 16u * i)) - 1u);
        count += __popc(lead_bits16(p) & valid);
        const uint32_t hb = ((((p.x >> 7) & 0x01010101u) * 0x01020408u) >> 24 & 0xFu) | (((((p.y >> 7) & 0x01010101u) * 0x01020408u) >> 24 & 0xFu) << 4) |
                            (((((p.z >> 7) & 0x01010101u) * 0x01020408u) >> 24 & 0xFu) << 8) | (((((p.w >> 7) & 0x01010101u) * 0x01020408u) >> 24 & 0xFu) << 12);
        high |= hb & valid;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) count += __shfl_xor_sync(0xffffffffu, count, d);
    if (__any_sync(0xffffffffu, high != 0u) && lane == 0) atomicAnd(ascii_flag, 0u);
    if (lane == 0) bytes_out[warp + 1] = 4ull * count;
}

// One warp per string: writes the packed symbols of string i at symbols + sym_offs[i] / 4.
__global__ void __launch_bounds__(256) utf8_pack_kernel(const uint8_t *data, const void *offs, uint32_t offs32, uint64_t n_strings,
                                                        const uint64_t *sym_offs, uint32_t *symbols) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_strings) return;
    const uint64_t o0 = tape_offset(offs, offs32, warp), o1 = tape_offset(offs, offs32, warp + 1);
    const uint8_t *s = data + o0;
    const uint32_t len = o1 < o0 ? 0u : (uint32_t)(o1 - o0);
    uint32_t *out = symbols + (sym_offs[warp] >> 2);
    uint32_t written = 0;
    for (uint32_t base = 0; base < len; base += 32) {
        const uint32_t p = base + lane;
        const uint32_t b = p < len ? ldg1(s + p) : 0x80u;
        const bool lead = p < len && (b & 0xC0u) != 0x80u;
        const uint32_t mask = __ballot_sync(0xffffffffu, lead);
        if (lead) {
            uint32_t sym = b;
            for (uint32_t k = 1; k < 4 && p + k < len; ++k) {
                const uint32_t c = ldg1(s + p + k);
                if ((c & 0xC0u) != 0x80u) break;
                sym = (sym << 8) | c;
            }
            out[written + __popc(mask & ((1u << lane) - 1u))] = sym;
        }
        written += __popc(mask);
    }
}

// ---- Batch-wide alphabet compaction for cross products. Every query meets every candidate, so the dictionary is one per
// BATCH, in global memory: plain ASCII keeps its value, every other symbol of either tape is numbered 129, 130, ... as the
// survey pass (which also counts symbols and looks for non-ASCII bytes, like utf8_count_kernel) first meets it. A batch with
// at most 127 distinct multi-byte symbols — text in one or two alphabetic scripts — is then translated to byte tapes with
// ordinary n + 1 offsets and runs on the byte cross-product kernels; anything wider keeps the 32-bit wavefront path.
// The harness calls with a few hundred to a few thousand short strings per side, so this path is about launches, not
// bytes: one memset (every field starts as 0xFF..), one survey, one scan and one translate launch for BOTH tapes.
constexpr uint32_t BATCH_DICT_SLOTS = 1024, BATCH_DICT_CAP = 127, BATCH_DICT_BASE = 128;
constexpr uint32_t RENAME_EMPTY_KEY = 0xFFFFFFFFu; // no packed symbol (either packing): a lead byte with continuation bytes (at most 0xBF) or zeros around it
struct BatchDict {
    uint32_t ascii;                  // stays 0xFFFFFFFF while no byte >= 0x80 has been seen
    uint32_t distinct_minus_1;       // symbols numbered so far, minus one (starts at 0xFFFFFFFF); above BATCH_DICT_CAP the batch does not fit
    uint32_t offsets_ok;             // stays 0xFFFFFFFF while every string lies inside its tape's span and runs forwards
    uint32_t pad;
    uint64_t span[4];                // first and last offset of tape A, of tape B (written by the survey: the host sizes its buffers from them)
    uint32_t keys[BATCH_DICT_SLOTS]; // packed symbols, RENAME_EMPTY_KEY where free
    uint8_t ids[BATCH_DICT_SLOTS];
};
// One byte tape as the kernels below see it.
struct TapeArg {
    const uint8_t *data;
    const void *offs;
    uint32_t offs32;
    uint64_t n;
};

__device__ __forceinline__ void batch_dict_insert(BatchDict *d, uint32_t key) {
    if (*reinterpret_cast<volatile uint32_t *>(&d->distinct_minus_1) + 1u > BATCH_DICT_CAP) return; // already too wide: the host falls back
    uint32_t h = (key * 0x9E3779B1u) >> 22;
    for (uint32_t probes = 0; probes < BATCH_DICT_SLOTS; ++probes, h = (h + 1u) & (BATCH_DICT_SLOTS - 1u)) {
        uint32_t at = reinterpret_cast<volatile uint32_t *>(d->keys)[h];
        if (at == RENAME_EMPTY_KEY) {
            at = atomicCAS(&d->keys[h], RENAME_EMPTY_KEY, key);
            if (at == RENAME_EMPTY_KEY) {
                const uint32_t number = atomicAdd(&d->distinct_minus_1, 1u) + 2u;
                d->ids[h] = (uint8_t)min(BATCH_DICT_BASE + number, 255u);
                return;
            }
        }
        if (at == key) return;
    }
    atomicAdd(&d->distinct_minus_1, BATCH_DICT_SLOTS); // table full: reads as "too wide"
}
__device__ __forceinline__ uint32_t batch_dict_lookup(const BatchDict *d, uint32_t key) {
    uint32_t h = (key * 0x9E3779B1u) >> 22;
    for (uint32_t probes = 0; probes < BATCH_DICT_SLOTS; ++probes, h = (h + 1u) & (BATCH_DICT_SLOTS - 1u)) {
        const uint32_t at = __ldg(&d->keys[h]);
        if (at == key) return __ldg(&d->ids[h]);
        if (at == RENAME_EMPTY_KEY) break;
    }
    return BATCH_DICT_BASE; // not reachable after a survey of the same tape
}

// The symbol that starts at byte p of a string (p is a lead byte): lead byte and up to three continuation bytes, big-endian.
__device__ __forceinline__ uint32_t symbol_at(const uint8_t *s, uint32_t p, uint32_t len, uint32_t lead_byte) {
    uint32_t sym = lead_byte;
    for (uint32_t k = 1; k < 4 && p + k < len; ++k) {
        const uint32_t c = ldg1(s + p + k);
        if ((c & 0xC0u) != 0x80u) break;
        sym = (sym << 8) | c;
    }
    return sym;
}

// One warp per string of either tape, a byte per lane: counts[i + 1] = 4 * symbols of string i (the unit of the 32-bit
// fallback path), the ASCII flag is cleared when a byte >= 0x80 is seen, every symbol that is not a plain ASCII byte enters
// the dictionary.
__global__ void __launch_bounds__(256) utf8_survey_kernel(TapeArg ta, uint64_t *a_counts, TapeArg tb, uint64_t *b_counts, BatchDict *dict) {
    uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= ta.n + tb.n) return;
    const bool second = warp >= ta.n;
    const TapeArg &t = second ? tb : ta;
    uint64_t *counts = second ? b_counts : a_counts;
    if (second) warp -= ta.n;
    const uint64_t o0 = tape_offset(t.offs, t.offs32, warp), o1 = tape_offset(t.offs, t.offs32, warp + 1);
    if (warp == 0 && lane == 0) counts[0] = 0, dict->span[second ? 2 : 0] = o0;
    if (warp + 1 == t.n && lane == 0) dict->span[second ? 3 : 1] = o1;
    // this pass reads the strings before anything has validated the offsets: a string that runs backwards, spans 4 GiB or
    // leaves [first offset, last offset] of its tape is not read, and the host fails the call with SWB_INVALID_ARG
    const uint64_t first = tape_offset(t.offs, t.offs32, 0), last = tape_offset(t.offs, t.offs32, t.n);
    const bool ok = o0 >= first && o1 >= o0 && o1 <= last && ((o1 - o0) >> 32) == 0u;
    if (!ok && lane == 0) atomicAnd(&dict->offsets_ok, 0u);
    const uint8_t *s = t.data + o0;
    const uint32_t len = ok ? (uint32_t)(o1 - o0) : 0u;
    uint32_t count = 0;
    bool high = false;
    for (uint32_t base = 0; base < len; base += 32) {
        const uint32_t p = base + lane;
        const uint32_t b = p < len ? ldg1(s + p) : 0x80u;
        const bool lead = p < len && (b & 0xC0u) != 0x80u;
        high = high || (p < len && b >= 0x80u);
        count += __popc(__ballot_sync(0xffffffffu, lead));
        if (lead) {
            const uint32_t sym = symbol_at(s, p, len, b);
            if (sym >= 0x80u) batch_dict_insert(dict, sym);
        }
    }
    if (__any_sync(0xffffffffu, high) && lane == 0) atomicAnd(&dict->ascii, 0u);
    if (lane == 0) counts[warp + 1] = 4ull * count;
}

// Inclusive scan in place of one or two small arrays, one block each (a few thousand entries: not worth a device-wide scan).
constexpr int SMALL_SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SMALL_SCAN_THREADS) utf8_small_scan_kernel(uint64_t *a, uint64_t n_a, uint64_t *b, uint64_t n_b) {
    uint64_t *v = blockIdx.x == 0 ? a : b;
    const uint64_t n = blockIdx.x == 0 ? n_a : n_b;
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < n; base += SMALL_SCAN_THREADS) {
        const uint64_t i = base + threadIdx.x;
        uint64_t x = i < n ? v[i] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t up = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += up;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t up = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += up;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint64_t before = carry + (warp ? warp_sums[warp - 1] : 0);
        if (i < n) v[i] = x + before;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sums[31];
        __syncthreads();
    }
}

// One warp per string of either tape: the ids of string i at out + sym_offs[i] / 4, and the byte tape's offsets next to them.
__global__ void __launch_bounds__(256) utf8_translate_kernel(TapeArg ta, const uint64_t *a_sym_offs, uint8_t *a_out, uint64_t *a_out_offs, TapeArg tb,
                                                             const uint64_t *b_sym_offs, uint8_t *b_out, uint64_t *b_out_offs, const BatchDict *dict) {
    uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= ta.n + tb.n) return;
    const bool second = warp >= ta.n;
    const TapeArg &t = second ? tb : ta;
    const uint64_t *sym_offs = second ? b_sym_offs : a_sym_offs;
    uint8_t *out = second ? b_out : a_out;
    uint64_t *out_offs = second ? b_out_offs : a_out_offs;
    if (second) warp -= ta.n;
    const uint64_t o0 = tape_offset(t.offs, t.offs32, warp), o1 = tape_offset(t.offs, t.offs32, warp + 1);
    const uint8_t *s = t.data + o0;
    const uint32_t len = o1 < o0 ? 0u : (uint32_t)(o1 - o0);
    uint8_t *dst = out + (sym_offs[warp] >> 2);
    if (lane == 0) {
        out_offs[warp] = sym_offs[warp] >> 2;
        if (warp + 1 == t.n) out_offs[t.n] = sym_offs[t.n] >> 2;
    }
    uint32_t written = 0;
    for (uint32_t base = 0; base < len; base += 32) {
        const uint32_t p = base + lane;
        const uint32_t b = p < len ? ldg1(s + p) : 0x80u;
        const bool lead = p < len && (b & 0xC0u) != 0x80u;
        const uint32_t mask = __ballot_sync(0xffffffffu, lead);
        if (lead) {
            const uint32_t sym = symbol_at(s, p, len, b);
            dst[written + __popc(mask & ((1u << lane) - 1u))] = (uint8_t)(sym < 0x80u ? sym : batch_dict_lookup(dict, sym));
        }
        written += __popc(mask);
    }
}

// Per-pair alphabet compaction for element-wise pairs. The bit-vector kernels index their match tables by BYTE, and an
// edit distance only ever asks whether pattern symbol i equals text symbol j. So each pair's code points are renamed to
// bytes and the renamed pair, which has exactly the distance of the original, runs on the byte kernels.
//   ASCII_IDENTITY: a one-byte symbol below 0x80 keeps its value — no dictionary work at all for it. Every other symbol of
//     the pair's shorter string (in bytes) is numbered 129, 130, ... in the order a per-warp open-addressing dictionary
//     (shared memory) first meets it; the other string is translated through the same dictionary and what it holds that
//     the shorter string does not becomes 128, which equals no id. Room for 127 distinct multi-byte symbols.
//   otherwise (second tier): every symbol goes through the dictionary, ids 1..255, 0 for "not in the shorter string".
// ONE pass over the byte tapes does everything — symbol split, dictionary, translation. A round covers 512 bytes of a
// string, a re-aligned 16-byte piece per lane. A lane marks its lead bytes (a symbol starts at every byte that is not
// 10xxxxxx and takes up to three continuation bytes: the oracle's rule); one warp prefix sum gives it its place in the
// staging row of ids and in the list of the round's dictionary symbols; plain ASCII goes straight to the row, the rest to
// the list, and the list is then worked off ONE SYMBOL PER LANE, all 32 lanes busy whatever the mix of scripts: probe,
// insert (atomicCAS), number the newcomers of the step with a ballot, write the id into the row. The row leaves as aligned
// 32-bit stores. There is no symbol count / scan pass: renamed string i is written at the byte offset of the original
// string i, and the output tape has TWO strings per input string — the renamed string and the gap behind it — so a
// PairSpace with idx_shift = 1 addresses it (out_offs[2 i] = start, out_offs[2 i + 1] = end).
// `flags`: bit 0 is raised by a pair with more distinct symbols than ids (the host goes to the next tier: the 255-id
// variant, then the 32-bit wavefront kernel), bit 1 by a string whose offsets run backwards or leave the tape's span (the
// call fails with SWB_INVALID_ARG; the pair is renamed to two empty strings and nothing outside the span is read).
constexpr int RENAME_WARPS = 4;
constexpr uint32_t RENAME_TABLE = 512;         // dictionary slots: at most 255 are ever numbered, so probes are short
constexpr uint32_t RENAME_OVERFLOW = 1u, RENAME_BAD_OFFSETS = 2u;

// bit k: byte k of the word is >= 0x80
__device__ __forceinline__ uint32_t high_bits4(uint32_t w) { return ((((w >> 7) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu; }

struct RenameShared {
    uint32_t keys[RENAME_TABLE];       // the dictionary: packed symbols
    uint32_t list[512];                // the round's dictionary symbols: byte position in the round | place in the row << 16
    uint32_t ext[32];                  // the lanes' symbol-end bits (see below)
    uint8_t key_id[RENAME_TABLE];      // id of the symbol in slot h
    alignas(16) uint8_t raw[512 + 16]; // the round's bytes
    alignas(16) uint8_t row[512 + 16]; // ids of the round's symbols, in order
};

template <bool ASCII_IDENTITY>
__global__ void __launch_bounds__(32 * RENAME_WARPS)
    utf8_rename_kernel(const uint8_t *a_data, const void *a_offs, uint32_t a32, uint64_t a_span, const uint8_t *b_data, const void *b_offs, uint32_t b32,
                       uint64_t b_span, uint64_t n_pairs, uint8_t *a_out, uint8_t *b_out, uint64_t *a_oo, uint64_t *b_oo, uint32_t *flags) {
    constexpr uint32_t CAP = ASCII_IDENTITY ? 127u : 255u, ID_BASE = ASCII_IDENTITY ? 128u : 0u; // ids ID_BASE + 1 .. ID_BASE + CAP; ID_BASE = "absent"
    __shared__ __align__(16) RenameShared shared[RENAME_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t pair = (uint64_t)blockIdx.x * RENAME_WARPS + warp;
    if (pair >= n_pairs) return;
    RenameShared &sh = shared[warp];
    const uint64_t a_base = tape_offset(a_offs, a32, 0), b_base = tape_offset(b_offs, b32, 0);
    uint64_t ab0 = tape_offset(a_offs, a32, pair), ab1 = tape_offset(a_offs, a32, pair + 1), bb0 = tape_offset(b_offs, b32, pair),
             bb1 = tape_offset(b_offs, b32, pair + 1);
    if (ab0 < a_base || ab1 < ab0 || ab1 - a_base > a_span || ((ab1 - ab0) >> 32) != 0u || bb0 < b_base || bb1 < bb0 || bb1 - b_base > b_span ||
        ((bb1 - bb0) >> 32) != 0u) {
        if (lane == 0) atomicOr(flags, RENAME_BAD_OFFSETS);
        ab0 = ab1 = a_base, bb0 = bb1 = b_base;
    }
    const uint32_t a_len = (uint32_t)(ab1 - ab0), b_len = (uint32_t)(bb1 - bb0); // bytes
#pragma unroll
    for (uint32_t i = 0; i < RENAME_TABLE / 128; ++i)
        reinterpret_cast<uint4 *>(sh.keys)[i * 32 + lane] = make_uint4(RENAME_EMPTY_KEY, RENAME_EMPTY_KEY, RENAME_EMPTY_KEY, RENAME_EMPTY_KEY);
    __syncwarp();
    bool failed = false;
    uint32_t n_ids = 0; // symbols numbered so far (the same in every lane)
    const bool a_first = a_len <= b_len;
    uint32_t n_symbols[2];

    // Pass 0 renames the shorter string, whose symbols enter the dictionary; pass 1 translates the other one. One body for
    // both (the loop is not unrolled): the code stays small enough for the instruction cache.
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const bool dict = pass == 0, is_a = a_first == dict;
        const uint8_t *str = is_a ? a_data + ab0 : b_data + bb0;
        const uint32_t len = is_a ? a_len : b_len;
        uint8_t *out = is_a ? a_out + (ab0 - a_base) : b_out + (bb0 - b_base);
        PieceStream st;
        st.open(str, len, 0u);
        uint32_t nsym = 0;
#pragma unroll 1
        for (uint32_t base = 0; 16u * base < len; base += 32) {
            const uint32_t i = base + (uint32_t)lane;
            const bool live = 16u * i < len;
            uint4 p = make_uint4(0, 0, 0, 0);
            if (live) p = st.piece(i);
            // the four bytes after the piece (continuation bytes of a symbol that starts near its end): the next lane's
            uint32_t ahead = __shfl_down_sync(0xffffffffu, p.x, 1);
            if (lane == 31) ahead = 16u * (i + 1u) < len ? st.piece(i + 1u).x : 0u;
            const uint32_t remaining = live ? len - 16u * i : 0u; // bytes of the string from this piece's first byte on
            const uint32_t valid = remaining >= 16u ? 0xFFFFu : ((1u << remaining) - 1u);
            // bit j of ext: byte j of the 20 ends the symbol before it (a lead byte, or the end of the string)
            const uint32_t ext = lead_bits16(p) | (lead_bits4(ahead) << 16) | (remaining >= 20u ? 0u : 0xFFFFFFFFu << remaining);
            const uint32_t leads = ext & valid;
            const uint32_t high = high_bits4(p.x) | (high_bits4(p.y) << 4) | (high_bits4(p.z) << 8) | (high_bits4(p.w) << 12);
            const uint32_t plain = ASCII_IDENTITY ? (leads & ~high & (ext >> 1)) : 0u; // one-byte symbols below 0x80
            const uint32_t listed = leads & ~plain;
            const uint32_t mine = __popc(leads) | (__popc(listed) << 16);
            uint32_t upto = mine; // inclusive scan over lanes of both counts at once
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, upto, d);
                if (lane >= d) upto += up;
            }
            const uint32_t totals = __shfl_sync(0xffffffffu, upto, 31);
            const uint32_t total = totals & 0xFFFFu, n_listed = totals >> 16;
            const uint32_t k = (upto - mine) & 0xFFFFu, s = (upto - mine) >> 16;
            // the bytes and the symbol ends, for the lanes that work the list off
            *reinterpret_cast<uint4 *>(sh.raw + 16 * lane) = p;
            if (lane == 31) *reinterpret_cast<uint32_t *>(sh.raw + 512) = ahead;
            sh.ext[lane] = ext;
            if (ASCII_IDENTITY && plain == 0xFFFFu && (k & 3u) == 0u) {
                // sixteen one-byte symbols: the ids are the bytes
                uint32_t *d32 = reinterpret_cast<uint32_t *>(sh.row + k);
                d32[0] = p.x, d32[1] = p.y, d32[2] = p.z, d32[3] = p.w;
            } else if (leads) {
                // straight-line, predicated: plain ASCII to its place in the row, everything else to the list
                const uint32_t w[4] = {p.x, p.y, p.z, p.w};
                const uint32_t here = 16u * (uint32_t)lane + (k << 16);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t before = (1u << j) - 1u;
                    const uint32_t kj = k + __popc(leads & before);
                    if (ASCII_IDENTITY && ((plain >> j) & 1u)) sh.row[kj] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
                    if ((listed >> j) & 1u) sh.list[s + __popc(listed & before)] = here + (uint32_t)j + (__popc(leads & before) << 16);
                }
            }
            __syncwarp();
            // the round's dictionary symbols, one per lane
            for (uint32_t q0 = 0; q0 < n_listed; q0 += 32) {
                const uint32_t q = q0 + (uint32_t)lane;
                const bool active = q < n_listed;
                const uint32_t entry = active ? sh.list[q] : 0u;
                const uint32_t pos = entry & 0xFFFFu, at_row = entry >> 16;
                const uint32_t *word = reinterpret_cast<const uint32_t *>(sh.raw + (pos & ~3u));
                const uint32_t four = __funnelshift_r(word[0], word[1], 8u * (pos & 3u));                    // bytes pos .. pos + 3
                const uint32_t conts = __ffs(((sh.ext[pos >> 4] >> ((pos & 15u) + 1u)) & 7u) | 8u) - 1u; // continuation bytes taken: 0..3
                const uint32_t key = four & (0xFFFFFFFFu >> (24u - 8u * conts));
                uint32_t h = (key * 0x9E3779B1u) >> 23;
                bool won = false, found = false;
                if (active) {
#pragma unroll 1
                    for (uint32_t probes = 0; probes < RENAME_TABLE; ++probes, h = (h + 1u) & (RENAME_TABLE - 1u)) {
                        uint32_t at = sh.keys[h];
                        if (dict && at == RENAME_EMPTY_KEY) {
                            at = atomicCAS(&sh.keys[h], RENAME_EMPTY_KEY, key);
                            if (at == RENAME_EMPTY_KEY) {
                                won = true;
                                break;
                            }
                        }
                        if (at == key) {
                            found = true;
                            break;
                        }
                        if (at == RENAME_EMPTY_KEY) break; // not in the dictionary
                    }
                }
                if (dict) { // the step's newcomers take the next numbers, in lane order
                    const uint32_t winners = __ballot_sync(0xffffffffu, won);
                    if (won) {
                        const uint32_t number = n_ids + __popc(winners & ((1u << lane) - 1u)) + 1u;
                        if (number > CAP) failed = true;
                        sh.key_id[h] = (uint8_t)(ID_BASE + number);
                    }
                    n_ids += __popc(winners);
                    if (active && !won && !found) failed = true; // table full (only after the ids ran out)
                    __syncwarp();
                    found = found || won;
                }
                if (active) sh.row[at_row] = found ? sh.key_id[h] : (uint8_t)ID_BASE;
            }
            __syncwarp();
            // the row leaves as aligned 32-bit words; the bytes before the first and after the last whole word one by one
            uint8_t *dst = out + nsym;
            const uint32_t head = min(total, (uint32_t)((4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u));
            const uint32_t words = (total - head) >> 2, tail = head + 4u * words;
            if ((uint32_t)lane < head) dst[lane] = sh.row[lane];
            const uint32_t *row32 = reinterpret_cast<const uint32_t *>(sh.row);
            uint32_t *dst32 = reinterpret_cast<uint32_t *>(dst + head);
            for (uint32_t w = lane; w < words; w += 32) dst32[w] = __funnelshift_r(row32[w], row32[w + 1], 8u * head);
            if (tail + (uint32_t)lane < total) dst[tail + lane] = sh.row[tail + lane];
            __syncwarp();
            nsym += total;
        }
        n_symbols[is_a ? 0 : 1] = nsym;
    }
    if (lane == 0) {
        a_oo[2 * pair] = ab0 - a_base, a_oo[2 * pair + 1] = ab0 - a_base + n_symbols[0];
        b_oo[2 * pair] = bb0 - b_base, b_oo[2 * pair + 1] = bb0 - b_base + n_symbols[1];
    }
    if (__any_sync(0xffffffffu, failed) && lane == 0) atomicOr(flags, RENAME_OVERFLOW);
}

} // namespace swb
