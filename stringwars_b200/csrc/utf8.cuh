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

// Per-pair alphabet compaction for element-wise pairs. The bit-vector kernels index their match tables by BYTE, and an
// edit distance only ever asks whether pattern symbol i equals text symbol j. So each pair's code points are renamed to
// bytes: the distinct symbols of the pair's shorter string get the number of the slot they occupy in a 255-slot
// open-addressing dictionary (shared memory, one per warp), the other string is translated through the same dictionary
// and everything it holds that the shorter string does not becomes 0 (which equals no id). The renamed pair has exactly
// the distance of the original and runs on the byte kernels.
// ONE pass over the byte tapes does everything — symbol split, dictionary, translation: a round covers 512 bytes of a
// string (a re-aligned 16-byte piece per lane); lanes mark their lead bytes, a warp prefix sum numbers the symbols, the
// packed symbols go to a shared-memory staging row, and the dictionary work then runs one symbol per lane, all 32 lanes
// busy, with coalesced byte stores. There is no symbol count / scan pass: renamed string i is written at the byte offset
// of the original string i, and the output tape has TWO strings per input string — the renamed string and the gap behind
// it — so a PairSpace with idx_shift = 1 addresses it (out_offs[2 i] = start, out_offs[2 i + 1] = end).
// A pair whose dictionary fills up (more than ~250 distinct code points in its shorter string) raises `overflow`; the
// host then keeps the batch on the 32-bit wavefront kernel.
constexpr int RENAME_WARPS = 4;
constexpr uint32_t RENAME_SLOTS = 256;        // slot 0 is never used: ids are 1..255
constexpr uint32_t RENAME_EMPTY = 0xFFFFFFFFu; // no packed UTF-8 symbol: continuation bytes are never 0xFF

__device__ __forceinline__ uint32_t rename_slot(uint32_t key) { return 1u + ((key * 0x9E3779B1u) >> 8) % 255u; }

__global__ void __launch_bounds__(32 * RENAME_WARPS)
    utf8_rename_kernel(const uint8_t *a_data, const void *a_offs, uint32_t a32, const uint8_t *b_data, const void *b_offs, uint32_t b32, uint64_t n_pairs,
                       uint8_t *a_out, uint8_t *b_out, uint64_t *a_oo, uint64_t *b_oo, uint32_t *overflow) {
    __shared__ uint32_t s_keys[RENAME_WARPS][RENAME_SLOTS];
    __shared__ uint32_t s_syms[RENAME_WARPS][512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t pair = (uint64_t)blockIdx.x * RENAME_WARPS + warp;
    if (pair >= n_pairs) return;
    uint32_t *keys = s_keys[warp], *syms = s_syms[warp];
    const uint64_t a_base = tape_offset(a_offs, a32, 0), b_base = tape_offset(b_offs, b32, 0);
    const uint64_t ab0 = tape_offset(a_offs, a32, pair), ab1 = tape_offset(a_offs, a32, pair + 1), bb0 = tape_offset(b_offs, b32, pair),
                   bb1 = tape_offset(b_offs, b32, pair + 1);
    const uint32_t a_len = ab1 < ab0 ? 0u : (uint32_t)(ab1 - ab0), b_len = bb1 < bb0 ? 0u : (uint32_t)(bb1 - bb0); // bytes
    for (uint32_t i = lane; i < RENAME_SLOTS; i += 32) keys[i] = RENAME_EMPTY;
    __syncwarp();
    bool failed = false;

    // Renames one string; DICT: its symbols enter the dictionary. Returns the number of symbols (same in every lane).
    auto rename = [&](const uint8_t *str, uint32_t len, uint8_t *out, auto dict_tag) -> uint32_t {
        constexpr bool DICT = decltype(dict_tag)::value;
        PieceStream st;
        st.open(str, len, 0u);
        uint32_t nsym = 0;
        for (uint32_t base = 0; 16u * base < len; base += 32) {
            const uint32_t i = base + (uint32_t)lane;
            const bool live = 16u * i < len;
            uint4 p = make_uint4(0, 0, 0, 0);
            uint32_t ahead = 0; // the four bytes after the piece (continuation bytes of a symbol that starts near its end)
            if (live) {
                p = st.piece(i);
                if (16u * (i + 1u) < len) ahead = st.piece(i + 1u).x;
            }
            const uint32_t remaining = live ? len - 16u * i : 0u; // bytes of the string from this piece's first byte on
            const uint32_t valid = remaining >= 16u ? 0xFFFFu : ((1u << remaining) - 1u);
            uint32_t leads = live ? (lead_bits16(p) & valid) : 0u;
            const uint32_t mine = __popc(leads);
            uint32_t upto = mine; // inclusive scan over lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, upto, d);
                if (lane >= d) upto += up;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, upto, 31);
            uint32_t k = upto - mine;
            const uint32_t words[5] = {p.x, p.y, p.z, p.w, ahead};
            while (leads) { // stage this piece's packed symbols (lead byte + up to three continuation bytes, big-endian)
                const uint32_t j = __ffs(leads) - 1u;
                leads &= leads - 1u;
                const uint32_t four = __funnelshift_r(words[j >> 2], words[(j >> 2) + 1u], 8u * (j & 3u)); // bytes j .. j + 3
                uint32_t sym = four & 0xFFu;
                const uint32_t room = remaining - j; // bytes of the string from byte j on (>= 1)
                if (room > 1u && ((four >> 8) & 0xC0u) == 0x80u) {
                    sym = (sym << 8) | ((four >> 8) & 0xFFu);
                    if (room > 2u && ((four >> 16) & 0xC0u) == 0x80u) {
                        sym = (sym << 8) | ((four >> 16) & 0xFFu);
                        if (room > 3u && ((four >> 24) & 0xC0u) == 0x80u) sym = (sym << 8) | (four >> 24);
                    }
                }
                syms[k++] = sym;
            }
            __syncwarp();
            for (uint32_t q = lane; q < total; q += 32) { // one symbol per lane: dictionary lookup / insertion
                const uint32_t key = syms[q];
                uint32_t h = rename_slot(key), id = 0;
                for (uint32_t probes = 0; probes < 255u; ++probes, h = h == 255u ? 1u : h + 1u) {
                    uint32_t at = keys[h];
                    if (DICT && at == RENAME_EMPTY) at = atomicCAS(&keys[h], RENAME_EMPTY, key) == RENAME_EMPTY ? key : keys[h];
                    if (at == key) {
                        id = h;
                        break;
                    }
                    if (at == RENAME_EMPTY) break; // not in the dictionary: stays 0
                }
                if (DICT && id == 0u) failed = true; // dictionary full
                out[nsym + q] = (uint8_t)id;
            }
            __syncwarp();
            nsym += total;
        }
        return nsym;
    };
    const bool dict_is_a = a_len <= b_len;
    uint32_t na, nb;
    if (dict_is_a) {
        na = rename(a_data + ab0, a_len, a_out + (ab0 - a_base), std::true_type{});
        nb = rename(b_data + bb0, b_len, b_out + (bb0 - b_base), std::false_type{});
    } else {
        nb = rename(b_data + bb0, b_len, b_out + (bb0 - b_base), std::true_type{});
        na = rename(a_data + ab0, a_len, a_out + (ab0 - a_base), std::false_type{});
    }
    if (lane == 0) {
        a_oo[2 * pair] = ab0 - a_base, a_oo[2 * pair + 1] = ab0 - a_base + na;
        b_oo[2 * pair] = bb0 - b_base, b_oo[2 * pair + 1] = bb0 - b_base + nb;
    }
    if (__any_sync(0xffffffffu, failed) && lane == 0) atomicOr(overflow, 1u);
}

} // namespace swb
