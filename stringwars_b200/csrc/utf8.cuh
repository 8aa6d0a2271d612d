// utf8.cuh — device-side split of UTF-8 tapes into 32-bit symbols for the codepoint-level engine.
//
// `LevenshteinDistancesUtf8` (StringWars similarities/bench.rs:386-399) and the rapidfuzz baseline over
// `str::chars()` (bench.rs:437) compare Unicode scalar values. Only EQUALITY of symbols matters for an edit
// distance, so the device does not decode to scalar values: a symbol is the big-endian packing of a lead byte and
// up to three continuation bytes, which is a bijection with scalar values on valid UTF-8. A symbol starts at
// every byte that is not 10xxxxxx; stray leading continuation bytes are ignored (same rule as the CPU oracle).
#pragma once
#include "common.cuh"

namespace swb {

// One warp per string: bytes_out[i + 1] = 4 * (number of symbols of string i); bytes_out[0] = 0.
// `ascii_flag` is cleared (atomicAnd) when any byte >= 0x80 is seen.
__device__ __forceinline__ uint64_t tape_offset(const void *offs, uint32_t offs32, uint64_t i) {
    return offs32 ? (uint64_t) reinterpret_cast<const uint32_t *>(offs)[i] : reinterpret_cast<const uint64_t *>(offs)[i];
}

__global__ void __launch_bounds__(256) utf8_count_kernel(const uint8_t *data, const void *offs, uint32_t offs32, uint64_t n_strings,
                                                         uint64_t *bytes_out, uint32_t *ascii_flag) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) bytes_out[0] = 0;
    if (warp >= n_strings) return;
    const uint64_t o0 = tape_offset(offs, offs32, warp), o1 = tape_offset(offs, offs32, warp + 1);
    const uint8_t *s = data + o0;
    const uint32_t len = o1 < o0 ? 0u : (uint32_t)(o1 - o0);
    uint32_t count = 0;
    bool high = false;
    for (uint32_t p = lane; p < len; p += 32) {
        const uint32_t b = ldg1(s + p);
        count += (b & 0xC0u) != 0x80u;
        high |= b >= 0x80u;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) count += __shfl_xor_sync(0xffffffffu, count, d);
    if (__any_sync(0xffffffffu, high) && lane == 0) atomicAnd(ascii_flag, 0u);
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

// Per-pair alphabet compaction: the bit-vector kernels index their match tables by BYTE, and an edit distance only ever
// asks whether pattern symbol i equals text symbol j. So each pair's 32-bit symbols are renamed to bytes: the distinct
// symbols of the pair's SHORTER string get ids 1, 2, ... in hash-table order, the other string is translated through the
// same dictionary and everything it holds that the shorter string does not becomes 0 (which equals no id). The renamed
// pair has exactly the distance of the original, and runs on the byte kernels (16+ TCUPS instead of the 32-bit
// wavefront's 2). One warp per pair; the dictionary is an open-addressing table of REMAP_SLOTS keys in shared memory.
// A pair whose shorter string has more than 255 distinct symbols raises `overflow` (the host then keeps the batch on the
// 32-bit wavefront kernel).
constexpr int REMAP_SLOTS = 512;
constexpr int REMAP_WARPS = 4;
constexpr uint32_t REMAP_EMPTY = 0xFFFFFFFFu; // no packed UTF-8 symbol: continuation bytes are never 0xFF

__device__ __forceinline__ uint32_t remap_hash(uint32_t key) { return (key * 0x9E3779B1u) >> (32 - 9); }

__global__ void __launch_bounds__(32 * REMAP_WARPS)
    utf8_remap_kernel(const uint32_t *a_sym, const uint64_t *a_so, const uint32_t *b_sym, const uint64_t *b_so, uint64_t n_pairs, uint8_t *a_out,
                      uint8_t *b_out, uint64_t *a_oo, uint64_t *b_oo, uint32_t *overflow) {
    __shared__ uint32_t s_keys[REMAP_WARPS][REMAP_SLOTS];
    __shared__ uint8_t s_ids[REMAP_WARPS][REMAP_SLOTS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t pair = (uint64_t)blockIdx.x * REMAP_WARPS + warp;
    if (pair >= n_pairs) return;
    uint32_t *keys = s_keys[warp];
    uint8_t *ids = s_ids[warp];
    const uint64_t a0 = a_so[pair] >> 2, a1 = a_so[pair + 1] >> 2, b0 = b_so[pair] >> 2, b1 = b_so[pair + 1] >> 2;
    const uint32_t la = (uint32_t)(a1 - a0), lb = (uint32_t)(b1 - b0);
    if (lane == 0) {
        a_oo[pair] = a0, b_oo[pair] = b0;
        if (pair + 1 == n_pairs) a_oo[n_pairs] = a1, b_oo[n_pairs] = b1;
    }
    const bool dict_is_a = la <= lb;
    const uint32_t *dict = dict_is_a ? a_sym + a0 : b_sym + b0;
    const uint32_t dict_len = dict_is_a ? la : lb;
    for (int i = lane; i < REMAP_SLOTS; i += 32) keys[i] = REMAP_EMPTY;
    __syncwarp();
    bool failed = false;
    for (uint32_t k = lane; k < dict_len; k += 32) {
        const uint32_t key = dict[k];
        uint32_t h = remap_hash(key);
        int probes = 0;
        for (; probes < REMAP_SLOTS; ++probes, h = (h + 1u) & (REMAP_SLOTS - 1)) {
            const uint32_t prev = atomicCAS(&keys[h], REMAP_EMPTY, key);
            if (prev == REMAP_EMPTY || prev == key) break;
        }
        failed |= probes == REMAP_SLOTS;
    }
    __syncwarp();
    // ids in slot order: 1 + number of occupied slots before this one
    uint32_t distinct = 0;
    for (int r = 0; r < REMAP_SLOTS / 32; ++r) {
        const bool occupied = keys[r * 32 + lane] != REMAP_EMPTY;
        const uint32_t mask = __ballot_sync(0xffffffffu, occupied);
        const uint32_t id = distinct + __popc(mask & ((1u << lane) - 1u)) + 1u;
        ids[r * 32 + lane] = (uint8_t)id;
        distinct += __popc(mask);
    }
    __syncwarp();
    if (__any_sync(0xffffffffu, failed) || distinct > 255u) {
        if (lane == 0) atomicOr(overflow, 1u);
        return;
    }
    auto translate = [&](const uint32_t *sym, uint32_t len, uint8_t *out) {
        for (uint32_t k = lane; k < len; k += 32) {
            const uint32_t key = sym[k];
            uint32_t h = remap_hash(key);
            uint32_t id = 0;
            for (int probes = 0; probes < REMAP_SLOTS; ++probes, h = (h + 1u) & (REMAP_SLOTS - 1)) {
                const uint32_t at = keys[h];
                if (at == key) {
                    id = ids[h];
                    break;
                }
                if (at == REMAP_EMPTY) break;
            }
            out[k] = (uint8_t)id;
        }
    };
    translate(a_sym + a0, la, a_out + a0);
    translate(b_sym + b0, lb, b_out + b0);
}

} // namespace swb
