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

} // namespace swb
