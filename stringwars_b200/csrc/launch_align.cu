// launch_align.cu — launchers of the scalar anti-diagonal wavefront kernels (align_wavefront.cuh): Levenshtein in
// wavefront mode, code-point symbols, Needleman-Wunsch and Smith-Waterman with linear / affine gaps, unary costs or the
// 32 x 32 class table. One translation unit per kernel family (engine_internal.h).
#include "engine_internal.h"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "align_wavefront.cuh"

namespace swb {

template <int G, int R, bool LOCAL, bool AFFINE, bool UNARY, typename BoundT, typename SymT = uint8_t>
static swb_status_t launch_wavefront_t(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin,
                                       uint64_t count, const AlignCosts &cs, uint32_t max_m, uint32_t max_n,
                                       cudaStream_t stream, const char **err) {
    constexpr int PPB = WF_THREADS / G;
    constexpr bool SHIFTED = !LOCAL && !AFFINE && UNARY;
    const bool single_strip = max_m <= (uint32_t)(G * R);
    const uint32_t slot_rows = round16(max_m * (uint32_t)sizeof(SymT) + 16u) + 16u, slot_cols = round16(max_n * (uint32_t)sizeof(SymT) + 16u) + 16u;
    const uint32_t bound_cap = (SHIFTED && single_strip) ? 0u : max_n + 1u; // shifted single-strip cells need no boundary row
    const uint32_t bound_bytes = round16(bound_cap * (uint32_t)sizeof(BoundT) * (AFFINE ? 2u : 1u));
    const size_t smem = 2304 + (size_t)PPB * (slot_rows + slot_cols + bound_bytes);
    if (smem > (size_t)ctx.max_smem_optin) {
        if constexpr (G == 32 && std::is_same<BoundT, int32_t>::value) {
            // Too long to stage: the LONG variant streams both strings from global memory and parks the boundary rows in
            // a global scratch row per pair; sub-batches keep that scratch under 1 GiB whatever the class holds.
            auto long_kernel = align_wavefront_kernel<G, R, LOCAL, AFFINE, UNARY, BoundT, SymT, true>;
            const uint64_t stride = (((uint64_t)max_n + 1u) * sizeof(BoundT) * (AFFINE ? 2u : 1u) + 15u) & ~(uint64_t)15u;
            const uint64_t per_launch = std::max<uint64_t>(PPB, std::min<uint64_t>(count, (1ull << 30) / stride) / PPB * PPB);
            SWB_CUDA(ctx.hscratch.reserve(std::min<uint64_t>(count, per_launch) * stride));
            for (uint64_t done = 0; done < count; done += per_launch) {
                const uint64_t now = std::min<uint64_t>(per_launch, count - done);
                ProfileBracket bracket(eng, ctx.device, stream);
                long_kernel<<<(unsigned)((now + PPB - 1) / PPB), WF_THREADS, 2304, stream>>>(ps, begin + done, now, cs, eng->d_tables[dev_index], 0u, 0u, 0u,
                                                                                         (uint8_t *)ctx.hscratch.ptr, stride);
                g_launches++;
                SWB_CUDA(cudaGetLastError());
            }
            eng->last_kernel = "align_wavefront<G=32,R=16,long>";
            return SWB_OK;
        }
        return fail(err, SWB_TOO_LONG, "pair of %u x %u bytes needs %zu B of shared memory per block (limit %d)", max_m, max_n,
                    smem, ctx.max_smem_optin);
    }
    auto kernel = align_wavefront_kernel<G, R, LOCAL, AFFINE, UNARY, BoundT, SymT>;
    SWB_CUDA(allow_smem((const void *)kernel, ctx.device, smem));
    const uint64_t blocks = (count + PPB - 1) / PPB;
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        kernel<<<(unsigned)blocks, WF_THREADS, smem, stream>>>(ps, begin, count, cs, eng->d_tables[dev_index], slot_rows, slot_cols, bound_cap, nullptr, 0ull);
    }
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

template <int G, int R>
static swb_status_t launch_wavefront_gr(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin,
                                        uint64_t count, const AlignCosts &cs, bool local, bool affine, bool unary, bool narrow,
                                        uint32_t max_m, uint32_t max_n, cudaStream_t stream, const char **err) {
#define SWB_WF(L, A, U, T) return launch_wavefront_t<G, R, L, A, U, T>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err)
    if (ps.len_shift == 2) { // decoded UTF-8: 32-bit symbols, unit costs only
        if (narrow) return launch_wavefront_t<G, R, false, false, true, int16_t, uint32_t>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);
        return launch_wavefront_t<G, R, false, false, true, int32_t, uint32_t>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);
    }
    if (!unary) { // general 32x32 substitution table: int32 boundary rows only
        if (!local && !affine) SWB_WF(false, false, false, int32_t);
        if (!local && affine) SWB_WF(false, true, false, int32_t);
        if (local && !affine) SWB_WF(true, false, false, int32_t);
        SWB_WF(true, true, false, int32_t);
    }
    if (!local && !affine && narrow) SWB_WF(false, false, true, int16_t);
    if (!local && !affine && !narrow) SWB_WF(false, false, true, int32_t);
    if (!local && affine && narrow) SWB_WF(false, true, true, int16_t);
    if (!local && affine && !narrow) SWB_WF(false, true, true, int32_t);
    if (local && !affine && narrow) SWB_WF(true, false, true, int16_t);
    if (local && !affine && !narrow) SWB_WF(true, false, true, int32_t);
    if (local && affine && narrow) SWB_WF(true, true, true, int16_t);
    SWB_WF(true, true, true, int32_t);
#undef SWB_WF
}

swb_status_t launch_wavefront(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin, uint64_t count,
                              uint32_t max_m, uint32_t max_n, cudaStream_t stream, const char **err, bool same_shape) {
    AlignCosts cs;
    const bool is_lev = eng->kind == ENGINE_LEVENSHTEIN || eng->kind == ENGINE_LEVENSHTEIN_UTF8;
    cs.match = eng->match, cs.mismatch = eng->mismatch;
    cs.ident = is_lev ? 1 : 0;
    cs.negate = is_lev ? 1 : 0;
    cs.first = eng->convention == SWB_GAP_BIO && !is_lev ? eng->open + eng->extend : eng->open;
    cs.next = eng->extend;
    const bool local = eng->kind == ENGINE_SW;
    const bool affine = cs.first != cs.next;
    // int16 boundary rows when every reachable |score| (also in the shifted H - gap*(i+j) form) stays inside int16
    int maxabs = std::max(std::abs(cs.first), std::abs(cs.next));
    if (eng->unary) maxabs = std::max(maxabs, std::max(std::abs(cs.match), std::abs(cs.mismatch)));
    else maxabs = std::max(maxabs, 128);
    const bool narrow = (uint64_t)(max_m + max_n + 2u) * 3u * (uint64_t)maxabs < 30000u;
    // Packed s16x2 variants (two pairs per lane group; launch_packed.cu decides whether the batch qualifies).
    if (same_shape && eng->unary && ps.len_shift == 0 && count >= 2) {
        const swb_status_t pst = launch_wavefront_packed_any(eng, ctx, dev_index, ps, begin, count, cs, local, affine, max_m, max_n, stream, err);
        if (pst != SWB_TOO_LONG) return pst;
    }
    // Narrowest group that covers the pattern in one strip; if the class's longest text makes the block's staged
    // strings overflow shared memory (lopsided pairs), widen the group: fewer pairs per block, same kernel.
    swb_status_t st = SWB_TOO_LONG;
#define SWB_GR(G, R, NAME)                                                                                                              \
    if (st == SWB_TOO_LONG && (max_m <= (uint32_t)(G) * (R) || (G) == 32)) {                                                              \
        eng->last_kernel = NAME;                                                                                                        \
        st = launch_wavefront_gr<G, R>(eng, ctx, dev_index, ps, begin, count, cs, local, affine, eng->unary, narrow, max_m, max_n, stream, err); \
    }
    SWB_GR(1, 16, "align_wavefront<G=1,R=16>")
    SWB_GR(2, 16, "align_wavefront<G=2,R=16>")
    SWB_GR(4, 16, "align_wavefront<G=4,R=16>")
    SWB_GR(8, 16, "align_wavefront<G=8,R=16>")
    SWB_GR(16, 16, "align_wavefront<G=16,R=16>")
    SWB_GR(32, 16, "align_wavefront<G=32,R=16>")
#undef SWB_GR
    if (st == SWB_TOO_LONG && narrow) // int16 boundary rows have no global-memory variant: the int32 one does
        st = launch_wavefront_gr<32, 16>(eng, ctx, dev_index, ps, begin, count, cs, local, affine, eng->unary, false, max_m, max_n, stream, err);
    return st;
}

} // namespace swb
