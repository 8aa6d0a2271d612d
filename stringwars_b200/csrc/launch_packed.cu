// launch_packed.cu — launchers of the packed s16x2 wavefront kernels (two pairs per lane group, DPX VIADDMNMX.S16x2).
#include "engine_internal.h"

#include <algorithm>
#include <cstdlib>

#include "align_wavefront_packed.cuh"
#include "align_wavefront_packed_general.cuh"

namespace swb {

// Two pairs per lane group in packed s16x2 registers (align_wavefront_packed.cuh). SWB_TOO_LONG = does not fit smem.
template <int G, int R>
static swb_status_t launch_wavefront_packed(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin, uint64_t count,
                                            const AlignCosts &cs, uint32_t max_m, uint32_t max_n, cudaStream_t stream, const char **err) {
    constexpr int GPB = WF_THREADS / G;
    const bool single_strip = max_m <= (uint32_t)(G * R);
    const uint32_t slot_rows = round16(max_m + 16u) + 16u, slot_cols = round16(max_n + 16u) + 16u;
    const uint32_t bound_cap = single_strip ? 0u : max_n + 1u;
    const size_t smem = 2304 + (size_t)GPB * (2u * slot_rows + 2u * slot_cols + round16(bound_cap * 4u));
    if (smem > (size_t)ctx.max_smem_optin) return fail(err, SWB_TOO_LONG, "packed wavefront: %zu B of shared memory per block", smem);
    SWB_CUDA(allow_smem((const void *)align_wavefront_packed_kernel<G, R>, ctx.device, smem));
    const uint64_t blocks = ((count + 1) / 2 + GPB - 1) / GPB;
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        align_wavefront_packed_kernel<G, R><<<(unsigned)blocks, WF_THREADS, smem, stream>>>(ps, begin, count, cs, eng->d_tables[dev_index], slot_rows, slot_cols, bound_cap);
    }
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

template <int G, int R, bool LOCAL, bool AFFINE>
static swb_status_t launch_wavefront_packed_general(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin,
                                                    uint64_t count, const AlignCosts &cs, uint32_t max_m, uint32_t max_n, cudaStream_t stream,
                                                    const char **err) {
    constexpr int GPB = WF_THREADS / G;
    const uint32_t slot_rows = round16(max_m + 16u) + 16u, slot_cols = round16(max_n + 16u) + 16u;
    const uint32_t bound_cap = max_n + 1u;
    const size_t smem = 2304 + (size_t)GPB * (2u * slot_rows + 2u * slot_cols + round16(bound_cap * 4u) * (AFFINE ? 2u : 1u));
    if (smem > (size_t)ctx.max_smem_optin) return fail(err, SWB_TOO_LONG, "packed wavefront: %zu B of shared memory per block", smem);
    SWB_CUDA(allow_smem((const void *)align_wavefront_packed_general_kernel<G, R, LOCAL, AFFINE>, ctx.device, smem));
    const uint64_t blocks = ((count + 1) / 2 + GPB - 1) / GPB;
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        align_wavefront_packed_general_kernel<G, R, LOCAL, AFFINE><<<(unsigned)blocks, WF_THREADS, smem, stream>>>(
            ps, begin, count, cs, eng->d_tables[dev_index], slot_rows, slot_cols, bound_cap);
    }
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

swb_status_t launch_wavefront_packed_any(swb_engine *eng, DeviceCtx &ctx, int dev_index, const PairSpace &ps, uint64_t begin, uint64_t count,
                                         const AlignCosts &cs, bool local, bool affine, uint32_t max_m, uint32_t max_n, cudaStream_t stream,
                                         const char **err) {
#ifdef SWB_EXPERIMENTS
    static const bool packed_off = getenv("SWB_NO_PACKED") != nullptr;
    if (packed_off) return SWB_TOO_LONG;
#endif
    swb_status_t pst = SWB_TOO_LONG;
    // Global, linear gaps: cells are kept shifted (H - gap*(i+j)), costs and scores must be >= 0 and inside 15 bits.
    const int cm_s = cs.match - 2 * cs.next, cx_s = cs.mismatch - 2 * cs.next;
    if (!local && !affine) {
        if (!(cx_s >= 0 && cm_s >= cx_s && (uint64_t)(std::min(max_m, max_n) + 1u) * (uint64_t)cm_s < 32000u)) return SWB_TOO_LONG;
#define SWB_PK(G, NAME)                                                                                                 \
    if (pst == SWB_TOO_LONG && (max_m <= (uint32_t)(G) * 16u || (G) == 32)) {                                             \
        eng->last_kernel = NAME;                                                                                        \
        pst = launch_wavefront_packed<G, 16>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);     \
    }
        SWB_PK(1, "align_wavefront_packed<G=1,R=16>")
        SWB_PK(2, "align_wavefront_packed<G=2,R=16>")
        SWB_PK(4, "align_wavefront_packed<G=4,R=16>")
        SWB_PK(8, "align_wavefront_packed<G=8,R=16>")
        SWB_PK(16, "align_wavefront_packed<G=16,R=16>")
        SWB_PK(32, "align_wavefront_packed<G=32,R=16>")
#undef SWB_PK
        return pst;
    }
    // Affine gaps and / or local alignment: mismatch <= 0 <= match, gaps <= 0, every reachable score inside int16 with
    // room for the gap tracks' floor (-32000): H >= 3*first + (m+n)*min(next, mismatch), H <= min(m,n)*match. Each
    // single gap cost must itself stay far from the floor: the kernel forms floor + next / floor + first with 16-bit
    // saturating-free adds, so |cost| is capped at 700 (int8 costs, as the reference's, are always inside).
    if (!(cs.mismatch <= 0 && cs.match >= 0 && cs.first <= 0 && cs.next <= 0 && cs.first >= -700 && cs.next >= -700 && cs.mismatch >= -700 &&
          3 * (int64_t)cs.first + (int64_t)(max_m + max_n) * std::min(cs.next, cs.mismatch) > -28000 && (int64_t)std::min(max_m, max_n) * cs.match < 32000))
        return SWB_TOO_LONG;
#define SWB_PG(G, NAME)                                                                                                                             \
    if (pst == SWB_TOO_LONG && (max_m <= (uint32_t)(G) * 16u || (G) == 32)) {                                                                         \
        eng->last_kernel = NAME;                                                                                                                    \
        if (local && affine) pst = launch_wavefront_packed_general<G, 16, true, true>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);  \
        else if (local) pst = launch_wavefront_packed_general<G, 16, true, false>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);      \
        else pst = launch_wavefront_packed_general<G, 16, false, true>(eng, ctx, dev_index, ps, begin, count, cs, max_m, max_n, stream, err);                  \
    }
    SWB_PG(1, "align_wavefront_packed_general<G=1,R=16>")
    SWB_PG(2, "align_wavefront_packed_general<G=2,R=16>")
    SWB_PG(4, "align_wavefront_packed_general<G=4,R=16>")
    SWB_PG(8, "align_wavefront_packed_general<G=8,R=16>")
    SWB_PG(16, "align_wavefront_packed_general<G=16,R=16>")
    SWB_PG(32, "align_wavefront_packed_general<G=32,R=16>")
#undef SWB_PG
    return pst;
}

} // namespace swb
