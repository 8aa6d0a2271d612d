// launch_lev.cu — launchers of the bit-vector Levenshtein kernels (one translation unit so nvcc builds the kernel
// families in parallel; see engine_internal.h).
#include "engine_internal.h"

#include <algorithm>
#include <cstdlib>

#include "lev_bitpar.cuh"
#include "lev_bitpar_coop.cuh"
#include "lev_bitpar_cross.cuh"
#ifdef SWB_EXPERIMENTS
#include "lev_bitpar_ws.cuh"
#endif
#ifdef SWB_EXPERIMENTS
#include "lev_bitpar_pipe.cuh"
#endif
#include "lev_short.cuh"

namespace swb {

#ifdef SWB_EXPERIMENTS
// Warp-specialised form of the thread-per-pair kernel (lev_bitpar_ws.cuh): builder warps fill double-buffered tables while
// compute warps run the column loop. Persistent blocks, three per SM.
template <int K>
static swb_status_t launch_bitpar_ws(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint64_t begin, uint64_t count, cudaStream_t stream,
                                     const char **err, uint32_t n_cap) {
    const size_t smem = WsConfig<K>::SMEM;
    SWB_CUDA(allow_smem((const void *)lev_bitpar_ws_kernel<K>, ctx.device, smem));
    const uint64_t groups = (count + WS_COMPUTE_THREADS - 1) / WS_COMPUTE_THREADS;
    const unsigned blocks = (unsigned)std::min<uint64_t>(groups, (uint64_t)ctx.sm_count * 3);
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        lev_bitpar_ws_kernel<K><<<blocks, WS_THREADS, smem, stream>>>(ps, begin, count, n_cap, (uint32_t)groups);
    }
    g_launches++;
    eng->last_kernel = K == 1 ? "lev_bitpar_ws<K=1>" : "lev_bitpar_ws<K=2>";
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}
#endif

template <int K>
static swb_status_t launch_bitpar(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint64_t begin, uint64_t count,
                                  cudaStream_t stream, const char **err, uint32_t n_cap = 0xFFFFFFFFu, const SpecArgs &sa = SpecArgs{}) {
#ifdef SWB_EXPERIMENTS
    // SWB_WS=1: the warp-specialised form (lev_bitpar_ws.cuh). Measured slower than this kernel (32.0-32.6 vs 34.4 TCUPS at
    // 64x64, profiles/r2_ws_k2.txt), so it is not dispatched by the shipped library.
    if constexpr (K == 2) {
        static const bool ws_on = getenv("SWB_WS") != nullptr;
        if (ws_on && count >= 32768) return launch_bitpar_ws<K>(eng, ctx, ps, begin, count, stream, err, n_cap);
    }
#endif
    constexpr int T = BitparConfig<K>::THREADS;
    const size_t smem = BitparConfig<K>::SMEM;
    SWB_CUDA(allow_smem((const void *)lev_bitpar_kernel<K>, ctx.device, smem));
    const uint64_t blocks = (count + (uint64_t)T * BitparConfig<K>::PAIRS_PER_THREAD - 1) / ((uint64_t)T * BitparConfig<K>::PAIRS_PER_THREAD);
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        lev_bitpar_kernel<K><<<(unsigned)blocks, T, smem, stream>>>(ps, begin, count, n_cap, sa);
    }
    g_launches++;
    static const char *names[] = {"", "lev_bitpar<K=1>", "lev_bitpar<K=2>", "", "lev_bitpar<K=4>", "", "", "", "lev_bitpar<K=8>"};
    eng->last_kernel = names[K];
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

#ifdef SWB_EXPERIMENTS
// Software-pipelined form of the thread-per-pair kernel (lev_bitpar_pipe.cuh) for batches whose every pattern has exactly
// 32 K symbols: persistent threads, two table buffers each, the next pair's tables built inside the current pair's columns.
// SWB_PIPE=1 dispatches it. Bit-exact, measured slower than lev_bitpar_kernel (30.3 vs 34.4 TCUPS at 64x64): both the ALU and
// the shared-memory pipe are near saturation, and three warps per scheduler absorb the bursts worse than seven do.
constexpr int PIPE_BLOCKS_PER_SM = 6;     // 32 KB of tables per block
constexpr uint64_t PIPE_MIN_PAIRS_PER_THREAD = 6; // below that the un-overlapped first build is not amortised
template <int K>
static bool pipe_worth(const DeviceCtx &ctx, uint64_t count) {
    static const bool on = getenv("SWB_PIPE") != nullptr;
    return on && count >= (uint64_t)ctx.sm_count * PIPE_BLOCKS_PER_SM * PipeConfig<K>::THREADS * PIPE_MIN_PAIRS_PER_THREAD;
}
template <int K>
static swb_status_t launch_bitpar_pipe(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint64_t begin, uint64_t count, cudaStream_t stream,
                                       const char **err, uint32_t n_cap) {
    const size_t smem = PipeConfig<K>::SMEM;
    SWB_CUDA(allow_smem((const void *)lev_bitpar_pipe_kernel<K>, ctx.device, smem));
    const unsigned blocks = (unsigned)std::min<uint64_t>((count + PipeConfig<K>::THREADS - 1) / PipeConfig<K>::THREADS, (uint64_t)ctx.sm_count * PIPE_BLOCKS_PER_SM);
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        lev_bitpar_pipe_kernel<K><<<blocks, PipeConfig<K>::THREADS, smem, stream>>>(ps, begin, count, n_cap);
    }
    g_launches++;
    eng->last_kernel = "lev_bitpar_pipe<K=2>";
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}
#endif

// Patterns of at most 32 bytes with the pattern's match masks derived in registers (lev_short.cuh): no tables.
template <int W, bool LONG_PAT = false>
static swb_status_t launch_short(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint64_t begin, uint64_t count,
                                 cudaStream_t stream, const char **err, uint32_t n_cap = 0xFFFFFFFFu, const SpecArgs &sa = SpecArgs{}) {
    const uint64_t blocks = (count + SHORT_THREADS - 1) / SHORT_THREADS;
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        lev_short_kernel<W, LONG_PAT><<<(unsigned)blocks, SHORT_THREADS, 0, stream>>>(ps, begin, count, n_cap, sa);
    }
    g_launches++;
    if (LONG_PAT) eng->last_kernel = W == 2 ? "lev_short<W=2,long>" : (W == 4 ? "lev_short<W=4,long>" : "lev_short<W=6,long>");
    else eng->last_kernel = W == 2 ? "lev_short<W=2>" : (W == 4 ? "lev_short<W=4>" : "lev_short<W=6>");
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

// Lane-group bit-vector kernel: text streamed in shared-memory chunks, patterns beyond 32*G*K rows in row strips
// through a global scratch row per pair — no length limit.
template <int G, int K>
static swb_status_t launch_bitpar_coop(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint64_t begin, uint64_t count,
                                       uint32_t max_m, uint32_t max_n, const char *name, cudaStream_t stream, const char **err) {
    constexpr int PPB = BPC_THREADS / G;
    constexpr uint32_t CAP = 32u * G * K;
    const uint32_t chunk_cols = std::min<uint32_t>(1024u, std::max<uint32_t>(16u, round16(max_n)));
    const size_t smem = (size_t)32 * BPC_THREADS * K * 4 + (size_t)PPB * (CAP + 32u + chunk_cols + 32u);
    uint8_t *scratch = nullptr;
    uint64_t scratch_stride = 0;
    if (max_m > CAP) { // row strips: one byte of horizontal-delta bits per text column and pair
        scratch_stride = round16(max_n) + 16u;
        SWB_CUDA(ctx.hscratch.reserve(count * scratch_stride));
        scratch = (uint8_t *)ctx.hscratch.ptr;
    }
    const uint64_t blocks = (count + PPB - 1) / PPB;
    if (scratch) {
        SWB_CUDA(allow_smem((const void *)lev_bitpar_coop_kernel<G, K, true>, ctx.device, smem));
        ProfileBracket bracket(eng, ctx.device, stream);
        lev_bitpar_coop_kernel<G, K, true><<<(unsigned)blocks, BPC_THREADS, smem, stream>>>(ps, begin, count, chunk_cols, scratch, scratch_stride);
    } else {
        SWB_CUDA(allow_smem((const void *)lev_bitpar_coop_kernel<G, K, false>, ctx.device, smem));
        ProfileBracket bracket(eng, ctx.device, stream);
        lev_bitpar_coop_kernel<G, K, false><<<(unsigned)blocks, BPC_THREADS, smem, stream>>>(ps, begin, count, chunk_cols, scratch, scratch_stride);
    }
    g_launches++;
    eng->last_kernel = name;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

// Cross product with table reuse: one thread per query, candidates swept (lev_bitpar_cross.cuh).
template <int K>
static swb_status_t launch_bitpar_cross(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint32_t n_queries, uint32_t cands_per_block,
                                        const char *name, uint32_t n_cap, const SpecArgs &sa, cudaStream_t stream, const char **err) {
    constexpr int T = CrossConfig<K>::THREADS;
    const uint32_t chunk = 1024;
    const size_t smem = (size_t)32 * T * K * 4 + (((size_t)T * (CROSS_TILE + 1) * 4 + 15) & ~(size_t)15) + chunk + 32;
    SWB_CUDA(allow_smem((const void *)lev_bitpar_cross_kernel<K>, ctx.device, smem));
    const dim3 grid((n_queries + T - 1) / T, (unsigned)((ps.n_c + cands_per_block - 1) / cands_per_block));
    {
        ProfileBracket bracket(eng, ctx.device, stream);
        lev_bitpar_cross_kernel<K><<<grid, T, smem, stream>>>(ps, n_queries, cands_per_block, chunk, n_cap, sa);
    }
    g_launches++;
    eng->last_kernel = name;
    SWB_CUDA(cudaGetLastError());
    return SWB_OK;
}

swb_status_t launch_lev_bitpar(swb_engine *eng, DeviceCtx &ctx, int K, const PairSpace &ps, uint64_t begin, uint64_t count, uint32_t n_cap,
                               const SpecArgs &sa, cudaStream_t stream, const char **err) {
    // K < 0 selects the register kernel of -K pattern words (-K - 10: the longer string is the pattern)
    switch (K) {
    case -12: return launch_short<2, true>(eng, ctx, ps, begin, count, stream, err, n_cap, sa);
    case -14: return launch_short<4, true>(eng, ctx, ps, begin, count, stream, err, n_cap, sa);
    case -16: return launch_short<6, true>(eng, ctx, ps, begin, count, stream, err, n_cap, sa);
    case -2: return launch_short<2>(eng, ctx, ps, begin, count, stream, err, n_cap, sa);
    case -4: return launch_short<4>(eng, ctx, ps, begin, count, stream, err, n_cap, sa);
    case -6: return launch_short<6>(eng, ctx, ps, begin, count, stream, err, n_cap, sa);
    case 102: // every pattern exactly 64 symbols (the shape the experimental pipelined kernel takes)
#ifdef SWB_EXPERIMENTS
        if (pipe_worth<2>(ctx, count) && !sa.report) return launch_bitpar_pipe<2>(eng, ctx, ps, begin, count, stream, err, n_cap);
#endif
        return launch_bitpar<2>(eng, ctx, ps, begin, count, stream, err, n_cap, sa);
    case 1: return launch_bitpar<1>(eng, ctx, ps, begin, count, stream, err, n_cap, sa);
    case 2: return launch_bitpar<2>(eng, ctx, ps, begin, count, stream, err, n_cap, sa);
    case 4: return launch_bitpar<4>(eng, ctx, ps, begin, count, stream, err, n_cap, sa);
    case 8: return launch_bitpar<8>(eng, ctx, ps, begin, count, stream, err, n_cap, sa);
    }
    return fail(err, SWB_INVALID_ARG, "no bit-vector kernel of %d words", K);
}

swb_status_t launch_lev_cross(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint32_t n_queries, uint32_t cands_per_block,
                              uint32_t max_q, uint32_t n_cap, const SpecArgs &sa, cudaStream_t stream, const char **err) {
    if (max_q <= 32) return launch_bitpar_cross<1>(eng, ctx, ps, n_queries, cands_per_block, "lev_bitpar_cross<K=1>", n_cap, sa, stream, err);
    if (max_q <= 64) return launch_bitpar_cross<2>(eng, ctx, ps, n_queries, cands_per_block, "lev_bitpar_cross<K=2>", n_cap, sa, stream, err);
    if (max_q <= 128) return launch_bitpar_cross<4>(eng, ctx, ps, n_queries, cands_per_block, "lev_bitpar_cross<K=4>", n_cap, sa, stream, err);
    return launch_bitpar_cross<8>(eng, ctx, ps, n_queries, cands_per_block, "lev_bitpar_cross<K=8>", n_cap, sa, stream, err);
}

swb_status_t launch_lev_class(swb_engine *eng, DeviceCtx &ctx, const PairSpace &ps, uint64_t begin, uint64_t count, uint32_t max_m,
                              uint32_t max_n, cudaStream_t stream, const char **err, uint32_t min_m) {
#ifdef SWB_EXPERIMENTS
    // SWB_EXP (bit mask, experiments only): 1: <=256 B via lane groups <4,2>; 2: <=1024 B via <8,4>; 4: <=512 B via <4,4>;
    // 8: patterns <= 32 B through the table kernel (K = 1) instead of the register kernel
    static const int exp_mask = getenv("SWB_EXP") ? atoi(getenv("SWB_EXP")) : 0;
    if ((exp_mask & 1) && max_m > 128 && max_m <= 256) return launch_bitpar_coop<4, 2>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=4,K=2>", stream, err);
    if ((exp_mask & 4) && max_m > 256 && max_m <= 512) return launch_bitpar_coop<4, 4>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=4,K=4>", stream, err);
    if ((exp_mask & 2) && max_m > 512 && max_m <= 1024) return launch_bitpar_coop<8, 4>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=8,K=4>", stream, err);
    if ((exp_mask & 8) && max_m <= 32) return launch_bitpar<1>(eng, ctx, ps, begin, count, stream, err);
#endif
    // register kernel up to 24 pattern bytes, thread-per-pair table kernels up to 256, lane-group kernels beyond
    if (max_n <= 8) return launch_short<2, true>(eng, ctx, ps, begin, count, stream, err);
    if (max_n <= 16) return launch_short<4, true>(eng, ctx, ps, begin, count, stream, err);
    if (max_n <= SHORT_MAX_M) return launch_short<6, true>(eng, ctx, ps, begin, count, stream, err);
    if (max_m <= 8) return launch_short<2>(eng, ctx, ps, begin, count, stream, err);
    if (max_m <= 16) return launch_short<4>(eng, ctx, ps, begin, count, stream, err);
    if (max_m <= SHORT_MAX_M) return launch_short<6>(eng, ctx, ps, begin, count, stream, err);
    if (max_m <= 32) return launch_bitpar<1>(eng, ctx, ps, begin, count, stream, err);
#ifdef SWB_EXPERIMENTS
    if (max_m == 64 && min_m == 64 && pipe_worth<2>(ctx, count)) return launch_bitpar_pipe<2>(eng, ctx, ps, begin, count, stream, err, 0xFFFFFFFFu);
#else
    (void)min_m;
#endif
    if (max_m <= 64) return launch_bitpar<2>(eng, ctx, ps, begin, count, stream, err);
    if (max_m <= 128) return launch_bitpar<4>(eng, ctx, ps, begin, count, stream, err);
    if (max_m <= 256) return launch_bitpar<8>(eng, ctx, ps, begin, count, stream, err);
    if (max_m <= 512) return launch_bitpar_coop<8, 2>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=8,K=2>", stream, err);
    if (max_m <= 1024) return launch_bitpar_coop<16, 2>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=16,K=2>", stream, err);
    if (max_m <= 2048) return launch_bitpar_coop<32, 2>(eng, ctx, ps, begin, count, max_m, max_n, "lev_bitpar_coop<G=32,K=2>", stream, err);
    return launch_bitpar_coop<32, 4>(eng, ctx, ps, begin, count, max_m, max_n, max_m <= 4096 ? "lev_bitpar_coop<G=32,K=4>" : "lev_bitpar_coop<G=32,K=4,strips>", stream, err);
}

} // namespace swb
