// multi_gpu.cu — the exchange BASELINE.json's north_star names, inside the C ABI: "the batch of pairs shards trivially
// across the 8 GPUs of one box with a single NCCL scatter/gather over NVLink".
//
// `swb_compute_pairs_scatter` (include/stringwars_b200.h): the batch is resident on the scope's first device (the
// root). One single-process NCCL communicator per scope (ncclCommInitAll over the scope's devices; libnccl is loaded
// with dlopen so that the library itself carries no link-time dependency and single-GPU users never touch it).
//
//   root                                   device d > 0
//   ----                                   ------------
//   cut_search_kernel: D*C byte-balanced cut points found on the device (binary search in the offsets), one small readback
//   rebase_offsets_kernel: u32 offsets relative to each shard's first byte (4 B per string on the wire instead of 8)
//   for chunk c: ncclGroupStart                         (the C chunks of a shard travel in C groups)
//     ncclSend(bytes of A, bytes of B, u32 offsets) ──► ncclRecv into the shard's buffers, event(d, c)
//   ncclGroupEnd
//   its own shard: kernels on the tapes in place        per chunk: wait event(d, c), schedule + kernels
//   ncclRecv(results of every shard into `out`)   ◄──── ncclSend(results)        (ONE grouped gather)
//
// A chunk's kernel therefore runs while the next chunk is still on NVLink; the root's shard never moves.
// There is no CPU path here: without NCCL the call fails with SWB_NCCL_ERROR.
#include "engine_internal.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <dlfcn.h>
#include <mutex>

namespace swb {

// ------------------------------------------------------------------------------------------ NCCL, loaded at run time
// The subset of nccl.h this file calls (NCCL 2.x ABI; values of ncclDataType_t are stable: ncclUint8 = 1).
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclUint8 = 1;

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mutex;

static bool load_nccl(const char **why) {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.handle) return true;
    const char *names[] = {getenv("SWB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    }
    if (!h) {
        *why = dlerror();
        return false;
    }
    NcclApi api;
    api.handle = h;
#define SWB_SYM(field, name)                                            \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name)); \
    if (!api.field) {                                                   \
        *why = "symbol " name " not found in libnccl";                  \
        return false;                                                   \
    }
    SWB_SYM(CommInitAll, "ncclCommInitAll")
    SWB_SYM(CommDestroy, "ncclCommDestroy")
    SWB_SYM(GroupStart, "ncclGroupStart")
    SWB_SYM(GroupEnd, "ncclGroupEnd")
    SWB_SYM(Send, "ncclSend")
    SWB_SYM(Recv, "ncclRecv")
    SWB_SYM(GetErrorString, "ncclGetErrorString")
    SWB_SYM(GetVersion, "ncclGetVersion")
#undef SWB_SYM
    g_nccl = api;
    return true;
}

struct NcclState {
    std::vector<ncclComm_t> comms;
    uint64_t *h_table = nullptr; // pinned: cut table read back from the root
    uint64_t *d_table = nullptr; // root device
    DeviceBuffer rel_a, rel_b;   // root: u32 offsets relative to their shard's first byte
    cudaEvent_t ev_root = nullptr;
};

constexpr uint32_t MAX_PARTS = 8 * 64;

void nccl_state_destroy(swb_scope *scope) {
    NcclState *ns = (NcclState *)scope->nccl;
    if (!ns) return;
    for (size_t i = 0; i < ns->comms.size(); ++i)
        if (ns->comms[i]) {
            cudaSetDevice(scope->devs[i].device);
            g_nccl.CommDestroy(ns->comms[i]);
        }
    ns->comms.clear();
    cudaSetDevice(scope->devs[0].device);
    if (ns->h_table) cudaFreeHost(ns->h_table);
    if (ns->d_table) cudaFree(ns->d_table);
    if (ns->ev_root) cudaEventDestroy(ns->ev_root);
    ns->rel_a.release(), ns->rel_b.release();
    delete ns;
    scope->nccl = nullptr;
}

#define SWB_NCCL(call)                                                                                         \
    do {                                                                                                       \
        ncclResult_t r_ = (call);                                                                              \
        if (r_ != 0) return fail(err, SWB_NCCL_ERROR, "%s failed: %s", #call, g_nccl.GetErrorString(r_));      \
    } while (0)

// The scatter call's per-scope state; the NCCL communicator is only created when the NCCL transport is used.
static swb_status_t nccl_state(swb_scope *scope, bool want_comm, NcclState **out, const char **err) {
    const int D = (int)scope->devs.size();
    NcclState *ns = (NcclState *)scope->nccl;
    if (!ns) {
        ns = new NcclState();
        scope->nccl = ns;
        cudaError_t e = cudaSetDevice(scope->devs[0].device);
        if (e == cudaSuccess) e = cudaHostAlloc(&ns->h_table, 3 * (MAX_PARTS + 1) * sizeof(uint64_t), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaMalloc(&ns->d_table, 3 * (MAX_PARTS + 1) * sizeof(uint64_t));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ns->ev_root, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            nccl_state_destroy(scope);
            return fail(err, SWB_CUDA_ERROR, "scatter state setup failed: %s", cudaGetErrorString(e));
        }
    }
    if (want_comm && ns->comms.empty()) {
        const char *why = "";
        if (!load_nccl(&why)) return fail(err, SWB_NCCL_ERROR, "libnccl could not be loaded (%s); set SWB_NCCL_LIB to its path", why ? why : "?");
        std::vector<int> devlist(D);
        for (int i = 0; i < D; ++i) devlist[i] = scope->devs[i].device;
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < i; ++j)
                if (devlist[i] == devlist[j]) return fail(err, SWB_INVALID_ARG, "an NCCL scatter needs distinct devices; device %d appears twice in the scope", devlist[i]);
        ns->comms.assign(D, nullptr);
        ncclResult_t r = g_nccl.CommInitAll(ns->comms.data(), D, devlist.data());
        if (r != 0) {
            ns->comms.clear();
            return fail(err, SWB_NCCL_ERROR, "ncclCommInitAll over %d devices failed: %s", D, g_nccl.GetErrorString(r));
        }
    }
    *out = ns;
    return SWB_OK;
}

// Direct peer access between the root and every other device of the scope (both directions), enabled once.
static swb_status_t enable_peers(swb_scope *scope, const char **err) {
    DeviceCtx &root = scope->devs[0];
    for (size_t d = 1; d < scope->devs.size(); ++d) {
        DeviceCtx &ctx = scope->devs[d];
        if (ctx.device == root.device) continue;
        for (int dir = 0; dir < 2; ++dir) {
            const int from = dir ? root.device : ctx.device, to = dir ? ctx.device : root.device;
            int can = 0;
            SWB_CUDA(cudaDeviceCanAccessPeer(&can, from, to));
            if (!can) return fail(err, SWB_UNSUPPORTED, "device %d cannot access device %d's memory", from, to);
            SWB_CUDA(cudaSetDevice(from));
            cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
            else if (e != cudaSuccess) return fail(err, SWB_CUDA_ERROR, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", from, to, cudaGetErrorString(e));
        }
    }
    return SWB_OK;
}

// ------------------------------------------------------------------------------------------ root-side kernels

// table[t] = first pair of part t (parts + 1 entries), then the A and B offsets at those pairs: parts of equal BYTES.
__global__ void cut_search_kernel(PairSpace ps, uint32_t parts, uint64_t *table) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > parts) return;
    const uint64_t n = ps.n_pairs;
    const uint64_t a0 = a_off(ps, 0), b0 = b_off(ps, 0);
    const uint64_t total = (a_off(ps, n) - a0) + (b_off(ps, n) - b0);
    uint64_t cut = n;
    if (t < parts) {
        const uint64_t target = total / parts * t;
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if ((a_off(ps, mid) - a0) + (b_off(ps, mid) - b0) < target) lo = mid + 1;
            else hi = mid;
        }
        cut = lo;
    }
    table[t] = cut;
    table[parts + 1 + t] = a_off(ps, cut);
    table[2 * (parts + 1) + t] = b_off(ps, cut);
}

// Chunk k = blockIdx.y (parts of the non-root shards): its count + 1 offsets, relative to the first byte of the shard the
// chunk belongs to, written at rel[cut_k - cut_first + (k - k_first)] (every chunk owns its closing offset).
__global__ void rebase_offsets_kernel(PairSpace ps, const uint64_t *table, uint32_t parts, uint32_t chunks_per_shard, uint32_t k_first,
                                      uint32_t *rel_a, uint32_t *rel_b) {
    const uint32_t k = k_first + blockIdx.y;
    const uint64_t cut0 = table[k], cut1 = table[k + 1], first = table[k_first];
    const uint32_t shard_first_part = k / chunks_per_shard * chunks_per_shard;
    const uint64_t a_base = table[parts + 1 + shard_first_part], b_base = table[2 * (parts + 1) + shard_first_part];
    const uint64_t dst = cut0 - first + (k - k_first);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= cut1 - cut0; i += (uint64_t)gridDim.x * blockDim.x) {
        rel_a[dst + i] = (uint32_t)(a_off(ps, cut0 + i) - a_base);
        rel_b[dst + i] = (uint32_t)(b_off(ps, cut0 + i) - b_base);
    }
}

} // namespace swb

using namespace swb;

static swb_status_t root_pointer(const swb_scope *scope, const void *p, const char *what, const char **err) {
    cudaPointerAttributes attr;
    if (!p || cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(err, SWB_INVALID_ARG, "%s is not a CUDA pointer", what);
    }
    if (attr.type == cudaMemoryTypeManaged) return SWB_OK;
    if (attr.type != cudaMemoryTypeDevice || attr.device != scope->devs[0].device)
        return fail(err, SWB_INVALID_ARG, "%s must be device memory of the scope's first device (%d)", what, scope->devs[0].device);
    return SWB_OK;
}

extern "C" swb_status_t swb_compute_pairs_scatter(swb_engine_t *engine, const swb_tape_view_t *a, const swb_tape_view_t *b, void *out,
                                                  swb_result_type_t out_type, double stats[4], const char **err) {
    swb_status_t st = check_engine_alive(engine, err);
    if (st != SWB_OK) return st;
    if (!a || !b || !out) return fail(err, SWB_INVALID_ARG, "NULL argument");
    if (a->n != b->n) return fail(err, SWB_INVALID_ARG, "tapes differ in length: %zu vs %zu", a->n, b->n);
    if ((a->offset_bits != 64 && a->offset_bits != 32) || (b->offset_bits != 64 && b->offset_bits != 32)) return fail(err, SWB_INVALID_ARG, "offset_bits must be 64 or 32");
    if (out_type != SWB_RESULT_64 && out_type != SWB_RESULT_32 && out_type != SWB_RESULT_16) return fail(err, SWB_INVALID_ARG, "unknown result type %d", (int)out_type);
    const uint64_t n = a->n;
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (n == 0) return SWB_OK;
    if (n >= 0xFFFFFFFFull) return fail(err, SWB_INVALID_ARG, "more than 2^32-1 pairs in one call");
    swb_scope *scope = engine->scope;
    const int D = (int)scope->devs.size();
    if ((st = root_pointer(scope, a->data, "a.data", err)) != SWB_OK || (st = root_pointer(scope, a->offsets, "a.offsets", err)) != SWB_OK ||
        (st = root_pointer(scope, b->data, "b.data", err)) != SWB_OK || (st = root_pointer(scope, b->offsets, "b.offsets", err)) != SWB_OK ||
        (st = root_pointer(scope, out, "out", err)) != SWB_OK)
        return st;
    const uint32_t out_narrow = (uint32_t)out_type;
    const size_t elt = result_bytes(out_narrow);
    const uint32_t offs32 = (a->offset_bits == 32 ? 1u : 0u) | (b->offset_bits == 32 ? 2u : 0u);
    const uint32_t aw = a->offset_bits / 8, bw = b->offset_bits / 8;
    DeviceCtx &root = scope->devs[0];
    SWB_CUDA(cudaSetDevice(root.device));
    const PairSpace whole = make_pair_space(a->data, a->offsets, b->data, b->offsets, n, 0, 0, out, out_narrow, offs32);
    if (D == 1) {
        SWB_CUDA(cudaEventRecord(root.ev_t0, root.compute_stream));
        if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) return fail(err, SWB_UNSUPPORTED, "the scatter call runs the code-point engine only on multi-device scopes");
        if ((st = run_device_batch(engine, 0, whole, root.compute_stream, err)) != SWB_OK) return st;
        SWB_CUDA(cudaEventRecord(root.ev_t1, root.compute_stream));
        SWB_CUDA(cudaStreamSynchronize(root.compute_stream));
        float ms = 0;
        SWB_CUDA(cudaEventElapsedTime(&ms, root.ev_t0, root.ev_t1));
        if (stats) stats[2] = ms;
        return SWB_OK;
    }
    const bool use_nccl = scope->transport == SWB_TRANSPORT_NCCL;
    NcclState *ns = nullptr;
    if ((st = nccl_state(scope, use_nccl, &ns, err)) != SWB_OK) return st;
    if (!use_nccl) {
        if (!scope->peers_enabled && (st = enable_peers(scope, err)) != SWB_OK) return st;
        scope->peers_enabled = true;
    }
    // One chunk buffer root -> device d (arrives on d's communication stream), either transport.
    auto send_chunk = [&](DeviceCtx &ctx, int d, void *dst, const void *src, size_t bytes) -> swb_status_t {
        if (use_nccl) {
            SWB_NCCL(g_nccl.Send(src, bytes, kNcclUint8, d, ns->comms[0], root.comm_stream));
            SWB_NCCL(g_nccl.Recv(dst, bytes, kNcclUint8, 0, ns->comms[d], ctx.comm_stream));
        } else {
            SWB_CUDA(cudaMemcpyPeerAsync(dst, ctx.device, src, root.device, bytes, ctx.comm_stream));
        }
        return SWB_OK;
    };

    // total bytes -> chunks per shard (about 64 MiB each: long enough for NVLink to run at rate, short enough that
    // the first kernel starts early), then the byte-balanced cut table
    uint64_t ends[4] = {0, 0, 0, 0};
    SWB_CUDA(cudaMemcpyAsync(&ends[0], a->offsets, aw, cudaMemcpyDeviceToHost, root.comm_stream));
    SWB_CUDA(cudaMemcpyAsync(&ends[1], (const uint8_t *)a->offsets + n * aw, aw, cudaMemcpyDeviceToHost, root.comm_stream));
    SWB_CUDA(cudaMemcpyAsync(&ends[2], b->offsets, bw, cudaMemcpyDeviceToHost, root.comm_stream));
    SWB_CUDA(cudaMemcpyAsync(&ends[3], (const uint8_t *)b->offsets + n * bw, bw, cudaMemcpyDeviceToHost, root.comm_stream));
    SWB_CUDA(cudaStreamSynchronize(root.comm_stream));
    if (ends[1] < ends[0] || ends[3] < ends[2]) return fail(err, SWB_INVALID_ARG, "offsets are not monotonic");
    const uint64_t total_bytes = (ends[1] - ends[0]) + (ends[3] - ends[2]);
    const uint32_t C = (uint32_t)std::min<uint64_t>(16, std::max<uint64_t>(1, total_bytes / D / (64ull << 20)));
    const uint32_t parts = (uint32_t)D * C;
    cut_search_kernel<<<(parts + 1 + 63) / 64, 64, 0, root.comm_stream>>>(whole, parts, ns->d_table);
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    SWB_CUDA(cudaMemcpyAsync(ns->h_table, ns->d_table, 3 * (parts + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, root.comm_stream));
    SWB_CUDA(cudaStreamSynchronize(root.comm_stream));
    const uint64_t *cut = ns->h_table, *a_at = ns->h_table + (parts + 1), *b_at = ns->h_table + 2 * (parts + 1);
    for (uint32_t k = 0; k < parts; ++k)
        if (cut[k + 1] < cut[k] || a_at[k + 1] < a_at[k] || b_at[k + 1] < b_at[k]) return fail(err, SWB_INVALID_ARG, "offsets are not monotonic");
    for (int d = 1; d < D; ++d)
        if (a_at[(d + 1) * C] - a_at[d * C] >= (1ull << 32) || b_at[(d + 1) * C] - b_at[d * C] >= (1ull << 32))
            return fail(err, SWB_TOO_LONG, "a shard of 4 GiB or more cannot travel with 32-bit offsets: use smaller batches");

    // receive buffers
    for (int d = 1; d < D; ++d) {
        DeviceCtx &ctx = scope->devs[d];
        const uint64_t n_d = cut[(d + 1) * C] - cut[d * C];
        if (n_d == 0) continue;
        SWB_CUDA(cudaSetDevice(ctx.device));
        SWB_CUDA(ctx.rx_a.reserve(a_at[(d + 1) * C] - a_at[d * C] + 64));
        SWB_CUDA(ctx.rx_b.reserve(b_at[(d + 1) * C] - b_at[d * C] + 64));
        SWB_CUDA(ctx.rx_a_offs.reserve((n_d + C + 1) * sizeof(uint32_t)));
        SWB_CUDA(ctx.rx_b_offs.reserve((n_d + C + 1) * sizeof(uint32_t)));
        SWB_CUDA(ctx.rx_out.reserve(n_d * elt));
        while (ctx.ev_chunks.size() < C + 1) {
            cudaEvent_t ev;
            SWB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx.ev_chunks.push_back(ev);
        }
    }
    SWB_CUDA(cudaSetDevice(root.device));
    const uint64_t n_remote = n - cut[C];
    SWB_CUDA(ns->rel_a.reserve((n_remote + parts + 1) * sizeof(uint32_t)));
    SWB_CUDA(ns->rel_b.reserve((n_remote + parts + 1) * sizeof(uint32_t)));
    uint32_t *rel_a = (uint32_t *)ns->rel_a.ptr, *rel_b = (uint32_t *)ns->rel_b.ptr;

    const auto wall0 = std::chrono::steady_clock::now();
    SWB_CUDA(cudaEventRecord(root.ev_t0, root.comm_stream));
    if (n_remote) {
        const dim3 grid((unsigned)std::min<uint64_t>(64, (n_remote / (parts - C) + 256) / 256), parts - C);
        rebase_offsets_kernel<<<grid, 256, 0, root.comm_stream>>>(whole, ns->d_table, parts, C, C, rel_a, rel_b);
        g_launches++;
        SWB_CUDA(cudaGetLastError());
    }
    if (!use_nccl) { // copy engines: the copies run on the destinations' streams, after the root has rebuilt the offsets
        SWB_CUDA(cudaEventRecord(ns->ev_root, root.comm_stream));
        for (int d = 1; d < D; ++d) {
            SWB_CUDA(cudaSetDevice(scope->devs[d].device));
            SWB_CUDA(cudaStreamWaitEvent(scope->devs[d].comm_stream, ns->ev_root, 0));
        }
        SWB_CUDA(cudaSetDevice(root.device));
    }
    // scatter: chunk c of every remote shard in one group
    uint64_t sent = 0;
    for (uint32_t c = 0; c < C; ++c) {
        if (use_nccl) SWB_NCCL(g_nccl.GroupStart());
        for (int d = 1; d < D; ++d) {
            DeviceCtx &ctx = scope->devs[d];
            const uint32_t k = d * C + c, k0 = d * C;
            const uint64_t cnt = cut[k + 1] - cut[k];
            if (cnt == 0) continue;
            const uint64_t ab = a_at[k + 1] - a_at[k], bb = b_at[k + 1] - b_at[k];
            uint8_t *rx_a = (uint8_t *)ctx.rx_a.ptr + (a_at[k0] & 15u) + (a_at[k] - a_at[k0]);
            uint8_t *rx_b = (uint8_t *)ctx.rx_b.ptr + (b_at[k0] & 15u) + (b_at[k] - b_at[k0]);
            const uint64_t rel_src = cut[k] - cut[C] + (k - C), rel_dst = cut[k] - cut[k0] + c;
            if (ab && (st = send_chunk(ctx, d, rx_a, a->data + a_at[k], ab)) != SWB_OK) return st;
            if (bb && (st = send_chunk(ctx, d, rx_b, b->data + b_at[k], bb)) != SWB_OK) return st;
            if ((st = send_chunk(ctx, d, (uint32_t *)ctx.rx_a_offs.ptr + rel_dst, rel_a + rel_src, (cnt + 1) * 4)) != SWB_OK) return st;
            if ((st = send_chunk(ctx, d, (uint32_t *)ctx.rx_b_offs.ptr + rel_dst, rel_b + rel_src, (cnt + 1) * 4)) != SWB_OK) return st;
            sent += ab + bb + 2 * (cnt + 1) * 4;
        }
        if (use_nccl) SWB_NCCL(g_nccl.GroupEnd());
        for (int d = 1; d < D; ++d) {
            DeviceCtx &ctx = scope->devs[d];
            if (cut[d * C + c + 1] == cut[d * C + c]) continue;
            SWB_CUDA(cudaSetDevice(ctx.device));
            SWB_CUDA(cudaEventRecord(ctx.ev_chunks[c], ctx.comm_stream));
        }
    }
    // the root's own shard, in place
    SWB_CUDA(cudaSetDevice(root.device));
    if (cut[C] > 0) {
        PairSpace ps = whole;
        ps.n_pairs = cut[C];
        if (engine->kind == ENGINE_LEVENSHTEIN_UTF8)
            st = run_device_batch_utf8(engine, 0, ps, ps.n_pairs, ps.n_pairs, a_at[C] - a_at[0], b_at[C] - b_at[0], root.compute_stream, err);
        else
            st = run_device_batch(engine, 0, ps, root.compute_stream, err);
        if (st != SWB_OK) return st;
    }
    // remote chunks: kernels as the data lands
    for (uint32_t c = 0; c < C; ++c)
        for (int d = 1; d < D; ++d) {
            DeviceCtx &ctx = scope->devs[d];
            const uint32_t k = d * C + c, k0 = d * C;
            const uint64_t cnt = cut[k + 1] - cut[k];
            if (cnt == 0) continue;
            SWB_CUDA(cudaSetDevice(ctx.device));
            SWB_CUDA(cudaStreamWaitEvent(ctx.compute_stream, ctx.ev_chunks[c], 0));
            const uint64_t rel_dst = cut[k] - cut[k0] + c;
            PairSpace ps = make_pair_space((const uint8_t *)ctx.rx_a.ptr + (a_at[k0] & 15u), (const uint32_t *)ctx.rx_a_offs.ptr + rel_dst,
                                           (const uint8_t *)ctx.rx_b.ptr + (b_at[k0] & 15u), (const uint32_t *)ctx.rx_b_offs.ptr + rel_dst, cnt, 0, 0,
                                           (uint8_t *)ctx.rx_out.ptr + (cut[k] - cut[k0]) * elt, out_narrow, 3u);
            if (engine->kind == ENGINE_LEVENSHTEIN_UTF8)
                st = run_device_batch_utf8(engine, d, ps, cnt, cnt, a_at[k + 1] - a_at[k], b_at[k + 1] - b_at[k], ctx.compute_stream, err);
            else
                st = run_device_batch(engine, d, ps, ctx.compute_stream, err);
            if (st != SWB_OK) return st;
        }
    // gather: every remote shard's results in ONE group
    uint64_t gathered = 0;
    for (int d = 1; d < D; ++d) {
        DeviceCtx &ctx = scope->devs[d];
        if (cut[(d + 1) * C] == cut[d * C]) continue;
        SWB_CUDA(cudaSetDevice(ctx.device));
        SWB_CUDA(cudaEventRecord(ctx.ev_chunks[C], ctx.compute_stream));
        SWB_CUDA(cudaStreamWaitEvent(ctx.comm_stream, ctx.ev_chunks[C], 0));
    }
    if (use_nccl) SWB_NCCL(g_nccl.GroupStart());
    for (int d = 1; d < D; ++d) {
        DeviceCtx &ctx = scope->devs[d];
        const uint64_t n_d = cut[(d + 1) * C] - cut[d * C];
        if (n_d == 0) continue;
        if (use_nccl) {
            SWB_NCCL(g_nccl.Send(ctx.rx_out.ptr, n_d * elt, kNcclUint8, 0, ns->comms[d], ctx.comm_stream));
            SWB_NCCL(g_nccl.Recv((uint8_t *)out + cut[d * C] * elt, n_d * elt, kNcclUint8, d, ns->comms[0], root.comm_stream));
        } else {
            SWB_CUDA(cudaMemcpyPeerAsync((uint8_t *)out + cut[d * C] * elt, root.device, ctx.rx_out.ptr, ctx.device, n_d * elt, ctx.comm_stream));
        }
        gathered += n_d * elt;
    }
    if (use_nccl) SWB_NCCL(g_nccl.GroupEnd());
    SWB_CUDA(cudaSetDevice(root.device));
    SWB_CUDA(cudaEventRecord(ns->ev_root, root.compute_stream));
    SWB_CUDA(cudaStreamWaitEvent(root.comm_stream, ns->ev_root, 0));
    SWB_CUDA(cudaEventRecord(root.ev_t1, root.comm_stream));
    for (int d = 0; d < D; ++d) {
        SWB_CUDA(cudaSetDevice(scope->devs[d].device));
        SWB_CUDA(cudaStreamSynchronize(scope->devs[d].comm_stream));
        SWB_CUDA(cudaStreamSynchronize(scope->devs[d].compute_stream));
    }
    if (stats) {
        float ms = 0;
        SWB_CUDA(cudaSetDevice(root.device));
        SWB_CUDA(cudaEventElapsedTime(&ms, root.ev_t0, root.ev_t1));
        // copy-engine transport: the last copies end on the other devices' streams, so the root's event does not see them
        if (!use_nccl) ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
        stats[0] = (double)sent, stats[1] = (double)gathered, stats[2] = ms, stats[3] = (double)C;
    }
    return SWB_OK;
}

// ------------------------------------------------------------------------------------------ peer loads instead of a scatter

namespace swb {
struct PeerState {
    uint64_t *h_table = nullptr, *d_table = nullptr;
    std::vector<char> enabled;
};
void peer_state_destroy(swb_scope *scope) {
    PeerState *st = (PeerState *)scope->peer;
    if (!st) return;
    cudaSetDevice(scope->devs[0].device);
    if (st->h_table) cudaFreeHost(st->h_table);
    if (st->d_table) cudaFree(st->d_table);
    delete st;
    scope->peer = nullptr;
}
} // namespace swb

extern "C" swb_status_t swb_compute_pairs_peer(swb_engine_t *engine, const swb_tape_view_t *a, const swb_tape_view_t *b, void *out,
                                               swb_result_type_t out_type, double stats[4], const char **err) {
    swb_status_t st = check_engine_alive(engine, err);
    if (st != SWB_OK) return st;
    if (!a || !b || !out) return fail(err, SWB_INVALID_ARG, "NULL argument");
    if (a->n != b->n) return fail(err, SWB_INVALID_ARG, "tapes differ in length: %zu vs %zu", a->n, b->n);
    if ((a->offset_bits != 64 && a->offset_bits != 32) || (b->offset_bits != 64 && b->offset_bits != 32)) return fail(err, SWB_INVALID_ARG, "offset_bits must be 64 or 32");
    if (out_type != SWB_RESULT_64 && out_type != SWB_RESULT_32 && out_type != SWB_RESULT_16) return fail(err, SWB_INVALID_ARG, "unknown result type %d", (int)out_type);
    if (engine->kind == ENGINE_LEVENSHTEIN_UTF8) return fail(err, SWB_UNSUPPORTED, "the peer-load call runs the byte engines only");
    const uint64_t n = a->n;
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (n == 0) return SWB_OK;
    if (n >= 0xFFFFFFFFull) return fail(err, SWB_INVALID_ARG, "more than 2^32-1 pairs in one call");
    swb_scope *scope = engine->scope;
    const int D = (int)scope->devs.size();
    if ((st = root_pointer(scope, a->data, "a.data", err)) != SWB_OK || (st = root_pointer(scope, a->offsets, "a.offsets", err)) != SWB_OK ||
        (st = root_pointer(scope, b->data, "b.data", err)) != SWB_OK || (st = root_pointer(scope, b->offsets, "b.offsets", err)) != SWB_OK ||
        (st = root_pointer(scope, out, "out", err)) != SWB_OK)
        return st;
    const uint32_t out_narrow = (uint32_t)out_type;
    const size_t elt = result_bytes(out_narrow);
    const uint32_t offs32 = (a->offset_bits == 32 ? 1u : 0u) | (b->offset_bits == 32 ? 2u : 0u);
    const uint32_t aw = a->offset_bits / 8, bw = b->offset_bits / 8;
    DeviceCtx &root = scope->devs[0];
    PeerState *ps_state = (PeerState *)scope->peer;
    if (!ps_state) {
        ps_state = new PeerState();
        ps_state->enabled.assign(D, 0);
        scope->peer = ps_state;
        SWB_CUDA(cudaSetDevice(root.device));
        SWB_CUDA(cudaHostAlloc(&ps_state->h_table, 3 * (MAX_PARTS + 1) * sizeof(uint64_t), cudaHostAllocDefault));
        SWB_CUDA(cudaMalloc(&ps_state->d_table, 3 * (MAX_PARTS + 1) * sizeof(uint64_t)));
    }
    if (!scope->peers_enabled && (st = enable_peers(scope, err)) != SWB_OK) return st;
    scope->peers_enabled = true;
    SWB_CUDA(cudaSetDevice(root.device));
    const PairSpace whole = make_pair_space(a->data, a->offsets, b->data, b->offsets, n, 0, 0, out, out_narrow, offs32);
    // D byte-balanced shards, cut on the device (the offsets never visit the host)
    cut_search_kernel<<<1, 64, 0, root.comm_stream>>>(whole, (uint32_t)D, ps_state->d_table);
    g_launches++;
    SWB_CUDA(cudaGetLastError());
    SWB_CUDA(cudaMemcpyAsync(ps_state->h_table, ps_state->d_table, 3 * (D + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, root.comm_stream));
    SWB_CUDA(cudaStreamSynchronize(root.comm_stream));
    const uint64_t *cut = ps_state->h_table, *a_at = cut + (D + 1), *b_at = cut + 2 * (D + 1);
    for (int d = 0; d < D; ++d)
        if (cut[d + 1] < cut[d] || a_at[d + 1] < a_at[d] || b_at[d + 1] < b_at[d]) return fail(err, SWB_INVALID_ARG, "offsets are not monotonic");
    const auto t0 = std::chrono::steady_clock::now();
    uint64_t remote_read = 0, remote_written = 0;
    for (int d = 0; d < D; ++d) {
        const uint64_t cnt = cut[d + 1] - cut[d];
        if (cnt == 0) continue;
        DeviceCtx &ctx = scope->devs[d];
        SWB_CUDA(cudaSetDevice(ctx.device));
        PairSpace ps = make_pair_space(a->data, (const uint8_t *)a->offsets + cut[d] * aw, b->data, (const uint8_t *)b->offsets + cut[d] * bw, cnt, 0, 0,
                                       (uint8_t *)out + cut[d] * elt, out_narrow, offs32);
        if ((st = run_device_batch(engine, d, ps, ctx.compute_stream, err)) != SWB_OK) return st;
        if (d > 0) remote_read += (a_at[d + 1] - a_at[d]) + (b_at[d + 1] - b_at[d]) + (cnt + 1) * (aw + bw), remote_written += cnt * elt;
    }
    for (int d = 0; d < D; ++d) {
        SWB_CUDA(cudaSetDevice(scope->devs[d].device));
        SWB_CUDA(cudaStreamSynchronize(scope->devs[d].compute_stream));
    }
    if (stats) {
        stats[0] = (double)remote_read, stats[1] = (double)remote_written;
        stats[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return SWB_OK;
}
