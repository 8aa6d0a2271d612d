// align_wavefront.cuh — anti-diagonal wavefront DP for Levenshtein / Needleman-Wunsch / Smith-Waterman, sm_100a.
//
// This is the kernel BASELINE.json's north_star names: a sub-warp GROUP of G lanes owns one string pair; both
// strings are staged into shared memory with coalesced 16-byte read-only loads; lane l keeps R consecutive DP
// rows in registers and walks the columns skewed by l, so that at step t the group is working on one
// anti-diagonal band of the matrix. The only inter-lane traffic is the bottom cell of each lane's strip,
// handed to lane l+1 with one `__shfl_up_sync` per step; the three-way max/min itself maps to DPX
// VIADDMNMX/VIMNMX3 integer instructions. Integer DP, no tensor cores.
//
// One recurrence covers every engine of similarities/bench.rs (maximising form):
//     H[i][j] = max(H[i-1][j-1] + S(a_i, b_j),  E[i][j],  F[i][j])          (clamped at 0 when LOCAL)
//     E[i][j] = max(E[i-1][j] + next, H[i-1][j] + first)    gap running down a column
//     F[i][j] = max(F[i][j-1] + next, H[i][j-1] + first)    gap running along a row
// Linear gaps (first == next) collapse E/F into H+gap. Levenshtein is S = -(a != b), gap = -1, distance = -H.
// Patterns longer than G*R rows are processed in row strips; the strip's bottom row is parked in a per-pair
// shared-memory boundary row (int16 when scores fit, else int32) and read back as the next strip's top row.
//
// Reference semantics: `LevenshteinDistances` (bench.rs:382-393), `NeedlemanWunschScores` /
// `SmithWatermanScores` with `unary_class_costs` tables (bench.rs:98-108, 655-674, 982-985).
#pragma once
#include "common.cuh"

namespace swb {

constexpr int WF_THREADS = 128;
constexpr int NEG_BIG = -(1 << 28);

struct AlignCosts {
    int match, mismatch; // unary fast path: S = classes equal ? match : mismatch
    int first, next;     // gap costs (<= 0 in maximising form): length-k gap = first + (k-1)*next
    int unary;           // 1: compare classes, 0: look up the 32x32 table
    int negate;          // 1: write -H (Levenshtein distance)
};

struct AlignTables {
    uint8_t byte_to_class[256];
    int8_t costs[32 * 32];   // costs[class(a)][class(b)]
    int8_t costs_t[32 * 32]; // transposed, used when the kernel swapped the pair so rows = shorter string
};

template <typename BoundT>
__device__ __forceinline__ int bound_load(const BoundT *p) {
    return (int)*p;
}

// G lanes per pair, R rows per lane. LOCAL: Smith-Waterman. AFFINE: Gotoh E/F tracks.
template <int G, int R, bool LOCAL, bool AFFINE, typename BoundT>
__global__ void __launch_bounds__(WF_THREADS)
    align_wavefront_kernel(PairSpace ps, uint64_t begin, uint64_t count, AlignCosts cs, const AlignTables *__restrict__ tables,
                           uint32_t slot_rows /* bytes reserved for the row string */, uint32_t slot_cols, uint32_t bound_cap) {
    constexpr int PAIRS_PER_BLOCK = WF_THREADS / G;
    extern __shared__ __align__(16) uint8_t smem[];
    // layout: [tables 2304 B][per pair: rows slot | cols slot | boundary H (| boundary E)]
    uint8_t *s_class = smem;                                   // 256
    int8_t *s_costs = reinterpret_cast<int8_t *>(smem + 256);  // 1024 + 1024 (transposed)
    const uint32_t bound_bytes = ((bound_cap * (uint32_t)sizeof(BoundT) * (AFFINE ? 2u : 1u)) + 15u) & ~15u;
    const uint32_t pair_bytes = slot_rows + slot_cols + bound_bytes;
    const int tid = threadIdx.x;
    const int group = tid / G, lane = tid % G;
    uint8_t *my = smem + 2304 + (size_t)group * pair_bytes;
    uint8_t *s_rows = my, *s_cols = my + slot_rows;
    BoundT *s_bh = reinterpret_cast<BoundT *>(my + slot_rows + slot_cols);
    BoundT *s_be = s_bh + bound_cap;

    // The tables are tiny (2304 B): always stage them; the class map is needed by the unary path too.
    for (int i = tid; i < 2304 / 4; i += WF_THREADS)
        reinterpret_cast<uint32_t *>(smem)[i] = reinterpret_cast<const uint32_t *>(tables)[i];

    const uint64_t local = (uint64_t)blockIdx.x * PAIRS_PER_BLOCK + group;
    const bool have = local < count;
    uint32_t m = 0, n = 0, mis_r = 0, mis_c = 0;
    bool swapped = false;
    uint64_t out_index = 0;
    if (have) {
        const PairView pv = pair_at(ps, begin + local);
        swapped = pv.la > pv.lb; // rows = shorter string: fewer strips, shorter fill/drain
        const uint8_t *rows = swapped ? pv.b : pv.a, *cols = swapped ? pv.a : pv.b;
        m = swapped ? pv.lb : pv.la;
        n = swapped ? pv.la : pv.lb;
        out_index = pv.out_index;
        mis_r = stage_string(s_rows, rows, m, lane, G);
        mis_c = stage_string(s_cols, cols, n, lane, G);
    }
    __syncthreads();
    const uint8_t *rows_s = s_rows + mis_r, *cols_s = s_cols + mis_c;
    const int8_t *cost_tab = s_costs + (swapped ? 1024 : 0);

    // "-infinity" for the gap tracks: far below any reachable score, representable in BoundT.
    const int e_floor = sizeof(BoundT) == 2 ? -32000 : NEG_BIG;
    // Top boundary row H[0][j] (and E[0][j] = -inf) for j = 0..n.
    for (uint32_t j = lane; j <= n && have; j += G) {
        int h = LOCAL ? 0 : (j == 0 ? 0 : cs.first + (int)(j - 1) * cs.next);
        s_bh[j] = (BoundT)h;
        if (AFFINE) s_be[j] = (BoundT)e_floor;
    }
    __syncwarp();

    // Warp-uniform trip counts (groups of a warp may hold pairs of different sizes).
    uint32_t n_max = n, m_max = m;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, d));
        m_max = max(m_max, __shfl_xor_sync(0xffffffffu, m_max, d));
    }
    const uint32_t strips = (m_max + G * R - 1) / (G * R);
    const uint32_t my_strips = (m + G * R - 1) / (G * R);

    int best = 0;       // LOCAL: running maximum
    int result = 0;     // GLOBAL: H[m][n]

    for (uint32_t strip = 0; strip < strips; ++strip) {
        const uint32_t row0 = strip * G * R + lane * R; // 0-based index of this lane's first row
        const bool strip_live = have && strip < my_strips;
        // Row symbols (classes) of this lane, and left boundary H[i][0].
        int rc[R];    // class of row symbol, or a sentinel no column class equals beyond the pattern
        int Hl[R];    // H[i][j-1]
        int Fl[R];    // F[i][j-1] (AFFINE)
        uint32_t valid_mask = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t i = row0 + r; // 0-based row
            const bool ok = strip_live && i < m;
            rc[r] = ok ? (int)s_class[rows_s[i]] : 0x7FFF;
            if (ok) valid_mask |= 1u << r;
            Hl[r] = LOCAL ? 0 : cs.first + (int)i * cs.next; // H[i+1][0]
            Fl[r] = e_floor;
        }
        // Cell above-left of this lane's first row at column 0: H[row0][0].
        int diag_top = (LOCAL || row0 == 0) ? 0 : cs.first + (int)(row0 - 1) * cs.next;
        int send_h = 0, send_e = e_floor;
        const bool park = strip + 1 < my_strips; // bottom row needed by a following strip
        const uint32_t steps = n_max + G - 1;
        for (uint32_t t = 0; t < steps; ++t) {
            const int recv_h = __shfl_up_sync(0xffffffffu, send_h, 1, G);
            int recv_e = e_floor;
            if (AFFINE) recv_e = __shfl_up_sync(0xffffffffu, send_e, 1, G);
            const uint32_t j = t - lane + 1; // 1-based column handled by this lane now
            if (strip_live && t >= (uint32_t)lane && j <= n) {
                const int cc = (int)s_class[cols_s[j - 1]];
                int up_h = lane == 0 ? bound_load(s_bh + j) : recv_h;
                int up_e = e_floor;
                if (AFFINE) up_e = lane == 0 ? bound_load(s_be + j) : recv_e;
                int diag = diag_top;
                diag_top = up_h;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    int s;
                    if (cs.unary) s = rc[r] == cc ? cs.match : cs.mismatch;
                    else s = (int)cost_tab[(rc[r] & 31) * 32 + cc];
                    int h;
                    if (AFFINE) {
                        const int e = max(up_e + cs.next, up_h + cs.first);
                        const int f = max(Fl[r] + cs.next, Hl[r] + cs.first);
                        h = max(max(diag + s, e), f);
                        if (LOCAL) h = max(h, 0);
                        up_e = e;
                        Fl[r] = f;
                    } else {
                        h = max(diag + s, max(Hl[r], up_h) + cs.next);
                        if (LOCAL) h = max(h, 0);
                    }
                    if (LOCAL && ((valid_mask >> r) & 1u)) best = max(best, h);
                    diag = Hl[r];
                    Hl[r] = h;
                    up_h = h;
                }
                send_h = up_h;
                send_e = up_e;
                if (park && lane == G - 1) {
                    s_bh[j] = (BoundT)up_h;
                    if (AFFINE) s_be[j] = (BoundT)max(up_e, e_floor);
                }
            }
        }
        // H[m][n] sits in the strip that holds row m-1, after that lane has consumed column n.
        if (have && !LOCAL && m > 0 && strip == (m - 1) / (G * R)) {
            const uint32_t within = (m - 1) % (G * R);
            if ((uint32_t)lane == within / R) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if ((uint32_t)r == within % R) result = Hl[r];
                ps.out[out_index] = (uint64_t)(int64_t)(cs.negate ? -result : result);
            }
        }
        __syncwarp();
    }
    if (LOCAL) {
#pragma unroll
        for (int d = G / 2; d >= 1; d >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, d, G));
        if (have && lane == 0) ps.out[out_index] = (uint64_t)(int64_t)best;
    }
    if (have && !LOCAL && lane == 0) {
        // degenerate shapes: one side empty -> pure gap
        if (m == 0) {
            int h = n == 0 ? 0 : cs.first + (int)(n - 1) * cs.next;
            ps.out[out_index] = (uint64_t)(int64_t)(cs.negate ? -h : h);
        }
    }
}

} // namespace swb
