// align_wavefront.cuh — anti-diagonal wavefront DP for Levenshtein / Needleman-Wunsch / Smith-Waterman, sm_100a.
//
// This is the kernel BASELINE.json's north_star names: a sub-warp GROUP of G lanes owns one string pair; both
// strings are staged into shared memory with coalesced 16-byte read-only loads; lane l keeps R consecutive DP
// rows in registers and walks the columns skewed by l, so that at step t the group is working on one
// anti-diagonal band of the matrix. The only inter-lane traffic is the bottom cell of each lane's strip,
// handed to lane l+1 with one `__shfl_up_sync` per step; the three-way max itself maps to the DPX
// VIMNMX / VIADDMNMX integer instructions. Integer DP, no tensor cores.
//
// One recurrence covers every engine of similarities/bench.rs (maximising form):
//     H[i][j] = max(H[i-1][j-1] + S(a_i, b_j),  E[i][j],  F[i][j])          (clamped at 0 when LOCAL)
//     E[i][j] = max(E[i-1][j] + next, H[i-1][j] + first)    gap running down a column
//     F[i][j] = max(F[i][j-1] + next, H[i][j-1] + first)    gap running along a row
// Linear gaps (first == next) collapse E/F into H+gap. Levenshtein is S = -(a != b), gap = -1, distance = -H.
//
// Issue-slot economy (the kernel is issue-bound, profiles/r1_*): for the global, linear-gap, unary-cost case
// (Levenshtein, NW "linear") cells are kept as H'[i][j] = H[i][j] - gap*(i+j). Then
//     H'[i][j] = max(H'[i-1][j-1] + (S - 2*gap), H'[i-1][j], H'[i][j-1]),   H'[i][0] = H'[0][j] = 0,
// i.e. one compare, one select, one VIMNMX and one VIADDMNMX per cell and no boundary arithmetic; the row
// registers are ping-ponged between two arrays (step loop unrolled by two) so no register copies are issued.
// Patterns longer than G*R rows are processed in row strips; the strip's bottom row is parked in a per-pair
// shared-memory boundary row (int16 when scores fit, else int32) and read back as the next strip's top row.
//
// Reference semantics: `LevenshteinDistances` (bench.rs:382-393), `NeedlemanWunschScores` /
// `SmithWatermanScores` with `unary_class_costs` tables (bench.rs:98-108, 655-674, 982-985).
#pragma once
#include "align_costs.h"
#include "common.cuh"

namespace swb {

constexpr int WF_THREADS = 128;
constexpr int NEG_BIG = -(1 << 28);


// G lanes per pair, R rows per lane. LOCAL: Smith-Waterman. AFFINE: Gotoh E/F tracks. UNARY: match/mismatch
// constants instead of the 32x32 table. SymT: uint8_t symbols, or uint32_t for decoded UTF-8 (codepoint-level
// Levenshtein, bench.rs:437; compared raw, `cs.ident` must be set).
// LONG: pairs too long for shared memory (the reference runs NW on multi-kilobyte XLSum lines,
// similarities/README.md:19-23). Nothing is staged: a lane reads its R row symbols from global memory once per strip and one
// column symbol per step (32 consecutive bytes per warp and step, L1-resident), and the boundary rows between strips live
// in a global scratch row per pair (`bound_g`, `bound_stride` bytes apart) — the same recurrence with no length limit.
template <int G, int R, bool LOCAL, bool AFFINE, bool UNARY, typename BoundT, typename SymT = uint8_t, bool LONG = false>
__global__ void __launch_bounds__(WF_THREADS)
    align_wavefront_kernel(PairSpace ps, uint64_t begin, uint64_t count, AlignCosts cs, const AlignTables *__restrict__ tables,
                           uint32_t slot_rows /* bytes reserved for the row string */, uint32_t slot_cols, uint32_t bound_cap,
                           uint8_t *bound_g, uint64_t bound_stride) {
    constexpr int PAIRS_PER_BLOCK = WF_THREADS / G;
    constexpr bool SHIFTED = !LOCAL && !AFFINE && UNARY; // cells hold H - gap*(i+j)
    extern __shared__ __align__(16) uint8_t smem[];
    // layout: [tables 2304 B][per pair: rows slot | cols slot | boundary H (| boundary E)]
    uint8_t *s_class = smem;                                  // 256
    int8_t *s_costs = reinterpret_cast<int8_t *>(smem + 256); // 1024 + 1024 (transposed)
    const uint32_t bound_bytes = ((bound_cap * (uint32_t)sizeof(BoundT) * (AFFINE ? 2u : 1u)) + 15u) & ~15u;
    const uint32_t pair_bytes = slot_rows + slot_cols + bound_bytes;
    const int tid = threadIdx.x;
    const int group = tid / G, lane = tid % G;
    uint8_t *my = smem + 2304 + (size_t)group * pair_bytes;
    uint8_t *s_rows = my, *s_cols = my + slot_rows;
    BoundT *s_bh = reinterpret_cast<BoundT *>(my + slot_rows + slot_cols);
    BoundT *s_be = s_bh + bound_cap;
    const SymT *rows_g = nullptr, *cols_g = nullptr;

    const bool need_tables = !cs.ident || !UNARY;
    if (need_tables)
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
        if (LONG) {
            rows_g = reinterpret_cast<const SymT *>(rows), cols_g = reinterpret_cast<const SymT *>(cols);
            s_bh = reinterpret_cast<BoundT *>(bound_g + local * bound_stride);
            s_be = s_bh + (n + 1u);
        } else {
            mis_r = stage_string(s_rows, rows, m * (uint32_t)sizeof(SymT), lane, G);
            mis_c = stage_string(s_cols, cols, n * (uint32_t)sizeof(SymT), lane, G);
        }
    }
    __syncthreads();
    SymT *rows_s = reinterpret_cast<SymT *>(s_rows + mis_r), *cols_s = reinterpret_cast<SymT *>(s_cols + mis_c);
    const int8_t *cost_tab = s_costs + (swapped ? 1024 : 0);
    const bool map_classes = need_tables && sizeof(SymT) == 1;
    // Symbols -> classes, once, in place (the class map folds bytes into <= 32 classes, bench.rs:99-102).
    if (map_classes && !LONG) {
        for (uint32_t i = lane; i < m; i += G) rows_s[i] = (SymT)s_class[(uint8_t)rows_s[i]];
        for (uint32_t j = lane; j < n; j += G) cols_s[j] = (SymT)s_class[(uint8_t)cols_s[j]];
    }
    // one symbol of the row / column string as the cells compare it (LONG: straight from global memory, mapped on the way)
    auto row_symbol = [&](uint32_t i) -> int {
        if (!LONG) return (int)rows_s[i];
        const SymT v = __ldg(rows_g + i);
        return map_classes ? (int)s_class[(uint8_t)v] : (int)v;
    };
    auto col_symbol = [&](uint32_t j) -> int {
        if (!LONG) return (int)cols_s[j];
        const SymT v = __ldg(cols_g + j);
        return map_classes ? (int)s_class[(uint8_t)v] : (int)v;
    };

    // Warp-uniform trip counts (groups of a warp may hold pairs of different sizes).
    uint32_t n_max = n, m_max = m;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, d));
        m_max = max(m_max, __shfl_xor_sync(0xffffffffu, m_max, d));
    }
    const uint32_t strips = (m_max + G * R - 1) / (G * R);
    const uint32_t my_strips = (m + G * R - 1) / (G * R);

    // "-infinity" for the gap tracks: far below any reachable score, representable in BoundT.
    const int e_floor = sizeof(BoundT) == 2 ? -32000 : NEG_BIG;
    const int gap = cs.next;
    const int c_match = SHIFTED ? cs.match - 2 * gap : cs.match;
    const int c_mismatch = SHIFTED ? cs.mismatch - 2 * gap : cs.mismatch;
    // The boundary row lives in shared memory only when it is not identically computable: shifted cells have
    // an all-zero top row, so a single strip never touches it.
    const bool use_bound = !SHIFTED || strips > 1 || LONG;
    if (use_bound) {
        for (uint32_t j = lane; j <= n && have; j += G) {
            const int h = (LOCAL || SHIFTED) ? 0 : (j == 0 ? 0 : cs.first + (int)(j - 1) * cs.next);
            s_bh[j] = (BoundT)h;
            if (AFFINE) s_be[j] = (BoundT)e_floor;
        }
    }
    __syncwarp();

    int best = 0; // LOCAL: running maximum

    for (uint32_t strip = 0; strip < strips; ++strip) {
        const uint32_t row0 = strip * G * R + lane * R; // 0-based index of this lane's first row
        const bool strip_live = have && strip < my_strips;
        const uint32_t n_live = strip_live ? n : 0u;
        int rc[R];        // class of each row symbol, or a sentinel no column class equals beyond the pattern
        int HA[R], HB[R]; // H[i][j-1], ping-pong
        int Fl[R];        // F[i][j-1] (AFFINE)
        uint32_t valid_mask = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t i = row0 + r; // 0-based row
            const bool ok = strip_live && i < m;
            rc[r] = ok ? row_symbol(i) : -1; // -1 is no byte, no class and no packed UTF-8 symbol
            if (ok) valid_mask |= 1u << r;
            HA[r] = HB[r] = (LOCAL || SHIFTED) ? 0 : cs.first + (int)i * cs.next; // H[i+1][0]
            Fl[r] = e_floor;
        }
        // Cell above-left of this lane's first row at column 0: H[row0][0].
        int diag_top = (LOCAL || SHIFTED || row0 == 0) ? 0 : cs.first + (int)(row0 - 1) * cs.next;
        int send_e = e_floor;
        const bool park = strip + 1 < my_strips; // bottom row needed by a following strip

        // One skewed step: reads Hold (column j-1), writes Hnew (column j). GUARDED=false is the steady state in
        // which every lane of the warp is inside its matrix: no predicate, no branch, no register copies.
        auto cells = [&](const int(&Hold)[R], int(&Hnew)[R], uint32_t j0, int recv_h, int recv_e) {
            const int cc = col_symbol(j0);
            int up_h = recv_h, up_e = recv_e;
            if (lane == 0) {
                up_h = use_bound ? (int)s_bh[j0 + 1] : 0;
                if (AFFINE) up_e = (int)s_be[j0 + 1];
            }
            int diag = diag_top;
            diag_top = up_h;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int s;
                if (UNARY) s = rc[r] == cc ? c_match : c_mismatch;
                else s = (int)cost_tab[(rc[r] & 31) * 32 + cc];
                int h;
                if (AFFINE) {
                    const int e = max(up_e + cs.next, up_h + cs.first);
                    const int f = max(Fl[r] + cs.next, Hold[r] + cs.first);
                    h = max(max(diag + s, e), f);
                    if (LOCAL) h = max(h, 0);
                    up_e = e;
                    Fl[r] = f;
                } else if (SHIFTED) {
                    // (diag+s, left) first: that part is independent of the row above, so the serial chain
                    // down the lane's strip is ONE max per cell
                    h = max(max(diag + s, Hold[r]), up_h);
                } else {
                    h = max(max(diag + s, Hold[r] + gap), up_h + gap);
                    if (LOCAL) h = max(h, 0);
                }
                if (LOCAL && ((valid_mask >> r) & 1u)) best = max(best, h);
                diag = Hold[r];
                Hnew[r] = h;
                up_h = h;
            }
            send_e = up_e;
            if (park && lane == G - 1) {
                s_bh[j0 + 1] = (BoundT)up_h;
                if (AFFINE) s_be[j0 + 1] = (BoundT)max(up_e, e_floor);
            }
        };
        // The shared-memory boundary row is updated in place: lane 0 reads column j long before lane G-1 parks the
        // new value of column j. The per-step shuffle orders the two in execution; `__syncwarp()` makes that an
        // explicit memory ordering too (shuffles are not memory barriers). Only on the paths that use the row.
        auto step_guarded = [&](const int(&Hold)[R], int(&Hnew)[R], uint32_t t) {
            if (use_bound) __syncwarp();
            const int recv_h = __shfl_up_sync(0xffffffffu, Hold[R - 1], 1, G);
            int recv_e = e_floor;
            if (AFFINE) recv_e = __shfl_up_sync(0xffffffffu, send_e, 1, G);
            const uint32_t j0 = t - (uint32_t)lane; // 0-based column handled by this lane now
            if (j0 < n_live) cells(Hold, Hnew, j0, recv_h, recv_e);
        };
        auto step_steady = [&](const int(&Hold)[R], int(&Hnew)[R], uint32_t t) {
            if (use_bound) __syncwarp();
            const int recv_h = __shfl_up_sync(0xffffffffu, Hold[R - 1], 1, G);
            int recv_e = e_floor;
            if (AFFINE) recv_e = __shfl_up_sync(0xffffffffu, send_e, 1, G);
            cells(Hold, Hnew, t - (uint32_t)lane, recv_h, recv_e);
        };

        // steps [0, t_lo): fill; [t_lo, t_hi): every lane of the warp active; [t_hi, steps): drain. All even.
        const uint32_t steps = (n_max + G - 1 + 1u) & ~1u; // surplus steps are inactive for every lane
        uint32_t n_min = n_live;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) n_min = min(n_min, __shfl_xor_sync(0xffffffffu, n_min, d));
        const uint32_t t_lo = min(steps, (uint32_t)((G - 1 + 1) & ~1));
        const uint32_t t_hi = max(t_lo, min(steps, n_min & ~1u));
        for (uint32_t t = 0; t < t_lo; t += 2) {
            step_guarded(HA, HB, t);
            step_guarded(HB, HA, t + 1);
        }
        for (uint32_t t = t_lo; t < t_hi; t += 2) {
            step_steady(HA, HB, t);
            step_steady(HB, HA, t + 1);
        }
        for (uint32_t t = t_hi; t < steps; t += 2) {
            step_guarded(HA, HB, t);
            step_guarded(HB, HA, t + 1);
        }
        // H[m][n] sits in the strip that holds row m-1; the lane's last active step t = lane+n-1 wrote HB when t is even.
        if (have && !LOCAL && m > 0 && strip == (m - 1) / (G * R)) {
            const uint32_t within = (m - 1) % (G * R);
            if ((uint32_t)lane == within / R) {
                const bool in_b = n > 0 && (((uint32_t)lane + n - 1u) & 1u) == 0u;
                int result = 0;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if ((uint32_t)r == within % R) result = in_b ? HB[r] : HA[r];
                if (SHIFTED) result += gap * (int)(m + n);
                store_result(ps, out_index, cs.negate ? -result : result);
            }
        }
        __syncwarp();
    }
    if (LOCAL) {
#pragma unroll
        for (int d = G / 2; d >= 1; d >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, d, G));
        if (have && lane == 0) store_result(ps, out_index, best);
    }
    if (have && !LOCAL && lane == 0 && m == 0) { // one side empty -> pure gap
        const int h = n == 0 ? 0 : cs.first + (int)(n - 1) * cs.next;
        store_result(ps, out_index, cs.negate ? -h : h);
    }
}

} // namespace swb
