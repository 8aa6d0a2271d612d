// align_wavefront_packed_general.cuh — the packed s16x2 (DPX) wavefront for the cases the shifted kernel does not
// cover: affine gaps (Gotoh E/F tracks) and local alignment (Smith-Waterman), unary costs, sm_100a.
//
// Same mapping as align_wavefront_packed.cuh (a group of G lanes carries TWO string pairs of the same shape in the
// 16-bit halves of every register, R rows per lane, bottom cells handed down with `__shfl_up_sync`, row strips with
// a packed boundary row in shared memory) and the same recurrence as align_wavefront.cuh, transcribed to packed
// signed arithmetic: every add is VIADDMNMX.S16x2 (max(a + b, c), with c = -32768 when only the sum is wanted —
// a plain 32-bit add would carry between the halves once values are negative), every max VIMNMX.S16x2:
//     ne = min.u16x2(rows2 ^ cols2, 1)            0 match / 1 mismatch per half
//     s  = select(ne * 0xFFFF, mismatch2, match2)  IMAD + LOP3
//     E  = max(E_up + next, H_up + first)          2 DPX          (AFFINE; gap running down a column)
//     F  = max(F_left + next, H_left + first)      2 DPX          (AFFINE; gap running along a row)
//     H  = max(H_diag + s, E, F [, 0])             2 (+1) DPX
// = 9-10 ALU-pipe instructions per TWO cells against 8-10 per cell in the scalar kernel. The host routes a batch here
// when every pair has the same shape, costs are unary with mismatch <= 0 <= match and gaps <= 0, and every reachable
// score fits int16 with room for the "-infinity" of the gap tracks (engine.cu, launch_wavefront).
//
// Reference semantics: `NeedlemanWunschScores` / `SmithWatermanScores` with `unary_class_costs` tables, linear (-2,-2)
// and affine (-5,-1) gaps (StringWars similarities/bench.rs:98-108, 655-674, 982-985).
#pragma once
#include "align_wavefront.cuh"

namespace swb {

namespace pk {
__device__ __forceinline__ uint32_t both(int v) { return ((uint32_t)v & 0xFFFFu) * 0x00010001u; } // v in both halves
__device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return __viaddmax_s16x2(a, b, 0x80008000u); }
__device__ __forceinline__ uint32_t addmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
__device__ __forceinline__ int lo(uint32_t p) { return (int)(int16_t)(p & 0xFFFFu); }
__device__ __forceinline__ int hi(uint32_t p) { return (int)(int16_t)(p >> 16); }
} // namespace pk

template <int G, int R, bool LOCAL, bool AFFINE>
// four resident blocks (<= 128 registers) for the local variants: measured +10 % (affine) / +1.5 % (linear); the global
// affine variant would spill at 128 registers (measured -5 %)
__global__ void __launch_bounds__(WF_THREADS, (LOCAL ? 4 : 3))
    align_wavefront_packed_general_kernel(PairSpace ps, uint64_t begin, uint64_t count, AlignCosts cs, const AlignTables *__restrict__ tables,
                                          uint32_t slot_rows, uint32_t slot_cols, uint32_t bound_cap) {
    constexpr int GROUPS_PER_BLOCK = WF_THREADS / G; // each group owns TWO pairs
    extern __shared__ __align__(16) uint8_t smem[];
    // layout: [tables 2304 B][per group: rows A | rows B | cols A | cols B | packed boundary H (| packed boundary E)]
    uint8_t *s_class = smem;
    const uint32_t bound_bytes = (bound_cap * 4u + 15u) & ~15u;
    const uint32_t group_bytes = 2u * slot_rows + 2u * slot_cols + bound_bytes * (AFFINE ? 2u : 1u);
    const int tid = threadIdx.x, group = tid / G, lane = tid % G;
    uint8_t *my = smem + 2304 + (size_t)group * group_bytes;
    uint8_t *s_rows[2] = {my, my + slot_rows};
    uint8_t *s_cols[2] = {my + 2 * slot_rows, my + 2 * slot_rows + slot_cols};
    uint32_t *s_bh = reinterpret_cast<uint32_t *>(my + 2 * slot_rows + 2 * slot_cols);
    uint32_t *s_be = reinterpret_cast<uint32_t *>(my + 2 * slot_rows + 2 * slot_cols + bound_bytes);

    const bool need_tables = !cs.ident;
    if (need_tables)
        for (int i = tid; i < 2304 / 4; i += WF_THREADS)
            reinterpret_cast<uint32_t *>(smem)[i] = reinterpret_cast<const uint32_t *>(tables)[i];

    const uint64_t first = ((uint64_t)blockIdx.x * GROUPS_PER_BLOCK + group) * 2u;
    const bool have = first < count;
    const bool have_b = first + 1 < count; // an odd batch: the last group computes its only pair in both halves
    uint32_t m = 0, n = 0;
    uint32_t mis_r[2] = {0, 0}, mis_c[2] = {0, 0};
    uint64_t out_index[2] = {0, 0};
    if (have) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const PairView pv = pair_at(ps, begin + first + ((half && have_b) ? 1u : 0u));
            const bool swapped = pv.la > pv.lb; // rows = shorter string (unary costs are symmetric)
            m = swapped ? pv.lb : pv.la, n = swapped ? pv.la : pv.lb; // identical for both pairs (host-checked)
            out_index[half] = pv.out_index;
            mis_r[half] = stage_string(s_rows[half], swapped ? pv.b : pv.a, m, lane, G);
            mis_c[half] = stage_string(s_cols[half], swapped ? pv.a : pv.b, n, lane, G);
        }
    }
    __syncthreads();
    uint8_t *rows_s[2] = {s_rows[0] + mis_r[0], s_rows[1] + mis_r[1]};
    uint8_t *cols_s[2] = {s_cols[0] + mis_c[0], s_cols[1] + mis_c[1]};
    if (need_tables) { // symbols -> classes, once, in place
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            for (uint32_t i = lane; i < m; i += G) rows_s[half][i] = s_class[rows_s[half][i]];
            for (uint32_t j = lane; j < n; j += G) cols_s[half][j] = s_class[cols_s[half][j]];
        }
    }

    uint32_t n_max = n, m_max = m;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, d));
        m_max = max(m_max, __shfl_xor_sync(0xffffffffu, m_max, d));
    }
    const uint32_t strips = (m_max + G * R - 1) / (G * R);
    const uint32_t my_strips = (m + G * R - 1) / (G * R);

    constexpr int E_FLOOR = -32000; // "-infinity" of the gap tracks: below every reachable score, above int16's wrap
    const uint32_t floor2 = pk::both(E_FLOOR);
    const uint32_t first2 = pk::both(cs.first), next2 = pk::both(cs.next);
    const uint32_t match2 = pk::both(cs.match), mismatch2 = pk::both(cs.mismatch);
    for (uint32_t j = lane; j <= n && have; j += G) {
        s_bh[j] = pk::both(LOCAL ? 0 : (j == 0 ? 0 : cs.first + (int)(j - 1) * cs.next)); // H[0][j]
        if (AFFINE) s_be[j] = floor2;
    }
    __syncwarp();

    uint32_t best = 0u; // LOCAL: packed running maxima (>= 0)

    for (uint32_t strip = 0; strip < strips; ++strip) {
        const uint32_t row0 = strip * G * R + lane * R;
        const bool strip_live = have && strip < my_strips;
        const uint32_t n_live = strip_live ? n : 0u;
        uint32_t rc[R];        // packed row symbols: A in the low half, B in the high half; 0xFFFF = beyond the pattern
        uint32_t HA[R], HB[R];  // packed H[i][j-1], ping-pong
        uint32_t Fl[R];         // packed F[i][j-1] (AFFINE)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t i = row0 + r;
            const bool ok = strip_live && i < m;
            rc[r] = ok ? ((uint32_t)rows_s[0][i] | ((uint32_t)rows_s[1][i] << 16)) : 0xFFFFFFFFu;
            HA[r] = HB[r] = pk::both(LOCAL ? 0 : cs.first + (int)i * cs.next); // H[i+1][0]
            Fl[r] = floor2;
        }
        uint32_t diag_top = pk::both((LOCAL || row0 == 0) ? 0 : cs.first + (int)(row0 - 1) * cs.next); // H[row0][0]
        uint32_t send_e = floor2;
        const bool park = strip + 1 < my_strips;

        auto cells = [&](const uint32_t(&Hold)[R], uint32_t(&Hnew)[R], uint32_t j0, uint32_t recv_h, uint32_t recv_e) {
            const uint32_t cc = (uint32_t)cols_s[0][j0] | ((uint32_t)cols_s[1][j0] << 16);
            uint32_t up_h = recv_h, up_e = recv_e;
            if (lane == 0) {
                up_h = s_bh[j0 + 1];
                if (AFFINE) up_e = s_be[j0 + 1];
            }
            uint32_t diag = diag_top;
            diag_top = up_h;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t ne = __vminu2(rc[r] ^ cc, 0x00010001u);
                const uint32_t mask = ne * 0xFFFFu; // 0x0000 / 0xFFFF per half (no carry: 1 * 0xFFFF fits the half)
                const uint32_t s = (mask & mismatch2) | (~mask & match2);
                uint32_t h;
                if (AFFINE) {
                    const uint32_t e = pk::addmax(up_e, next2, pk::add(up_h, first2));
                    const uint32_t f = pk::addmax(Fl[r], next2, pk::add(Hold[r], first2));
                    h = pk::max2(pk::addmax(diag, s, e), f);
                    up_e = e;
                    Fl[r] = f;
                } else {
                    h = pk::max2(pk::addmax(diag, s, pk::add(Hold[r], next2)), pk::add(up_h, next2));
                }
                if (LOCAL) {
                    h = pk::max2(h, 0u);
                    // rows beyond the pattern hold cells derived from real ones by non-positive increments (the host
                    // checks mismatch <= 0 and gaps <= 0), so they never exceed the true maximum: no row mask
                    best = pk::max2(best, h);
                }
                diag = Hold[r];
                Hnew[r] = h;
                up_h = h;
            }
            send_e = up_e;
            if (park && lane == G - 1) {
                s_bh[j0 + 1] = up_h;
                if (AFFINE) s_be[j0 + 1] = pk::max2(up_e, floor2);
            }
        };
        // The boundary row is updated in place (lane 0 reads column j long before lane G-1 parks it); the shuffle
        // orders the two in execution, `__syncwarp()` makes it a memory ordering too.
        auto step_guarded = [&](const uint32_t(&Hold)[R], uint32_t(&Hnew)[R], uint32_t t) {
            __syncwarp();
            const uint32_t recv_h = __shfl_up_sync(0xffffffffu, Hold[R - 1], 1, G);
            uint32_t recv_e = floor2;
            if (AFFINE) recv_e = __shfl_up_sync(0xffffffffu, send_e, 1, G);
            const uint32_t j0 = t - (uint32_t)lane;
            if (j0 < n_live) cells(Hold, Hnew, j0, recv_h, recv_e);
        };
        auto step_steady = [&](const uint32_t(&Hold)[R], uint32_t(&Hnew)[R], uint32_t t) {
            __syncwarp();
            const uint32_t recv_h = __shfl_up_sync(0xffffffffu, Hold[R - 1], 1, G);
            uint32_t recv_e = floor2;
            if (AFFINE) recv_e = __shfl_up_sync(0xffffffffu, send_e, 1, G);
            cells(Hold, Hnew, t - (uint32_t)lane, recv_h, recv_e);
        };
        const uint32_t steps = (n_max + G - 1 + 1u) & ~1u;
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
                uint32_t packed = 0;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if ((uint32_t)r == within % R) packed = in_b ? HB[r] : HA[r];
                const int res_a = pk::lo(packed), res_b = pk::hi(packed);
                store_result(ps, out_index[0], cs.negate ? -res_a : res_a);
                if (have_b) store_result(ps, out_index[1], cs.negate ? -res_b : res_b);
            }
        }
        __syncwarp();
    }
    if (LOCAL) {
#pragma unroll
        for (int d = G / 2; d >= 1; d >>= 1) best = pk::max2(best, __shfl_xor_sync(0xffffffffu, best, d, G));
        if (have && lane == 0) {
            store_result(ps, out_index[0], pk::lo(best));
            if (have_b) store_result(ps, out_index[1], pk::hi(best));
        }
    }
    if (have && !LOCAL && lane == 0 && m == 0) { // one side empty -> pure gap
        const int h = n == 0 ? 0 : cs.first + (int)(n - 1) * cs.next;
        store_result(ps, out_index[0], cs.negate ? -h : h);
        if (have_b) store_result(ps, out_index[1], cs.negate ? -h : h);
    }
}

} // namespace swb
