// align_wavefront_packed.cuh — the anti-diagonal wavefront with TWO pairs per lane group, packed s16x2 (DPX), sm_100a.
//
// The scalar wavefront kernel (align_wavefront.cuh) is ALU-pipe bound at 4 integer instructions per cell
// (compare, select, VIMNMX, VIADDMNMX; profiles/r1_c_wavefront_*). Blackwell keeps Hopper's DPX instructions,
// including the packed forms VIMNMX.S16x2 / VIADDMNMX.S16x2 / VIMNMX.U16x2 (verified in SASS), so this kernel
// carries two independent string pairs in the low and high 16-bit halves of every register: one group of G lanes
// advances both DP matrices with the same shuffles and the same instruction stream — 4 ALU + 1 FMA-pipe
// instruction per TWO cells:
//     x  = rows2 ^ cols2                      LOP3          (per half: 0 iff the symbols match)
//     ne = min.u16x2(x, 0x00010001)           VIMNMX.U16x2  (per half: 0 match / 1 mismatch)
//     s  = c_match2 - ne * (c_match - c_mismatch)   IMAD    (FMA pipe; halves stay >= 0, no borrow)
//     h  = max.s16x2(max(diag + s, left), up)  VIADDMNMX.S16x2 + VIMNMX.S16x2
// It applies to the global, linear-gap, unary-cost case (Levenshtein in wavefront mode, NW "linear"), where cells
// are kept as H - gap*(i+j) >= 0 (see align_wavefront.cuh), on batches whose pairs all have the same shape so the
// two halves share every loop bound; the host checks that scores fit 15 bits and routes anything else to the
// scalar kernel. The boundary row between row strips is a packed 2 x int16 word per column — BASELINE.json
// config 5's "int16 score tile in smem".
#pragma once
#include "align_wavefront.cuh"

namespace swb {

template <int G, int R>
__global__ void __launch_bounds__(WF_THREADS)
    align_wavefront_packed_kernel(PairSpace ps, uint64_t begin, uint64_t count, AlignCosts cs, const AlignTables *__restrict__ tables,
                                  uint32_t slot_rows, uint32_t slot_cols, uint32_t bound_cap) {
    constexpr int GROUPS_PER_BLOCK = WF_THREADS / G; // each group owns TWO pairs
    extern __shared__ __align__(16) uint8_t smem[];
    // layout: [tables 2304 B][per group: rows A | rows B | cols A | cols B | packed boundary row]
    uint8_t *s_class = smem;
    const uint32_t bound_bytes = (bound_cap * 4u + 15u) & ~15u;
    const uint32_t group_bytes = 2u * slot_rows + 2u * slot_cols + bound_bytes;
    const int tid = threadIdx.x, group = tid / G, lane = tid % G;
    uint8_t *my = smem + 2304 + (size_t)group * group_bytes;
    uint8_t *s_rows[2] = {my, my + slot_rows};
    uint8_t *s_cols[2] = {my + 2 * slot_rows, my + 2 * slot_rows + slot_cols};
    uint32_t *s_bh = reinterpret_cast<uint32_t *>(my + 2 * slot_rows + 2 * slot_cols);

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
            const bool swapped = pv.la > pv.lb; // rows = shorter string (costs are symmetric in the unary case)
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
    const int gap = cs.next;
    const uint32_t c_match = (uint32_t)(cs.match - 2 * gap), c_mismatch = (uint32_t)(cs.mismatch - 2 * gap); // shifted costs, >= 0
    const uint32_t c_match2 = c_match * 0x00010001u, c_delta = c_match - c_mismatch;
    const bool use_bound = strips > 1;
    if (use_bound)
        for (uint32_t j = lane; j <= n && have; j += G) s_bh[j] = 0u;
    __syncwarp();

    for (uint32_t strip = 0; strip < strips; ++strip) {
        const uint32_t row0 = strip * G * R + lane * R;
        const bool strip_live = have && strip < my_strips;
        const uint32_t n_live = strip_live ? n : 0u;
        uint32_t rc[R];         // packed row symbols: A in the low half, B in the high half; 0xFFFF = beyond the pattern
        uint32_t HA[R], HB[R];  // packed H[i][j-1], ping-pong
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t i = row0 + r;
            const bool ok = strip_live && i < m;
            rc[r] = ok ? ((uint32_t)rows_s[0][i] | ((uint32_t)rows_s[1][i] << 16)) : 0xFFFFFFFFu;
            HA[r] = HB[r] = 0u;
        }
        uint32_t diag_top = 0u;
        const bool park = strip + 1 < my_strips;

        auto cells = [&](const uint32_t(&Hold)[R], uint32_t(&Hnew)[R], uint32_t j0, uint32_t recv) {
            const uint32_t cc = (uint32_t)cols_s[0][j0] + ((uint32_t)cols_s[1][j0] << 16);
            uint32_t up = recv;
            if (lane == 0) up = use_bound ? s_bh[j0 + 1] : 0u;
            uint32_t diag = diag_top;
            diag_top = up;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t ne = __vminu2(rc[r] ^ cc, 0x00010001u);
                const uint32_t s = c_match2 - ne * c_delta;
                const uint32_t h = __vmaxs2(__viaddmax_s16x2(diag, s, Hold[r]), up);
                diag = Hold[r];
                Hnew[r] = h;
                up = h;
            }
            if (park && lane == G - 1) s_bh[j0 + 1] = up;
        };
        auto step_guarded = [&](const uint32_t(&Hold)[R], uint32_t(&Hnew)[R], uint32_t t) {
            if (use_bound) __syncwarp();
            const uint32_t recv = __shfl_up_sync(0xffffffffu, Hold[R - 1], 1, G);
            const uint32_t j0 = t - (uint32_t)lane;
            if (j0 < n_live) cells(Hold, Hnew, j0, recv);
        };
        auto step_steady = [&](const uint32_t(&Hold)[R], uint32_t(&Hnew)[R], uint32_t t) {
            if (use_bound) __syncwarp();
            const uint32_t recv = __shfl_up_sync(0xffffffffu, Hold[R - 1], 1, G);
            cells(Hold, Hnew, t - (uint32_t)lane, recv);
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
        if (have && m > 0 && strip == (m - 1) / (G * R)) {
            const uint32_t within = (m - 1) % (G * R);
            if ((uint32_t)lane == within / R) {
                const bool in_b = n > 0 && (((uint32_t)lane + n - 1u) & 1u) == 0u;
                uint32_t packed = 0;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if ((uint32_t)r == within % R) packed = in_b ? HB[r] : HA[r];
                const int shift = gap * (int)(m + n);
                const int res_a = (int)(packed & 0xFFFFu) + shift, res_b = (int)(packed >> 16) + shift;
                store_result(ps, out_index[0], cs.negate ? -res_a : res_a);
                if (have_b) store_result(ps, out_index[1], cs.negate ? -res_b : res_b);
            }
        }
        __syncwarp();
    }
    if (have && lane == 0 && m == 0) { // one side empty -> pure gap
        const int h = n == 0 ? 0 : cs.first + (int)(n - 1) * cs.next;
        store_result(ps, out_index[0], cs.negate ? -h : h);
        if (have_b) store_result(ps, out_index[1], cs.negate ? -h : h);
    }
}

} // namespace swb
