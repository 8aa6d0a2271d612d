// Micro-benchmark: the K=2 (64-row) bit-vector column step in registers only (no table lookups), in several
// instruction selections, to see which one the SM actually runs fastest at the kernel's occupancy (6 warps/SMSP).
// Prints cycles per column per warp-slot and columns/cycle/SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int LUT> __device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT)); return r; }
__device__ __forceinline__ uint32_t mad_lo(uint32_t x, uint32_t mul, uint32_t add) {
    uint32_t r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(mul), "r"(add)); return r; }
__device__ __forceinline__ void mul_wide(uint32_t x, uint32_t mul, uint32_t &lo, uint32_t &hi) {
    uint64_t r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x), "r"(mul)); lo = (uint32_t)r; hi = (uint32_t)(r >> 32); }
__device__ __forceinline__ uint32_t funnel1(uint32_t lo, uint32_t hi) { // (hi:lo << 1) >> 32
    uint32_t r; asm("shf.l.wrap.b32 %0, %1, %2, 1;" : "=r"(r) : "r"(lo), "r"(hi)); return r; }

template <int V>
__device__ __forceinline__ void step(uint32_t h0, uint32_t l0, uint32_t h1, uint32_t l1, uint32_t &Pv0, uint32_t &Pv1, uint32_t &Mv0,
                                     uint32_t &Mv1, uint32_t one, uint32_t two) {
    if (V == 0) { // Myers blocks, Xv/Xh, shifts as written in C (previous kernel)
        uint32_t hi[2] = {h0, h1}, lo[2] = {l0, l1}, Pv[2] = {Pv0, Pv1}, Mv[2] = {Mv0, Mv1};
        uint32_t hp_in = 1u, hn_in = 0u;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const uint32_t eqc = w == 0 ? (hi[w] & lo[w]) : ((hi[w] & lo[w]) | hn_in);
            const uint32_t xv = (hi[w] & lo[w]) | Mv[w];
            const uint32_t xh = (((eqc & Pv[w]) + Pv[w]) ^ Pv[w]) | eqc;
            const uint32_t ph = Mv[w] | ~(xh | Pv[w]);
            const uint32_t mh = Pv[w] & xh;
            const uint32_t phs = ph * 2u + hp_in;
            const uint32_t mhs = mh * 2u + hn_in;
            Pv[w] = mhs | ~(xv | phs);
            Mv[w] = phs & xv;
            if (w + 1 < 2) hp_in = ph >> 31, hn_in = mh >> 31;
        }
        Pv0 = Pv[0]; Pv1 = Pv[1]; Mv0 = Mv[0]; Mv1 = Mv[1];
        return;
    }
    const uint32_t t0 = lop3<0x80>(h0, l0, Pv0), x0 = lop3<0xEA>(h0, l0, Mv0);
    const uint32_t t1 = lop3<0x80>(h1, l1, Pv1), x1 = lop3<0xEA>(h1, l1, Mv1);
    uint32_t s0, s1;
    asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(s0), "=r"(s1) : "r"(t0), "r"(t1), "r"(Pv0), "r"(Pv1));
    const uint32_t d0 = lop3<0xBE>(s0, Pv0, x0), d1 = lop3<0xBE>(s1, Pv1, x1);
    const uint32_t ph0 = lop3<0xF1>(Mv0, d0, Pv0), ph1 = lop3<0xF1>(Mv1, d1, Pv1);
    const uint32_t mh0 = Pv0 & d0, mh1 = Pv1 & d1;
    uint32_t phs0, phs1, mhs0, mhs1, c, unused;
    if (V == 1) { // IMAD.WIDE carries
        phs0 = mad_lo(ph0, two, one); mul_wide(ph0, two, unused, c); phs1 = mad_lo(ph1, two, c);
        mul_wide(mh0, two, mhs0, c); mhs1 = mad_lo(mh1, two, c);
    } else if (V == 2) { // funnel shifts for the upper word
        phs0 = mad_lo(ph0, two, one); phs1 = funnel1(ph0, ph1);
        mhs0 = mad_lo(mh0, two, 0u); mhs1 = funnel1(mh0, mh1);
    } else if (V == 3) { // carry = x >> 31 (SHF), shifted = IMAD
        phs0 = mad_lo(ph0, two, one); phs1 = mad_lo(ph1, two, ph0 >> 31);
        mhs0 = mad_lo(mh0, two, 0u); mhs1 = mad_lo(mh1, two, mh0 >> 31);
    } else if (V == 5) { // carry = mul.hi(x, 2) (IMAD.HI), shifted = IMAD
        uint32_t c0, c1;
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(c0) : "r"(ph0), "r"(two));
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(c1) : "r"(mh0), "r"(two));
        phs0 = mad_lo(ph0, two, one); phs1 = mad_lo(ph1, two, c0);
        mhs0 = mad_lo(mh0, two, 0u); mhs1 = mad_lo(mh1, two, c1);
    } else { // V == 4: everything left to the compiler (C shifts)
        phs0 = (ph0 << 1) | 1u; phs1 = (ph1 << 1) | (ph0 >> 31);
        mhs0 = mh0 << 1; mhs1 = (mh1 << 1) | (mh0 >> 31);
    }
    Pv0 = lop3<0xF1>(mhs0, d0, phs0); Pv1 = lop3<0xF1>(mhs1, d1, phs1);
    Mv0 = phs0 & d0; Mv1 = phs1 & d1;
}

template <int V>
__global__ void bench(uint32_t *out, uint32_t seed, uint32_t one, uint32_t two, long long *cycles, int cols) {
    uint32_t Pv0 = ~0u, Pv1 = ~0u, Mv0 = 0, Mv1 = 0;
    uint32_t e[8];
    for (int i = 0; i < 8; ++i) e[i] = (threadIdx.x + seed) * (2654435761u + 977u * i);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < cols; j += 4) {
        step<V>(e[0], e[1], e[2], e[3], Pv0, Pv1, Mv0, Mv1, one, two);
        step<V>(e[4], e[5], e[6], e[7], Pv0, Pv1, Mv0, Mv1, one, two);
        step<V>(e[1], e[2], e[3], e[4], Pv0, Pv1, Mv0, Mv1, one, two);
        step<V>(e[5], e[6], e[7], e[0], Pv0, Pv1, Mv0, Mv1, one, two);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = Pv0 ^ Pv1 ^ Mv0 ^ Mv1;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}


// ---- the same step fed by real nibble-table lookups from shared memory (the kernel's column loop without its pair setup)
template <int V>
__global__ void bench_lookup(uint32_t *out, uint32_t seed, uint32_t one, uint32_t two, long long *cycles, int cols, uint32_t opaque0) {
    extern __shared__ __align__(16) unsigned char tbl[]; // [32][T][2] u32
    const int T = blockDim.x, tid = threadIdx.x;
    for (int e = 0; e < 32; ++e) {
        uint2 v = make_uint2((tid + seed + e) * 2654435761u, (tid + e) * 40503u + seed);
        *reinterpret_cast<uint2 *>(tbl + (size_t)e * T * 8 + tid * 8) = v;
    }
    __syncthreads();
    const uint32_t estride = T * 8; // power of two for T = 128 .. 1024
    const uint32_t base_hi = (uint32_t)__cvta_generic_to_shared(tbl) + tid * 8, base_lo = base_hi + 16 * estride;
    const uint32_t mul0 = estride + opaque0, mul4 = estride / 16 + opaque0, mul8 = estride / 256 + opaque0;
    uint32_t Pv0 = ~0u, Pv1 = ~0u, Mv0 = 0, Mv1 = 0;
    uint32_t word = (tid + seed) * 2654435761u;
    long long t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < cols; j += 4) {
        word = word * 1664525u + 1013904223u;
        uint32_t ah[4], al[4];
        if (V == 0) { // IMAD addressing
            const uint32_t w12 = __umulhi(word, 1u << 20), w24 = __umulhi(word, 1u << 8);
            ah[0] = mad_lo(word & 0xF0u, mul4, base_hi), al[0] = mad_lo(word & 0xFu, mul0, base_lo);
            ah[1] = mad_lo(w12 & 0xFu, mul0, base_hi), al[1] = mad_lo(word & 0xF00u, mul8, base_lo);
            ah[2] = mad_lo(w12 & 0xF00u, mul8, base_hi), al[2] = mad_lo(w12 & 0xF0u, mul4, base_lo);
            ah[3] = mad_lo(w24 & 0xF0u, mul4, base_hi), al[3] = mad_lo(w24 & 0xFu, mul0, base_lo);
        } else { // shifts + masks written in C
            const uint32_t wh = (word >> 4) & 0x0F0F0F0Fu, wl = word & 0x0F0F0F0Fu;
#pragma unroll
            for (int k = 0; k < 4; ++k) ah[k] = base_hi + ((wh >> (8 * k)) & 0xFFu) * estride, al[k] = base_lo + ((wl >> (8 * k)) & 0xFFu) * estride;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t h0, h1, l0, l1;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(h0), "=r"(h1) : "r"(ah[k]));
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(l0), "=r"(l1) : "r"(al[k]));
            step<2>(h0, l0, h1, l1, Pv0, Pv1, Mv0, Mv1, one, two);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = Pv0 ^ Pv1 ^ Mv0 ^ Mv1;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V>
void run_lookup(const char *name, int warps) {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
    const int cols = 4096;
    const size_t smem = (size_t)32 * warps * 32 * 8;
    cudaFuncSetAttribute(bench_lookup<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench_lookup<V><<<1, warps * 32, smem>>>(out, 1, 1, 2, cyc, cols, 0);
    bench_lookup<V><<<1, warps * 32, smem>>>(out, 1, 1, 2, cyc, cols, 0);
    long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps/SM %2d: %.2f cycles/column/warp-per-SMSP-slot, %.3f warp-columns/cycle/SM (%s)\n", name, warps,
           (double)c / cols / (warps / 4.0), (double)cols * warps / c, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

template <int V>
void run(const char *name, int warps) {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
    const int cols = 4096;
    bench<V><<<1, warps * 32>>>(out, 1, 1, 2, cyc, cols);
    bench<V><<<1, warps * 32>>>(out, 1, 1, 2, cyc, cols);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    uint32_t chk; cudaMemcpy(&chk, out, 4, cudaMemcpyDeviceToHost);
    printf("%-34s warps/SM %2d: %.2f cycles/column/warp-per-SMSP-slot, %.3f warp-columns/cycle/SM  [chk %08x]\n", name, warps,
           (double)c / cols / (warps / 4.0), (double)cols * warps / c, chk);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int warps : {4, 8, 24, 32}) {
        run<0>("V0 Myers blocks (old)", warps);
        run<1>("V1 Hyyro + IMAD.WIDE carries", warps);
        run<2>("V2 Hyyro + funnel shifts", warps);
        run<3>("V3 Hyyro + SHF carry + IMAD", warps);
        run<4>("V4 Hyyro + C shifts", warps);
        run<5>("V5 Hyyro + IMAD.HI carry + IMAD", warps);
    }
    for (int warps : {4, 8, 16, 24}) {
        run_lookup<0>("lookup + V2, IMAD addressing", warps);
        run_lookup<1>("lookup + V2, C addressing", warps);
    }
    return 0;
}
