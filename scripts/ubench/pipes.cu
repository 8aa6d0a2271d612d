// Micro-benchmark: issue throughput of the integer instructions the bit-vector kernels are built from, alone and
// mixed, on one SM sub-partition set (B200, sm_100a). Prints warp-instructions per cycle per SM.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; run: ./pipes
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256
template <int MODE>
__global__ void bench(uint32_t *out, uint32_t seed, uint32_t one, uint32_t two, long long *cycles) {
    uint32_t a = threadIdx.x + seed, b = a * 3u, c = a ^ 0x1234u, d = a + 77u, e = b + 5u, f = c + 9u, g = d ^ e, h = f ^ a;
    uint32_t lo, hi;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int r = 0; r < REP / 8; ++r) {
            if (MODE == 0) { // LOP3 x8 (4 independent chains x2)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d) : "r"(e), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(e), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d) : "r"(e), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(e), "r"(f));
            } else if (MODE == 1) { // IMAD x8
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(two), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(two), "r"(f));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(g) : "r"(two), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b) : "r"(two), "r"(f));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(two), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(two), "r"(f));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(g) : "r"(two), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b) : "r"(two), "r"(f));
            } else if (MODE == 2) { // IMAD.WIDE x8
                uint64_t w0, w1, w2, w3;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w0) : "r"(a), "r"(two));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w1) : "r"(d), "r"(two));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w2) : "r"(g), "r"(two));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w3) : "r"(b), "r"(two));
                a = (uint32_t)w0 ^ (uint32_t)(w0 >> 32); d = (uint32_t)w1 ^ (uint32_t)(w1 >> 32);
                g = (uint32_t)w2 ^ (uint32_t)(w2 >> 32); b = (uint32_t)w3 ^ (uint32_t)(w3 >> 32);
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w0) : "r"(a), "r"(two));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w1) : "r"(d), "r"(two));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w2) : "r"(g), "r"(two));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w3) : "r"(b), "r"(two));
                a = (uint32_t)w0 ^ (uint32_t)(w0 >> 32); d = (uint32_t)w1 ^ (uint32_t)(w1 >> 32);
                g = (uint32_t)w2 ^ (uint32_t)(w2 >> 32); b = (uint32_t)w3 ^ (uint32_t)(w3 >> 32);
            } else if (MODE == 3) { // IMAD.HI x8
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(c));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(d) : "r"(f));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(g) : "r"(c));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(b) : "r"(f));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(c));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(d) : "r"(f));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(g) : "r"(c));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(b) : "r"(f));
            } else if (MODE == 4) { // 4 LOP3 + 4 IMAD interleaved
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(two), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(e) : "r"(two), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(two), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(e) : "r"(two), "r"(f));
            } else if (MODE == 5) { // 6 LOP3 + 2 IMAD.WIDE
                uint64_t w0, w1;
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w0) : "r"(d), "r"(two));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(h), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w1) : "r"(e), "r"(two));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(h), "r"(c));
                d = (uint32_t)(w0 >> 32) + (uint32_t)w1; e = (uint32_t)(w1 >> 32) + (uint32_t)w0;
            } else if (MODE == 6) { // IADD3 x8
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(c));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(d) : "r"(f));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(g) : "r"(c));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b) : "r"(f));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(c));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(d) : "r"(f));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(g) : "r"(c));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b) : "r"(f));
            } else if (MODE == 7) { // 4 LOP3 + 4 SHF (funnel)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(d) : "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(e) : "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(d) : "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(e) : "r"(f));
            } else if (MODE == 9) { // dp4a x8
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(two), "r"(c));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(two), "r"(f));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(g) : "r"(two), "r"(c));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(b) : "r"(two), "r"(f));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(two), "r"(c));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(two), "r"(f));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(g) : "r"(two), "r"(c));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(b) : "r"(two), "r"(f));
            } else if (MODE == 10) { // 4 LOP3 + 4 dp4a
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(two), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(e) : "r"(two), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(two), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(e) : "r"(two), "r"(f));
            } else if (MODE == 11) { // 4 LOP3 + 4 IMAD.HI
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(d) : "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(e) : "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(d) : "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(e) : "r"(f));
            } else if (MODE == 12) { // 6 LOP3 + 2 IMAD.HI
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(d) : "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(h), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(e) : "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(h), "r"(c));
            } else if (MODE == 13) { // 6 LOP3 + 2 IMAD
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(two), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(h), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(e) : "r"(two), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(h), "r"(c));
            } else if (MODE == 8) { // 5 LOP3 + 3 IMAD
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(two), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(h), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(e) : "r"(two), "r"(f));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(two), "r"(f));
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps) {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
    bench<MODE><<<1, warps * 32>>>(out, 1, 1, 2, cyc);
    bench<MODE><<<1, warps * 32>>>(out, 1, 1, 2, cyc);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double inst = 64.0 * REP * warps;
    printf("%-28s warps/SM %2d: %.3f warp-inst/cycle/SM (%.3f per SMSP)\n", name, warps, inst / c, inst / c / 4);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int warps : {8, 24}) {
        run<0>("LOP3", warps); run<1>("IMAD", warps); run<2>("IMAD.WIDE(+1 LOP3 each)", warps); run<3>("IMAD.HI", warps);
        run<4>("LOP3:IMAD 4:4", warps); run<5>("LOP3:IMAD.WIDE 6:2 (+2 IADD)", warps); run<6>("IADD", warps);
        run<7>("LOP3:SHF 4:4", warps); run<8>("LOP3:IMAD 5:3", warps);
        run<9>("DP4A", warps); run<10>("LOP3:DP4A 4:4", warps); run<11>("LOP3:IMAD.HI 4:4", warps); run<12>("LOP3:IMAD.HI 6:2", warps); run<13>("LOP3:IMAD 6:2", warps);
    }
    return 0;
}
