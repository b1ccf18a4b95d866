// Microbenchmark (development tool): issue rate of the integer instructions the encoder is made of, alone
// and in pairs, to learn which ones share a pipe on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(x) x x x x x x x x
#define REP32(x) REP8(x) REP8(x) REP8(x) REP8(x)

// each test: 8 independent chains per thread (ILP), 32 instructions per chain step
#define CHAINS uint32_t a0 = seed * threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
#define OP8(fmt) \
    asm volatile(fmt : "+r"(a0) : "r"(k)); asm volatile(fmt : "+r"(a1) : "r"(k)); asm volatile(fmt : "+r"(a2) : "r"(k)); asm volatile(fmt : "+r"(a3) : "r"(k)); \
    asm volatile(fmt : "+r"(a4) : "r"(k)); asm volatile(fmt : "+r"(a5) : "r"(k)); asm volatile(fmt : "+r"(a6) : "r"(k)); asm volatile(fmt : "+r"(a7) : "r"(k));

#define S_LOP  "lop3.b32 %0, %0, %1, 0x55555555, 0x96;"
#define S_SHF  "shf.l.wrap.b32 %0, %0, %1, %1;"
#define S_PRMT "prmt.b32 %0, %0, %1, 0x5432;"
#define S_IADD "add.u32 %0, %0, %1;"

#define S_IMAD "mad.lo.u32 %0, %0, %1, %1;"
#define S_IMADHI "mul.hi.u32 %0, %0, %1;"
#define S_DP2A "dp2a.lo.u32.u32 %0, %0, %1, %1;"
#define S_FLO  "bfind.u32 %0, %0;"
#define S_POPC "popc.b32 %0, %0;"
#define S_SEL  "{.reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, 0x1234567, %0, p;}"
#define C8(f) a0 = f(a0, k); a1 = f(a1, k); a2 = f(a2, k); a3 = f(a3, k); a4 = f(a4, k); a5 = f(a5, k); a6 = f(a6, k); a7 = f(a7, k); \
    asm volatile("" : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7));
#define S_SHR "shr.u32 %0, %0, %1;"

template <int T>
__global__ void __launch_bounds__(256) kern(uint32_t *out, int iters, uint32_t seed, uint32_t k)
{
    CHAINS
    for (int it = 0; it < iters; it++) {
        if (T == 0) { REP8(OP8(S_LOP)) }
        if (T == 1) { REP8(OP8(S_SHF)) }
        if (T == 2) { REP8(OP8(S_PRMT)) }
        if (T == 3) { REP8(OP8(S_IADD)) }
        if (T == 4) { REP8(C8(__vadd2)) }
        if (T == 5) { REP8(OP8(S_IMAD)) }
        if (T == 6) { REP8(OP8(S_IMADHI)) }
        if (T == 7) { REP8(OP8(S_DP2A)) }
        if (T == 8) { REP8(OP8(S_FLO)) }
        if (T == 9) { REP8(OP8(S_POPC)) }
        if (T == 10) { REP8(OP8(S_SEL)) }
        if (T == 11) { REP8(C8(__vmaxu2)) }
        if (T == 12) { REP8(OP8(S_LOP) OP8(S_IMAD)) }     // pairs: 16 instructions per REP
        if (T == 13) { REP8(OP8(S_LOP) OP8(S_DP2A)) }
        if (T == 14) { REP8(OP8(S_IMAD) OP8(S_DP2A)) }
        if (T == 15) { REP8(OP8(S_LOP) OP8(S_IMADHI)) }
        if (T == 16) { REP8(OP8(S_LOP) OP8(S_PRMT)) }
        if (T == 17) { REP8(OP8(S_LOP) C8(__vadd2)) }
        if (T == 18) { REP8(OP8(S_LOP) OP8(S_SHF)) }
        if (T == 19) { REP8(OP8(S_LOP) OP8(S_FLO)) }
        if (T == 20) { REP8(OP8(S_IMAD) C8(__vadd2)) }
        if (T == 21) { REP8(OP8(S_LOP) C8(__vmaxu2)) }
        if (T == 22) { REP8(OP8(S_SHR)) }
        if (T == 23) { REP8(OP8(S_IMAD) OP8(S_IMADHI)) }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

// shared memory: plain LDS / STS / RED.OR with a given address stride (words per lane, in 1/8)
template <int T>
__global__ void __launch_bounds__(256) smem(uint32_t *out, int iters, uint32_t stride_q8)
{
    __shared__ uint32_t s[8192];
    for (int i = threadIdx.x; i < 8192; i += 256) s[i] = i;
    __syncthreads();
    uint32_t t = threadIdx.x, acc = 0;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(s);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            uint32_t w = (((t & 31) * stride_q8) >> 3) + 97 * u + (t >> 5) * 700 + (it & 3);
            uint32_t addr = base + ((w & 8191) << 2);
            if (T == 0) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); acc += v; }
            if (T == 1) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(acc + u) : "memory"); }
            if (T == 2) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(addr), "r"(acc + u) : "memory"); }
            if (T == 3) { uint32_t v; asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(v) : "r"(addr), "r"(u) : "memory"); acc += v; }
            if (T == 4) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr & ~7u)); acc += v.x + v.y; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + s[t];
}

template <class F> float time_it(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main()
{
    uint32_t *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    int dev_clk; cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
    const int iters = 2048;
    const char *names[] = {"LOP3", "SHF", "PRMT", "IADD", "VIADD.16x2", "IMAD", "IMAD.HI", "IDP.2A", "FLO", "POPC", "SETP+SEL", "VIMNMX.16x2",
                           "LOP3+IMAD", "LOP3+IDP", "IMAD+IDP", "LOP3+IMAD.HI", "LOP3+PRMT", "LOP3+VIADD", "LOP3+SHF", "LOP3+FLO", "IMAD+VIADD", "LOP3+VIMNMX", "SHR", "IMAD+IMAD.HI"};
    const int per_rep[] = {8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 16, 8, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 8, 16};
#define RUN(T) { float ms = time_it([&] { kern<T><<<148 * 8, 256>>>(d, iters, 12345u, 3u); }); \
      double inst = (double)iters * 8 * per_rep[T] * 8 /*warps per cta*/ * 8 /*ctas per sm*/; \
      printf("%-14s %.3f ms  %.2f warp-inst/cycle/SM (at %.0f MHz nominal)\n", names[T], ms, inst / (ms * 1e-3 * dev_clk * 1e3), dev_clk / 1e3); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15) RUN(16) RUN(17) RUN(18) RUN(19) RUN(20) RUN(21) RUN(22) RUN(23)
    const char *sn[] = {"LDS.32", "STS.32", "RED.OR", "ATOMS.OR", "LDS.64"};
#define RUNS(T, st) { float ms = time_it([&] { smem<T><<<148 * 8, 256>>>(d, 512, st); }); \
      double inst = 512.0 * 8 * 8 * 8; printf("%-9s stride %5.2f words: %.3f ms  %.3f warp-inst/cycle/SM\n", sn[T], st / 8.0, ms, inst / (ms * 1e-3 * dev_clk * 1e3)); }
    for (uint32_t st : {8u, 12u, 16u, 264u}) { RUNS(0, st) RUNS(1, st) RUNS(2, st) RUNS(3, st) RUNS(4, st) }
    return 0;
}
