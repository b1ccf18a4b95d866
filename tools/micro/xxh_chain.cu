// Microbenchmark (development tool): cycles per XXH32 round of ONE dependency chain, for several ways to write the round.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xxh_chain xxh_chain.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define P1 0x9E3779B1u
#define P2 0x85EBCA77u
template <int T> __global__ void k(uint32_t *out, long long *cyc, const uint32_t *in, int n)
{
    __shared__ uint32_t s[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) s[i] = in[i];
    __syncwarp();
    uint32_t v = threadIdx.x, v2 = v + 77;
    long long t0 = clock64();
    for (int it = 0; it < n; it++) {
#pragma unroll 16
        for (int i = 0; i < 64; i++) {
            uint32_t x = s[(i * 4 + (threadIdx.x & 3)) & 1023];
            if (T == 0) v = __funnelshift_l(v + x * P2, v + x * P2, 13) * P1;
            if (T == 1) { uint32_t y = x * P2; uint32_t t = v + y; v = __funnelshift_l(t, t, 13) * P1; }
            if (T == 2) { uint32_t z = x * (P2 * (P1 << 13)); uint32_t t = v + x * P2; v = v * (P1 << 13) + z + (t >> 19) * P1; }
            if (T == 3) { uint32_t t = v + x * P2; v = t * (P1 << 13) + __umulhi(t, 1u << 13) * P1; }
            if (T == 4) { v = __funnelshift_l(v + x * P2, v + x * P2, 13) * P1; v2 = __funnelshift_l(v2 + x * P2, v2 + x * P2, 13) * P1; }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = v ^ v2;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    uint32_t *o, *in; long long *c, h;
    cudaMalloc(&o, 128); cudaMalloc(&in, 4096); cudaMalloc(&c, 8); cudaMemset(in, 0x5A, 4096);
    const int n = 4096;
#define RUN(T, name) k<T><<<1, 32>>>(o, c, in, n); k<T><<<1, 32>>>(o, c, in, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("%-40s %.2f cycles per round\n", name, (double)h / (64.0 * n));
    RUN(0, "imad, shf, imad") RUN(1, "mul off chain: iadd, shf, imad") RUN(2, "split rotate: iadd, shr, imad | imad") RUN(3, "imad, imad.hi, imad") RUN(4, "two chains interleaved (per pair)")
    return 0;
}
