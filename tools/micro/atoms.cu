// Microbenchmark: shared-memory atomicOr vs plain store throughput with the access pattern of the bit packer.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, int iters, uint32_t stride_q8)
{
    __shared__ uint32_t s[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) s[i] = 0;
    __syncthreads();
    uint32_t t = threadIdx.x;
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
        uint32_t w = ((t * stride_q8) >> 3) + (it & 7);   // thread t -> word ~ t*stride/8
        uint32_t v = (t + it) | 1u;
        if (MODE == 0) { s[w & 4095] = v; }
        else if (MODE == 1) { atomicOr(&s[w & 4095], v); }
        else if (MODE == 2) { atomicOr(&s[w & 4095], v); atomicOr(&s[(w + 1) & 4095], v >> 3); }
        else if (MODE == 3) { uint32_t m = __match_any_sync(0xffffffffu, w); acc += __reduce_or_sync(m, v); }
        else if (MODE == 4) { acc += __shfl_up_sync(0xffffffffu, v, 1) + __shfl_down_sync(0xffffffffu, w, 1); }
    }
    __syncthreads();
    out[blockIdx.x * 256 + t] = s[t] + acc;
}
template <int MODE> void run(const char *name, uint32_t stride_q8)
{
    uint32_t *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int iters = 4096;
    k<MODE><<<148 * 8, 256>>>(d, iters, stride_q8);
    cudaEventRecord(a);
    k<MODE><<<148 * 8, 256>>>(d, iters, stride_q8);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double warp_ops = 148.0 * 8 * 8 * iters;            // warp-level loop iterations
    double cyc = ms * 1e-3 * 1.9e9;                      // approx cycles
    printf("%-28s stride %4.2f words/thread: %.3f ms, %.2f cycles per warp-iteration per SM\n", name, stride_q8 / 8.0, ms,
           cyc / (warp_ops / 148.0));
    cudaFree(d);
}
int main()
{
    for (uint32_t st : {8u, 11u, 16u, 24u}) {
        run<0>("plain STS", st);
        run<1>("1x atomicOr", st);
        run<2>("2x atomicOr", st);
        run<3>("match_any+redux.or", st);
        run<4>("2x shfl", st);
    }
    return 0;
}
