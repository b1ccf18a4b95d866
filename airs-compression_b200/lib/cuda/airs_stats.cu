/*
 * airs_stats.cu - residual statistics per job, for choosing encoder parameters
 * (include/airs_cuda.h: airs_cuda_residual_stats).
 *
 * The reference leaves the Golomb parameter to the user (lib/cmp.h:64-112, the
 * examples use fixed values).  What the parameter should be follows from the
 * zig-zag mapped residuals (ref map_to_unsigned, lib/compress/encoder.c:274-286) the
 * encoder will see: a Golomb code with parameter g is shortest for a
 * geometric source whose mean is about g / ln 2.  This kernel delivers, per
 * job, the sum, the maximum and a log2 histogram of the mapped residuals of
 * the job's first frame under its primary preprocessing (NONE: the samples,
 * DIFF and everything else: first differences, ref lib/compress/preprocess.c:
 * 268-290) - one pass over the samples at HBM speed.
 */
#include <cuda_runtime.h>

#include "airs_device.cuh"
#include "../../../include/airs_cuda.h"

extern "C" int airs_internal_fail(int code, const char *text);
extern "C" int airs_internal_check_device(void);

namespace {

constexpr uint32_t kStatThreads = 256;

__device__ __forceinline__ uint32_t mapped(uint32_t r16)
{
	const uint32_t r = r16 & 0xFFFFu;
	return ((r << 1) ^ (0u - (r >> 15))) & 0xFFFFu;
}

__device__ __forceinline__ uint32_t sample16(const uint8_t *src, uint32_t dtype, uint32_t i)
{
	if (dtype == AIRS_DTYPE_I16_IN_I32)
		return __ldg(reinterpret_cast<const uint32_t *>(src) + i) & 0xFFFFu;
	return __ldg(reinterpret_cast<const uint16_t *>(src) + i);
}

/* grid-stride over jobs, one CTA per job at a time */
__global__ void __launch_bounds__(kStatThreads) stats_kernel(const uint8_t *src_base, const airs_job *jobs, uint32_t n_jobs,
							    airs_stats *out)
{
	__shared__ unsigned long long s_sum;
	__shared__ uint32_t s_max, s_hist[kStatThreads / 32][20]; /* one histogram per warp: shared-memory atomics */
	uint32_t *const wh = s_hist[threadIdx.x >> 5];

	for (uint32_t j = blockIdx.x; j < n_jobs; j += gridDim.x) {
		const airs_job &job = jobs[j];
		const uint32_t dtype = job.dtype;
		const uint32_t stride = dtype == AIRS_DTYPE_I16_IN_I32 ? 4u : 2u;
		const bool ok = src_base && dtype <= AIRS_DTYPE_U16 && job.src_size && job.src_size % stride == 0 && job.n_frames;
		const uint32_t n = ok ? job.src_size / stride : 0u;
		const uint8_t *src = src_base + job.src_offset;
		const bool diff = job.params.primary_preprocessing != CMP_PREPROCESS_NONE;
		for (uint32_t i = threadIdx.x; i < (kStatThreads / 32) * 20; i += kStatThreads)
			(&s_hist[0][0])[i] = 0;
		if (threadIdx.x == 0) {
			s_sum = 0;
			s_max = 0;
		}
		__syncthreads();
		unsigned long long sum = 0;
		uint32_t mx = 0;
		const bool vec = dtype != AIRS_DTYPE_I16_IN_I32 && ((uintptr_t)src & 15u) == 0;
		const uint32_t n8 = vec ? n / 8u : 0u;
		for (uint32_t p = threadIdx.x; p < n8; p += kStatThreads) { /* 8 samples per 16-byte load */
			const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src) + p);
			const uint32_t w[4] = {q.x, q.y, q.z, q.w};
			uint32_t prev = (diff && p) ? __ldg(reinterpret_cast<const uint16_t *>(src) + 8u * p - 1u) : 0u;
#pragma unroll
			for (int k = 0; k < 8; k++) {
				const uint32_t x = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
				const uint32_t m = mapped(diff ? x - prev : x);
				prev = x;
				sum += m;
				mx = max(mx, m);
				atomicAdd(wh + (32u - (uint32_t)__clz((int)m)), 1u); /* bins: 0, 1, 2-3, 4-7, .. */
			}
		}
		for (uint32_t i = 8u * n8 + threadIdx.x; i < n; i += kStatThreads) {
			const uint32_t x = sample16(src, dtype, i);
			const uint32_t prev = (diff && i) ? sample16(src, dtype, i - 1u) : 0u;
			const uint32_t m = mapped(diff ? x - prev : x);
			sum += m;
			mx = max(mx, m);
			atomicAdd(wh + (32u - (uint32_t)__clz((int)m)), 1u);
		}
		/* warp reductions, then one atomic per warp and value */
		for (uint32_t d = 16; d; d >>= 1) {
			sum += __shfl_down_sync(0xFFFFFFFFu, sum, d);
			mx = max(mx, __shfl_down_sync(0xFFFFFFFFu, mx, d));
		}
		if ((threadIdx.x & 31u) == 0) {
			atomicAdd(&s_sum, sum);
			atomicMax(&s_max, mx);
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			out[j].sum_mapped = s_sum;
			out[j].n_samples = n;
			out[j].max_mapped = s_max;
		}
		if (threadIdx.x < 17) {
			uint32_t c = 0;
			for (uint32_t w = 0; w < kStatThreads / 32; w++)
				c += s_hist[w][threadIdx.x];
			out[j].log2_hist[threadIdx.x] = c;
		}
		__syncthreads();
	}
}

} /* namespace */

extern "C" int airs_cuda_residual_stats(const void *src, const struct airs_job *jobs, uint32_t n_jobs,
					struct airs_stats *stats, void *stream)
{
	if (!jobs || !stats)
		return airs_internal_fail(AIRS_E_ARGUMENT, "jobs and stats must be non-NULL");
	if (n_jobs == 0)
		return AIRS_OK;
	int rc = airs_internal_check_device();
	if (rc != AIRS_OK)
		return rc;
	int dev = 0, sms = 148;
	if (cudaGetDevice(&dev) == cudaSuccess)
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const uint32_t cap = 8u * (uint32_t)sms;
	stats_kernel<<<n_jobs < cap ? n_jobs : cap, kStatThreads, 0, (cudaStream_t)stream>>>(
		(const uint8_t *)src, jobs, n_jobs, stats);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		return airs_internal_fail(AIRS_E_CUDA, cudaGetErrorString(e));
	return AIRS_OK;
}

/* Golomb parameter for a mean mapped residual: the code is shortest near g = mean * ln 2
 * (Golomb 1966; Gallager / van Voorhis 1975 for the exact rule); integer arithmetic so that the host
 * and any other caller agree: g = max(1, (mean_num * 45426 / mean_den) >> 16) with 45426 / 65536 = ln 2 */
extern "C" uint32_t airs_cuda_golomb_param_for_mean(uint64_t sum_mapped, uint32_t n_samples)
{
	if (n_samples == 0)
		return 1;
	/* sum < 2^16 * 2^32: the product fits 64 bits for n up to 2^32 / ... use 128-bit free form */
	const unsigned long long q = sum_mapped / n_samples, r = sum_mapped % n_samples;
	unsigned long long g = (q * 45426ull + (r * 45426ull) / n_samples) >> 16;
	if (g < 1)
		g = 1;
	if (g > 65535)
		g = 65535;
	return (uint32_t)g;
}
