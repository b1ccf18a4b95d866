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
	const uint32_t v = __ldg(reinterpret_cast<const uint16_t *>(src) + i);
	return (dtype & AIRS_DTYPE_BE) ? ((v << 8) | (v >> 8)) & 0xFFFFu : v;
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
		const bool ok = src_base && (dtype <= AIRS_DTYPE_U16 || dtype == AIRS_DTYPE_I16_BE || dtype == AIRS_DTYPE_U16_BE) && job.src_size && job.src_size % stride == 0 && job.n_frames;
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
		const bool vec = dtype != AIRS_DTYPE_I16_IN_I32 && !(dtype & AIRS_DTYPE_BE) && ((uintptr_t)src & 15u) == 0;
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

/* Parameter search: the exact code bits of a job's first frame under each candidate encoder (ref
 * cmp_encoder_encode_s16, encoder.c:327-378: the length of every code word, nothing is written).  One CTA per job at
 * a time; the candidates' constants in shared memory, a running sum per candidate and thread in registers. */
constexpr uint32_t kCandRegs = 8; /* candidates per pass over the samples */

__global__ void __launch_bounds__(kStatThreads) candidate_kernel(const uint8_t *src_base, const airs_job *jobs, uint32_t n_jobs,
								const airs_candidate *cand, uint32_t n_cand, unsigned long long *bits)
{
	__shared__ EncConst s_enc[AIRS_MAX_CANDIDATES];
	__shared__ uint32_t s_ok[AIRS_MAX_CANDIDATES];
	__shared__ unsigned long long s_tot[AIRS_MAX_CANDIDATES];

	if (threadIdx.x < n_cand) {
		const airs_candidate c = cand[threadIdx.x];
		s_ok[threadIdx.x] = !airs_failed(airs_encoder_check(c.encoder_type, c.g, c.outlier)) && c.encoder_type != CMP_ENCODER_UNCOMPRESSED;
		airs_enc_const(&s_enc[threadIdx.x], c.encoder_type, s_ok[threadIdx.x] ? c.g : 1u, c.outlier);
	}
	for (uint32_t j = blockIdx.x; j < n_jobs; j += gridDim.x) {
		const airs_job &job = jobs[j];
		const uint32_t dtype = job.dtype;
		const uint32_t stride = dtype == AIRS_DTYPE_I16_IN_I32 ? 4u : 2u;
		const bool ok = src_base && (dtype <= AIRS_DTYPE_U16 || dtype == AIRS_DTYPE_I16_BE || dtype == AIRS_DTYPE_U16_BE) && job.src_size && job.src_size % stride == 0 && job.n_frames;
		const uint32_t n = ok ? job.src_size / stride : 0u;
		const uint8_t *src = src_base + job.src_offset;
		const bool diff = job.params.primary_preprocessing != CMP_PREPROCESS_NONE;
		if (threadIdx.x < AIRS_MAX_CANDIDATES)
			s_tot[threadIdx.x] = 0;
		__syncthreads();
		for (uint32_t c0 = 0; c0 < n_cand; c0 += kCandRegs) {
			uint32_t acc[kCandRegs];
#pragma unroll
			for (uint32_t k = 0; k < kCandRegs; k++)
				acc[k] = 0;
			for (uint32_t i = threadIdx.x; i < n; i += kStatThreads) {
				const uint32_t x = sample16(src, dtype, i);
				const uint32_t prev = (diff && i) ? sample16(src, dtype, i - 1u) : 0u;
				const uint32_t m = mapped(diff ? x - prev : x);
#pragma unroll
				for (uint32_t k = 0; k < kCandRegs; k++) {
					if (c0 + k < n_cand) { /* (the same for all threads) */
						const EncConst &e = s_enc[c0 + k];
						uint32_t cw, cwlen, raw, rawlen;
						if (e.type == CMP_ENCODER_GOLOMB_ZERO)
							airs_encode_mapped<CMP_ENCODER_GOLOMB_ZERO>(e, m, cw, cwlen, raw, rawlen);
						else
							airs_encode_mapped<CMP_ENCODER_GOLOMB_MULTI>(e, m, cw, cwlen, raw, rawlen);
						acc[k] += cwlen + rawlen; /* <= 48 bits a sample, 2^16 samples a thread at most */
					}
				}
			}
#pragma unroll
			for (uint32_t k = 0; k < kCandRegs; k++) {
				unsigned long long v = acc[k];
				for (uint32_t d = 16; d; d >>= 1)
					v += __shfl_down_sync(0xFFFFFFFFu, v, d);
				if ((threadIdx.x & 31u) == 0 && c0 + k < n_cand)
					atomicAdd(&s_tot[c0 + k], v);
			}
		}
		__syncthreads();
		if (threadIdx.x < n_cand)
			bits[(size_t)j * n_cand + threadIdx.x] = (ok && s_ok[threadIdx.x]) ? s_tot[threadIdx.x] : ~0ull;
		__syncthreads();
	}
}

} /* namespace */

extern "C" int airs_cuda_candidate_bits(const void *src, const struct airs_job *jobs, uint32_t n_jobs,
					const struct airs_candidate *cand, uint32_t n_cand, uint64_t *bits, void *stream)
{
	if (!jobs || !cand || !bits)
		return airs_internal_fail(AIRS_E_ARGUMENT, "jobs, cand and bits must be non-NULL");
	if (n_cand == 0 || n_cand > AIRS_MAX_CANDIDATES)
		return airs_internal_fail(AIRS_E_ARGUMENT, "1 .. AIRS_MAX_CANDIDATES candidates");
	if (n_jobs == 0)
		return AIRS_OK;
	int rc = airs_internal_check_device();
	if (rc != AIRS_OK)
		return rc;
	int dev = 0, sms = 148;
	if (cudaGetDevice(&dev) == cudaSuccess)
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const uint32_t cap = 8u * (uint32_t)sms;
	candidate_kernel<<<n_jobs < cap ? n_jobs : cap, kStatThreads, 0, (cudaStream_t)stream>>>(
		(const uint8_t *)src, jobs, n_jobs, cand, n_cand, reinterpret_cast<unsigned long long *>(bits));
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		return airs_internal_fail(AIRS_E_CUDA, cudaGetErrorString(e));
	return AIRS_OK;
}

extern "C" uint32_t airs_cuda_param_candidates(const struct airs_stats *st, uint32_t encoder_type, struct airs_candidate *out,
					       uint32_t max)
{
	static const uint32_t eighths[] = {8, 4, 6, 7, 9, 10, 12, 16}; /* g0 first, then its neighbourhood */
	static const uint32_t outliers[] = {8, 4, 16};                /* multiples of g */
	uint32_t n = 0;
	if (!st || !out || (encoder_type != CMP_ENCODER_GOLOMB_ZERO && encoder_type != CMP_ENCODER_GOLOMB_MULTI))
		return 0;
	const uint32_t g0 = airs_cuda_golomb_param_for_mean(st->sum_mapped, st->n_samples);
	for (uint32_t a = 0; a < sizeof(eighths) / sizeof(eighths[0]); a++) {
		unsigned long long g = ((unsigned long long)g0 * eighths[a] + 4u) / 8u;
		g = g < 1 ? 1 : g > 65535 ? 65535 : g;
		for (uint32_t b = 0; b < (encoder_type == CMP_ENCODER_GOLOMB_MULTI ? 3u : 1u); b++) {
			airs_candidate c = {encoder_type, (uint32_t)g, encoder_type == CMP_ENCODER_GOLOMB_MULTI ? (uint32_t)g * outliers[b] : 0u, 0u};
			if (airs_failed(airs_encoder_check(c.encoder_type, c.g, c.outlier)))
				continue;
			bool seen = false;
			for (uint32_t k = 0; k < n; k++)
				seen = seen || (out[k].g == c.g && out[k].outlier == c.outlier);
			if (!seen && n < max)
				out[n++] = c;
		}
	}
	return n;
}

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
