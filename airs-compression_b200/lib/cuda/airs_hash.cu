/*
 * airs_hash.cu - airs_cuda_hash_streams: the 64-bit hash of include/airs_stream_hash.h over every
 * stream a batch produced, one warp per stream, so that whole workloads can be compared with the
 * CPU reference through 12 bytes per stream (verification aid; no reference counterpart).
 */
#include <cuda_runtime.h>

#include "../../../include/airs_stream_hash.h"
#include "airs_device.cuh"
#include "airs_launch.h"

/* hash of the r bytes at p, by one warp; the result is valid on lane 0 */
__device__ __forceinline__ uint64_t warp_hash(const uint8_t *p, uint32_t r, uint32_t lane)
{
	const uint32_t mis = (uint32_t)((uintptr_t)p & 7u);
	const uint64_t *q = reinterpret_cast<const uint64_t *>(p - mis);
	const uint32_t n_words = (r + 7u) / 8u;
	uint64_t h = 0;

	for (uint32_t i = lane; i < n_words; i += 32u) {
		const uint32_t valid = min(8u, r - 8u * i); /* bytes of word i inside the stream */
		uint64_t w = q[i] >> (8u * mis);
		if (mis && mis + valid > 8u)
			w |= q[i + 1u] << (64u - 8u * mis);
		if (valid < 8u)
			w &= (1ull << (8u * valid)) - 1ull;
		h += airs_hash_term(w, i);
	}
#pragma unroll
	for (int d = 16; d; d >>= 1) {
		const uint32_t lo = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)h, d);
		const uint32_t hi = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)(h >> 32), d);
		h += ((uint64_t)hi << 32) | lo;
	}
	return h + airs_hash_mix((uint64_t)r);
}

__global__ void __launch_bounds__(128) airs_hash_kernel(AirsLaunch b, uint64_t *hashes)
{
	const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;

	if (k >= b.n_results)
		return;
	const uint32_t j = b.result_job[k]; /* written by airs_plan_kernel */
	const uint32_t r = b.results[k];
	uint64_t h = 0;
	if (j < b.n_jobs && !airs_failed(r) && b.dst) {
		const airs_job &job = b.jobs[j];
		const uint8_t *p = b.layout == AIRS_LAYOUT_CONCAT
					   ? b.dst + b.out_offsets[k]
					   : b.dst + job.dst_offset + (uint64_t)(k - job.first_result) * job.dst_frame_stride;
		h = warp_hash(p, r, lane);
	}
	if (lane == 0)
		hashes[k] = h;
}

__global__ void __launch_bounds__(128) airs_hash_ranges_kernel(const uint8_t *base, const uint64_t *offsets, const uint32_t *sizes,
							       uint32_t n, uint64_t *hashes)
{
	const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;

	if (k >= n)
		return;
	const uint32_t r = sizes[k];
	const uint64_t h = airs_failed(r) ? 0 : warp_hash(base + offsets[k], r, lane);
	if (lane == 0)
		hashes[k] = h;
}

extern "C" cudaError_t airs_launch_hash(const AirsLaunch *b, uint64_t *hashes, cudaStream_t stream)
{
	if (b->n_results == 0)
		return cudaSuccess;
	airs_hash_kernel<<<(b->n_results + 3u) / 4u, 128, 0, stream>>>(*b, hashes);
	return cudaGetLastError();
}

extern "C" cudaError_t airs_launch_hash_ranges(const uint8_t *base, const uint64_t *offsets, const uint32_t *sizes, uint32_t n,
					       uint64_t *hashes, cudaStream_t stream)
{
	if (n == 0)
		return cudaSuccess;
	airs_hash_ranges_kernel<<<(n + 3u) / 4u, 128, 0, stream>>>(base, offsets, sizes, n, hashes);
	return cudaGetLastError();
}
