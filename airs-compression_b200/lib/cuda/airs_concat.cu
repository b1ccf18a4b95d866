/*
 * airs_concat.cu - the two-phase CONCAT layout: the batch is encoded like a SLOTS
 * batch into temporary slots (every kernel of airs_kernels.cu at full speed, the
 * warp-per-job kernel included), one scan turns the stream sizes into offsets,
 * and a copy kernel lays the streams out back to back - what
 * `airspace -c --stdout f1 f2 ...` produces (ref programs/airspacecli.c:131-202).
 *
 * The single-phase path (airs_encode_kernel: size every frame, look-back scan,
 * encode at the final place) stays as the path that needs no temporary memory;
 * it is also what runs when the temporary slots or the destination turn out too
 * small, because only it follows the reference loop through a destination that
 * overflows (a frame that does not fit fails, its context goes on from there).
 */
#include <cuda_runtime.h>

#include "airs_device.cuh"
#include "airs_launch.h"

namespace {

constexpr uint32_t kScanThreads = 1024;
constexpr uint32_t kBigStream = 16384; /* streams from this size on are copied by a whole CTA */

__device__ __forceinline__ uint64_t align16(uint64_t v)
{
	return (v + 15u) & ~(uint64_t)15u;
}

/* what the two scans add up: the temporary bytes of a job, the stream bytes of a frame */
__device__ __forceinline__ uint64_t job_bytes(const AirsConcat &c, uint32_t j)
{
	const uint64_t b = (uint64_t)c.jobs[j].n_frames * align16(c.jobs[j].dst_capacity);
	return b > c.tmp_size ? c.tmp_size + 1u : b; /* one of these alone closes the gate; no overflow of the sum */
}

__device__ __forceinline__ uint64_t frame_bytes(const AirsConcat &c, uint32_t k)
{
	const uint32_t r = c.results[k];
	return airs_failed(r) ? 0u : r;
}

/* exclusive scan of one value per thread over the CTA; the total goes to `total` */
__device__ uint64_t cta_exclusive_scan(uint64_t v, uint64_t *warp_sums, uint64_t &total)
{
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint64_t inc = v;
	for (uint32_t d = 1; d < 32u; d <<= 1) {
		const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
		if (lane >= d)
			inc += t;
	}
	if (lane == 31u)
		warp_sums[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		uint64_t w = lane < blockDim.x / 32u ? warp_sums[lane] : 0u, winc = w;
		for (uint32_t d = 1; d < 32u; d <<= 1) {
			const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, winc, d);
			if (lane >= d)
				winc += t;
		}
		warp_sums[lane] = winc - w;
		if (lane == 31u)
			warp_sums[32] = winc;
	}
	__syncthreads();
	total = warp_sums[32];
	const uint64_t r = warp_sums[warp] + inc - v;
	__syncthreads();
	return r;
}

/* three-kernel scan over n values (WHAT = 0: jobs, 1: frames): sums of tiles of 1024, a scan
 * over those sums by one CTA, then the tiles again with their offsets */
template <int WHAT>
__global__ void __launch_bounds__(kScanThreads) concat_tile_sums_kernel(AirsConcat c, uint32_t n, uint64_t *sums)
{
	__shared__ uint64_t warp_sums[33];
	if (WHAT == 1 && (*c.flag || c.flag[1]))
		return;
	const uint32_t i = blockIdx.x * kScanThreads + threadIdx.x;
	const uint64_t v = i < n ? (WHAT == 0 ? job_bytes(c, i) : frame_bytes(c, i)) : 0u;
	uint64_t total;
	(void)cta_exclusive_scan(v, warp_sums, total);
	if (threadIdx.x == 0)
		sums[blockIdx.x] = total;
}

template <int WHAT>
__global__ void __launch_bounds__(kScanThreads) concat_top_kernel(AirsConcat c, uint32_t n_tiles, uint64_t *sums)
{
	__shared__ uint64_t warp_sums[33];
	if (WHAT == 1 && (*c.flag || c.flag[1]))
		return;
	uint64_t carry = 0;
	for (uint32_t t0 = 0; t0 < n_tiles; t0 += kScanThreads) {
		const uint32_t i = t0 + threadIdx.x;
		const uint64_t v = i < n_tiles ? sums[i] : 0u;
		uint64_t total;
		const uint64_t ex = cta_exclusive_scan(v, warp_sums, total);
		if (i < n_tiles)
			sums[i] = carry + ex;
		carry += total;
	}
	if (threadIdx.x == 0) {
		sums[n_tiles] = carry; /* the grand total */
		/* temporary slots: 32 bytes of slack, the copy reads whole 16-byte groups around the last
		 * bytes of a stream.  Streams: a destination that cannot hold them all is the business of
		 * the single-phase path, which follows the reference loop through the overflow. */
		if (WHAT == 0 ? (carry + 32u > c.tmp_size || !c.tmp || !c.dst) : carry + (c.base ? *c.base : 0u) > c.dst_size)
			atomicExch(c.flag, 1u);
		if (WHAT == 1)
			*c.n_big = 0;
	}
}

/* jobs: the job table rewritten onto the temporary slots (frames of a job one after the other,
 * capacities rounded up to 16 bytes) */
__global__ void __launch_bounds__(kScanThreads) concat_slots_kernel(AirsConcat c, const uint64_t *sums)
{
	__shared__ uint64_t warp_sums[33];
	const uint32_t j = blockIdx.x * kScanThreads + threadIdx.x;
	const uint64_t v = j < c.n_jobs ? job_bytes(c, j) : 0u;
	uint64_t total;
	const uint64_t off = sums[blockIdx.x] + cta_exclusive_scan(v, warp_sums, total);
	if (j < c.n_jobs) {
		airs_job job = c.jobs[j];
		job.dst_offset = off;
		job.dst_frame_stride = align16(job.dst_capacity);
		job.first_result -= c.result_base; /* a slice of a larger batch: results / out_offsets start at its first frame */
		c.slot_jobs[j] = job;
	}
}

/* frames: out_offsets[k] = bytes of the streams in front of frame k (a frame that failed has
 * none); long streams are listed for the CTA-wide copy */
__global__ void __launch_bounds__(kScanThreads) concat_offsets_kernel(AirsConcat c, const uint64_t *sums, uint32_t n_tiles)
{
	__shared__ uint64_t warp_sums[33];
	if (*c.flag || c.flag[1])
		return;
	const uint32_t k = blockIdx.x * kScanThreads + threadIdx.x;
	const uint64_t v = k < c.n_results ? frame_bytes(c, k) : 0u;
	uint64_t total;
	const uint64_t base = c.base ? *c.base : 0u; /* a slice of a larger batch starts where the slice before ended */
	const uint64_t off = base + sums[blockIdx.x] + cta_exclusive_scan(v, warp_sums, total);
	if (k < c.n_results) {
		c.out_offsets[k] = off;
		if (v >= kBigStream)
			c.big_list[atomicAdd(c.n_big, 1u)] = k;
	}
	if (k == 0)
		c.out_offsets[c.n_results] = base + sums[n_tiles];
}

/* 16 bytes from byte position p (any alignment) of a 16-byte aligned source */
__device__ __forceinline__ uint4 load_shifted(const uint8_t *src, uint64_t p)
{
	const uint4 *s4 = reinterpret_cast<const uint4 *>(src + (p & ~(uint64_t)15u));
	const uint32_t sh = (uint32_t)(p & 15u);
	const uint4 a = __ldg(s4);
	if (sh == 0u)
		return a;
	const uint4 b = __ldg(s4 + 1);
	const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
	const uint32_t i = sh >> 2, bs = 8u * (sh & 3u);
	uint32_t o[4];
#pragma unroll
	for (uint32_t k = 0; k < 4u; k++) {
		uint32_t lo = 0, hi = 0;
#pragma unroll
		for (uint32_t q = 0; q < 4u; q++) { /* w[i + k], w[i + k + 1] without dynamic register indexing */
			lo = i == q ? w[q + k] : lo;
			hi = i == q ? w[q + k + 1] : hi;
		}
		o[k] = __funnelshift_r(lo, hi, bs);
	}
	return make_uint4(o[0], o[1], o[2], o[3]);
}

/* bytes [a, b) of a stream from its temporary slot (16-byte aligned) to its place (any
 * alignment), by `n_thr` threads of which this is number `thr`.  Whole 16-byte groups of the
 * destination are stored as such; the bytes of [a, b) in front of the first and behind the
 * last group travel one by one (at the ends of a stream they share their groups with the
 * neighbour streams). */
__device__ __forceinline__ void copy_piece(const uint8_t *src, uint8_t *dst, uint32_t a, uint32_t b, uint32_t thr,
					   uint32_t n_thr)
{
	const uint32_t head = min(b - a, (uint32_t)((16u - ((uintptr_t)(dst + a) & 15u)) & 15u));
	const uint32_t groups = (b - a - head) / 16u;
	const uint32_t g0 = a + head;
	for (uint32_t i = a + thr; i < g0; i += n_thr)
		dst[i] = src[i];
	uint32_t g = thr;
	for (; g + 3u * n_thr < groups; g += 4u * n_thr) { /* four groups in flight per thread */
		uint4 q[4];
#pragma unroll
		for (uint32_t u = 0; u < 4u; u++)
			q[u] = load_shifted(src, g0 + 16u * (uint64_t)(g + u * n_thr));
#pragma unroll
		for (uint32_t u = 0; u < 4u; u++)
			*reinterpret_cast<uint4 *>(dst + g0 + 16u * (uint64_t)(g + u * n_thr)) = q[u];
	}
	for (; g < groups; g += n_thr)
		*reinterpret_cast<uint4 *>(dst + g0 + 16u * (uint64_t)g) = load_shifted(src, g0 + 16u * (uint64_t)g);
	for (uint32_t i = g0 + 16u * groups + thr; i < b; i += n_thr)
		dst[i] = src[i];
}

__device__ __forceinline__ bool stream_of(const AirsConcat &c, uint32_t k, const uint8_t *&src, uint8_t *&dst, uint32_t &r)
{
	r = c.results[k];
	if (airs_failed(r) || r == 0u)
		return false;
	const uint32_t j = c.result_job[k]; /* written by airs_plan_kernel */
	if (j >= c.n_jobs)
		return false;
	const airs_job &job = c.slot_jobs[j];
	const uint32_t f = k - job.first_result;
	if (f >= job.n_frames)
		return false;
	src = c.tmp + job.dst_offset + (uint64_t)f * job.dst_frame_stride;
	dst = c.dst + c.out_offsets[k];
	return true;
}

/* the streams go from their temporary slots to their places in the concatenation: short ones
 * by one warp each, long ones (listed by concat_offsets_kernel) by one CTA each, 8 KiB pieces
 * dealt out to its warps */
__global__ void __launch_bounds__(256) concat_gather_kernel(AirsConcat c)
{
	if (*c.flag || c.flag[1])
		return;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, wpc = blockDim.x / 32u;
	const uint32_t n_warps = gridDim.x * wpc;
	const uint8_t *src;
	uint8_t *dst;
	uint32_t r;

	for (uint32_t k = blockIdx.x * wpc + warp; k < c.n_results; k += n_warps)
		if (stream_of(c, k, src, dst, r) && r < kBigStream)
			copy_piece(src, dst, 0u, r, lane, 32u);
	const uint32_t n_big = *c.n_big;
	for (uint32_t i = blockIdx.x; i < n_big; i += gridDim.x) {
		if (!stream_of(c, c.big_list[i], src, dst, r))
			continue;
		/* pieces start at multiples of 8 KiB of the DESTINATION's 16-byte grid */
		const uint32_t lead = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
		for (uint32_t p = warp;; p += wpc) {
			const uint32_t a = p == 0u ? 0u : lead + (p << 13);
			if (a >= r)
				break;
			copy_piece(src, dst, a, min(r, lead + ((p + 1u) << 13)), lane, 32u);
		}
	}
}

} /* namespace */

static inline uint32_t tiles_of(uint32_t n)
{
	return n ? (n + kScanThreads - 1u) / kScanThreads : 1u; /* (a grid of 0 CTAs is not a launch; the kernels check their index) */
}

extern "C" size_t airs_concat_scratch_bytes(uint32_t n_jobs, uint32_t n_results)
{
	/* tile sums of both scans (+ totals), the list of long streams and its counter */
	return 8 * ((size_t)tiles_of(n_jobs) + tiles_of(n_results) + 4) + 4 * ((size_t)n_results + 4) + 64;
}

extern "C" cudaError_t airs_launch_concat_slots(const AirsConcat *c, cudaStream_t stream)
{
	const uint32_t nt = tiles_of(c->n_jobs);
	uint64_t *sums = c->sums;
	concat_tile_sums_kernel<0><<<nt, kScanThreads, 0, stream>>>(*c, c->n_jobs, sums);
	concat_top_kernel<0><<<1, kScanThreads, 0, stream>>>(*c, nt, sums);
	concat_slots_kernel<<<nt, kScanThreads, 0, stream>>>(*c, sums);
	return cudaGetLastError();
}

extern "C" cudaError_t airs_launch_concat_gather(const AirsConcat *c, unsigned int grid, cudaStream_t stream)
{
	const uint32_t nt = tiles_of(c->n_results);
	uint64_t *sums = c->sums + tiles_of(c->n_jobs) + 2;
	concat_tile_sums_kernel<1><<<nt, kScanThreads, 0, stream>>>(*c, c->n_results, sums);
	concat_top_kernel<1><<<1, kScanThreads, 0, stream>>>(*c, nt, sums);
	concat_offsets_kernel<<<nt, kScanThreads, 0, stream>>>(*c, sums, nt);
	concat_gather_kernel<<<grid, 256, 0, stream>>>(*c);
	return cudaGetLastError();
}
