/*
 * airs_raw.cu - airs_raw_kernel: single frames under the UNCOMPRESSED encoder (ref cmp_encoder_encode_s16,
 * encoder.c:331-334: every residual as 16 raw bits), without preprocessing or with differences, whose slot
 * is large enough for the stream.  Nothing about such a stream depends on the data: sample i lies at byte
 * header + 2 i, so every 8-byte group of the output is made by one thread from the four samples it holds - a copy
 * with a byte swap at memory speed; frames are cut into chunks of 32 KiB of stream that are dealt round robin
 * over the CTAs (airs_plan_kernel numbers the chunks; a chunk finds its frame by bisection over the records).  (The CTA-per-job kernel
 * pushes the same samples through its bit packer: 0.38 of the roofline on 512 frames of 2 MiB.)
 *
 *   NONE + UNCOMPRESSED: 16 header bytes (ref cmp_hdr_serialize, header.c:24-67: no extended header), group g of
 *                        the stream = samples 4 (g - 2) .. 4 (g - 2) + 3, big endian.
 *   DIFF + UNCOMPRESSED: 22 header bytes, group g = the differences of samples 4 (g - 3) + 1 .. 4 (g - 2)
 *                        (ref preprocess.c:268-290; the first "difference" is the sample itself).
 * The groups that hold header bytes and the incomplete last group are written two bytes at a time by the first
 * threads of the frame's first CTA.  XXH32 trailers by the checksum kernels behind it, as for every stream.
 *
 * airs_plan_kernel lists the jobs (their records are FastJob records, kept in the front of the tile-extension
 * array, which the tile jobs fill from the back).  (Done by the CTAs of airs_fast_kernel behind their short jobs,
 * to save the launch, the call spilled that kernel's encoder: config 3 0.83 -> 0.91 ms.)
 */
#include <cuda_runtime.h>

#include "airs_fast.cuh"
#include "airs_launch.h"

namespace {

constexpr uint32_t kRawThreads = 256;
constexpr uint32_t kRawChunk = AIRS_RAW_CHUNK; /* 8-byte groups of a frame a CTA takes at a time (32 KiB of stream) */

__device__ __forceinline__ uint32_t be_pair(uint32_t w) /* two samples of a word -> their big-endian bytes in stream order */
{
	return __byte_perm(w, 0, 0x2301);
}

/* 16-bit unit u (bytes 2 u, 2 u + 1) of the stream behind the header: residual (u - hdr / 2), big endian, as stored */
__device__ __forceinline__ uint32_t unit_at(const uint16_t *x, uint32_t i, bool diff)
{
	const uint32_t v = __ldg(x + i), p = (diff && i) ? (uint32_t)__ldg(x + i - 1u) : 0u;
	const uint32_t r = (v - p) & 0xFFFFu;
	return (r >> 8) | ((r & 0xFFu) << 8);
}

/* the whole 8-byte groups [g_lo, g_hi) of the frame, group g by thread `thr` of `n_thr` */
__device__ __forceinline__ void raw_groups(const uint16_t *x, uint32_t n, uint8_t *dst, bool diff, uint32_t g_lo, uint32_t g_hi, uint32_t thr,
					   uint32_t n_thr)
{
	const uint2 *x2 = reinterpret_cast<const uint2 *>(x);
	uint2 *out = reinterpret_cast<uint2 *>(dst);
	const uint32_t lane = threadIdx.x & 31u;

	constexpr uint32_t kAhead = 4; /* groups a thread has in flight */
	if (!diff) { /* group g = quad g - 2 */
		for (uint32_t g = g_lo + thr; g < g_hi; g += kAhead * n_thr) {
			uint2 q[kAhead];
#pragma unroll
			for (uint32_t u = 0; u < kAhead; u++)
				q[u] = g + u * n_thr < g_hi ? __ldg(x2 + (g + u * n_thr - 2u)) : make_uint2(0, 0);
#pragma unroll
			for (uint32_t u = 0; u < kAhead; u++)
				if (g + u * n_thr < g_hi)
					out[g + u * n_thr] = make_uint2(be_pair(q[u].x), be_pair(q[u].y));
		}
		return;
	}
	/* group g = differences of samples 4 (g - 3) + 1 .. 4 (g - 2): quad g - 3 from the lane in front, one sample of quad g - 2 */
	for (uint32_t g0 = g_lo + thr - lane; g0 < g_hi; g0 += kAhead * n_thr) { /* whole warps, for the shuffles */
		uint2 own[kAhead], first[kAhead];
#pragma unroll
		for (uint32_t u = 0; u < kAhead; u++) {
			const uint32_t g = g0 + u * n_thr + lane;
			own[u] = first[u] = make_uint2(0, 0); /* (only the first sample of `own` is used: the quad may reach beyond the frame) */
			if (g < g_hi)
				own[u] = 4u * (g - 2u) + 3u < n ? __ldg(x2 + (g - 2u)) : make_uint2(__ldg(x + 4u * (g - 2u)), 0u);
			if (lane == 0 && g < g_hi)
				first[u] = __ldg(x2 + (g - 3u)); /* (g >= 3: the groups in front hold header bytes) */
		}
#pragma unroll
		for (uint32_t u = 0; u < kAhead; u++) {
			const uint32_t g = g0 + u * n_thr + lane;
			uint2 prev = make_uint2(__shfl_up_sync(0xFFFFFFFFu, own[u].x, 1), __shfl_up_sync(0xFFFFFFFFu, own[u].y, 1));
			if (lane == 0)
				prev = first[u];
			if (g < g_hi) {
				/* samples a0 a1 a2 a3 (quad g - 3) and b0 (quad g - 2): a1 - a0, a2 - a1, a3 - a2, b0 - a3 */
				const uint32_t lo = __vsub2(__byte_perm(prev.x, prev.y, 0x5432), prev.x);     /* (a1 - a0) | (a2 - a1) << 16 */
				const uint32_t hi = __vsub2(__byte_perm(prev.y, own[u].x, 0x5432), prev.y);   /* (a3 - a2) | (b0 - a3) << 16 */
				out[g] = make_uint2(be_pair(lo), be_pair(hi));
			}
		}
	}
}

} /* namespace */

__global__ void __launch_bounds__(kRawThreads) airs_raw_kernel(AirsLaunch b)
{
	if (b.ticket[AIRS_TICKET_INVALID] || (b.gate && (*b.gate != 0u) != (b.gate_want != 0u)))
		return;
	/* airs_plan_kernel: jobs << 40 | chunks of kRawChunk groups; the jobs' records in the order of their first chunks */
	const uint64_t counts = *reinterpret_cast<const uint64_t *>(b.ticket + AIRS_TICKET_RAW);
	const uint32_t n_chunks = (uint32_t)(counts & ((1ull << 40) - 1u)), n_raw = (uint32_t)(counts >> 40);
	if (n_chunks == 0u)
		return;
	const FastJob *recs = reinterpret_cast<const FastJob *>(b.tile_ext);
	uint32_t k = 0xFFFFFFFFu, base = 0, cnt = 0; /* the job of the chunk before */

	for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
		if (k == 0xFFFFFFFFu || c < base || c >= base + cnt) { /* (uniform over the CTA) the last record whose first chunk is <= c */
			/* a search by the whole CTA: 256 probes a round (the first chunks of the records ascend) */
			uint32_t lo = 0, hi = n_raw; /* invariant: recs[lo].tile_base <= c < recs[hi].tile_base (hi = n_raw: none) */
			while (hi - lo > 1u) {
				const uint32_t span = hi - lo;
				auto probe_at = [&](uint32_t t) -> uint32_t { /* probe t of this round, ascending in t; >= hi: none */
					return span > kRawThreads ? lo + (uint32_t)(((uint64_t)span * (t + 1u)) / (kRawThreads + 1u)) : lo + t + 1u;
				};
				const uint32_t idx = probe_at(threadIdx.x);
				const uint32_t cnt = (uint32_t)__syncthreads_count(idx < hi && __ldg(&recs[idx].tile_base) <= c);
				const uint32_t nlo = cnt ? probe_at(cnt - 1u) : lo, nhi = (cnt < kRawThreads && probe_at(cnt) < hi) ? probe_at(cnt) : hi;
				lo = nlo;
				hi = nhi;
			}
			k = lo;
			base = __ldg(&recs[k].tile_base);
			cnt = __ldg(&recs[k].n_tiles);
		}
		const FastJob &r = recs[k];
		const uint32_t n = r.n, flags = r.flags;
		const bool diff = (flags & AIRS_FJ_PRE_DIFF) != 0u;
		const uint32_t hdr = diff ? CMP_HDR_SIZE + 6u : CMP_HDR_SIZE, end = hdr + 2u * n; /* bytes in front of the trailer */
		const uint32_t g_lo = (hdr + 7u) / 8u, g_hi = end / 8u;
		const uint16_t *x = reinterpret_cast<const uint16_t *>((uintptr_t)r.src);
		uint8_t *dst = reinterpret_cast<uint8_t *>((uintptr_t)r.dst);
		const uint32_t ci = c - base; /* this chunk of the frame: the groups [ci, ci + 1) * kRawChunk, from g_lo on */
		const uint32_t c_lo = max(g_lo, ci * kRawChunk), c_hi = min(g_hi, (ci + 1u) * kRawChunk);
		if (c_hi > c_lo)
			raw_groups(x, n, dst, diff, c_lo, c_hi, threadIdx.x, kRawThreads);
		if (ci != 0u)
			continue;
		/* the frame's first chunk: header, the 16-bit units next to it and behind the last whole group, the result */
		const uint32_t checksum = (flags & AIRS_FJ_CHECKSUM) ? 1u : 0u, size = end + 4u * checksum;
		if (threadIdx.x < hdr) {
			const uint64_t id = r.identifier;
			uint32_t v;
			switch (threadIdx.x) {
			case 0: v = 0x80u | (CMP_VERSION_NUMBER >> 8); break;
			case 1: v = CMP_VERSION_NUMBER & 0xFFu; break;
			case 2: v = size >> 16; break;
			case 3: v = size >> 8; break;
			case 4: v = size; break;
			case 5: v = (2u * n) >> 16; break;
			case 6: v = (2u * n) >> 8; break;
			case 7: v = 2u * n; break;
			case 8: case 9: case 10: case 11: case 12: case 13: v = (uint32_t)(id >> (8u * (13u - threadIdx.x))); break;
			case 14: v = 0; break; /* sequence number of a fresh context */
			case 15: v = ((diff ? CMP_PREPROCESS_DIFF : CMP_PREPROCESS_NONE) << 4) | (checksum << 3) | CMP_ENCODER_UNCOMPRESSED; break;
			default: v = 0; break; /* extended header of an uncompressed pass: no model, no encoder parameters */
			}
			dst[threadIdx.x] = (uint8_t)v;
		}
		{ /* units in [hdr, 8 g_lo) and in [8 max(g_hi, g_lo), end): at most 3 + 3 */
			const uint32_t head_end = min(8u * g_lo, end), tail_begin = max(8u * g_hi, head_end);
			const uint32_t n_head = (head_end - hdr) / 2u, n_tail = (end - tail_begin) / 2u;
			if (threadIdx.x >= 32u && threadIdx.x < 32u + n_head + n_tail) {
				const uint32_t t = threadIdx.x - 32u;
				const uint32_t byte = t < n_head ? hdr + 2u * t : tail_begin + 2u * (t - n_head);
				*reinterpret_cast<uint16_t *>(dst + byte) = (uint16_t)unit_at(x, (byte - hdr) / 2u, diff);
			}
		}
		if (threadIdx.x == 0)
			b.results[r.first_result] = size;
	}
}

extern "C" cudaError_t airs_launch_raw(const AirsLaunch *b, unsigned int grid, cudaStream_t stream)
{
	airs_raw_kernel<<<grid, kRawThreads, 0, stream>>>(*b);
	return cudaGetLastError();
}
