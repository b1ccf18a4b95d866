/*
 * airs_iwt.cu - the multi-level 5/3 integer wavelet transform of whole frames, in front of the warp encoders.
 *
 * Reference: iwt_multi_level_decomposition_i16 / iwt_single_level_i16, lib/compress/preprocess.c:140-221 - one
 * sequential in-place sweep per level (stride 1, 2, 4, .. < n): the odd multiples of the stride become details
 * d[i] = x[i] - floor((x[i - s] + x[i + s]) / 2), the even ones approximations a[i] = x[i] + floor((d[i - s] + d[i + s]) / 4),
 * with one-sided forms at the edges, every store truncated to 16 bits.  Level l + 1 only reads the approximations of
 * level l, so a coefficient of level L depends on the samples at most 2 (2^L - 1) positions away.
 *
 * airs_plan_kernel lists the frames (IwtRec) and numbers their tiles of AIRS_IWT_TILE samples over the whole batch.
 *   airs_iwt_kernel:      every CTA takes a run of consecutive tiles, each with a halo of 512 samples on either side
 *                         (recomputed, never exchanged: 510 positions is as far as eight levels look).  Levels 1-3
 *                         run in registers - a lane holds 8 consecutive samples, neighbours by shuffle - and give seven
 *                         of eight coefficients; levels 4-6 run the same way on the approximations at every 8th
 *                         position (shared memory), levels 7-8 as sweeps over every 64th; the tile's coefficients
 *                         leave as 16-byte stores to the work buffer.
 *   airs_iwt_tail_kernel: what is left - levels 9 and up act on every 256th coefficient only, at most 8192 of them for the
 *                         frames this path takes (AIRS_IWT_MAX_SAMPLES) - by one CTA per frame, in shared memory at once.
 * The sweeps in shared memory work on 16-bit elements with one pad word behind every 32 words, so that the lanes of
 * every level (lane k at element 2 s k) fall on 32 different banks.
 * The coefficients lie in the work buffer as the reference leaves them; airs_fast_kernel / airs_tile_kernel code them
 * from there like samples without preprocessing (AIRS_FJ_IWT).
 */
#include <cuda_runtime.h>

#include "airs_device.cuh"
#include "airs_fast.cuh"
#include "airs_launch.h"

namespace {

constexpr uint32_t kThreads = 256;
constexpr uint32_t kTile = AIRS_IWT_TILE;
constexpr uint32_t kHalo = 512;          /* >= 2 (2^8 - 1) and a multiple of 2 * 128 */
constexpr uint32_t kLevels = 8;          /* levels of the tiled kernel: strides 1 .. 128 */
constexpr uint32_t kTailStride = 1u << kLevels;
constexpr uint32_t kBuf = kTile + 2u * kHalo;
constexpr uint32_t kBufWords = kBuf / 2u + kBuf / 64u + 2u;
static_assert(AIRS_IWT_MAX_SAMPLES / kTailStride <= kBuf, "the tail of the longest frame fits the buffer");
static_assert(kHalo >= 2u * ((1u << kLevels) - 1u) && kHalo % (2u << (kLevels - 1u)) == 0, "halo");

/* halfword index of element e in the padded buffer */
__device__ __forceinline__ uint32_t phys(uint32_t e)
{
	return e + ((e >> 6) << 1);
}

/* One level at stride s over the elements lo <= i < hi of a sequence of n elements, element i at hb[phys(i - lo)];
 * lo is a multiple of 2 s.  Elements whose neighbours lie outside [lo, hi) but inside the sequence are left alone
 * (they belong to the halo).  All details first (they read untouched even multiples), then all approximations. */
__device__ __forceinline__ void level_sweep(int16_t *hb, uint32_t s, uint32_t lo, uint32_t hi, uint32_t n)
{
	const uint32_t s2 = 2u * s;

	for (uint32_t i = lo + s + s2 * threadIdx.x; i < hi; i += s2 * kThreads) {
		const uint32_t e = i - lo;
		if (i + s < hi)
			hb[phys(e)] = (int16_t)(hb[phys(e)] - (((int32_t)hb[phys(e - s)] + hb[phys(e + s)]) >> 1));
		else if (i + s >= n) /* the last one has no right neighbour */
			hb[phys(e)] = (int16_t)(hb[phys(e)] - hb[phys(e - s)]);
	}
	__syncthreads();
	for (uint32_t i = lo + s2 * threadIdx.x; i < hi; i += s2 * kThreads) {
		const bool has_l = i >= s, has_r = i + s < n;
		if ((has_l && i - s < lo) || (has_r && i + s >= hi))
			continue;
		const uint32_t e = i - lo;
		int32_t t = 0;
		if (has_l && has_r)
			t = ((int32_t)hb[phys(e - s)] + hb[phys(e + s)]) >> 2;
		else if (has_r)
			t = (int32_t)hb[phys(e + s)] >> 1;
		else if (has_l)
			t = (int32_t)hb[phys(e - s)] >> 1;
		hb[phys(e)] = (int16_t)(hb[phys(e)] + t);
	}
	__syncthreads();
}

__device__ __forceinline__ bool gate_closed(const AirsLaunch &b)
{
	return b.ticket[AIRS_TICKET_INVALID] != 0u || (b.gate && (*b.gate != 0u) != (b.gate_want != 0u));
}

__device__ __forceinline__ int32_t w16(int32_t v) /* what a store to an int16_t keeps */
{
	return (int32_t)(int16_t)v;
}

/* the incomplete last pack of a frame: cnt (1..7) samples from sample `first` on, the rest 0 */
__device__ __noinline__ uint4 load_partial_pack(const uint16_t *s16, uint32_t first, uint32_t cnt)
{
	uint32_t w[4] = {0, 0, 0, 0};
	for (uint32_t i = 0; i < cnt; i++)
		w[i >> 1] |= (uint32_t)__ldg(s16 + first + i) << (16u * (i & 1u));
	return make_uint4(w[0], w[1], w[2], w[3]);
}

/*
 * Levels 1-3 of one pack (8 consecutive samples from position i0 = 8 p on) in registers, the packs of a warp
 * side by side over its lanes.  What a level needs of the neighbouring packs comes by shuffle (one per level
 * and direction); lanes 0, 1 and 31 end up with values that depend on packs the warp does not hold - the caller
 * uses lanes 2-29.  EDGE: the pack or one of its neighbours touches an end of the frame (one-sided forms).
 * Returns the level-3 approximation of position i0; out = the pack's coefficients with position i0 left zero.
 */
template <bool EDGE>
__device__ __forceinline__ int32_t pack_levels(const uint4 v, int32_t i0, int32_t n, uint4 &out)
{
	int32_t x[8];
	x[0] = (int32_t)(v.x << 16) >> 16; x[1] = (int32_t)v.x >> 16;
	x[2] = (int32_t)(v.y << 16) >> 16; x[3] = (int32_t)v.y >> 16;
	x[4] = (int32_t)(v.z << 16) >> 16; x[5] = (int32_t)v.z >> 16;
	x[6] = (int32_t)(v.w << 16) >> 16; x[7] = (int32_t)v.w >> 16;
	const uint32_t full = 0xFFFFFFFFu;
	/* detail at offset k from its two neighbours l, r at distance s (r missing beyond the frame) */
#define AIRS_DETAIL(c_, l_, r_, k_, s_) ((!EDGE || i0 + (k_) + (s_) < n) ? w16((c_) - (((l_) + (r_)) >> 1)) : w16((c_) - (l_)))
	/* approximation at offset k from the details at distance s on either side */
#define AIRS_APPROX(c_, l_, r_, k_, s_)                                                                           \
	(!EDGE ? w16((c_) + (((l_) + (r_)) >> 2))                                                                  \
	       : w16((c_) + ((i0 + (k_) >= (s_)) ? ((i0 + (k_) + (s_) < n) ? ((l_) + (r_)) >> 2 : (l_) >> 1)       \
						  : ((i0 + (k_) + (s_) < n) ? (r_) >> 1 : 0))))
	const int32_t x8 = __shfl_down_sync(full, x[0], 1);
	const int32_t d1_1 = AIRS_DETAIL(x[1], x[0], x[2], 1, 1), d1_3 = AIRS_DETAIL(x[3], x[2], x[4], 3, 1);
	const int32_t d1_5 = AIRS_DETAIL(x[5], x[4], x[6], 5, 1), d1_7 = AIRS_DETAIL(x[7], x[6], x8, 7, 1);
	const int32_t d1_m1 = __shfl_up_sync(full, d1_7, 1);
	const int32_t a1_0 = AIRS_APPROX(x[0], d1_m1, d1_1, 0, 1), a1_2 = AIRS_APPROX(x[2], d1_1, d1_3, 2, 1);
	const int32_t a1_4 = AIRS_APPROX(x[4], d1_3, d1_5, 4, 1), a1_6 = AIRS_APPROX(x[6], d1_5, d1_7, 6, 1);
	const int32_t a1_8 = __shfl_down_sync(full, a1_0, 1);
	const int32_t d2_2 = AIRS_DETAIL(a1_2, a1_0, a1_4, 2, 2), d2_6 = AIRS_DETAIL(a1_6, a1_4, a1_8, 6, 2);
	const int32_t d2_m2 = __shfl_up_sync(full, d2_6, 1);
	const int32_t a2_0 = AIRS_APPROX(a1_0, d2_m2, d2_2, 0, 2), a2_4 = AIRS_APPROX(a1_4, d2_2, d2_6, 4, 2);
	const int32_t a2_8 = __shfl_down_sync(full, a2_0, 1);
	const int32_t d3_4 = AIRS_DETAIL(a2_4, a2_0, a2_8, 4, 4);
	const int32_t d3_m4 = __shfl_up_sync(full, d3_4, 1);
	const int32_t a3_0 = AIRS_APPROX(a2_0, d3_m4, d3_4, 0, 4);
#undef AIRS_DETAIL
#undef AIRS_APPROX
	out.x = (uint32_t)d1_1 << 16;
	out.y = ((uint32_t)d2_2 & 0xFFFFu) | ((uint32_t)d1_3 << 16);
	out.z = ((uint32_t)d3_4 & 0xFFFFu) | ((uint32_t)d1_5 << 16);
	out.w = ((uint32_t)d2_6 & 0xFFFFu) | ((uint32_t)d1_7 << 16);
	return a3_0;
}

constexpr uint32_t kPackLanes = 28; /* packs a warp finishes per step: lanes 2 .. 29 */
constexpr uint32_t kRegLevels = 3;  /* levels in registers; the others act on every 8th coefficient, in shared memory */
constexpr uint32_t kApxWords = kBuf / 16u + kBuf / 512u + 8u; /* (a whole pack behind an incomplete last one) */

__global__ void __launch_bounds__(kThreads) airs_iwt_kernel(AirsLaunch b)
{
	__shared__ uint4 outb[kTile / 8u];    /* the tile's coefficients, pack by pack */
	__shared__ uint32_t apx[kApxWords];   /* level-3 approximations of the tile and its halos (every 8th position), padded */
	__shared__ uint32_t s_first;
	int16_t *hb = reinterpret_cast<int16_t *>(apx);
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

	if (gate_closed(b))
		return;
	const uint64_t counts = *reinterpret_cast<const uint64_t *>(b.ticket + AIRS_TICKET_IWT);
	const uint32_t n_recs = (uint32_t)(counts >> 40), n_tiles = (uint32_t)(counts & ((1ull << 40) - 1u));
	if (n_tiles == 0u)
		return;
	const IwtRec *recs = reinterpret_cast<const IwtRec *>(b.iwt_recs);
	/* this CTA's run of tiles; the record of its first one by bisection (tile_base ascends with the record index) */
	const uint32_t run = (n_tiles + gridDim.x - 1u) / gridDim.x;
	const uint32_t g0 = blockIdx.x * run, g1 = min(g0 + run, n_tiles);
	if (g0 >= g1)
		return;
	if (tid == 0) {
		uint32_t lo = 0, hi = n_recs; /* the last record with tile_base <= g0 */
		while (hi - lo > 1u) {
			const uint32_t mid = (lo + hi) / 2u;
			if (__ldg(&recs[mid].tile_base) <= g0)
				lo = mid;
			else
				hi = mid;
		}
		s_first = lo;
	}
	__syncthreads();
	uint32_t r = s_first;
	IwtRec rec = recs[r];

	for (uint32_t g = g0; g < g1; g++) {
		while (g >= rec.tile_base + rec.n_tiles) /* (every listed frame has at least one tile) */
			rec = recs[++r];
		const uint32_t n = rec.n;
		const uint32_t t0 = (g - rec.tile_base) * kTile, t1 = min(t0 + kTile, n);
		const uint32_t lo = t0 >= kHalo ? t0 - kHalo : 0u, hi = min(t0 + kTile + kHalo, n);
		/* in packs of 8 positions */
		const int32_t P_lo = (int32_t)(lo / 8u), P_hi = (int32_t)((hi + 7u) / 8u), P_n = (int32_t)((n + 7u) / 8u);
		const int32_t P_t0 = (int32_t)(t0 / 8u), P_t1 = (int32_t)((t1 + 7u) / 8u);
		const uint4 *src4 = reinterpret_cast<const uint4 *>((uintptr_t)rec.src);

		/* levels 1-3: warps over runs of 28 packs of [P_lo, P_hi), two more packs in front and two behind; the
		 * samples of a warp's next run travel while it works on one */
#define AIRS_LOAD_PACK(p_)                                                                                     \
	(((p_) >= 0 && 8 * (p_) + 8 <= (int32_t)n) ? __ldg(src4 + (p_))                                         \
	 : ((p_) >= 0 && 8 * (p_) < (int32_t)n)   ? load_partial_pack(reinterpret_cast<const uint16_t *>((uintptr_t)rec.src), 8u * (uint32_t)(p_), n - 8u * (uint32_t)(p_)) \
						    : make_uint4(0, 0, 0, 0))
		constexpr int32_t kStep = (int32_t)(kPackLanes * (kThreads / 32u));
		int32_t c = P_lo + (int32_t)(kPackLanes * warp);
		uint4 v = make_uint4(0, 0, 0, 0);
		if (c < P_hi)
			v = AIRS_LOAD_PACK(c - 2 + (int32_t)lane);
		for (; c < P_hi; c += kStep) {
			const int32_t p = c - 2 + (int32_t)lane;
			uint4 vn = make_uint4(0, 0, 0, 0);
			if (c + kStep < P_hi)
				vn = AIRS_LOAD_PACK(p + kStep);
			const bool edge = c - 2 <= 0 || 8 * (c + 30) + 8 > (int32_t)n; /* (for the whole warp) */
			uint4 o;
			const int32_t a3 = edge ? pack_levels<true>(v, 8 * p, (int32_t)n, o) : pack_levels<false>(v, 8 * p, (int32_t)n, o);
			if (lane >= 2u && lane < 2u + kPackLanes && p < P_hi) {
				hb[phys((uint32_t)(p - P_lo))] = (int16_t)a3;
				if (p >= P_t0 && p < P_t1)
					outb[p - P_t0] = o;
			}
			v = vn;
		}
#undef AIRS_LOAD_PACK
		__syncthreads();
		/* levels 4-6 the same way on the approximations: to y[k] = w[8 k] they are the levels 1-3.  At most 144 packs
		 * of 8: one run per warp, all of them read before any is written back in place */
		{
			const int32_t Q_lo = P_lo / 8, Q_hi = (P_hi + 7) / 8;
			const int32_t c2 = Q_lo + (int32_t)(kPackLanes * warp), q = c2 - 2 + (int32_t)lane;
			static_assert(kBuf / 64u <= kPackLanes * (kThreads / 32u), "one run of packs per warp");
			uint4 y = make_uint4(0, 0, 0, 0), o = y;
			int32_t a6 = 0;
			const bool have = c2 < Q_hi; /* (for the whole warp) */
			uint32_t *yw = apx + (phys((uint32_t)(8 * q - P_lo)) >> 1);
			const bool inside = have && q >= Q_lo && q < Q_hi;
			if (inside) /* (elements behind the region or the frame: never read by a coefficient that counts) */
				y = make_uint4(yw[0], yw[1], yw[2], yw[3]);
			if (have) {
				const bool edge = c2 - 2 <= 0 || 8 * (c2 + 30) + 8 > P_n;
				a6 = edge ? pack_levels<true>(y, 8 * q, P_n, o) : pack_levels<false>(y, 8 * q, P_n, o);
			}
			__syncthreads();
			if (inside && lane >= 2u && lane < 2u + kPackLanes) {
				yw[0] = o.x | ((uint32_t)a6 & 0xFFFFu);
				yw[1] = o.y;
				yw[2] = o.z;
				yw[3] = o.w;
			}
			__syncthreads();
		}
		/* levels 7-8 on every 64th position: sweeps in shared memory */
		for (uint32_t l = 2u * kRegLevels, s = 8; l < kLevels && 8u * s < n; l++, s <<= 1)
			level_sweep(hb, s, (uint32_t)P_lo, (uint32_t)P_hi, (uint32_t)P_n);
		for (uint32_t k = tid; k < (uint32_t)(P_t1 - P_t0); k += kThreads)
			reinterpret_cast<uint16_t *>(outb)[8u * k] = (uint16_t)hb[phys((uint32_t)(P_t0 - P_lo) + k)];
		__syncthreads();
		{
			uint4 *w4 = reinterpret_cast<uint4 *>((uintptr_t)rec.work) + P_t0;
			const uint32_t full_packs = (t1 - t0) / 8u;
			for (uint32_t k = tid; k < full_packs; k += kThreads)
				w4[k] = outb[k];
			if (tid < ((t1 - t0) & 7u)) /* the incomplete last pack of the frame */
				reinterpret_cast<uint16_t *>((uintptr_t)rec.work)[t0 + 8u * full_packs + tid] = reinterpret_cast<const uint16_t *>(outb)[8u * full_packs + tid];
		}
		__syncthreads();
	}
}

/* levels 9 and up of every listed frame longer than 256 samples: the decimated sequence y[k] = w[256 k] of
 * ceil(n / 256) elements goes through ALL its levels in shared memory (to it, the levels 1, 2, .. are what the
 * strides 256, 512, .. are to w, edges included) */
__global__ void __launch_bounds__(kThreads) airs_iwt_tail_kernel(AirsLaunch b)
{
	__shared__ uint32_t buf[kBufWords];
	int16_t *hb = reinterpret_cast<int16_t *>(buf);
	const uint32_t tid = threadIdx.x;

	if (gate_closed(b))
		return;
	const uint32_t n_recs = (uint32_t)(*reinterpret_cast<const uint64_t *>(b.ticket + AIRS_TICKET_IWT) >> 40);
	const IwtRec *recs = reinterpret_cast<const IwtRec *>(b.iwt_recs);

	for (uint32_t r = blockIdx.x; r < n_recs; r += gridDim.x) {
		const IwtRec rec = recs[r];
		if (rec.n <= kTailStride)
			continue;
		const uint32_t m = (rec.n + kTailStride - 1u) / kTailStride;
		int16_t *w = reinterpret_cast<int16_t *>((uintptr_t)rec.work);
		for (uint32_t k = tid; k < m; k += kThreads)
			hb[phys(k)] = w[(size_t)k * kTailStride];
		__syncthreads();
		for (uint32_t s = 1; s < m; s <<= 1)
			level_sweep(hb, s, 0u, m, m);
		for (uint32_t k = tid; k < m; k += kThreads)
			w[(size_t)k * kTailStride] = hb[phys(k)];
		__syncthreads();
	}
}

} /* namespace */

extern "C" cudaError_t airs_launch_iwt(const AirsLaunch *b, cudaStream_t stream)
{
	static thread_local int cached_dev = -1, sms = 0;
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (dev != cached_dev) {
		if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
			return e;
		cached_dev = dev;
	}
	/* (no CTA waits for another one: the grids need not be resident) */
	airs_iwt_kernel<<<(unsigned int)sms * 8u, kThreads, 0, stream>>>(*b);
	airs_iwt_tail_kernel<<<(unsigned int)sms * 4u, kThreads, 0, stream>>>(*b);
	return cudaGetLastError();
}
