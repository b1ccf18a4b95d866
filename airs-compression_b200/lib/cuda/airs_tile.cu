/*
 * airs_tile.cu - airs_tile_kernel: long single-frame jobs (chunks of more than 32768 samples without model:
 * BASELINE configs 1, 4 and 5 in their 2 MiB cut) spread over ALL resident warps, 1024 samples ("tile") at a
 * time, whatever the number of jobs - one 1 Mi-sample cmp_compress_u16() call uses the whole GPU.
 *
 * ONE WARP PER TILE, no block barrier anywhere (an earlier version gave a tile to a whole CTA and pipelined
 * three barriers per tile around a look-back warp: 45 % of the issue slots, barrier stalls on top of the list).
 * Tiles are numbered in stream order over all tile jobs and dealt out round robin over the warps of the grid
 * (all resident: every tile a warp waits for is held by a warp that is running and works on it in the same
 * iteration); a warp then works like airs_fast_kernel on a short job:
 *   two units of airs_fastcore.cuh (code words in registers, warp scan, strings OR-ed into the warp's own
 *   staging words) at TILE-LOCAL bit positions -> the tile's bit count and its last seven bits are published ->
 *   the tile's start in its stream = the bit counts of all tiles in front of it in its frame, summed over
 *   three levels of descriptors in global memory (look_back()) -> the staging words leave shifted to their
 *   place: five shared-memory words and four funnel shifts per 16-byte group, byte-swapped 128-bit stores.
 * The job of a tile is found two tiles ahead, its record and first samples one ahead, so the only thing a warp
 * ever waits for is the look-back - while the other 23 warps of its SM encode.
 * Seams: tiles meet at bit granularity.  A tile writes the bytes [start / 8, end / 8) of the stream; the
 * start % 8 bits of its first byte that belong to its predecessor come out of the "tails" ring.  The tile that
 * ends a frame also writes the last, zero-padded byte, the 22 header bytes (ref cmp_hdr_serialize,
 * header.c:24-67; the first tile starts behind them) and the result.
 *
 * Reference being replaced: the per-sample loop of compress_engine (cmp.c:296-312) with
 * bitstream_add_bits32 (bitstream_writer.h:124-158), for one frame by many warps.
 */
#include <cuda_runtime.h>

#include "airs_fastcore.cuh"
#include "airs_launch.h"

namespace {

using namespace fastcore;

#ifndef AIRS_TILE_ABLATE
#define AIRS_TILE_ABLATE 0
#endif
#ifndef AIRS_TILE_NAP
#define AIRS_TILE_NAP 200 /* ns between two polls of a window that is not complete */
#endif
constexpr uint32_t kTWarps = AIRS_TILE_THREADS / 32;
constexpr uint32_t kTile = AIRS_TILE_SAMPLES;            /* 1024 samples */
constexpr uint32_t kTileUnits = kTile / kUnit;
constexpr uint32_t kTileWords = kTile * 48 / 32;         /* a tile at 48 bits per sample */
constexpr uint32_t kPad = 8;                             /* words in front of the area: word -1 takes the carried bits */
constexpr uint32_t kRing = AIRS_TILE_RING;               /* descriptors of the last kRing tiles (far more than are in flight) */
constexpr uint32_t kHdrBits = 8u * (CMP_HDR_SIZE + 6u);
static_assert(kTile % kUnit == 0 && kTileUnits >= 1, "a tile is whole units");

constexpr uint64_t kFlagA = 1ull << 62;
constexpr uint32_t kValBits = 28; /* bits of 1024 tiles: less than 2^26 */

/* the staging words of one warp: MSB-first 32-bit words of the tile under construction at tile-local bit
 * positions, all zero when idle; strings that end in word 0 or 1 OR zeros below the area (pad) */
struct TileArea {
	alignas(16) uint32_t pad[kPad];
	uint32_t stg[kTileWords + 8];
};
struct TileWarp {
	TileArea area[2]; /* [k % 2]: the tile being encoded, the tile waiting for its place */
};

__device__ __forceinline__ uint64_t ld_desc(const uint64_t *p)
{
	uint64_t v;
	asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_desc(uint64_t *p, uint64_t v)
{
	asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint64_t make_desc(uint32_t tag, uint32_t value)
{
	return kFlagA | ((uint64_t)(tag + 1u) << kValBits) | value;
}

/* a descriptor that has been published for entry `tag` of its ring? */
__device__ __forceinline__ bool desc_ready(uint64_t v, uint32_t tag)
{
	return (uint32_t)(v >> kValBits) == tag + 1u && (v >> 62) != 0;
}

/* one entry of a look-back window: polled until it is there, then kept */
struct Probe {
	const uint64_t *p; /* nullptr: this lane has no entry in the window, or has it already */
	uint64_t v;        /* the answer to the last request */
	uint32_t tag, val;
	__device__ __forceinline__ void set(const uint64_t *ring_, uint32_t mask, uint32_t idx, bool wanted)
	{
		p = wanted ? ring_ + (idx & mask) : nullptr;
		tag = idx;
		val = 0;
		v = 0;
	}
	__device__ __forceinline__ void request() /* (no waiting here: the answer is looked at by check()) */
	{
		if (p)
			v = ld_desc(p);
	}
	__device__ __forceinline__ void check()
	{
		if (p && desc_ready(v, tag)) {
			val = (uint32_t)v & ((1u << kValBits) - 1u);
			p = nullptr;
		}
	}
	__device__ __forceinline__ void poll()
	{
		request();
		check();
	}
};

/* the five windows of a tile's look-back (see look_back()) and the tail of the tile in front of it */
struct LookBack {
	Probe t_own, t_first, b_own, b_first, q_mid;
	const uint64_t *tail_p;
	uint64_t tail_v;
};

/* The holder of the last tile of a block of 32 (1024) tiles publishes the block's bit count: it polls the 31
 * entries in front of its own.  They belong to tiles that other warps count at about the same time (tickets
 * are drawn in order): nothing they wait for depends on this warp. */
static __device__ __noinline__ void publish_block_sums(uint64_t *ring, uint32_t T, uint32_t my_bits, uint32_t lane)
{
	uint64_t *l1 = ring + 2u * kRing, *l2 = l1 + kRing / 32u;
	const uint32_t B = T >> 5, Q = B >> 5;
	Probe pr;
	uint32_t spins = 0;
	pr.set(ring, kRing - 1u, 32u * B + lane, lane < 31u);
	for (;;) {
		pr.poll();
		if (__all_sync(kFull, pr.p == nullptr))
			break;
		__nanosleep(AIRS_TILE_NAP);
		if (++spins > (1u << 22))
			__trap();
	}
	const uint32_t block = __reduce_add_sync(kFull, pr.val) + my_bits;
	if (lane == 0)
		st_desc(l1 + (B & (kRing / 32u - 1u)), make_desc(B, block));
	if ((T & 1023u) != 1023u)
		return;
	pr.set(l1, kRing / 32u - 1u, 32u * Q + lane, lane < 31u);
	for (;;) {
		pr.poll();
		if (__all_sync(kFull, pr.p == nullptr))
			break;
		__nanosleep(AIRS_TILE_NAP);
		if (++spins > (1u << 22))
			__trap();
	}
	const uint32_t big = __reduce_add_sync(kFull, pr.val) + block;
	if (lane == 0)
		st_desc(l2 + (Q & (kRing / 1024u - 1u)), make_desc(Q, big));
}

/*
 * Where the tile dealt at position T (number tidx of its frame) starts in its stream, in bits behind the header.
 * All rings are indexed by DEALING POSITION: the tiles of a frame are consecutive in it, every tile a tile waits
 * for has a lower one, and the positions in flight lie within a few rows of each other (the rings wrap).
 *
 * All warps work on one "row" of consecutive tiles at the same time, so a look-back that walks from count to
 * count until it meets a finished prefix walks half a row (thousands of tiles, a round trip per window).
 * Instead the counts are summed over THREE LEVELS: tiles, blocks of 32 tiles (ring l1) and blocks of 1024 tiles
 * (ring l2); the sums of a block are published by the holder of its last tile right when that tile is counted
 * (publish_block_sums()).  A tile's start is at most five windows of 32 entries - tiles of its own block,
 * blocks of its own 1024-block, whole 1024-blocks, and the same again upwards from the frame's first tile -
 * whose entries do not depend on any look-back: all are requested at once and polled until they are there.
 */
__device__ __forceinline__ void look_back_begin(LookBack &lb, const uint64_t *ring, const uint64_t *tails, uint32_t T, uint32_t tidx,
						uint32_t lane)
{
	const uint64_t *l1 = ring + 2u * kRing, *l2 = l1 + kRing / 32u;
	const uint32_t T0 = T - tidx;
	const uint32_t B = T >> 5, B0 = T0 >> 5, Q = B >> 5, Q0 = B0 >> 5;
	const bool any = tidx != 0u;

	/* tiles of my block in front of me, in my frame */
	lb.t_own.set(ring, kRing - 1u, 32u * B + lane, any && lane < (T & 31u) && 32u * B + lane >= T0);
	/* the frame's first tiles, up to the end of their block (another block than mine) */
	lb.t_first.set(ring, kRing - 1u, T0 + lane, any && B > B0 && T0 + lane < 32u * (B0 + 1u));
	/* blocks of my 1024-block in front of mine, behind the frame's first block */
	lb.b_own.set(l1, kRing / 32u - 1u, 32u * Q + lane, any && lane < (B & 31u) && 32u * Q + lane > B0);
	/* the blocks behind the frame's first one, up to the end of their 1024-block (another one than mine) */
	lb.b_first.set(l1, kRing / 32u - 1u, B0 + 1u + lane, any && Q > Q0 && B0 + 1u + lane < 32u * (Q0 + 1u));
	/* whole 1024-blocks between the two */
	lb.q_mid.set(l2, kRing / 1024u - 1u, Q0 + 1u + lane, any && Q0 + 1u + lane < Q);
	lb.t_own.request();
	lb.t_first.request();
	lb.b_own.request();
	lb.b_first.request();
	lb.q_mid.request();
	lb.tail_p = any ? tails + ((T - 1u) & (kRing - 1u)) : nullptr;
	lb.tail_v = any ? ld_desc(lb.tail_p) : 0ull;
}

/* ... the answers; what is still missing is polled.  Returns the bits in front of the tile (behind the header) */
__device__ __forceinline__ uint32_t look_back_finish(LookBack &lb)
{
	uint32_t spins = 0;
	lb.t_own.check();
	lb.t_first.check();
	lb.b_own.check();
	lb.b_first.check();
	lb.q_mid.check();
	while (!__all_sync(kFull, lb.t_own.p == nullptr && lb.t_first.p == nullptr && lb.b_own.p == nullptr && lb.b_first.p == nullptr &&
					  lb.q_mid.p == nullptr)) {
		__nanosleep(AIRS_TILE_NAP); /* the issue slots of a spinning warp are taken from the warps that encode */
		lb.t_own.poll();
		lb.t_first.poll();
		lb.b_own.poll();
		lb.b_first.poll();
		lb.q_mid.poll();
		if (++spins > (1u << 22))
			__trap(); /* an entry that never arrives: fail the launch instead of hanging the device */
	}
	return __reduce_add_sync(kFull, lb.t_own.val + lb.t_first.val + lb.b_own.val + lb.b_first.val + lb.q_mid.val);
}

__device__ __forceinline__ uint4 ld_cg4(const uint4 *p) /* from the L2: the line may have been written by another SM */
{
	uint4 v;
	asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
	return v;
}

/* the units of one tile: code words, scans, strings staged from tile-local bit 0 on; returns the tile's bits.
 * model: the context's model (nullptr: none) - PRE = kPreModel takes the residuals against it and leaves the
 * updated model behind, init_model (primary pass of a context with model) leaves the samples behind */
template <bool MULTI, int PRE, bool BE>
__device__ __forceinline__ uint32_t tile_units(const Dbg &dbg, const FK &k, TileArea &ws, const uint8_t *src, uint32_t n, uint32_t first,
					       uint4 (&nx)[kRows], uint32_t front, uint32_t lane, uint4 *model, bool init_model, const ModelK &mk, bool be)
{
	const uint32_t stg_bit = 8u * (uint32_t)__cvta_generic_to_shared(ws.stg);
	const uint32_t n_whole = n / 8u;
	const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
	const uint4 zero4 = make_uint4(0, 0, 0, 0);
	uint32_t bits = 0;

#pragma unroll 1 /* (one copy of the unit per instantiation: six instantiations are inlined into the kernel) */
	for (uint32_t u = 0; u < kTileUnits; u++) {
		const uint32_t ufirst = first + u * kUnit;
		if (ufirst >= n)
			break;
		uint4 x[kRows], m[kRows];
#pragma unroll
		for (uint32_t j = 0; j < kRows; j++) {
			x[j] = nx[j];
			if (PRE == kPreModel) { /* (frames with a model are whole pieces) */
				const uint32_t p = ufirst / 8u + unit_piece(lane, j);
				m[j] = p < n_whole ? ld_cg4(model + p) : zero4;
			}
		}
		if (u + 1u < kTileUnits) { /* the next unit's samples travel while this one is encoded */
#pragma unroll
			for (uint32_t j = 0; j < kRows; j++) {
				const uint32_t p = (ufirst + kUnit) / 8u + unit_piece(lane, j);
				nx[j] = p < n_whole ? load_piece<BE>(src4, p, be) : zero4;
				if (p == n_whole && (n & 7u))
					nx[j] = load_partial_piece(reinterpret_cast<const uint16_t *>(src), 8u * p, n & 7u, BE && be);
			}
		}
		if (ufirst + kUnit <= n) {
			const uint32_t nv[kRows] = {8u, 8u};
			bits += encode_unit_pre<MULTI, PRE, false>(dbg, k, x, m, mk, front, nv, lane, stg_bit + bits);
		} else { /* the ragged end of the frame */
			uint32_t nv[kRows];
#pragma unroll
			for (uint32_t j = 0; j < kRows; j++) {
				const uint32_t p = ufirst + 8u * unit_piece(lane, j);
				nv[j] = p >= n ? 0u : min(8u, n - p);
			}
			bits += encode_unit_pre<MULTI, PRE, true>(dbg, k, x, m, mk, front, nv, lane, stg_bit + bits);
		}
		if (PRE == kPreModel || init_model) {
#pragma unroll
			for (uint32_t j = 0; j < kRows; j++) {
				const uint32_t p = ufirst / 8u + unit_piece(lane, j);
				if (p < n_whole)
					model[p] = PRE == kPreModel ? m[j] : x[j];
			}
		}
	}
	return bits;
}

/* what tile T of a job is: its frame and its place in it, the pass the frame takes (ref cmp.c:228-262: a primary
 * pass, then sec_iter secondary ones, again and again - nothing fails on this path, a frame that does not fit sends
 * its whole job to airs_encode_kernel) */
struct TileWhat {
	uint32_t f, tidx, phase, cyc;
};
template <bool FRAMES>
__device__ __forceinline__ TileWhat tile_what(uint32_t T, uint32_t rec, uint32_t lane)
{
	TileWhat w;
	const uint32_t jt = T - __shfl_sync(kFull, rec, 14);
	if (!FRAMES || __shfl_sync(kFull, rec, 22) <= 1u) { /* a single frame: no divisions */
		w.f = w.phase = w.cyc = 0;
		w.tidx = jt;
		return w;
	}
	const uint32_t tpf = __shfl_sync(kFull, rec, 23), s1 = __shfl_sync(kFull, rec, 28) + 1u;
	w.f = jt / tpf;
	w.tidx = jt - w.f * tpf;
	w.cyc = w.f / s1;
	w.phase = w.f - w.cyc * s1;
	return w;
}

} /* namespace */

/* FRAMES = false: a batch whose tile jobs are single frames (no model, no secondary passes: four instantiations of
 * the warp encoder, 96 registers, six CTAs per SM); FRAMES = true: contexts of several frames among them (six
 * instantiations, 128 registers, five CTAs per SM).  Both are launched; airs_plan_kernel leaves behind which one works. */
template <bool FRAMES, bool BE>
__global__ void __launch_bounds__(AIRS_TILE_THREADS, AIRS_TILE_CTAS_PER_SM) airs_tile_kernel(AirsLaunch b)
{
	__shared__ TileWarp wsh[kTWarps];
	const uint32_t lane = threadIdx.x & 31u;
	TileWarp &ws = wsh[threadIdx.x >> 5];

	if (b.ticket[AIRS_TICKET_INVALID] || (b.gate && (*b.gate != 0u) != (b.gate_want != 0u))) /* a bad job table; two-phase CONCAT: not the phase that runs */
		return;
	const uint64_t counts = *reinterpret_cast<const uint64_t *>(b.ticket + 10); /* airs_plan_kernel: tile jobs << 40 | tiles */
	const uint32_t n_tiles = (uint32_t)(counts & ((1ull << 40) - 1u)), n_tjobs = (uint32_t)(counts >> 40);
	if (n_tiles == 0u || (b.ticket[AIRS_TICKET_TILE_SHAPE + 3] > 1u) != FRAMES) /* (the largest number of frames of a tile job) */
		return;
	const FastJob *recs_end = reinterpret_cast<const FastJob *>(b.fast_jobs) + (b.n_jobs - 1u); /* slot s at recs_end - s */
	uint64_t *ring = b.tile_ring, *tails = b.tile_ring + kRing;

	for (uint32_t w = lane; w < 2u * (kPad + kTileWords + 8u); w += 32u)
		ws.area[0].pad[w] = 0;
	__syncwarp();
	Dbg dbg;
#ifdef AIRS_BOUNDS_CHECK
	dbg.lo = (uint32_t)__cvta_generic_to_shared(&ws);
	dbg.hi = dbg.lo + (uint32_t)sizeof(TileWarp);
#endif

	/* The job of a tile is found in steps spread over the tiles in front of it, so that no step waits for global
	 * memory: tickets are drawn three tiles ahead; the tile_base fields of the 32 jobs behind the job of the tile
	 * before (tile jobs sit at recs_end - slot with ascending tile_base, and a warp's tickets only grow) are
	 * requested two tiles ahead; the slot follows from a ballot one tile ahead, when the job's record and the
	 * tile's first samples are requested. */
	auto probe = [&](uint32_t slot, uint32_t T) -> uint32_t { /* tile_base of the 32 slots behind `slot` */
		const uint32_t s = slot + 1u + lane;
		return (T < n_tiles && s < n_tjobs) ? __ldg(&(recs_end - s)->tile_base) : 0xFFFFFFFFu;
	};
	auto resolve = [&](uint32_t &slot, uint32_t T, uint32_t first) { /* the slot of tile T, first probe given */
		if (T >= n_tiles)
			return;
		uint32_t v = first;
		for (;;) {
			const uint32_t c = (uint32_t)__popc(__ballot_sync(kFull, v <= T));
			slot += c;
			if (c < 32u)
				break;
			v = probe(slot, T); /* a jump over more than 32 jobs: keep probing */
		}
	};
	const TileExt *ext_end = reinterpret_cast<const TileExt *>(b.tile_ext) + (b.n_jobs - 1u);
	auto record = [&](uint32_t slot, uint32_t T) -> uint32_t { /* word `lane` of the record of tile T's job: FastJob, TileExt */
		if (T >= n_tiles)
			return 0u;
		return lane < 16u ? __ldg(reinterpret_cast<const uint32_t *>(recs_end - slot) + lane)
				  : __ldg(reinterpret_cast<const uint32_t *>(ext_end - slot) + (lane - 16u));
	};
	const uint4 zero4 = make_uint4(0, 0, 0, 0);
	uint4 nx[kRows] = {zero4, zero4};
	uint32_t nfront = 0;
	/* the samples of the first unit of tile T (job record in `rec`), and the sample in front of the tile */
	auto request = [&](uint32_t rec, uint32_t T) {
		if (T >= n_tiles)
			return;
		const TileWhat tw = tile_what<FRAMES>(T, rec, lane);
		const uint8_t *src = reinterpret_cast<const uint8_t *>((uintptr_t)(__shfl_sync(kFull, rec, 0) | (uint64_t)__shfl_sync(kFull, rec, 1) << 32)) +
				     (uint64_t)tw.f * (__shfl_sync(kFull, rec, 16) | (uint64_t)__shfl_sync(kFull, rec, 17) << 32);
		const uint32_t n = __shfl_sync(kFull, rec, 6), first = tw.tidx * kTile, n_whole = n / 8u;
		const bool be = BE && (__shfl_sync(kFull, rec, 8) & AIRS_FJ_BE) != 0u;
		const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
#pragma unroll
		for (uint32_t j = 0; j < kRows; j++) {
			const uint32_t p = first / 8u + unit_piece(lane, j);
			nx[j] = p < n_whole ? load_piece<BE>(src4, p, be) : zero4;
			if (p == n_whole && (n & 7u))
				nx[j] = load_partial_piece(reinterpret_cast<const uint16_t *>(src), 8u * p, n & 7u, BE && be);
		}
		/* lane 0: the sample in front of the tile (DIFF), in the upper half of a word */
		nfront = 0;
		if (lane == 0 && first != 0u && first < n) {
			const uint32_t v = __ldg(reinterpret_cast<const uint16_t *>(src) + first - 1u);
			nfront = (be ? ((v << 8) | (v >> 8)) & 0xFFFFu : v) << 16;
		}
	};

	/* Tiles are dealt out round robin over all warps of the grid (all resident): tile T belongs to warp T mod
	 * n_warps in its iteration T / n_warps.  A counter would hand tiles to whoever asks first - but a tile has to be
	 * asked for a few tiles ahead to hide the way to its samples, and a tile that sits reserved while its warp
	 * is still busy with earlier ones keeps EVERY tile behind it in the frame waiting (measured: 7 x slower).
	 *
	 * The order in which tiles are dealt is the order of their numbers - unless all tile jobs have the same number of
	 * frames and of tiles per frame (airs_plan_kernel leaves minima and maxima behind).  Then the tiles are dealt
	 * FRAME BY FRAME over all contexts (frame 0 of every context, then frame 1 of every context, ...): a tile of a
	 * frame with model waits for the same tile of the frame before, so of one context only a frame or two are in
	 * work at a time, and a batch of contexts dealt context by context would be encoded one context after the other.
	 * Either way every tile a tile waits for - the tiles in front of it in its frame, the same tile of the frame
	 * before - is dealt before it.  In the uniform order the job of a tile follows from its number by arithmetic.
	 * prologue: tile 0 in full, tile 1 probed */
	const uint32_t n_warps = gridDim.x * kTWarps, gw = blockIdx.x * kTWarps + (threadIdx.x >> 5);
	const uint32_t u_tpf = b.ticket[AIRS_TICKET_TILE_SHAPE + 1], u_frames = b.ticket[AIRS_TICKET_TILE_SHAPE + 3];
	const uint32_t u_row = n_tjobs * u_tpf, u_job = u_frames * u_tpf;
	/* (a frame with model waits for the count a row before it: the row has to fit the ring several times) */
	const bool uniform = ~b.ticket[AIRS_TICKET_TILE_SHAPE] == u_tpf && ~b.ticket[AIRS_TICKET_TILE_SHAPE + 2] == u_frames &&
			     (uint64_t)n_tjobs * u_tpf <= kRing / 4u;
	/* the d-th tile that is dealt: its number and (uniform order) the slot of its job */
	auto dealt = [&](uint32_t d, uint32_t &slot) -> uint32_t {
		if (!uniform || d >= n_tiles)
			return d;
		const uint32_t f = d / u_row, rem = d - f * u_row, sl = rem / u_tpf;
		slot = sl;
		return sl * u_job + f * u_tpf + (rem - sl * u_tpf);
	};
	uint32_t d0 = gw, d1 = gw + n_warps, d2 = gw + 2u * n_warps; /* dealing positions of this warp's next tiles */
	uint32_t slot0 = 0xFFFFFFFFu; /* "the slot in front of slot 0" */
	uint32_t t0 = dealt(d0, slot0);
	if (!uniform)
		resolve(slot0, t0, probe(slot0, t0));
	uint32_t rec = record(slot0, t0);
	uint32_t slot1 = slot0;
	uint32_t t1 = dealt(d1, slot1);
	uint32_t fb = uniform ? 0u : probe(slot1, t1); /* candidates for tile 1 */
	request(rec, t0);

	/* Software pipeline, iteration k: tile k is encoded into staging area k % 2 and counted; then tile k - 1, counted
	 * one iteration ago - the tiles in front of it have had a tile's time to be counted as well - gets its place and
	 * leaves area (k - 1) % 2.  A warp that waited for its place right behind its own code words would wait for the
	 * slowest warp of the device in every iteration, together with all other warps of its SM. */
	uint32_t prec = 0, pT = 0xFFFFFFFFu, pD = 0, pbits = 0; /* the pending tile: its job's record, its number, its dealing position, its bits */
	for (uint32_t par = 0;; par ^= 1u) {
		const bool have_cur = d0 < n_tiles, have_pend = pT != 0xFFFFFFFFu;
		if (!have_cur && !have_pend)
			break;
		uint32_t bits = 0;
		uint32_t nrec = 0;
		if (have_cur) {
			/* ---- requests whose answers are looked at behind this tile's code words */
			if (!uniform)
				resolve(slot1, t1, fb);           /* job of the next tile */
			nrec = record(slot1, t1);
			if (!uniform)
				fb = probe(slot1, d2);            /* candidates for the tile behind it */

#define AIRS_REC(i) __shfl_sync(kFull, rec, (i))
			const uint32_t T = t0;
			const TileWhat tw = tile_what<FRAMES>(T, rec, lane);
			const uint32_t tidx = tw.tidx;
			const bool sec = FRAMES && tw.phase != 0u; /* a secondary pass */
			const uint8_t *src = reinterpret_cast<const uint8_t *>((uintptr_t)(AIRS_REC(0) | (uint64_t)AIRS_REC(1) << 32)) +
					     (uint64_t)tw.f * (AIRS_REC(16) | (uint64_t)AIRS_REC(17) << 32);
			const uint32_t n = AIRS_REC(6), flags = AIRS_REC(8), flags2 = AIRS_REC(27), tpf = AIRS_REC(23);
			const uint32_t g = sec ? AIRS_REC(24) : AIRS_REC(10), outlier = sec ? AIRS_REC(25) : AIRS_REC(11);
			const uint32_t magic = sec ? AIRS_REC(26) : AIRS_REC(12), L = ((sec ? flags2 : flags) >> 8) & 15u;
			uint4 *model = reinterpret_cast<uint4 *>((uintptr_t)(AIRS_REC(20) | (uint64_t)AIRS_REC(21) << 32));
#undef AIRS_REC
			const bool multi = sec ? (flags2 & AIRS_TX_MULTI2) != 0u : (flags & AIRS_FJ_MULTI) != 0u;
			const int pre = sec ? ((flags2 & AIRS_TX_PRE2_MODEL) ? kPreModel : (flags2 & AIRS_TX_PRE2_DIFF) ? kPreDiff : kPreNone)
					    : ((flags & AIRS_FJ_PRE_DIFF) ? kPreDiff : kPreNone);
			const bool has_model = FRAMES && (flags2 & AIRS_TX_MODEL) != 0u;
			const FK kk = make_fk(multi, g, L, outlier, magic);
			const ModelK mk = make_model_k((flags2 >> 16) & 31u, (flags2 & AIRS_TX_SIGNED) != 0u);
			const uint32_t first = tidx * kTile;
			const bool last = first + kTile >= n;
			TileArea &ar = ws.area[par];

			/* ---- a context with model: the tile's slice of the model is what the same tile of the frame before left
			 * (its count is published behind its model stores) */
			if (has_model && tw.f != 0u) {
				const uint32_t Tp = d0 - (uniform ? u_row : tpf); /* (its dealing position) */
				const uint64_t *p = ring + (Tp & (kRing - 1u));
				uint32_t spins = 0;
				while (!desc_ready(ld_desc(p), Tp)) {
					__nanosleep(AIRS_TILE_NAP);
					if (++spins > (1u << 22))
						__trap();
				}
			}

			/* ---- the tile's code words, staged at tile-local bit positions */
			const bool init_model = has_model && !sec;
			const bool be = BE && (flags & AIRS_FJ_BE) != 0u;
			if (multi) {
				if (FRAMES && pre == kPreModel)
					bits = tile_units<true, kPreModel, BE>(dbg, kk, ar, src, n, first, nx, nfront, lane, model, false, mk, be);
				else if (pre == kPreDiff)
					bits = tile_units<true, kPreDiff, BE>(dbg, kk, ar, src, n, first, nx, nfront, lane, model, init_model, mk, be);
				else
					bits = tile_units<true, kPreNone, BE>(dbg, kk, ar, src, n, first, nx, nfront, lane, model, init_model, mk, be);
			} else {
				if (FRAMES && pre == kPreModel)
					bits = tile_units<false, kPreModel, BE>(dbg, kk, ar, src, n, first, nx, nfront, lane, model, false, mk, be);
				else if (pre == kPreDiff)
					bits = tile_units<false, kPreDiff, BE>(dbg, kk, ar, src, n, first, nx, nfront, lane, model, init_model, mk, be);
				else
					bits = tile_units<false, kPreNone, BE>(dbg, kk, ar, src, n, first, nx, nfront, lane, model, init_model, mk, be);
			}
			if (has_model)
				__threadfence(); /* the model stores of all lanes in front of the count */
			__syncwarp();
			if (lane == 0) {
				st_desc(ring + (d0 & (kRing - 1u)), make_desc(d0, bits));
				if (!last) { /* the last 7 bits of the tile, for the tile behind it (bits >= 1024 > 7) */
					const uint32_t wi = bits >> 5, sft = bits & 31u;
					const uint32_t before = wi ? ar.stg[wi - 1u] : 0u;
					const uint32_t last32 = sft ? __funnelshift_l(ar.stg[wi], before, sft) : before; /* the 32 bits that end at bit `bits` */
					st_desc(tails + (d0 & (kRing - 1u)), ((uint64_t)(d0 + 1u) << 8) | (last32 & 0x7Fu));
				}
			}
			if ((d0 & 31u) == 31u && AIRS_TILE_ABLATE < 1)
				publish_block_sums(ring, d0, bits, lane);
			/* ---- the next tile's first samples travel while the tile before this one leaves */
			request(nrec, t1);
		}

		if (have_pend) {
#define AIRS_REC(i) __shfl_sync(kFull, prec, (i))
			const uint32_t T = pT;
			const TileWhat tw = tile_what<FRAMES>(T, prec, lane);
			const uint32_t tidx = tw.tidx;
			const bool sec = FRAMES && tw.phase != 0u;
			uint8_t *dst = reinterpret_cast<uint8_t *>((uintptr_t)(AIRS_REC(2) | (uint64_t)AIRS_REC(3) << 32)) +
				       (uint64_t)tw.f * (AIRS_REC(18) | (uint64_t)AIRS_REC(19) << 32);
			const uint32_t n = AIRS_REC(6), cap_eff = AIRS_REC(7), flags = AIRS_REC(8), flags2 = AIRS_REC(27);
			const bool last = tidx * kTile + kTile >= n;
			uint32_t *stg = ws.area[par ^ 1u].stg;

			/* ---- where the tile starts; the bits of its first byte that belong to the tile in front */
#if AIRS_TILE_ABLATE >= 1 /* development: timing without the look-back (wrong streams) */
			const uint32_t excl = kHdrBits + tidx * 4096u;
#else
			/* (requesting the windows in front of the tile's code words and looking at them behind was measured: the 17
			 * registers the answers sit in spill the encoder: 1.33 -> 1.82 ms on 256 x 4 MiB) */
			LookBack lb;
			look_back_begin(lb, ring, tails, pD, tidx, lane);
			const uint32_t excl = kHdrBits + look_back_finish(lb);
#endif
			const uint32_t m = excl & 7u;
			if (m && tidx != 0u) {
				uint64_t v = lb.tail_v;
				uint32_t spins = 0;
				while ((uint32_t)(v >> 8) != pD) { /* the tag of the tile in front is its position + 1 */
					v = ld_desc(lb.tail_p);
					if (++spins > (1u << 22))
						__trap();
				}
				if (lane == 0)
					stg[-1] = (uint32_t)v & ((1u << m) - 1u); /* word -1: the bits in front of the tile's first bit */
			}
			__syncwarp();

			/* ---- the tile leaves its staging words, shifted to its place in the stream */
			const uint32_t a = (uint32_t)((uintptr_t)dst & 15u);
			uint8_t *base = dst - a;
			const uint32_t g0 = 8u * a + excl, g1 = g0 + pbits;
			const uint32_t B0 = g0 >> 3;
			uint32_t B1 = last ? (g1 + 7u) >> 3 : g1 >> 3;
			B1 = min(B1, a + cap_eff);
			const uint32_t s = g0 & 31u, W0 = g0 >> 5;
			if (B1 > B0 && AIRS_TILE_ABLATE < 2) {
				/* whole 16-byte groups: five staging words, four funnel shifts, one byte-swapped 128-bit store */
				const uint32_t Gfull0 = (B0 + 15u) >> 4, Gfull1 = B1 >> 4;
				for (uint32_t G = Gfull0 + lane; G < Gfull1; G += 32u) {
					const int32_t i0 = (int32_t)(4u * G) - (int32_t)W0; /* local word of the group's first word */
					uint32_t wv[5];
#pragma unroll
					for (int i = 0; i < 5; i++)
						wv[i] = stg[i0 - 1 + i];
					*reinterpret_cast<uint4 *>(base + 16u * G) =
						make_uint4(airs_bswap32(__funnelshift_r(wv[1], wv[0], s)), airs_bswap32(__funnelshift_r(wv[2], wv[1], s)),
							   airs_bswap32(__funnelshift_r(wv[3], wv[2], s)), airs_bswap32(__funnelshift_r(wv[4], wv[3], s)));
				}
				/* the bytes in front of the first and behind the last whole group (fewer than 16 each; all bytes of a
				 * tile without a whole group): one byte per lane */
				const uint32_t head_end = min(16u * Gfull0, B1);
				const uint32_t byte = lane < 16u ? B0 + lane : max(16u * Gfull1, head_end) + (lane - 16u);
				const bool mine = lane < 16u ? byte < head_end : (Gfull1 >= Gfull0 && byte < B1);
				if (mine) {
					const int32_t L = (int32_t)(8u * byte) - (int32_t)g0; /* tile-local bit of the byte's first bit: >= -7 */
					const int32_t wi = L >> 5;
					const uint32_t v = __funnelshift_l(stg[wi + 1], stg[wi], (uint32_t)L & 31u);
					base[byte] = (uint8_t)(v >> 24);
				}
			}
			__syncwarp();
			{ /* the area goes back to all zero: word -1 (the carried bits), then whole 16-byte groups */
				const uint32_t nv4 = ((pbits + 31u) / 32u + 1u + 3u) / 4u;
				uint4 *stg4 = reinterpret_cast<uint4 *>(stg);
				for (uint32_t v = lane; v < nv4; v += 32u)
					stg4[v] = make_uint4(0, 0, 0, 0);
				if (lane == 0)
					stg[-1] = 0;
			}
			if (last) { /* the frame is complete: header, result (ref cmp.c:321-337) */
				const uint32_t checksum = (flags & AIRS_FJ_CHECKSUM) ? 1u : 0u;
				const uint32_t size = ((excl + pbits + 7u) >> 3) + 4u * checksum;
				const uint64_t id = (((uint64_t)AIRS_REC(5) << 32 | AIRS_REC(4)) + tw.cyc) & 0xFFFFFFFFFFFFull; /* a reset per primary pass */
				const uint32_t id_lo = (uint32_t)id, id_hi = (uint32_t)(id >> 32);
				const uint32_t first_result = AIRS_REC(9) + tw.f, job = AIRS_REC(13);
				const uint32_t g = sec ? AIRS_REC(24) : AIRS_REC(10), outlier = sec ? AIRS_REC(25) : AIRS_REC(11);
				const bool multi = sec ? (flags2 & AIRS_TX_MULTI2) != 0u : (flags & AIRS_FJ_MULTI) != 0u;
				const uint32_t pre = sec ? ((flags2 & AIRS_TX_PRE2_MODEL) ? CMP_PREPROCESS_MODEL : (flags2 & AIRS_TX_PRE2_DIFF) ? CMP_PREPROCESS_DIFF : CMP_PREPROCESS_NONE)
							 : ((flags & AIRS_FJ_PRE_DIFF) ? CMP_PREPROCESS_DIFF : (flags & AIRS_FJ_IWT) ? CMP_PREPROCESS_IWT : CMP_PREPROCESS_NONE);
				const uint32_t rate = pre == CMP_PREPROCESS_MODEL ? (flags2 >> 16) & 31u : 0u;
				const bool frames = AIRS_REC(22) > 1u; /* a context of several frames */
				if (size <= cap_eff) {
					if (lane < CMP_HDR_SIZE + 6u) {
						const uint32_t enc = multi ? CMP_ENCODER_GOLOMB_MULTI : CMP_ENCODER_GOLOMB_ZERO;
						uint32_t v;
						switch (lane) {
						case 0: v = 0x80u | (CMP_VERSION_NUMBER >> 8); break;
						case 1: v = CMP_VERSION_NUMBER & 0xFFu; break;
						case 2: v = size >> 16; break;
						case 3: v = size >> 8; break;
						case 4: v = size; break;
						case 5: v = (2u * n) >> 16; break;
						case 6: v = (2u * n) >> 8; break;
						case 7: v = 2u * n; break;
						case 8: v = id_hi >> 8; break;
						case 9: v = id_hi; break;
						case 10: v = id_lo >> 24; break;
						case 11: v = id_lo >> 16; break;
						case 12: v = id_lo >> 8; break;
						case 13: v = id_lo; break;
						case 14: v = tw.phase; break; /* sequence number */
						case 15: v = (pre << 4) | (checksum << 3) | enc; break;
						case 16: v = rate; break;
						case 17: v = g >> 8; break;
						case 18: v = g; break;
						case 19: v = outlier >> 16; break;
						case 20: v = outlier >> 8; break;
						default: v = outlier; break;
						}
						dst[lane] = (uint8_t)v;
					}
					if (lane == 0)
						b.results[first_result] = size;
				} else if (lane == 0) {
					/* A frame that does not fit.  Stored raw instead, or - in a context of several frames - a failure after
					 * which the context goes on differently: airs_encode_kernel, which runs behind this kernel, redoes the
					 * whole job (once, whatever the number of its frames that do not fit) */
					if ((flags & AIRS_FJ_FALLBACK_OK) || frames) {
						if (atomicExch(&b.plans[job].pad[0], 1u) == 0u)
							b.big_list[atomicAdd(&b.ticket[2], 1u)] = job;
					} else {
						b.results[first_result] = AIRS_ERR(DST_TOO_SMALL);
					}
				}
			}
			__syncwarp();
#undef AIRS_REC
		}

		/* tile k becomes the pending one */
		prec = rec;
		pT = have_cur ? t0 : 0xFFFFFFFFu;
		pD = d0;
		pbits = bits;
		if (have_cur) {
			t0 = t1;
			rec = nrec;
			d0 = d1;
			d1 = d2;
			d2 += n_warps;
			t1 = dealt(d1, slot1);
		}
	}
}

/* resident CTAs of one variant of the kernel on the current device */
template <bool FRAMES, bool BE>
static cudaError_t tile_resident(int *out)
{
	int dev = 0, sms = 0, per_sm = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e == cudaSuccess)
		e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (e == cudaSuccess) {
		const size_t need = (size_t)AIRS_TILE_CTAS_PER_SM * (sizeof(TileWarp) * kTWarps + 1024);
		const int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
		e = cudaFuncSetAttribute(airs_tile_kernel<FRAMES, BE>, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
	}
	if (e == cudaSuccess)
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, airs_tile_kernel<FRAMES, BE>, AIRS_TILE_THREADS, 0);
	*out = sms * per_sm;
	return e;
}

/* Both variants are launched, each over the CTAs that are resident at the same time (tiles are dealt round robin
 * over the grid and wait for each other: a CTA that is not resident would never be waited for in vain); the variant
 * the batch does not need returns at once. */
extern "C" cudaError_t airs_launch_tile(const AirsLaunch *b, cudaStream_t stream)
{
	static thread_local int cached_dev = -1, grid[4] = {0, 0, 0, 0};
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (dev != cached_dev) {
		if ((e = tile_resident<false, false>(&grid[0])) != cudaSuccess || (e = tile_resident<true, false>(&grid[1])) != cudaSuccess ||
		    (e = tile_resident<false, true>(&grid[2])) != cudaSuccess || (e = tile_resident<true, true>(&grid[3])) != cudaSuccess)
			return e;
		if (grid[0] < 1 || grid[1] < 1 || grid[2] < 1 || grid[3] < 1)
			return cudaErrorLaunchOutOfResources;
		cached_dev = dev;
	}
	if (b->be_batch) { /* the variants that know big-endian containers (AIRS_BATCH_BIG_ENDIAN) */
		airs_tile_kernel<false, true><<<grid[2], AIRS_TILE_THREADS, 0, stream>>>(*b);
		airs_tile_kernel<true, true><<<grid[3], AIRS_TILE_THREADS, 0, stream>>>(*b);
	} else {
		airs_tile_kernel<false, false><<<grid[0], AIRS_TILE_THREADS, 0, stream>>>(*b);
		airs_tile_kernel<true, false><<<grid[1], AIRS_TILE_THREADS, 0, stream>>>(*b);
	}
	return cudaGetLastError();
}

/* AIRS_BOUNDS_CHECK builds: strings that would have been staged outside the warp's staging words since the last
 * call (-1: not such a build) */
extern "C" int airs_tile_bounds_violations(void)
{
#ifdef AIRS_BOUNDS_CHECK
	unsigned int v = 0, zero = 0;
	cudaMemcpyFromSymbol(&v, fastcore::airs_bounds_violations, sizeof(v));
	cudaMemcpyToSymbol(fastcore::airs_bounds_violations, &zero, sizeof(zero));
	return (int)v;
#else
	return -1;
#endif
}
