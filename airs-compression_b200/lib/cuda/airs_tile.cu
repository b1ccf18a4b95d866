/*
 * airs_tile.cu - airs_tile_kernel: long single-frame jobs (chunks of more than 32768 samples without model:
 * BASELINE configs 1, 4 and 5 in their 2 MiB cut) spread over ALL resident CTAs, 2048 samples ("tile") at a
 * time, whatever the number of jobs - one 1 Mi-sample cmp_compress_u16() call uses the whole GPU.
 *
 * The bit position at which a tile's code words start is the sum of the bit counts of all tiles in front of
 * it in its frame: a single-pass scan with decoupled look-back over 64-bit tile descriptors in global memory
 * (flag | tile id | bits: "aggregate" as soon as the tile's bits are counted, "inclusive prefix" once its own
 * start is known).  Tiles are handed out in stream order by an atomic ticket, so every tile a CTA waits for
 * is held by a CTA that is running.  Software pipeline per CTA (128 threads, 4 warps, one unit of
 * airs_fastcore.cuh per warp), iteration k:
 *   code words of tile k (registers) -> warp scans -> barrier -> thread 0 publishes the aggregate of tile k;
 *   all warps stage tile k at TILE-LOCAL bit positions in staging area k % 2 while warp 0 resolves the
 *   look-back of tile k - 1 (published one tile earlier: its predecessors have had a tile's time to answer)
 *   -> barrier -> tile k - 1 leaves its staging area shifted by its start position: five shared-memory
 *   words and four funnel shifts per 16-byte group, byte-swapped 128-bit stores.
 * Seams: tiles meet at bit granularity.  A tile writes the bytes [start / 8, end / 8) of the stream; the
 * start % 8 bits of its first byte that belong to its predecessor come out of a second ring ("tail": the
 * last 7 bits of every tile, published when the tile is staged).  The tile that ends a frame also writes the
 * last, zero-padded byte, the header (ref cmp_hdr_serialize, header.c:24-67) and the result.
 *
 * Reference being replaced: the per-sample loop of compress_engine (cmp.c:296-312) with
 * bitstream_add_bits32 (bitstream_writer.h:124-158), for one frame by many CTAs.
 */
#include <cuda_runtime.h>

#include "airs_fastcore.cuh"
#include "airs_launch.h"

namespace {

using namespace fastcore;

constexpr uint32_t kTWarps = AIRS_TILE_THREADS / 32;
constexpr uint32_t kTile = kTWarps * kUnit;             /* 2048 samples */
constexpr uint32_t kTileWords = kTile * 48 / 32;         /* a tile at 48 bits per sample */
constexpr uint32_t kPad = 8;                             /* words in front of an area: word -1 takes the carried bits */
constexpr uint32_t kRing = AIRS_TILE_RING;               /* descriptors of the last kRing tiles (far more than are in flight) */

constexpr uint64_t kFlagA = 1ull << 62, kFlagP = 2ull << 62;
constexpr uint32_t kValBits = 28;                        /* bits of a frame: less than 2^27 + header */

struct TileInfo {          /* one tile in flight (current or pending), shared memory */
	uint32_t rec[16];  /* its job's record (FastJob) */
	uint32_t T;        /* global tile id */
	uint32_t tidx;     /* index of the tile in its job */
	uint32_t bits;     /* code bits of the tile */
	uint32_t excl;     /* stream bits in front of the tile (header included), from the look-back */
	uint32_t valid;
};

struct TileShared {
	alignas(16) uint32_t stg[2][kPad + kTileWords + 8];
	TileInfo tile[3];  /* [k % 3]: current, pending and next tile */
	uint32_t wtot[2][kTWarps];
	uint32_t slot, slot_base, slot_tiles; /* the tile job the last ticket fell into */
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

__device__ __forceinline__ uint64_t make_desc(uint64_t flag, uint32_t T, uint32_t value)
{
	return flag | ((uint64_t)(T + 1u) << kValBits) | value;
}

/* warp 0, one iteration early: the descriptors of the 32 tiles in front of `ti` and the tail of the one right in
 * front of it are requested (no waiting here); look_back() looks at the answers behind the tile's code words */
__device__ __forceinline__ void look_back_early(const TileInfo &ti, const uint64_t *ring, const uint64_t *tails, uint32_t hdr_bits,
						uint32_t lane, uint64_t &v0, uint64_t &tail0)
{
	const uint32_t T = ti.T, tidx = ti.tidx;
	const int64_t j = (int64_t)T - 1 - lane;
	v0 = (tidx != 0u && j >= (int64_t)T - tidx) ? ld_desc(ring + ((uint32_t)j & (kRing - 1u))) : (kFlagP | hdr_bits);
	tail0 = tidx != 0u ? ld_desc(tails + ((T - 1u) & (kRing - 1u))) : 0ull;
}

/* warp 0: where tile `ti` starts in its stream (decoupled look-back), and the bits it shares its first byte with.
 * v0 / tail0: the answers to look_back_early(); only what they leave open is polled */
__device__ __forceinline__ void look_back(TileInfo &ti, uint64_t *ring, uint64_t *tails, uint32_t hdr_bits, uint32_t lane,
					  uint32_t *carry_word, uint64_t v0, uint64_t tail0, uint32_t *stats)
{
#ifdef AIRS_TILE_STATS
	uint32_t st_polls = 0, st_tail = 0, st_windows = 0;
#endif
	const uint32_t T = ti.T, tidx = ti.tidx;
	uint32_t excl = hdr_bits;

	if (tidx != 0u) {
		const int64_t lowest = (int64_t)T - tidx; /* first tile of the job; in front of it: the header */
		int64_t idx = (int64_t)T - 1;
		uint32_t sum = 0;
		bool early = true;
		for (;;) {
			const int64_t j = idx - lane;
			const bool real = j >= lowest;
			uint64_t v;
			bool ready;
			uint32_t nr, spins = 0;
			do { /* until every descriptor in front of the nearest prefix has been published */
				if (++spins > (1u << 24))
					__trap(); /* a predecessor that never answers: fail the launch instead of hanging the device */
#ifdef AIRS_TILE_STATS
				st_polls += early ? 0u : 1u;
#endif
				v = early ? v0 : (real ? ld_desc(ring + ((uint32_t)j & (kRing - 1u))) : (kFlagP | hdr_bits));
				early = false;
				ready = !real || (((uint32_t)(v >> kValBits) == (uint32_t)j + 1u) && (v >> 62) != 0);
				nr = __ballot_sync(kFull, !ready);
				const uint32_t pm = __ballot_sync(kFull, ready && (v >> 62) == 2u);
				const uint32_t first_nr = nr ? (uint32_t)__ffs((int)nr) - 1u : 32u;
				const uint32_t p = pm ? (uint32_t)__ffs((int)pm) - 1u : 32u;
				if (p < first_nr) { /* aggregates of the lanes in front of p, prefix of p */
					sum += __reduce_add_sync(kFull, lane <= p ? (uint32_t)v & ((1u << kValBits) - 1u) : 0u);
					nr = 0;
					idx = -1; /* done */
					break;
				}
			} while (nr);
			if (idx < 0)
				break;
			sum += __reduce_add_sync(kFull, (uint32_t)v & ((1u << kValBits) - 1u)); /* 32 aggregates, no prefix yet */
			idx -= 32;
#ifdef AIRS_TILE_STATS
			st_windows++;
#endif
		}
		excl = sum;
	}
	/* the predecessor's last bits that share this tile's first byte */
	uint32_t carry = 0;
	const uint32_t m = excl & 7u;
	if (m && tidx != 0u) {
		uint64_t v = tail0;
		uint32_t spins = 0;
		while ((uint32_t)(v >> 8) != T) { /* tag of tile T - 1 is T */
			v = ld_desc(tails + ((T - 1u) & (kRing - 1u)));
			if (++spins > (1u << 24))
				__trap();
#ifdef AIRS_TILE_STATS
			st_tail++;
#endif
		}
		carry = (uint32_t)v & ((1u << m) - 1u);
	}
#ifdef AIRS_TILE_STATS
	if (lane == 0) {
		atomicAdd(stats + 0, 1u);
		atomicAdd(stats + 1, st_polls);
		atomicAdd(stats + 2, st_tail);
		atomicAdd(stats + 3, st_windows);
	}
#endif
	if (lane == 0) {
		ti.excl = excl;
		*carry_word = carry; /* word -1 of the tile's staging area: the bits in front of its first bit */
		st_desc(ring + (T & (kRing - 1u)), make_desc(kFlagP, T, excl + ti.bits));
	}
}

/* Tile k of a job whose encoder / preprocessing are fixed at compile time: code words (registers), warp scan,
 * barrier B1, the tile's aggregate published, strings staged at tile-local bit positions in area `par`.
 * Contains a block barrier: all threads of the CTA call it (they all hold the same tile). */
template <bool MULTI, bool DIFF>
__device__ __forceinline__ void tile_encode(const Dbg &dbg, TileShared &sh, TileInfo &cur, uint64_t *ring, const FK &k,
					    const uint4 (&x)[kRows], uint32_t front, uint32_t par)
{
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t n = cur.rec[6], first = cur.tidx * kTile + warp * kUnit;
	UnitStrings<MULTI> s;
	uint32_t b;

	if (first + kUnit <= n) {
		const uint32_t nv[kRows] = {8u, 8u};
		b = unit_codes<MULTI, DIFF, false>(k, x, front, nv, lane, s);
	} else { /* the ragged end of the frame (or nothing at all) */
		uint32_t nv[kRows];
#pragma unroll
		for (uint32_t j = 0; j < kRows; j++) {
			const uint32_t p = first + 8u * (32u * j + lane);
			nv[j] = p >= n ? 0u : min(8u, n - p);
		}
		b = unit_codes<MULTI, DIFF, true>(k, x, front, nv, lane, s);
	}
	const uint32_t incl = unit_scan(b);
	if (lane == 31u)
		sh.wtot[par][warp] = (incl & 0xFFFFu) + (incl >> 16);
	__syncthreads(); /* B1: warp totals of the tile */

	uint32_t wpre = 0, tile_bits = 0;
#pragma unroll
	for (uint32_t w = 0; w < kTWarps; w++) {
		const uint32_t t = sh.wtot[par][w];
		wpre += w < warp ? t : 0u;
		tile_bits += t;
	}
	if (tid == 0) {
		cur.bits = tile_bits;
		st_desc(ring + (cur.T & (kRing - 1u)), make_desc(kFlagA, cur.T, tile_bits));
	}
	/* staged at tile-local bit positions: the shift to the stream position happens on the way out */
	const uint32_t tot0 = __shfl_sync(kFull, incl, 31) & 0xFFFFu, excl = incl - b;
	const uint32_t base = 8u * (uint32_t)__cvta_generic_to_shared(&sh.stg[par][kPad]) + wpre;
	const uint32_t pos[kRows] = {base + (excl & 0xFFFFu), base + tot0 + (excl >> 16)};
	unit_put<MULTI>(dbg, s, pos);
}

} /* namespace */

__global__ void __launch_bounds__(AIRS_TILE_THREADS, AIRS_TILE_CTAS_PER_SM) airs_tile_kernel(AirsLaunch b)
{
	__shared__ TileShared sh;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

	if (b.ticket[AIRS_TICKET_INVALID] || (b.gate && (*b.gate != 0u) != (b.gate_want != 0u))) /* a bad job table; two-phase CONCAT: not the phase that runs */
		return;
	const uint64_t counts = *reinterpret_cast<const uint64_t *>(b.ticket + 10); /* airs_plan_kernel: tile jobs << 40 | tiles */
	const uint32_t n_tiles = (uint32_t)(counts & ((1ull << 40) - 1u)), n_tjobs = (uint32_t)(counts >> 40);
	if (n_tiles == 0u)
		return;
	const FastJob *recs_end = reinterpret_cast<const FastJob *>(b.fast_jobs) + (b.n_jobs - 1u); /* slot s at recs_end - s */
	uint64_t *ring = b.tile_ring, *tails = b.tile_ring + kRing;

	for (uint32_t w = tid; w < 2u * (kPad + kTileWords + 8u); w += AIRS_TILE_THREADS)
		(&sh.stg[0][0])[w] = 0;
	if (tid < 3u)
		sh.tile[tid].valid = 0;
	if (tid == 0) {
		sh.slot = 0xFFFFFFFFu; /* "the slot in front of slot 0": the first search starts at slot 0 */
		sh.slot_base = 0;
		sh.slot_tiles = 0;
	}
	__syncthreads();

	/* Warp 1 finds the job of every tile, in steps spread over three iterations so that no step waits for
	 * global memory: tickets are drawn three tiles ahead; the tile_base fields of the 32 jobs behind the job of
	 * the tile before (tile jobs sit at recs_end - slot with ascending tile_base, and a CTA's tickets only grow)
	 * are requested two tiles ahead; the slot follows from a ballot one tile ahead, when the job's record is
	 * requested; the record reaches shared memory behind the code words of the tile in front. */
	uint32_t t1 = 0, t2 = 0, t3 = 0;   /* tickets of tiles k + 1, k + 2, k + 3 (warp 1, lane 0 draws) */
	uint32_t slot1 = 0xFFFFFFFFu;      /* slot of tile k + 1 once resolved; before: of the tile in front of it */
	uint32_t fb = 0xFFFFFFFFu;         /* tile_base of slot (slot of tile k) + 1 + lane, requested for tile k + 1 */
	uint32_t recw = 0;                 /* word `lane` of the record of tile k + 1 */
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
	auto publish_tile = [&](uint32_t q, uint32_t T, uint32_t word) { /* record word `lane` -> sh.tile[q] */
		TileInfo &ti = sh.tile[q];
		if (T < n_tiles) {
			if (lane < 16u)
				ti.rec[lane] = word;
			const uint32_t base = __shfl_sync(kFull, word, 14);
			if (lane == 0) {
				ti.T = T;
				ti.tidx = T - base;
				ti.valid = 1;
			}
		} else if (lane == 0) {
			ti.valid = 0;
		}
		__syncwarp();
	};
	if (warp == 1) { /* prologue: tile 0 in full, tile 1 probed, tile 2 drawn */
		/* the first three tiles of every CTA are dealt out round robin: a CTA must never hold consecutive tiles (the
		 * second one would be counted an iteration after the first, and every tile behind it would wait for that) */
		const uint32_t t0 = blockIdx.x;
		t1 = blockIdx.x + gridDim.x;
		t2 = blockIdx.x + 2u * gridDim.x;
		resolve(slot1, t0, probe(slot1, t0));
		const uint32_t w0 = (t0 < n_tiles && lane < 16u) ? __ldg(reinterpret_cast<const uint32_t *>(recs_end - slot1) + lane) : 0u;
		publish_tile(0, t0, w0);
		fb = probe(slot1, t1);
	}
	__syncthreads();

	const uint4 zero4 = make_uint4(0, 0, 0, 0);
	uint4 nx[kRows] = {zero4, zero4};
	uint32_t nfront = 0;
	/* the samples of this warp's unit of tile sh.tile[q], requested ahead */
	auto request = [&](uint32_t q) {
		const TileInfo &ti = sh.tile[q];
		if (!ti.valid)
			return;
		const uint8_t *src = reinterpret_cast<const uint8_t *>((uintptr_t)(ti.rec[0] | (uint64_t)ti.rec[1] << 32));
		const uint32_t n = ti.rec[6], first = ti.tidx * kTile + warp * kUnit, n_whole = n / 8u;
		const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
#pragma unroll
		for (uint32_t j = 0; j < kRows; j++) {
			const uint32_t p = first / 8u + 32u * j + lane;
			nx[j] = p < n_whole ? __ldg(src4 + p) : zero4;
			if (p == n_whole && (n & 7u))
				nx[j] = load_partial_piece(reinterpret_cast<const uint16_t *>(src), 8u * p, n & 7u);
		}
		/* lane 0: the sample in front of the unit (DIFF), in the upper half of a word */
		nfront = (lane == 0 && first != 0u && first < n) ? (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(src) + first - 1u) << 16 : 0u;
	};
	request(0);

	Dbg dbg;
#ifdef AIRS_BOUNDS_CHECK
	dbg.lo = (uint32_t)__cvta_generic_to_shared(&sh.stg[0][0]);
	dbg.hi = dbg.lo + (uint32_t)sizeof(sh.stg);
#endif

	for (uint32_t k = 0;; k++) {
		const uint32_t q = k % 3u, qp = (k + 2u) % 3u, qn = (k + 1u) % 3u, par = k & 1u;
		TileInfo &cur = sh.tile[q], &pend = sh.tile[qp];
		const bool have_cur = cur.valid != 0u, have_pend = k > 0u && pend.valid != 0u;
		if (!have_cur && !have_pend)
			break;

		/* ---- requests whose answers are looked at behind this tile's code words */
		uint64_t lb_v0 = 0, lb_tail0 = 0;
		if (warp == 0 && have_pend)
			look_back_early(pend, ring, tails, 8u * (CMP_HDR_SIZE + 6u), lane, lb_v0, lb_tail0);
		if (warp == 1) {
			resolve(slot1, t1, fb);                                /* job of tile k + 1 */
			recw = (t1 < n_tiles && lane < 16u) ? __ldg(reinterpret_cast<const uint32_t *>(recs_end - slot1) + lane) : 0u;
			t2 = __shfl_sync(kFull, t2, 0);                        /* (drawn one iteration ago) */
			fb = probe(slot1, t2);                                 /* candidates for tile k + 2 */
			if (lane == 0)
				t3 = 3u * gridDim.x + atomicAdd(&b.ticket[12], 1u); /* ticket of tile k + 3 */
		}

		/* ---- tile k: code words, aggregate, staging (tile_encode holds barrier B1) */
		if (have_cur) {
			const uint32_t flags = cur.rec[8];
			const bool multi = (flags & AIRS_FJ_MULTI) != 0u;
			const FK kk = make_fk(multi, cur.rec[10], (flags >> 8) & 15u, cur.rec[11], cur.rec[12]);
			uint4 x[kRows];
#pragma unroll
			for (uint32_t j = 0; j < kRows; j++)
				x[j] = nx[j];
			if (multi) {
				if (flags & AIRS_FJ_PRE_DIFF)
					tile_encode<true, true>(dbg, sh, cur, ring, kk, x, nfront, par);
				else
					tile_encode<true, false>(dbg, sh, cur, ring, kk, x, nfront, par);
			} else {
				if (flags & AIRS_FJ_PRE_DIFF)
					tile_encode<false, true>(dbg, sh, cur, ring, kk, x, nfront, par);
				else
					tile_encode<false, false>(dbg, sh, cur, ring, kk, x, nfront, par);
			}
		} else {
			__syncthreads(); /* B1 of a CTA that only has a tile left to send off */
		}
		if (warp == 0 && have_pend) /* (thread 0 wrote pend.bits two barriers ago) */
			look_back(pend, ring, tails, 8u * (CMP_HDR_SIZE + 6u), lane, &sh.stg[par ^ 1u][kPad - 1u], lb_v0, lb_tail0, b.ticket + 16);
		if (warp == 1) { /* the record of tile k + 1 has arrived */
			publish_tile(qn, t1, recw);
			t1 = t2;
			t2 = t3; /* lane 0; broadcast when it is needed */
		}
		__syncthreads(); /* B2: tile k staged, tile k - 1 placed, tile k + 1 known */

		if (have_cur && tid == 0) { /* the last 7 bits of tile k, for the tile behind it */
			const uint32_t e = cur.bits; /* >= 2048 > 7 unless the frame ends here (then nobody asks) */
			const uint32_t *stg = &sh.stg[par][kPad];
			const uint32_t wi = e >> 5, sft = e & 31u;
			const uint32_t before = wi ? stg[wi - 1u] : 0u;
			const uint32_t last32 = sft ? __funnelshift_l(stg[wi], before, sft) : before; /* the 32 bits that end at bit e */
			st_desc(tails + (cur.T & (kRing - 1u)), ((uint64_t)(cur.T + 1u) << 8) | (last32 & 0x7Fu));
		}
		request(qn); /* samples of tile k + 1 travel while tile k - 1 leaves */

		/* ---- tile k - 1 leaves its staging area, shifted to its place in the stream */
		if (have_pend) {
			const uint32_t parp = par ^ 1u;
			uint32_t *stg = &sh.stg[parp][kPad];
			uint8_t *dst = reinterpret_cast<uint8_t *>((uintptr_t)(pend.rec[2] | (uint64_t)pend.rec[3] << 32));
			const uint32_t n = pend.rec[6], cap_eff = pend.rec[7], flags = pend.rec[8];
			const uint32_t a = (uint32_t)((uintptr_t)dst & 15u);
			uint8_t *base = dst - a;
			const uint32_t g0 = 8u * a + pend.excl, g1 = g0 + pend.bits;
			const bool last = (pend.tidx + 1u) * kTile >= n;
			const uint32_t B0 = g0 >> 3;
			uint32_t B1 = last ? (g1 + 7u) >> 3 : g1 >> 3;
			B1 = min(B1, a + cap_eff);
			const uint32_t s = g0 & 31u, W0 = g0 >> 5;
			/* whole 16-byte groups: five staging words, four funnel shifts, one byte-swapped 128-bit store */
			const uint32_t Gfull0 = (B0 + 15u) >> 4, Gfull1 = B1 >> 4;
			for (uint32_t G = Gfull0 + tid; G < Gfull1; G += AIRS_TILE_THREADS) {
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
			 * tile without a whole group): one byte per lane of the last warp */
			if (warp == kTWarps - 1u) {
				const uint32_t head_end = min(16u * Gfull0, B1);
				uint32_t byte = lane < 16u ? B0 + lane : max(16u * Gfull1, head_end) + (lane - 16u);
				const bool mine = lane < 16u ? byte < head_end : (Gfull1 >= Gfull0 && byte < B1);
				if (mine) {
					const int32_t L = (int32_t)(8u * byte) - (int32_t)g0; /* tile-local bit of the byte's first bit: >= -7 */
					const int32_t wi = L >> 5;
					const uint32_t v = __funnelshift_l(stg[wi + 1], stg[wi], (uint32_t)L & 31u);
					base[byte] = (uint8_t)(v >> 24);
				}
			}
			__syncthreads(); /* B3: everybody has read what it needs */
			{ /* the area goes back to all zero: word -1 (the carried bits), then whole 16-byte groups */
				const uint32_t nv4 = ((pend.bits + 31u) / 32u + 1u + 3u) / 4u;
				uint4 *stg4 = reinterpret_cast<uint4 *>(stg);
				for (uint32_t v = tid; v < nv4; v += AIRS_TILE_THREADS)
					stg4[v] = make_uint4(0, 0, 0, 0);
				if (tid == 0)
					stg[-1] = 0;
			}
			if (last) { /* the frame is complete: header, result (ref cmp.c:321-337) */
				const uint32_t checksum = (flags & AIRS_FJ_CHECKSUM) ? 1u : 0u;
				const uint32_t size = ((pend.excl + pend.bits + 7u) >> 3) + 4u * checksum;
				if (size <= cap_eff) {
					if (tid < CMP_HDR_SIZE + 6u) {
						const uint32_t id_lo = pend.rec[4], id_hi = pend.rec[5], g = pend.rec[10], outlier = pend.rec[11];
						const uint32_t pre = (flags & AIRS_FJ_PRE_DIFF) ? CMP_PREPROCESS_DIFF : CMP_PREPROCESS_NONE;
						const uint32_t enc = (flags & AIRS_FJ_MULTI) ? CMP_ENCODER_GOLOMB_MULTI : CMP_ENCODER_GOLOMB_ZERO;
						uint32_t v;
						switch (tid) {
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
						case 14: v = 0; break;
						case 15: v = (pre << 4) | (checksum << 3) | enc; break;
						case 16: v = 0; break;
						case 17: v = g >> 8; break;
						case 18: v = g; break;
						case 19: v = outlier >> 16; break;
						case 20: v = outlier >> 8; break;
						default: v = outlier; break;
						}
						dst[tid] = (uint8_t)v;
					}
					if (tid == 0)
						b.results[pend.rec[9]] = size;
				} else if (tid == 0) {
					if (flags & AIRS_FJ_FALLBACK_OK) /* stored raw instead: airs_encode_kernel, which runs behind this kernel, redoes the job */
						b.big_list[atomicAdd(&b.ticket[2], 1u)] = pend.rec[13];
					else
						b.results[pend.rec[9]] = AIRS_ERR(DST_TOO_SMALL);
				}
			}
		}
		/* (no barrier: what iteration k + 1 writes before its barrier B1 - registers, requests - touches nothing that
		 * is read here; the tile structs and staging areas change hands behind B1) */
	}
}

extern "C" cudaError_t airs_launch_tile(const AirsLaunch *b, unsigned int grid, cudaStream_t stream)
{
	airs_tile_kernel<<<grid, AIRS_TILE_THREADS, 0, stream>>>(*b);
	return cudaGetLastError();
}

/* AIRS_BOUNDS_CHECK builds: strings that would have been staged outside the CTA's staging areas since the last
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

/* resident CTAs of airs_tile_kernel on the current device */
extern "C" cudaError_t airs_tile_resident_ctas(int *out)
{
	int dev = 0, sms = 0, per_sm = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e == cudaSuccess)
		e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (e == cudaSuccess) {
		const size_t need = (size_t)AIRS_TILE_CTAS_PER_SM * (sizeof(TileShared) + 1024);
		const int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
		e = cudaFuncSetAttribute(airs_tile_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
	}
	if (e == cudaSuccess)
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, airs_tile_kernel, AIRS_TILE_THREADS, 0);
	*out = sms * per_sm;
	return e;
}
