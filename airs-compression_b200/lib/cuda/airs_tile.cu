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

/* the job (slot) tile T belongs to: tile jobs sit at fast_jobs[n_jobs - 1 - slot] with ascending tile_base;
 * a CTA's tickets only grow, so the search goes forward from the slot of its last tile, 32 slots per step */
__device__ __forceinline__ void find_slot(TileShared &sh, const FastJob *recs_end, uint32_t n_tjobs, uint32_t T, uint32_t lane)
{
	uint32_t slot = sh.slot, base = sh.slot_base, tiles = sh.slot_tiles;
	if (T >= base + tiles) {
		for (;;) {
			const uint32_t s = slot + 1u + lane;
			const uint32_t b = s < n_tjobs ? __ldg(&(recs_end - s)->tile_base) : 0xFFFFFFFFu;
			const uint32_t c = (uint32_t)__popc(__ballot_sync(kFull, b <= T));
			slot += c;
			if (c < 32u)
				break;
		}
	}
	__syncwarp();
	if (lane == 0)
		sh.slot = slot;
}

/* warp 0: where tile `ti` starts in its stream (decoupled look-back), and the bits it shares its first byte with */
__device__ __forceinline__ void look_back(TileInfo &ti, uint64_t *ring, uint64_t *tails, uint32_t hdr_bits, uint32_t lane,
					  uint32_t *carry_word)
{
	const uint32_t T = ti.T, tidx = ti.tidx;
	uint32_t excl = hdr_bits;

	if (tidx != 0u) {
		const int64_t lowest = (int64_t)T - tidx; /* first tile of the job; in front of it: the header */
		int64_t idx = (int64_t)T - 1;
		uint32_t sum = 0;
		for (;;) {
			const int64_t j = idx - lane;
			const bool real = j >= lowest;
			uint64_t v;
			bool ready;
			uint32_t nr, spins = 0;
			do { /* until every descriptor in front of the nearest prefix has been published */
				if (++spins > (1u << 24))
					__trap(); /* a predecessor that never answers: fail the launch instead of hanging the device */
				v = real ? ld_desc(ring + ((uint32_t)j & (kRing - 1u))) : (kFlagP | hdr_bits);
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
		}
		excl = sum;
	}
	/* the predecessor's last bits that share this tile's first byte */
	uint32_t carry = 0;
	const uint32_t m = excl & 7u;
	if (m && tidx != 0u) {
		uint64_t v;
		uint32_t spins = 0;
		do {
			v = ld_desc(tails + ((T - 1u) & (kRing - 1u)));
			if (++spins > (1u << 24))
				__trap();
		} while ((uint32_t)(v >> 8) != T); /* tag of tile T - 1 is T */
		carry = (uint32_t)v & ((1u << m) - 1u);
	}
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

	if (b.gate && (*b.gate != 0u) != (b.gate_want != 0u)) /* two-phase CONCAT: not the phase that runs */
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

	/* fetches tile T (a ticket) into sh.tile[q]: warp-parallel; all lanes of the calling warp */
	auto fetch = [&](uint32_t q, uint32_t T) {
		TileInfo &ti = sh.tile[q];
		if (T < n_tiles) {
			find_slot(sh, recs_end, n_tjobs, T, lane);
			__syncwarp();
			const uint32_t slot = sh.slot;
			const uint32_t *r = reinterpret_cast<const uint32_t *>(recs_end - slot);
			if (lane < 16u)
				ti.rec[lane] = __ldg(r + lane);
			__syncwarp();
			if (lane == 0) {
				sh.slot_base = ti.rec[14];
				sh.slot_tiles = ti.rec[15];
				ti.T = T;
				ti.tidx = T - ti.rec[14];
				ti.valid = 1;
			}
		} else if (lane == 0) {
			ti.valid = 0;
		}
		__syncwarp();
	};

	/* prologue: the first tile of this CTA.  Warp 1 draws the tickets, always one iteration before it needs them */
	uint32_t t_next = 0;
	if (warp == 1) {
		if (lane == 0)
			t_next = atomicAdd(&b.ticket[12], 1u);
		t_next = __shfl_sync(kFull, t_next, 0);
		fetch(0, t_next);
		if (lane == 0)
			t_next = atomicAdd(&b.ticket[12], 1u);
		t_next = __shfl_sync(kFull, t_next, 0);
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
			look_back(pend, ring, tails, 8u * (CMP_HDR_SIZE + 6u), lane, &sh.stg[par ^ 1u][kPad - 1u]);
		if (warp == 1) { /* the tile after this one: ticket drawn one iteration ago */
			fetch(qn, t_next);
			if (lane == 0)
				t_next = atomicAdd(&b.ticket[12], 1u);
			t_next = __shfl_sync(kFull, t_next, 0);
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
			for (uint32_t G = (B0 >> 4) + tid; 16u * G < B1; G += AIRS_TILE_THREADS) {
				const int32_t i0 = (int32_t)(4u * G) - (int32_t)W0; /* local word of the group's first word */
				uint32_t wv[5];
#pragma unroll
				for (int i = 0; i < 5; i++)
					wv[i] = stg[i0 - 1 + i]; /* i0 >= -3: inside the pad */
				uint32_t o[4];
#pragma unroll
				for (int i = 0; i < 4; i++)
					o[i] = __funnelshift_r(wv[i + 1], wv[i], s);
				const uint32_t byte0 = 16u * G;
				if (byte0 >= B0 && byte0 + 16u <= B1) {
					*reinterpret_cast<uint4 *>(base + byte0) =
						make_uint4(airs_bswap32(o[0]), airs_bswap32(o[1]), airs_bswap32(o[2]), airs_bswap32(o[3]));
				} else {
#pragma unroll 1
					for (uint32_t kb = 0; kb < 16u; kb++)
						if (byte0 + kb >= B0 && byte0 + kb < B1)
							base[byte0 + kb] = (uint8_t)(o[kb >> 2] >> (24u - 8u * (kb & 3u)));
				}
			}
			__syncthreads(); /* B3: everybody has read what it needs */
			{
				const uint32_t nw = (pend.bits + 31u) / 32u + 1u;
				for (uint32_t w = tid; w < nw + 1u; w += AIRS_TILE_THREADS)
					stg[(int32_t)w - 1] = 0;
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
		__syncthreads(); /* B4: tile structs of iteration k are free */
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
