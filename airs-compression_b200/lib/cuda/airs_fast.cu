/*
 * airs_fast.cu - airs_fast_kernel: one WARP per short single-frame job (a chunk of up to
 * 32768 samples without model: BASELINE config 3 and the small-chunk cuts of config 4).
 *
 * The kernel is bound by the integer ALU pipe (LOP3 / SHF / PRMT / IADD3 / SEL share one pipe
 * of 64 lanes per clock and SM, tools/micro/pipes.cu), so everything here is written to need
 * few of those instructions per sample, and to be warp-synchronous: a warp owns its job, its
 * staging words and its position in the stream; there is no block barrier and no code that
 * runs on one thread while the others wait.
 *
 * Per unit of 512 samples ("rows" 0 and 1: lane l holds the pieces 2 l and 2 l + 1 of 8 samples,
 * see airs_fastcore.cuh; the next unit's loads are in flight while one is encoded):
 *   packed 16x2 residuals (none / diff, VIADD.16x2 + PRMT) -> packed zig-zag -> per sample
 *   the Golomb code word ARITHMETICALLY: quotient by one multiply-high (airs_fast.cuh), code
 *   word by one shift, one multiply-add and one three-input add, escape by two selects (ref
 *   cmp_encoder_encode_s16, encoder.c:327-378; golomb_encode, encoder.c:303-324) -> the code
 *   words of a pair of samples merged into one string of at most 64 bits -> one shuffle scan
 *   over the packed bit counts of both rows -> every string shifted into place by three funnel
 *   shifts and OR-ed into the warp's MSB-first staging words (shared-memory reductions; ref
 *   bitstream_add_bits32, bitstream_writer.h:124-158) -> complete 16-byte groups leave as
 *   byte-swapped coalesced 128-bit stores.
 * Any data goes through this one path - escapes cost nothing extra, which is what BASELINE
 * config 3 (escape heavy, mixed parameters) asks for.  A batch whose jobs mostly share one
 * encoder additionally gets the pair-table arm (airs_fast_table.cuh).
 *
 * Header (ref cmp_hdr_serialize, header.c:24-67) and the uncompressed fallback (cmp.c:363-392)
 * are handled by the same warp; XXH32 trailers by airs_checksum_kernel behind it.
 */
#include <cuda_runtime.h>

#include "airs_fastcore.cuh"
#include "airs_launch.h"

namespace {

using namespace fastcore;

constexpr uint32_t kFWarps = AIRS_FAST_THREADS / 32;
/* The staging words are drained when more than one unit at 48 bits per sample is staged (for most data: every
 * few units, so that the fixed cost of a drain is shared); the next unit then still fits, and so do the 22
 * header bytes behind up to 8 bytes of alignment in front of the first one */
constexpr uint32_t kStgWords = 2 * kUnitMaxBits / 32 + 8;

/* the staging words of one warp: MSB-first 32-bit words of the stream under construction, all
 * zero when idle; the pad in front absorbs the zeros that strings ending in word 0 or 1 OR
 * below the area */
struct FastWarp {
	alignas(16) uint32_t pad[4];
	uint32_t stg[kStgWords];
};

/* where the stream of the job stands: word 0 of the staging area is 32-bit word gw0 (a multiple of 4)
 * of the 16-byte aligned space that starts at base = dst - (dst & 15) and holds sbits bits (< 128 behind a
 * drain); [lo, hi) is the byte window of that space the job may write */
struct Out {
	uint8_t *base;
	uint32_t lo, hi;
	uint32_t gw0, sbits;
};

/* complete 16-byte groups of the staging area leave as coalesced stores; the partial group moves to the front */
__device__ __forceinline__ void drain(FastWarp &ws, Out &o, uint32_t lane)
{
	const uint32_t staged = o.sbits, nvec = staged >> 7, b0 = o.gw0 * 4u;
	uint4 *stg4 = reinterpret_cast<uint4 *>(ws.stg);

	__syncwarp();
	if (b0 >= o.lo && b0 + 16u * nvec <= o.hi) {
		uint4 *out = reinterpret_cast<uint4 *>(o.base + b0);
		for (uint32_t v = lane; v < nvec; v += 32u) {
			const uint4 q = stg4[v];
			stg4[v] = make_uint4(0, 0, 0, 0);
			out[v] = make_uint4(airs_bswap32(q.x), airs_bswap32(q.y), airs_bswap32(q.z), airs_bswap32(q.w));
		}
	} else { /* an edge of the window: the stream starts at dst % 16 == 8, or the capacity ends here */
		for (uint32_t v = lane; v < nvec; v += 32u) {
			const uint32_t b = b0 + 16u * v;
			const uint4 q = stg4[v];
			if (b >= o.lo && b + 16u <= o.hi) {
				*reinterpret_cast<uint4 *>(o.base + b) =
					make_uint4(airs_bswap32(q.x), airs_bswap32(q.y), airs_bswap32(q.z), airs_bswap32(q.w));
			} else {
				const uint8_t *s8 = reinterpret_cast<const uint8_t *>(stg4 + v);
#pragma unroll 1
				for (uint32_t k = 0; k < 16u; k++)
					if (b + k >= o.lo && b + k < o.hi)
						o.base[b + k] = s8[k ^ 3u];
			}
			stg4[v] = make_uint4(0, 0, 0, 0);
		}
	}
	__syncwarp();
	if (lane == 0 && nvec) {
		const uint4 carry = stg4[nvec];
		stg4[nvec] = make_uint4(0, 0, 0, 0);
		stg4[0] = carry;
	}
	__syncwarp();
	o.gw0 += nvec * 4u;
	o.sbits = staged & 127u;
}

/* the pass of one job through its units */
template <bool MULTI, bool DIFF, bool BE>
__device__ __forceinline__ void encode_units(const Dbg &dbg, const FK &k, FastWarp &ws, Out &o, const uint8_t *src, uint32_t n, uint32_t lane,
					     bool be)
{
	const uint32_t n_whole = n / 8u;         /* complete pieces */
	const uint32_t n_full_units = n / kUnit; /* units whose 64 pieces are all complete */
	const uint32_t stg_bit = 8u * (uint32_t)__cvta_generic_to_shared(ws.stg);
	const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
	const uint4 zero4 = make_uint4(0, 0, 0, 0);
	const uint32_t full_nv[kRows] = {8u, 8u};
	uint32_t front = 0;
	uint4 x[kRows], nx[kRows];

	/* the complete pieces of the next unit are requested while one unit is encoded */
#pragma unroll
	for (uint32_t j = 0; j < kRows; j++)
		nx[j] = unit_piece(lane, j) < n_whole ? load_piece<BE>(src4, unit_piece(lane, j), be) : zero4;
	uint32_t u = 0;
	for (; u < n_full_units; u++) {
#pragma unroll
		for (uint32_t j = 0; j < kRows; j++)
			x[j] = nx[j];
		const uint32_t p1 = (u + 1u) * kUnitPieces + unit_piece(lane, 0);
#pragma unroll
		for (uint32_t j = 0; j < kRows; j++)
			nx[j] = p1 + j < n_whole ? load_piece<BE>(src4, p1 + j, be) : zero4;
		o.sbits += encode_unit<MULTI, DIFF, false>(dbg, k, x, front, full_nv, lane, stg_bit + o.sbits);
		if (o.sbits > kUnitMaxBits)
			drain(ws, o, lane);
	}
	if (u * kUnit < n) { /* the ragged last unit */
		uint32_t nv[kRows];
#pragma unroll
		for (uint32_t j = 0; j < kRows; j++) {
			const uint32_t p = u * kUnitPieces + unit_piece(lane, j);
			nv[j] = 8u * p >= n ? 0u : min(8u, n - 8u * p);
			if (nv[j] != 0u && nv[j] != 8u)
				nx[j] = load_partial_piece(reinterpret_cast<const uint16_t *>(src), 8u * p, nv[j], BE && be);
		}
		o.sbits += encode_unit<MULTI, DIFF, true>(dbg, k, nx, front, nv, lane, stg_bit + o.sbits);
	}
	drain(ws, o, lane);
}

} /* namespace */

/* BE: the variant for batches that may hold big-endian containers (AIRS_BATCH_BIG_ENDIAN); the other one has no trace
 * of them */
template <bool BE>
__global__ void __launch_bounds__(AIRS_FAST_THREADS, AIRS_FAST_CTAS_PER_SM) airs_fast_kernel(AirsLaunch b)
{
	__shared__ FastWarp wsh[kFWarps];
	const uint32_t lane = threadIdx.x & 31u;
	FastWarp &ws = wsh[threadIdx.x >> 5];

	if (b.ticket[AIRS_TICKET_INVALID] || (b.gate && (*b.gate != 0u) != (b.gate_want != 0u))) /* a bad job table; two-phase CONCAT: not the phase that runs */
		return;
	const uint32_t n_fast = b.ticket[3]; /* entries of fast_jobs, written by airs_plan_kernel */
	if (n_fast == 0u)
		return;
	for (uint32_t w = lane; w < 4u + kStgWords; w += 32u)
		ws.pad[w] = 0;
	__syncwarp();
	Dbg dbg;
#ifdef AIRS_BOUNDS_CHECK
	dbg.lo = (uint32_t)__cvta_generic_to_shared(&ws);
	dbg.hi = dbg.lo + (uint32_t)sizeof(FastWarp);
#endif

	/* tickets are drawn two jobs ahead and the record of the next job one job ahead, so that neither
	 * the atomic's round trip nor the record's load is waited for */
	const FastJob *recs = reinterpret_cast<const FastJob *>(b.fast_jobs);
	uint32_t t = 0, t1 = 0, t2 = 0;
	if (lane == 0) {
		t = atomicAdd(&b.ticket[1], 1u);
		t1 = atomicAdd(&b.ticket[1], 1u);
	}
	t = __shfl_sync(kFull, t, 0);
	t1 = __shfl_sync(kFull, t1, 0);
	uint32_t rec = (t < n_fast && lane < 16u) ? reinterpret_cast<const uint32_t *>(recs + t)[lane] : 0u;

	while (t < n_fast) {
		if (lane == 0)
			t2 = atomicAdd(&b.ticket[1], 1u);
		const uint32_t nrec = (t1 < n_fast && lane < 16u) ? __ldg(reinterpret_cast<const uint32_t *>(recs + t1) + lane) : 0u;

#define AIRS_REC(i) __shfl_sync(kFull, rec, (i))
		const uint8_t *src = reinterpret_cast<const uint8_t *>((uintptr_t)(AIRS_REC(0) | (uint64_t)AIRS_REC(1) << 32));
		uint8_t *dst = reinterpret_cast<uint8_t *>((uintptr_t)(AIRS_REC(2) | (uint64_t)AIRS_REC(3) << 32));
		const uint32_t id_lo = AIRS_REC(4), id_hi = AIRS_REC(5);
		const uint32_t n = AIRS_REC(6), cap_eff = AIRS_REC(7), flags = AIRS_REC(8), first_result = AIRS_REC(9);
		const uint32_t g = AIRS_REC(10), outlier = AIRS_REC(11), L = (flags >> 8) & 15u;
		const uint32_t magic = AIRS_REC(12);
#undef AIRS_REC
		const bool multi = (flags & AIRS_FJ_MULTI) != 0u;
		const FK k = make_fk(multi, g, L, outlier, magic);

		const uint32_t a = (uint32_t)((uintptr_t)dst & 15u);
		const uint32_t checksum = (flags & AIRS_FJ_CHECKSUM) ? 1u : 0u;
		Out o;
		o.base = dst - a;
		o.lo = a;
		o.hi = a + cap_eff;
		o.gw0 = 0;
		o.sbits = 8u * a;

		/* the 22 header bytes travel through the staging words in front of the code words (size field zero,
		 * patched at the end), as five and a half big-endian words: ref cmp_hdr_serialize, header.c:24-67;
		 * fields cmp.c:265-279.  The area is all zero here, a is 0 or 8: plain stores */
		if (lane == 0) {
			uint32_t *h = ws.stg + a / 4u;
			const uint32_t pre = (flags & AIRS_FJ_PRE_DIFF) ? CMP_PREPROCESS_DIFF : (flags & AIRS_FJ_IWT) ? CMP_PREPROCESS_IWT : CMP_PREPROCESS_NONE;
			const uint32_t enc = multi ? CMP_ENCODER_GOLOMB_MULTI : CMP_ENCODER_GOLOMB_ZERO;
			h[0] = (0x8000u | CMP_VERSION_NUMBER) << 16;
			h[1] = (2u * n) & 0xFFFFFFu;
			h[2] = (id_hi << 16) | (id_lo >> 16);
			h[3] = (id_lo << 16) | (pre << 4) | (checksum << 3) | enc;
			h[4] = (g << 8) | ((outlier >> 16) & 0xFFu);
			h[5] = outlier << 16;
		}
		__syncwarp();
		o.sbits += 8u * (CMP_HDR_SIZE + 6u);

		const bool be = BE && (flags & AIRS_FJ_BE) != 0u;
		if (multi) {
			if (flags & AIRS_FJ_PRE_DIFF)
				encode_units<true, true, BE>(dbg, k, ws, o, src, n, lane, be);
			else
				encode_units<true, false, BE>(dbg, k, ws, o, src, n, lane, be);
		} else {
			if (flags & AIRS_FJ_PRE_DIFF)
				encode_units<false, true, BE>(dbg, k, ws, o, src, n, lane, be);
			else
				encode_units<false, false, BE>(dbg, k, ws, o, src, n, lane, be);
		}

		const uint32_t frame_bits = o.gw0 * 32u + o.sbits - 8u * a;
		const uint32_t payload_end = (frame_bits + 7u) >> 3; /* header + code bytes, zero padded (bitstream_writer.h:205-227) */
		const uint32_t size = payload_end + 4u * checksum;
		uint32_t result = size > cap_eff ? AIRS_ERR(DST_TOO_SMALL) : size;

		{ /* the last partial group: one byte per lane */
			const uint32_t nb = (o.sbits + 7u) >> 3, pos = o.gw0 * 4u + lane;
			if (lane < nb && pos >= o.lo && pos < o.hi)
				o.base[pos] = (uint8_t)(ws.stg[lane >> 2] >> (24u - 8u * (lane & 3u)));
			__syncwarp();
			if (lane < 4u)
				ws.stg[lane] = 0;
			__syncwarp();
		}
		if (!airs_failed(result)) {
			if (lane < 3u) /* the size field of the header that went out with the stream */
				dst[2u + lane] = (uint8_t)(size >> (16u - 8u * lane));
		} else if (flags & AIRS_FJ_FALLBACK_OK) {
			/* stored raw as a fresh primary pass, two more identifiers drawn (ref cmp.c:380-392): 16-byte
			 * header of a NONE + UNCOMPRESSED stream, the samples big endian (dst is 8-byte aligned) */
			const uint32_t raw_size = CMP_HDR_SIZE + 2u * n + 4u * checksum;
			const uint64_t id = ((((uint64_t)id_hi << 32) | id_lo) + 2u) & 0xFFFFFFFFFFFFull;
			uint32_t *out = reinterpret_cast<uint32_t *>(dst + CMP_HDR_SIZE);
			const uint32_t *in = reinterpret_cast<const uint32_t *>(src);
			for (uint32_t i = lane; i < n / 2u; i += 32u)
				out[i] = be ? __ldg(in + i) : airs_be_pair(__ldg(in + i));
			if (lane == 0 && (n & 1u)) {
				uint32_t xs = __ldg(reinterpret_cast<const uint16_t *>(src) + n - 1u);
				if (be)
					xs = ((xs << 8) | (xs >> 8)) & 0xFFFFu;
				dst[CMP_HDR_SIZE + 2u * (n - 1u)] = (uint8_t)(xs >> 8);
				dst[CMP_HDR_SIZE + 2u * (n - 1u) + 1u] = (uint8_t)xs;
			}
			if (lane < CMP_HDR_SIZE) {
				uint32_t v;
				switch (lane) {
				case 0: v = 0x80u | (CMP_VERSION_NUMBER >> 8); break;
				case 1: v = CMP_VERSION_NUMBER & 0xFFu; break;
				case 2: v = raw_size >> 16; break;
				case 3: v = raw_size >> 8; break;
				case 4: v = raw_size; break;
				case 5: v = (2u * n) >> 16; break;
				case 6: v = (2u * n) >> 8; break;
				case 7: v = 2u * n; break;
				case 14: v = 0; break;
				case 15: v = checksum << 3; break;
				default: v = (uint32_t)(id >> (8u * (13u - lane))); break;
				}
				dst[lane] = (uint8_t)v;
			}
			result = raw_size;
		}
		if (lane == 0)
			b.results[first_result] = result;

		t = t1;
		rec = nrec;
		t1 = __shfl_sync(kFull, t2, 0);
	}
}

extern "C" cudaError_t airs_launch_fast(const AirsLaunch *b, unsigned int grid, cudaStream_t stream)
{
	if (b->be_batch)
		airs_fast_kernel<true><<<grid, AIRS_FAST_THREADS, 0, stream>>>(*b);
	else
		airs_fast_kernel<false><<<grid, AIRS_FAST_THREADS, 0, stream>>>(*b);
	return cudaGetLastError();
}

/* AIRS_BOUNDS_CHECK builds: strings that would have been staged outside their warp's words since the last call
 * (-1: not such a build) */
extern "C" int airs_fast_bounds_violations(void)
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

/* resident CTAs of airs_fast_kernel on the current device */
extern "C" cudaError_t airs_fast_resident_ctas(int *out)
{
	int dev = 0, sms = 0, per_sm = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e == cudaSuccess)
		e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (e == cudaSuccess) {
		/* shared memory for the CTAs the kernel is compiled for (1 KiB per CTA is reserved by the system); the
		 * rest of the 228 KiB stays L1 */
		const size_t need = (size_t)AIRS_FAST_CTAS_PER_SM * (sizeof(FastWarp) * kFWarps + 1024);
		const int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
		e = cudaFuncSetAttribute(airs_fast_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
		if (e == cudaSuccess)
			e = cudaFuncSetAttribute(airs_fast_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
	}
	if (e == cudaSuccess) { /* (a job per warp by ticket: the grid need not be resident at once; the smaller of both counts) */
		int per_sm_be = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, airs_fast_kernel<false>, AIRS_FAST_THREADS, 0);
		if (e == cudaSuccess)
			e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_be, airs_fast_kernel<true>, AIRS_FAST_THREADS, 0);
		per_sm = per_sm_be < per_sm ? per_sm_be : per_sm;
	}
	*out = sms * per_sm;
	return e;
}
