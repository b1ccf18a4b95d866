/*
 * airs_kernels.cu - the sm_100a compression kernels.
 *
 * airs_plan_kernel    one thread per job: cmp_initialise validation and every
 *                     constant that follows from the parameters (airs_plan.cuh).
 * airs_encode_kernel  persistent CTAs of 128 threads; a CTA takes one job (one
 *                     compression context, lib/cmp.h:129-137) at a time from a
 *                     ticket counter and pushes its frames through in order.
 *
 * Per 2048-sample tile, 16 consecutive samples per thread:
 *   2 x 128-bit loads -> packed 16x2 residuals (VIADD.16x2: none / diff / IWT
 *   coefficient / model) -> packed zig-zag (PRMT sign replicate) -> Golomb /
 *   escape codeword and length per sample, branch-free, in registers -> the two
 *   codewords of a sample pair merged into one bit string -> warp-shuffle +
 *   REDUX block scan over bit counts -> every thread shifts its strings into
 *   place with funnel shifts and ORs 32-bit words of the MSB-first stream into
 *   shared memory (shared-memory atomicOr: 2.3 cycles per warp instruction,
 *   tools/micro/atoms.cu) -> coalesced byte-swapped stores.  Two block barriers
 *   per tile.  Header, padding and XXH32 trailer are written when the size is
 *   known.  Ragged, unaligned or i32-container tiles take tile_generic (same
 *   results, rolled loops).
 *
 * Reference being replaced: compress_engine and cmp_compress_generic
 * (lib/compress/cmp.c:213-393), preprocess.c:268-411, encoder.c:274-378,
 * bitstream_writer.h:124-227, header.c:24-67,137-163.  See DESIGN.md.
 */
#include <cuda_runtime.h>

#include "airs_device.cuh"
#include "airs_launch.h"
#include "airs_plan.cuh"
#include "airs_private.h"

namespace {

constexpr uint32_t kThreads = AIRS_THREADS; /* 128 */
constexpr uint32_t kSpt = 16;               /* samples per thread and tile */
constexpr uint32_t kTile = kThreads * kSpt; /* 2048 samples = 4 KiB of u16 */
constexpr uint32_t kWarps = kThreads / 32;
constexpr uint32_t kStgWords = kTile * 48 / 32 + 16;
constexpr uint64_t kMask48 = 0xFFFFFFFFFFFFull;
constexpr uint32_t kFull = 0xFFFFFFFFu;

/* one pass over one frame, written by thread 0 (cheap: copies from the plan) */
struct Pass {
	EncConst enc;
	const uint8_t *src;
	uint8_t *dst;
	uint16_t *work;
	uint64_t identifier;
	uint32_t pre;
	uint32_t n;
	uint32_t dtype;
	uint32_t hdr_len;
	uint32_t cap_eff;
	uint32_t trip;
	uint32_t model_mode; /* 0: none, 1: model := samples, 2: model update */
	uint32_t rate;
	uint32_t is_signed;
	uint32_t checksum;
	uint32_t seq;
	uint32_t err;
};

/* context state between frames: the mutable part of struct cmp_context */
struct CtxState {
	uint64_t identifier;
	uint64_t counter;
	uint32_t seq;
	uint32_t model_size;
};

struct Shared {
	uint32_t stg[kStgWords]; /* MSB-first 32-bit words of the stream being assembled (all zero when idle) */
	uint32_t wsum[kWarps];
	JobPlan plan;
	airs_job job;
	Pass pass;
	CtxState ctx;
	uint64_t offset;
	uint32_t ticket;
	uint32_t checksum;
};

/* byte window of the destination a pass may write, in the 4-byte aligned space
 * that starts at dst - (dst & 3) */
struct OutWin {
	uint8_t *base4;
	uint32_t lo, hi;
};

/* position of the stream under construction: stg[0] is word gw0 of the aligned
 * space and already holds sbits bits */
struct Cursor {
	uint32_t gw0, sbits;
};

/* -------------------------------------------------------------------------
 * thread-0 logic per frame (ref cmp.c:228-294); all heavy lifting is in the plan
 * ---------------------------------------------------------------------- */

__device__ void ctx_reset(CtxState &c) /* ref cmp_reset, cmp.c:452-465 */
{
	c.seq = 0;
	c.identifier = c.counter++ & kMask48;
	c.model_size = 0;
}

__device__ __noinline__ void plan_pass(Shared &sh, bool forced_raw, bool align_check)
{
	const JobPlan &pl = sh.plan;
	CtxState &c = sh.ctx;
	Pass &P = sh.pass;
	const uint32_t packed = pl.n * 2u;
	uint32_t sel;

	P.err = 0;
	if (c.seq == 0 || c.seq > pl.sec_iter) {
		ctx_reset(c);
		sel = 0;
		c.model_size = packed;
	} else {
		sel = 1;
		if ((pl.flags & AIRS_PF_MODEL) && packed != c.model_size) {
			P.err = AIRS_ERR(SRC_SIZE_MISMATCH);
			return;
		}
	}
	P.enc = pl.enc[sel];
	P.pre = pl.pre[sel];
	if (forced_raw) { /* ref cmp.c:383-386 */
		P.pre = CMP_PREPROCESS_NONE;
		P.enc.type = CMP_ENCODER_UNCOMPRESSED;
	}
	P.model_mode = 0;
	if (pl.flags & AIRS_PF_MODEL) {
		if (pl.model_err) {
			P.err = pl.model_err;
			return;
		}
		P.model_mode = c.seq == 0 ? 1u : 2u;
	}
	if (!P.dst) { /* ref bitstream_writer.h:65-68 */
		P.err = AIRS_ERR(DST_NULL);
		return;
	}
	if (align_check && ((uintptr_t)P.dst & 7u)) {
		P.err = AIRS_ERR(DST_UNALIGNED);
		return;
	}
	if (pl.orig_err) {
		P.err = pl.orig_err;
		return;
	}
	P.hdr_len = (P.pre == CMP_PREPROCESS_NONE && P.enc.type == CMP_ENCODER_UNCOMPRESSED) ? 16u : 22u;
	if (P.hdr_len > pl.cap_eff) { /* the placeholder header does not fit: header.c:62 */
		P.err = AIRS_ERR(DST_TOO_SMALL);
		return;
	}
	if (!forced_raw && pl.pre_err[sel]) {
		P.err = pl.pre_err[sel];
		return;
	}
	P.identifier = c.identifier;
	P.seq = c.seq;
}

__device__ __noinline__ void plan_frame(Shared &sh, const AirsLaunch &b, uint32_t frame)
{
	const airs_job &j = sh.job;
	const JobPlan &pl = sh.plan;
	Pass &P = sh.pass;

	P.err = pl.frame_err;
	if (P.err)
		return;
	P.src = b.src + j.src_offset + (uint64_t)frame * j.src_frame_stride;
	P.dst = b.dst ? b.dst + j.dst_offset + (uint64_t)frame * j.dst_frame_stride : nullptr;
	P.work = (b.work && j.work_size) ? (uint16_t *)(b.work + j.work_offset) : nullptr;
	P.dtype = j.dtype;
	P.n = pl.n;
	P.cap_eff = pl.cap_eff;
	P.trip = pl.trip;
	P.rate = pl.rate;
	P.is_signed = (pl.flags & AIRS_PF_SIGNED) ? 1u : 0u;
	P.checksum = (pl.flags & AIRS_PF_CHECKSUM) ? 1u : 0u;
	plan_pass(sh, false, b.layout == AIRS_LAYOUT_SLOTS);
}

/* header byte k (ref cmp_hdr_serialize, header.c:24-67; fields cmp.c:265-279) */
__device__ uint32_t header_byte(const Pass &P, uint32_t k, uint32_t size)
{
	const uint32_t orig = P.n * 2u;

	switch (k) {
	case 0: return 0x80u | (CMP_VERSION_NUMBER >> 8);
	case 1: return CMP_VERSION_NUMBER & 0xFF;
	case 2: return size >> 16;
	case 3: return size >> 8;
	case 4: return size;
	case 5: return orig >> 16;
	case 6: return orig >> 8;
	case 7: return orig;
	case 8: case 9: case 10: case 11: case 12: case 13:
		return (uint32_t)(P.identifier >> (8 * (13 - k)));
	case 14: return P.seq;
	case 15: return (P.pre << 4) | (P.checksum << 3) | P.enc.type;
	case 16: return P.pre == CMP_PREPROCESS_MODEL ? P.rate : 0;
	case 17: return P.enc.type != CMP_ENCODER_UNCOMPRESSED ? P.enc.g >> 8 : 0;
	case 18: return P.enc.type != CMP_ENCODER_UNCOMPRESSED ? P.enc.g : 0;
	case 19: return P.enc.type != CMP_ENCODER_UNCOMPRESSED ? P.enc.outlier >> 16 : 0;
	case 20: return P.enc.type != CMP_ENCODER_UNCOMPRESSED ? P.enc.outlier >> 8 : 0;
	default: return P.enc.type != CMP_ENCODER_UNCOMPRESSED ? P.enc.outlier : 0;
	}
}

/* -------------------------------------------------------------------------
 * sample access (generic path)
 * ---------------------------------------------------------------------- */

__device__ __forceinline__ uint32_t sample_at(const uint8_t *src, uint32_t dtype, uint32_t i)
{
	/* ref sample_read_i16, sample_reader.h:63-72 */
	if (dtype == AIRS_DTYPE_I16_IN_I32)
		return __ldg((const uint32_t *)src + i) & 0xFFFFu;
	return __ldg((const uint16_t *)src + i);
}

/* -------------------------------------------------------------------------
 * IWT over the work buffer (ref preprocess.c:140-221).  Per level: all detail
 * coefficients (odd multiples of the stride) from the untouched even
 * neighbours, barrier, then all approximations from the new details - the same
 * values the reference's in-place sequential sweep produces.
 * ---------------------------------------------------------------------- */
__device__ __forceinline__ int16_t wrap16(int32_t v)
{
	return (int16_t)(uint16_t)(uint32_t)v;
}

__device__ __noinline__ void iwt_global(const Pass &P)
{
	int16_t *w = (int16_t *)P.work;
	const uint64_t n = P.n;
	const uint32_t tid = threadIdx.x;

	for (uint64_t i = tid; i < n; i += kThreads)
		w[i] = (int16_t)sample_at(P.src, P.dtype, (uint32_t)i);
	__syncthreads();
	for (uint64_t s = 1; s < n; s <<= 1) {
		for (uint64_t i = s + 2 * s * tid; i < n; i += 2 * s * kThreads) {
			if (i + s < n)
				w[i] = wrap16(w[i] - wrap16(((int32_t)w[i - s] + w[i + s]) >> 1));
			else
				w[i] = wrap16(w[i] - w[i - s]);
		}
		__syncthreads();
		for (uint64_t i = 2 * s * tid; i < n; i += 2 * s * kThreads) {
			bool has_l = i >= s, has_r = i + s < n;
			if (has_l && has_r)
				w[i] = wrap16(w[i] + wrap16(((int32_t)w[i - s] + w[i + s]) >> 2));
			else if (has_r)
				w[i] = wrap16(w[i] + wrap16((int32_t)w[i + s] >> 1));
			else if (has_l)
				w[i] = wrap16(w[i] + wrap16((int32_t)w[i - s] >> 1));
		}
		__syncthreads();
	}
}

/* -------------------------------------------------------------------------
 * XXH32 of the big-endian samples (ref cmp_checksum, header.c:137-163): the
 * four lanes of the hash run on lanes 0-3 of the calling warp.
 * ---------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t pair_at(const Pass &P, bool al4, uint32_t i)
{
	if (P.dtype == AIRS_DTYPE_I16_IN_I32) {
		const uint32_t *p = (const uint32_t *)P.src;
		return (__ldg(p + i) & 0xFFFFu) | (__ldg(p + i + 1) << 16);
	}
	if (al4)
		return __ldg((const uint32_t *)((const uint16_t *)P.src + i));
	const uint16_t *p = (const uint16_t *)P.src;
	return (uint32_t)__ldg(p + i) | ((uint32_t)__ldg(p + i + 1) << 16);
}

__device__ __noinline__ uint32_t frame_checksum(const Pass &P)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t n = P.n, nbytes = n * 2u;
	const uint32_t stripes = n / 8u;
	const bool al4 = ((uintptr_t)P.src & 3u) == 0;
	const uint32_t seed = AIRS_CHECKSUM_SEED;
	uint32_t v = lane == 0 ? seed + AIRS_XP1 + AIRS_XP2 : lane == 1 ? seed + AIRS_XP2 : lane == 2 ? seed : seed - AIRS_XP1;

	if (lane < 4) {
		uint32_t s = 0;
		for (; s + 4 <= stripes; s += 4) {
			uint32_t w0 = pair_at(P, al4, (s + 0) * 8 + 2 * lane);
			uint32_t w1 = pair_at(P, al4, (s + 1) * 8 + 2 * lane);
			uint32_t w2 = pair_at(P, al4, (s + 2) * 8 + 2 * lane);
			uint32_t w3 = pair_at(P, al4, (s + 3) * 8 + 2 * lane);
			v = airs_xxh_round(v, airs_be_pair(w0));
			v = airs_xxh_round(v, airs_be_pair(w1));
			v = airs_xxh_round(v, airs_be_pair(w2));
			v = airs_xxh_round(v, airs_be_pair(w3));
		}
		for (; s < stripes; s++)
			v = airs_xxh_round(v, airs_be_pair(pair_at(P, al4, s * 8 + 2 * lane)));
	}
	uint32_t v1 = __shfl_sync(kFull, v, 1);
	uint32_t v2 = __shfl_sync(kFull, v, 2);
	uint32_t v3 = __shfl_sync(kFull, v, 3);
	uint32_t h = 0;
	if (lane == 0) {
		h = nbytes >= 16 ? airs_rotl(v, 1) + airs_rotl(v1, 7) + airs_rotl(v2, 12) + airs_rotl(v3, 18)
				 : seed + AIRS_XP5;
		h += nbytes;
		uint32_t i = stripes * 8;
		for (; i + 2 <= n; i += 2)
			h = airs_rotl(h + airs_be_pair(pair_at(P, false, i)) * AIRS_XP3, 17) * AIRS_XP4;
		if (i < n) {
			uint32_t sv = sample_at(P.src, P.dtype, i);
			h = airs_rotl(h + (sv >> 8) * AIRS_XP5, 11) * AIRS_XP1;
			h = airs_rotl(h + (sv & 0xFFu) * AIRS_XP5, 11) * AIRS_XP1;
		}
		h ^= h >> 15;
		h *= AIRS_XP2;
		h ^= h >> 13;
		h *= AIRS_XP3;
		h ^= h >> 16;
	}
	return __shfl_sync(kFull, h, 0);
}

/* -------------------------------------------------------------------------
 * building blocks shared by the fast and the generic tile
 * ---------------------------------------------------------------------- */

/* per-thread bit writer into the staging words: up to 32 bits per push */
struct Packer {
	uint32_t lo;   /* pending bits, right aligned (bits above `fill` are stale) */
	uint32_t fill; /* number of pending bits, < 32 between pushes */
	uint32_t wp;   /* staging word the pending bits belong to */
};

__device__ __forceinline__ void packer_open(Packer &p, uint32_t bitpos)
{
	p.lo = 0;
	p.fill = bitpos & 31u;
	p.wp = bitpos >> 5;
}

__device__ __forceinline__ void packer_push(Packer &p, uint32_t *stg, uint32_t code, uint32_t len)
{
	/* (hi:lo) = (lo << len) | code; ref bitstream_add_bits32, bitstream_writer.h:124-158 */
	const uint32_t hi = __funnelshift_lc(p.lo, 0u, len);
	p.lo = __funnelshift_lc(0u, p.lo, len) | code;
	p.fill += len;
	if (p.fill >= 32u) {
		atomicOr(&stg[p.wp], __funnelshift_r(p.lo, hi, p.fill));
		p.wp++;
		p.fill -= 32u;
	}
}

__device__ __forceinline__ void packer_close(Packer &p, uint32_t *stg)
{
	if (p.fill)
		atomicOr(&stg[p.wp], p.lo << (32u - p.fill));
}

/* exclusive scan of per-thread bit counts over the CTA; one barrier */
__device__ __forceinline__ uint32_t block_scan(Shared &sh, uint32_t tb, uint32_t &total)
{
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t incl = tb;

#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t t = __shfl_up_sync(kFull, incl, d);
		if (lane >= (uint32_t)d)
			incl += t;
	}
	if (lane == 31)
		sh.wsum[warp] = incl;
	__syncthreads();
	uint32_t ws = lane < kWarps ? sh.wsum[lane] : 0u;
	total = __reduce_add_sync(kFull, ws);
	uint32_t wpre = __reduce_add_sync(kFull, lane < warp ? ws : 0u);
	return wpre + incl - tb;
}

__device__ __forceinline__ void store_word(const OutWin &o, uint32_t gword, uint32_t v)
{
	uint64_t b = (uint64_t)gword * 4;

	if (b >= o.lo && b + 4 <= o.hi) {
		*(uint32_t *)(o.base4 + b) = airs_bswap32(v);
	} else {
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (b + k >= o.lo && b + k < o.hi)
				o.base4[b + k] = (uint8_t)(v >> (24 - 8 * k));
	}
}

/* after the packing barrier: full staged words leave as coalesced stores, the
 * staging area is zeroed behind them, the trailing partial word moves to
 * stg[0].  No barrier afterwards: the next tile touches the staging words only
 * after its own scan barrier, which every thread reaches after its copy-out. */
__device__ __forceinline__ void copy_out(Shared &sh, const OutWin &o, Cursor &c, uint32_t tile_bits)
{
	const uint32_t tid = threadIdx.x;
	const uint32_t staged = c.sbits + tile_bits;
	const uint32_t wfull = staged >> 5;

	for (uint32_t w = tid; w < wfull; w += kThreads) {
		uint32_t v = sh.stg[w];
		sh.stg[w] = 0;
		store_word(o, c.gw0 + w, v);
	}
	if (tid == 0 && wfull) { /* thread 0 zeroed stg[0] itself; stg[wfull] is nobody else's */
		uint32_t carry = sh.stg[wfull];
		sh.stg[wfull] = 0;
		sh.stg[0] = carry;
	}
	c.gw0 += wfull;
	c.sbits = staged & 31u;
}

/* -------------------------------------------------------------------------
 * generic tile: any container, any alignment, ragged end.  Rolled loops; the
 * codewords are computed twice (lengths for the scan, bits for the packer).
 * ---------------------------------------------------------------------- */

__device__ __forceinline__ uint32_t residual_at(const Pass &P, uint32_t i, uint32_t x)
{
	switch (P.pre) { /* ref preprocess.c:268-290,348-353,406-411 */
	case CMP_PREPROCESS_DIFF:
		return i ? x - sample_at(P.src, P.dtype, i - 1) : x;
	case CMP_PREPROCESS_IWT:
		return P.work[i];
	case CMP_PREPROCESS_MODEL:
		return x - P.work[i];
	default:
		return x;
	}
}

__device__ __forceinline__ void encode_any(const EncConst &e, uint32_t r, uint32_t &cw, uint32_t &cl,
					   uint32_t &rw, uint32_t &rl)
{
	switch (e.type) {
	case CMP_ENCODER_UNCOMPRESSED:
		airs_encode<CMP_ENCODER_UNCOMPRESSED>(e, r, cw, cl, rw, rl);
		break;
	case CMP_ENCODER_GOLOMB_ZERO:
		airs_encode<CMP_ENCODER_GOLOMB_ZERO>(e, r, cw, cl, rw, rl);
		break;
	default:
		airs_encode<CMP_ENCODER_GOLOMB_MULTI>(e, r, cw, cl, rw, rl);
		break;
	}
}

__device__ __noinline__ void tile_generic(Shared &sh, const OutWin &o, uint32_t a, Cursor &c, uint32_t base,
					  bool size_only)
{
	const Pass &P = sh.pass;
	const EncConst e = P.enc;
	const uint32_t tid = threadIdx.x;
	const uint32_t n = P.n;
	const uint32_t i0 = min(base + tid * kSpt, n);
	const uint32_t i1 = min(i0 + kSpt, n);
	const bool need_x = P.pre != CMP_PREPROCESS_IWT || P.model_mode;
	uint32_t tb = 0;

	for (uint32_t i = i0; i < i1; i++) {
		uint32_t x = need_x ? sample_at(P.src, P.dtype, i) : 0u;
		uint32_t cw, cl, rw, rl;
		encode_any(e, residual_at(P, i, x), cw, cl, rw, rl);
		tb += cl + rl;
	}
	uint32_t tile_bits;
	const uint32_t excl = block_scan(sh, tb, tile_bits);
	if (size_only) {
		const uint32_t staged = c.sbits + tile_bits;
		c.gw0 += staged >> 5;
		c.sbits = staged & 31u;
		__syncthreads();
		return;
	}
	Packer pk;
	packer_open(pk, c.sbits + excl);
	uint32_t cum = c.gw0 * 32u + c.sbits + excl - 8u * a; /* stream bits before this thread's samples */
	for (uint32_t i = i0; i < i1; i++) {
		uint32_t x = need_x ? sample_at(P.src, P.dtype, i) : 0u;
		uint32_t m = (P.pre == CMP_PREPROCESS_MODEL || P.model_mode == 2) ? (uint32_t)P.work[i] : 0u;
		uint32_t cw, cl, rw, rl;
		encode_any(e, residual_at(P, i, x), cw, cl, rw, rl);
		packer_push(pk, sh.stg, cw, cl);
		if (rl)
			packer_push(pk, sh.stg, rw, rl);
		cum += cl + rl;
		/* model := samples, or model update, while the reference's writer has not
		 * given up (ref cmp.c:300-311) */
		if (P.model_mode && cum < P.trip)
			P.work[i] = (uint16_t)(P.model_mode == 1 ? x : airs_model_update(x, m, P.rate, P.is_signed));
	}
	if (tb)
		packer_close(pk, sh.stg);
	__syncthreads();
	copy_out(sh, o, c, tile_bits);
}

/* -------------------------------------------------------------------------
 * fast tile: 2048 samples, 16-bit container, 16-byte aligned source (and work
 * buffer when used).  Everything lives in registers; all loops are unrolled.
 * ---------------------------------------------------------------------- */

/* packed 16x2 zig-zag: ref map_to_unsigned, encoder.c:274-286 */
__device__ __forceinline__ uint32_t zigzag2(uint32_t d)
{
	uint32_t sign; /* per 16-bit lane: 0xFFFF if negative */
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sign) : "r"(d), "r"(0u), "r"(0xBB99u));
	return ((d << 1) & 0xFFFEFFFEu) ^ sign;
}

__device__ __forceinline__ void load8(const uint16_t *p, uint32_t w[8], bool ro)
{
	const uint4 *q = (const uint4 *)p;
	uint4 a = ro ? __ldg(q) : q[0], b = ro ? __ldg(q + 1) : q[1];
	w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
	w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

template <int ENC, int PRE, int MODEL>
__device__ __noinline__ void tile_fast(Shared &sh, const OutWin &o, uint32_t a, Cursor &c, uint32_t base,
				       bool size_only)
{
	const Pass &P = sh.pass;
	const EncConst e = P.enc;
	const uint32_t tid = threadIdx.x;
	const uint32_t i0 = base + tid * kSpt;
	const uint16_t *src = (const uint16_t *)P.src;
	uint32_t w[8], mw[8], d[8];

	if (PRE != CMP_PREPROCESS_IWT || MODEL)
		load8(src + i0, w, true);
	if (PRE == CMP_PREPROCESS_MODEL || MODEL == 2)
		load8(P.work + i0, mw, false);

	/* packed residuals: ref preprocess.c:268-290,348-353,406-411 */
	if (PRE == CMP_PREPROCESS_NONE) {
#pragma unroll
		for (int k = 0; k < 8; k++)
			d[k] = w[k];
	} else if (PRE == CMP_PREPROCESS_DIFF) {
		uint32_t pw = i0 ? __ldg((const uint32_t *)(src + i0 - 2)) : 0u; /* [x(i0-2), x(i0-1)] */
#pragma unroll
		for (int k = 0; k < 8; k++) {
			uint32_t prev = __funnelshift_l(k ? w[k - 1] : pw, w[k], 16); /* [x(2k-1), x(2k)] */
			d[k] = __vsub2(w[k], prev);
		}
	} else if (PRE == CMP_PREPROCESS_IWT) {
		load8(P.work + i0, d, false);
	} else {
#pragma unroll
		for (int k = 0; k < 8; k++)
			d[k] = __vsub2(w[k], mw[k]);
	}

	/* codewords; the two samples of a word are merged into one string of pl bits:
	 * (ph:pc) = code_lo << len_hi | code_hi */
	uint32_t pc[8], ph[8], pl[8], tb = 0, mx = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) {
		if (ENC == CMP_ENCODER_UNCOMPRESSED) {
			pc[k] = __byte_perm(d[k], 0, 0x1032); /* first sample in the upper half */
			ph[k] = 0;
			pl[k] = 32;
		} else if (ENC == CMP_ENCODER_GOLOMB_ZERO) {
			const uint32_t z = zigzag2(d[k]);
			const uint32_t m0 = z & 0xFFFFu, m1 = z >> 16;
			uint32_t g0, g1, gl0, gl1;
			airs_golomb(e, m0 + 1u, g0, gl0);
			airs_golomb(e, m1 + 1u, g1, gl1);
			const bool e0 = m0 >= e.outlier, e1 = m1 >= e.outlier;
			const uint32_t c0 = e0 ? m0 : g0, l0 = e0 ? e.L + 17u : gl0;
			const uint32_t c1 = e1 ? m1 : g1, l1 = e1 ? e.L + 17u : gl1;
			pc[k] = __funnelshift_lc(0u, c0, l1) | c1;
			ph[k] = __funnelshift_lc(c0, 0u, l1);
			pl[k] = l0 + l1;
		} else {
			/* escapes carry a raw part: samples are 64-bit strings, the pair fits
			 * 64 bits or is flagged through mx */
			const uint32_t z = zigzag2(d[k]);
			uint32_t c0, l0, r0, q0, c1, l1, r1, q1;
			airs_encode_mapped<CMP_ENCODER_GOLOMB_MULTI>(e, z & 0xFFFFu, c0, l0, r0, q0);
			airs_encode_mapped<CMP_ENCODER_GOLOMB_MULTI>(e, z >> 16, c1, l1, r1, q1);
			const uint64_t s0 = ((uint64_t)c0 << q0) | r0, s1 = ((uint64_t)c1 << q1) | r1;
			const uint32_t sl0 = l0 + q0, sl1 = l1 + q1;
			const uint64_t s = sl1 < 64u ? (s0 << sl1) | s1 : 0;
			pc[k] = (uint32_t)s;
			ph[k] = (uint32_t)(s >> 32);
			pl[k] = sl0 + sl1;
		}
		tb += pl[k];
		mx = max(mx, pl[k]);
	}
	/* 0: every pair of the warp fits 32 bits, 1: 64 bits, 2: not even that (rare) */
	const uint32_t wide = __reduce_max_sync(kFull, mx > 64u ? 2u : (mx > 32u ? 1u : 0u));

	uint32_t tile_bits;
	const uint32_t excl = block_scan(sh, tb, tile_bits);
	if (size_only) {
		const uint32_t staged = c.sbits + tile_bits;
		c.gw0 += staged >> 5;
		c.sbits = staged & 31u;
		__syncthreads();
		return;
	}
	const uint32_t tile_end = c.gw0 * 32u + c.sbits + tile_bits - 8u * a; /* stream bits after this tile */

	if (__syncthreads_or(wide == 2u)) {
		/* a pair longer than 64 bits (multi-escape pile-up): redo the tile the slow way.
		 * Nothing has been staged yet. */
		tile_generic(sh, o, a, c, base, false);
		return;
	}

	Packer pk;
	packer_open(pk, c.sbits + excl);
	if (wide == 0u) {
#pragma unroll
		for (int k = 0; k < 8; k++)
			packer_push(pk, sh.stg, pc[k], pl[k]);
	} else {
#pragma unroll
		for (int k = 0; k < 8; k++) {
			const uint32_t hl = pl[k] > 32u ? pl[k] - 32u : 0u;
			packer_push(pk, sh.stg, ph[k], hl);
			packer_push(pk, sh.stg, pc[k], pl[k] - hl);
		}
	}
	packer_close(pk, sh.stg);

	/* model := samples, or model update (ref cmp.c:304-311) */
	if (MODEL) {
		if (tile_end < P.trip) {
			uint32_t nm[8];
#pragma unroll
			for (int k = 0; k < 8; k++) {
				if (MODEL == 1) {
					nm[k] = w[k];
				} else {
					uint32_t lo = airs_model_update(w[k] & 0xFFFFu, mw[k] & 0xFFFFu, P.rate, P.is_signed);
					uint32_t hi = airs_model_update(w[k] >> 16, mw[k] >> 16, P.rate, P.is_signed);
					nm[k] = lo | (hi << 16);
				}
			}
			uint4 *q = (uint4 *)(P.work + i0);
			q[0] = make_uint4(nm[0], nm[1], nm[2], nm[3]);
			q[1] = make_uint4(nm[4], nm[5], nm[6], nm[7]);
		} else {
			/* the stream overflows its capacity inside this tile: per-sample gate */
			uint32_t cum = c.gw0 * 32u + c.sbits + excl - 8u * a;
			for (uint32_t i = i0; i < i0 + kSpt; i++) {
				uint32_t x = sample_at(P.src, P.dtype, i);
				uint32_t m = P.work[i];
				uint32_t cw, cl, rw, rl;
				encode_any(e, residual_at(P, i, x), cw, cl, rw, rl);
				cum += cl + rl;
				if (cum < P.trip)
					P.work[i] = (uint16_t)(MODEL == 1 ? x : airs_model_update(x, m, P.rate, P.is_signed));
			}
		}
	}
	__syncthreads();
	copy_out(sh, o, c, tile_bits);
}

template <int ENC, int PRE>
__device__ __forceinline__ void tile_fast_model(Shared &sh, const OutWin &o, uint32_t a, Cursor &c, uint32_t base,
						bool size_only, uint32_t model_mode)
{
	if (model_mode == 0)
		tile_fast<ENC, PRE, 0>(sh, o, a, c, base, size_only);
	else if (model_mode == 1)
		tile_fast<ENC, PRE, 1>(sh, o, a, c, base, size_only);
	else
		tile_fast<ENC, PRE, 2>(sh, o, a, c, base, size_only);
}

template <int ENC>
__device__ __forceinline__ void tile_fast_pre(Shared &sh, const OutWin &o, uint32_t a, Cursor &c, uint32_t base,
					      bool size_only, uint32_t pre, uint32_t model_mode)
{
	switch (pre) {
	case CMP_PREPROCESS_NONE:
		tile_fast_model<ENC, CMP_PREPROCESS_NONE>(sh, o, a, c, base, size_only, model_mode);
		break;
	case CMP_PREPROCESS_DIFF:
		tile_fast_model<ENC, CMP_PREPROCESS_DIFF>(sh, o, a, c, base, size_only, model_mode);
		break;
	case CMP_PREPROCESS_IWT:
		tile_fast_model<ENC, CMP_PREPROCESS_IWT>(sh, o, a, c, base, size_only, model_mode);
		break;
	default: /* MODEL preprocessing only happens in secondary passes: model_mode == 2 */
		tile_fast<ENC, CMP_PREPROCESS_MODEL, 2>(sh, o, a, c, base, size_only);
		break;
	}
}

/* one pass over one frame; returns the stream size or an error (uniform over the CTA).
 * ref compress_engine, cmp.c:213-338 */
__device__ uint32_t encode_pass(Shared &sh, bool size_only, bool suppress)
{
	const Pass &P = sh.pass;
	const uint32_t tid = threadIdx.x;

	if (P.err)
		return P.err;

	const uint32_t a = (uint32_t)((uintptr_t)P.dst & 3u);
	OutWin o;
	o.base4 = P.dst - a;
	o.lo = a + P.hdr_len;
	o.hi = suppress ? o.lo : a + P.cap_eff; /* suppress: run for the model side effects only */
	Cursor c;
	c.gw0 = (8u * (a + P.hdr_len)) >> 5;
	c.sbits = (8u * (a + P.hdr_len)) & 31u;

	if (P.pre == CMP_PREPROCESS_IWT)
		iwt_global(P);

	const uint32_t n = P.n, pre = P.pre, enc = P.enc.type, model_mode = P.model_mode;
	const bool fast_ok = P.dtype != AIRS_DTYPE_I16_IN_I32 && ((uintptr_t)P.src & 15u) == 0 &&
			     (((uintptr_t)P.work & 15u) == 0 || (pre < CMP_PREPROCESS_IWT && !model_mode));
	for (uint32_t base = 0; base < n; base += kTile) {
		if (fast_ok && base + kTile <= n) {
			if (enc == CMP_ENCODER_UNCOMPRESSED)
				tile_fast_pre<CMP_ENCODER_UNCOMPRESSED>(sh, o, a, c, base, size_only, pre, model_mode);
			else if (enc == CMP_ENCODER_GOLOMB_ZERO)
				tile_fast_pre<CMP_ENCODER_GOLOMB_ZERO>(sh, o, a, c, base, size_only, pre, model_mode);
			else
				tile_fast_pre<CMP_ENCODER_GOLOMB_MULTI>(sh, o, a, c, base, size_only, pre, model_mode);
		} else {
			tile_generic(sh, o, a, c, base, size_only);
		}
	}
	__syncthreads();

	const uint32_t frame_bits = c.gw0 * 32u + c.sbits - 8u * a;
	const uint32_t payload_end = (frame_bits + 7u) >> 3; /* header + code bytes */
	const uint32_t size = payload_end + (P.checksum ? 4u : 0u);
	uint32_t result;
	if (size > P.cap_eff)
		result = AIRS_ERR(DST_TOO_SMALL);
	else if (size > CMP_HDR_MAX_COMPRESSED_SIZE)
		result = AIRS_ERR(HDR_CMP_SIZE_TOO_LARGE);
	else
		result = size;
	if (size_only)
		return result;

	/* checksum of the samples while another warp flushes the tail */
	if (P.checksum && !suppress && tid < 32) {
		uint32_t h = frame_checksum(P);
		if (tid == 0)
			sh.checksum = h;
	}
	if (tid == 32) { /* last partial word, zero padded (ref bitstream_writer.h:205-227) */
		uint32_t v = sh.stg[0];
		uint32_t nb = (c.sbits + 7u) >> 3;
		for (uint32_t k = 0; k < nb; k++) {
			uint64_t b = (uint64_t)c.gw0 * 4 + k;
			if (b >= o.lo && b < o.hi)
				o.base4[b] = (uint8_t)(v >> (24 - 8 * k));
		}
		sh.stg[0] = 0;
	}
	__syncthreads();

	if (P.checksum && !suppress && tid < 4) { /* trailer, big endian (ref cmp.c:314-319) */
		uint64_t b = (uint64_t)a + payload_end + tid;
		if (b < o.hi)
			o.base4[b] = (uint8_t)(sh.checksum >> (24 - 8 * tid));
	}
	if (!airs_failed(result) && !suppress && tid < P.hdr_len) /* header with the final size (ref cmp.c:329-334) */
		P.dst[tid] = (uint8_t)header_byte(P, tid, size);
	return result;
}

/* -------------------------------------------------------------------------
 * CONCAT layout: single-pass device-wide scan over stream sizes (decoupled
 * look-back over one 64-bit status word per frame: 2 flag bits | 62 value bits;
 * flag 1 = this frame's size, flag 2 = inclusive prefix).  Called by warp 0;
 * returns the byte offset of frame k.  Frames are published in result-index
 * order by CTAs that took their jobs from the ticket counter in order, so every
 * predecessor is running or done.
 * ---------------------------------------------------------------------- */
__device__ uint64_t lookback_offset(volatile uint64_t *st, uint32_t k, uint32_t my_size)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint64_t kValue = (1ull << 62) - 1;
	uint64_t excl = 0;

	if (lane == 0)
		st[k] = (1ull << 62) | my_size;
	for (int64_t idx = (int64_t)k - 1; idx >= 0; idx -= 32) {
		const int64_t j = idx - lane;
		uint64_t v;
		do {
			v = j >= 0 ? st[j] : (2ull << 62);
		} while (__any_sync(kFull, (v >> 62) == 0));
		const uint32_t incl = __ballot_sync(kFull, (v >> 62) == 2);
		if (incl) {
			const int first = __ffs((int)incl) - 1; /* nearest predecessor holding a prefix */
			excl += __reduce_add_sync(kFull, (int)lane < first ? (uint32_t)(v & kValue) : 0u);
			const uint64_t pv = v & kValue;
			const uint32_t plo = __shfl_sync(kFull, (uint32_t)pv, first);
			const uint32_t phi = __shfl_sync(kFull, (uint32_t)(pv >> 32), first);
			excl += ((uint64_t)phi << 32) | plo;
			break;
		}
		excl += __reduce_add_sync(kFull, (uint32_t)(v & kValue));
	}
	if (lane == 0)
		st[k] = (2ull << 62) | (excl + my_size);
	return excl;
}

} /* namespace */

__global__ void __launch_bounds__(128) airs_plan_kernel(AirsLaunch b)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;

	if (j >= b.n_jobs)
		return;
	JobPlan pl;
	airs_make_plan(pl, b.jobs[j], b.src, b.work);
	b.plans[j] = pl;
	if (b.init_results && !b.ctx_io)
		b.init_results[j] = pl.init_result;
}

__global__ void __launch_bounds__(AIRS_THREADS, 6) airs_encode_kernel(AirsLaunch b)
{
	__shared__ Shared sh;
	const uint32_t tid = threadIdx.x;

	for (uint32_t w = tid; w < kStgWords; w += kThreads)
		sh.stg[w] = 0;
	if (tid == 0)
		sh.ticket = atomicAdd(b.ticket, 1u);

	for (;;) {
		__syncthreads();
		const uint32_t job = sh.ticket;
		if (job >= b.n_jobs)
			break;
		/* plan (32 words) and job descriptor (30 words) into shared memory, coalesced */
		if (tid < 32)
			((uint32_t *)&sh.plan)[tid] = ((const uint32_t *)&b.plans[job])[tid];
		else if (tid < 62)
			((uint32_t *)&sh.job)[tid - 32] = ((const uint32_t *)&b.jobs[job])[tid - 32];
		__syncthreads();
		if (tid == 0) {
			sh.ticket = atomicAdd(b.ticket, 1u); /* next job, fetched while this one runs */
			CtxState &c = sh.ctx;
			if (b.ctx_io) { /* host shim: continue the caller's context */
				const airs_ctx_state &st = b.ctx_io[job];
				c.identifier = st.identifier;
				c.counter = st.counter;
				c.seq = st.seq;
				c.model_size = st.model_size;
			} else { /* fresh context: the cmp_reset at the end of cmp_initialise */
				c.counter = sh.job.identifier_base;
				ctx_reset(c);
			}
		}
		const uint32_t n_frames = sh.job.n_frames;
		const uint32_t first = sh.job.first_result;

		for (uint32_t f = 0; f < n_frames; f++) {
			if (tid == 0)
				plan_frame(sh, b, f);
			__syncthreads();
			uint32_t r;
			if (b.layout == AIRS_LAYOUT_CONCAT) {
				/* size first (exact, no output), then the offset from the scan, then one
				 * pass that writes; a frame that fails contributes no bytes */
				r = encode_pass(sh, true, false);
				if ((sh.plan.flags & AIRS_PF_FALLBACK_OK) && r == AIRS_ERR(DST_TOO_SMALL)) {
					__syncthreads();
					if (tid == 0) {
						ctx_reset(sh.ctx);
						plan_pass(sh, true, false);
					}
					__syncthreads();
					r = sh.pass.err ? sh.pass.err : sh.plan.raw_size;
				}
				const uint32_t k = first + f;
				if (tid < 32) {
					uint64_t off = lookback_offset(b.lookback, k, airs_failed(r) ? 0u : r);
					if (tid == 0) {
						sh.offset = off;
						b.out_offsets[k] = off;
						if (k + 1 == b.n_results)
							b.out_offsets[k + 1] = off + (airs_failed(r) ? 0u : r);
						sh.pass.dst = b.dst + off;
					}
				}
				__syncthreads();
				const bool fits = !airs_failed(r) && sh.offset + r <= b.dst_size;
				if (!sh.pass.err) {
					uint32_t r2 = encode_pass(sh, false, !fits);
					if (!airs_failed(r))
						r = fits ? r2 : AIRS_ERR(DST_TOO_SMALL);
				}
			} else {
				r = encode_pass(sh, false, false);
				if ((sh.plan.flags & AIRS_PF_FALLBACK_OK) && r == AIRS_ERR(DST_TOO_SMALL)) {
					/* store the frame raw as a fresh primary pass (ref cmp.c:380-392) */
					__syncthreads();
					if (tid == 0) {
						ctx_reset(sh.ctx);
						plan_pass(sh, true, true);
					}
					__syncthreads();
					r = encode_pass(sh, false, false);
				}
			}
			__syncthreads();
			if (tid == 0) {
				if (!airs_failed(r))
					sh.ctx.seq = (sh.ctx.seq + 1u) & 0xFFu;
				b.results[first + f] = r;
			}
		}
		if (tid == 0 && b.ctx_io) {
			airs_ctx_state &st = b.ctx_io[job];
			st.identifier = sh.ctx.identifier;
			st.counter = sh.ctx.counter;
			st.seq = sh.ctx.seq;
			st.model_size = sh.ctx.model_size;
		}
	}
}

extern "C" cudaError_t airs_launch_plan(const AirsLaunch *b, cudaStream_t stream)
{
	airs_plan_kernel<<<(b->n_jobs + 127) / 128, 128, 0, stream>>>(*b);
	return cudaGetLastError();
}

extern "C" cudaError_t airs_launch_encode(const AirsLaunch *b, unsigned int grid, cudaStream_t stream)
{
	airs_encode_kernel<<<grid, AIRS_THREADS, 0, stream>>>(*b);
	return cudaGetLastError();
}
