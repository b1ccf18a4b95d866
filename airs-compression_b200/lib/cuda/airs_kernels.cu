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

#ifdef AIRS_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[16];
#define PHASE_T(var) const long long var = clock64()
#define PHASE_ADD(i, a, b) do { if (threadIdx.x == 0) atomicAdd(&g_phase_clk[i], (unsigned long long)((b) - (a))); } while (0)
#else
#define PHASE_T(var)
#define PHASE_ADD(i, a, b)
#endif

namespace {

constexpr uint32_t kThreads = AIRS_THREADS; /* 128 */
constexpr uint32_t kSpt = AIRS_SPT;         /* samples per thread and tile */
constexpr uint32_t kPairs = kSpt / 2;       /* packed 16x2 words per thread */
constexpr uint32_t kVec = kSpt / 8;         /* 16-byte vectors per thread */
constexpr uint32_t kTile = kThreads * kSpt; /* 2048 samples = 4 KiB of u16 */
static_assert(kTile == 2048 && (kSpt == 8 || kSpt == 16), "tile geometry");
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
	uint32_t lut_sel;   /* which table of Shared::lut this pass uses */
	uint32_t lut_range; /* its half range R, 0: none */
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
	uint4 in[2][kThreads * (kVec + 1)]; /* cp.async ring: per thread its samples + the word before them */
	uint4 min[2][kThreads * kVec];      /* cp.async ring: per thread its slice of the model */
	uint32_t lut[2][256]; /* codeword tables of the primary / secondary encoder (build_lut) */
	uint32_t lut_range[2];
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
	P.lut_sel = sel;
	P.lut_range = sh.lut_range[sel];
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

/* per-thread bit writer into the staging words: up to 32 bits per push.  Branch
 * free: the word store is a predicated shared-memory reduction, so a push is a
 * short dependency chain (two funnel shifts, an OR, an add) whatever the data. */
struct Packer {
	uint32_t lo;   /* pending bits, right aligned (bits above `fill` are stale) */
	uint32_t fill; /* number of pending bits, < 32 between pushes */
	uint32_t wp;   /* shared-memory byte address of the word the pending bits belong to */
};

__device__ __forceinline__ void packer_open(Packer &p, const uint32_t *stg, uint32_t bitpos)
{
	p.lo = 0;
	p.fill = bitpos & 31u;
	p.wp = (uint32_t)__cvta_generic_to_shared(stg) + ((bitpos >> 5) << 2);
}

__device__ __forceinline__ void packer_push(Packer &p, uint32_t code, uint32_t len)
{
	/* (hi:lo) = (lo << len) | code; ref bitstream_add_bits32, bitstream_writer.h:124-158 */
	const uint32_t hi = __funnelshift_lc(p.lo, 0u, len);
	p.lo = __funnelshift_lc(0u, p.lo, len) | code;
	p.fill += len;
	/* the completed word if fill >= 32, else 0: OR-ing 0 is harmless and keeps the push free
	 * of branches (ptxas turns a predicated ATOMS into a branch) */
	const uint32_t word = p.fill >= 32u ? __funnelshift_r(p.lo, hi, p.fill) : 0u;
	asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(p.wp), "r"(word) : "memory");
	p.wp += (p.fill >> 5) << 2;
	p.fill &= 31u;
}

__device__ __forceinline__ void packer_close(Packer &p)
{
	const uint32_t word = p.fill ? p.lo << (32u - p.fill) : 0u;
	asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(p.wp), "r"(word) : "memory");
}

/* exclusive scan of per-thread bit counts over the CTA; one barrier.  `flag` (warp
 * uniform, 0/1) is OR-reduced over the CTA on the way (bit 31 of the warp sums). */
__device__ __forceinline__ uint32_t block_scan(Shared &sh, uint32_t tb, uint32_t &total, uint32_t flag = 0,
					       uint32_t *any = nullptr)
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
		sh.wsum[warp] = incl | (flag << 31);
	__syncthreads();
	uint32_t ws = lane < kWarps ? sh.wsum[lane] : 0u;
	if (any)
		*any = __reduce_or_sync(kFull, ws) >> 31;
	ws &= 0x7FFFFFFFu;
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

	if ((uint64_t)c.gw0 * 4 >= o.lo && ((uint64_t)c.gw0 + wfull) * 4 <= o.hi) {
		/* every word lies inside the window (the usual case): plain coalesced stores */
		uint32_t *out = (uint32_t *)o.base4 + c.gw0;
		for (uint32_t w = tid; w < wfull; w += kThreads) {
			uint32_t v = sh.stg[w];
			sh.stg[w] = 0;
			out[w] = airs_bswap32(v);
		}
	} else {
		for (uint32_t w = tid; w < wfull; w += kThreads) {
			uint32_t v = sh.stg[w];
			sh.stg[w] = 0;
			store_word(o, c.gw0 + w, v);
		}
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
	packer_open(pk, sh.stg, c.sbits + excl);
	uint32_t cum = c.gw0 * 32u + c.sbits + excl - 8u * a; /* stream bits before this thread's samples */
	for (uint32_t i = i0; i < i1; i++) {
		uint32_t x = need_x ? sample_at(P.src, P.dtype, i) : 0u;
		uint32_t m = (P.pre == CMP_PREPROCESS_MODEL || P.model_mode == 2) ? (uint32_t)P.work[i] : 0u;
		uint32_t cw, cl, rw, rl;
		encode_any(e, residual_at(P, i, x), cw, cl, rw, rl);
		packer_push(pk, cw, cl);
		packer_push(pk, rw, rl);
		cum += cl + rl;
		/* model := samples, or model update, while the reference's writer has not
		 * given up (ref cmp.c:300-311) */
		if (P.model_mode && cum < P.trip)
			P.work[i] = (uint16_t)(P.model_mode == 1 ? x : airs_model_update(x, m, P.rate, P.is_signed));
	}
	if (tb)
		packer_close(pk);
	__syncthreads();
	copy_out(sh, o, c, tile_bits);
}

/* -------------------------------------------------------------------------
 * fast path: full 2048-sample tiles of a 16-bit container, 16-byte aligned
 * source (and work buffer when used).  One loop over the tiles of a frame;
 * the loads of tile t+1 are issued before tile t is encoded; everything lives
 * in registers and all inner loops are unrolled.
 *
 * Codewords come from a 256-entry table in shared memory indexed by the
 * residual itself (r + R, |r| < R <= 128): entry = length << 26 | code bits,
 * built per job for both passes (build_lut).  This covers escapes as well, as
 * long as the whole code of the sample fits 25 bits.  A warp whose 512 samples
 * do not all hit the table computes its codewords arithmetically instead
 * (airs_golomb).
 * ---------------------------------------------------------------------- */

/* packed 16x2 zig-zag: ref map_to_unsigned, encoder.c:274-286 */
__device__ __forceinline__ uint32_t zigzag2(uint32_t d)
{
	uint32_t sign; /* per 16-bit lane: 0xFFFF if negative */
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sign) : "r"(d), "r"(0u), "r"(0xBB99u));
	return ((d << 1) & 0xFFFEFFFEu) ^ sign;
}

/* Ampere-style asynchronous copies global -> shared (LDGSTS): the loads of the next
 * tile are in flight while this one is encoded, without holding registers */
__device__ __forceinline__ void cp_async16(uint32_t smem, const void *g)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(g) : "memory");
}

__device__ __forceinline__ void cp_async4(uint32_t smem, const void *g)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem), "l"(g) : "memory");
}

__device__ __forceinline__ void cp_async_commit()
{
	asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int N>
__device__ __forceinline__ void cp_async_wait()
{
	asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr uint32_t kLutLenShift = 26;
constexpr uint32_t kLutCodeMask = (1u << kLutLenShift) - 1;

/* table of one encoder; returns the usable half range R (0: table unusable).
 * All threads call it; contains one barrier. */
__device__ __noinline__ uint32_t build_lut(const EncConst &e, uint32_t *lut)
{
	uint32_t bad = 0; /* bit i: an entry with |r| < 8 << i does not fit */

	if (e.type == CMP_ENCODER_UNCOMPRESSED)
		return 0; /* uniform */
	for (uint32_t idx = threadIdx.x; idx < 256; idx += kThreads) {
		const uint32_t r = (idx - 128u) & 0xFFFFu;
		uint32_t cw, cl, rw, rl;
		if (e.type == CMP_ENCODER_GOLOMB_ZERO)
			airs_encode<CMP_ENCODER_GOLOMB_ZERO>(e, r, cw, cl, rw, rl);
		else
			airs_encode<CMP_ENCODER_GOLOMB_MULTI>(e, r, cw, cl, rw, rl);
		const uint32_t len = cl + rl;
		const bool ok = len <= 25u;
		lut[idx] = ok ? (len << kLutLenShift) | (cw << rl) | rw : 0u;
		if (!ok) {
			const uint32_t dist = idx >= 128u ? idx - 127u : 128u - idx; /* r in [-R, R) <=> dist <= R */
			bad |= dist <= 8u ? 0x1Fu : dist <= 16u ? 0x1Eu : dist <= 32u ? 0x1Cu : dist <= 64u ? 0x18u : 0x10u;
		}
	}
	bad = __syncthreads_or(bad);
	return 128u >> __popc(bad); /* 128, 64, 32, 16, 8 or (all bad) 4 */
}

/* the arithmetic encoders on a packed pair of zig-zag mapped samples; only
 * called by warps that missed the table */
__device__ __forceinline__ void encode_pair_compute(const EncConst &e, uint32_t z, uint32_t &pc, uint32_t &ph,
						    uint32_t &pl)
{
	if (e.type == CMP_ENCODER_GOLOMB_ZERO) {
		uint32_t c0, l0, c1, l1, r, q;
		airs_encode_mapped<CMP_ENCODER_GOLOMB_ZERO>(e, z & 0xFFFFu, c0, l0, r, q);
		airs_encode_mapped<CMP_ENCODER_GOLOMB_ZERO>(e, z >> 16, c1, l1, r, q);
		pc = __funnelshift_lc(0u, c0, l1) | c1;
		ph = __funnelshift_lc(c0, 0u, l1);
		pl = l0 + l1;
	} else {
		/* escapes carry a raw part: samples are 64-bit strings, the pair fits 64 bits
		 * or is flagged through pl > 64 */
		uint32_t c0, l0, r0, q0, c1, l1, r1, q1;
		airs_encode_mapped<CMP_ENCODER_GOLOMB_MULTI>(e, z & 0xFFFFu, c0, l0, r0, q0);
		airs_encode_mapped<CMP_ENCODER_GOLOMB_MULTI>(e, z >> 16, c1, l1, r1, q1);
		const uint64_t s0 = ((uint64_t)c0 << q0) | r0, s1 = ((uint64_t)c1 << q1) | r1;
		const uint32_t sl0 = l0 + q0, sl1 = l1 + q1;
		const uint64_t s = sl1 < 64u ? (s0 << sl1) | s1 : 0;
		pc = (uint32_t)s;
		ph = (uint32_t)(s >> 32);
		pl = sl0 + sl1;
	}
}

__device__ __noinline__ void frame_fast(Shared &sh, const OutWin &o, uint32_t a, Cursor &c, uint32_t n_tiles,
					bool size_only)
{
	const Pass &P = sh.pass;
	const EncConst e = P.enc;
	const uint32_t tid = threadIdx.x;
	const uint32_t pre = P.pre, model_mode = P.model_mode, enc = e.type;
	const uint16_t *src = (const uint16_t *)P.src + tid * kSpt;
	uint16_t *work = P.work + tid * kSpt;
	const bool need_x = pre != CMP_PREPROCESS_IWT || model_mode;
	const bool need_m = pre == CMP_PREPROCESS_MODEL || model_mode == 2;
	/* table lookup constants: u = r + R per lane, hit <=> u < 2R */
	const uint32_t R = P.lut_range;
	const uint32_t rp = R * 0x00010001u, notmask = ~((2u * R - 1u) * 0x00010001u), imask = (2u * R - 1u) << 2;
	const char *lut = (const char *)(sh.lut[P.lut_sel] + (128u - R));
	/* per-thread slots of the two cp.async rings; a thread only ever reads what it copied
	 * itself, so no barrier is involved in the input staging */
	const uint32_t in_sm[2] = { (uint32_t)__cvta_generic_to_shared(&sh.in[0][tid * (kVec + 1)]),
				    (uint32_t)__cvta_generic_to_shared(&sh.in[1][tid * (kVec + 1)]) };
	const uint32_t m_sm[2] = { (uint32_t)__cvta_generic_to_shared(&sh.min[0][tid * kVec]),
				   (uint32_t)__cvta_generic_to_shared(&sh.min[1][tid * kVec]) };
	const bool diff = pre == CMP_PREPROCESS_DIFF;

#define AIRS_STAGE_TILE(t_)                                                                          \
	do {                                                                                         \
		const uint32_t st_ = (t_) & 1u, b_ = (t_) * kTile;                                   \
		if (need_x) {                                                                        \
			for (uint32_t v_ = 0; v_ < kVec; v_++)                                       \
				cp_async16(in_sm[st_] + 16 * v_, src + b_ + 8 * v_);                 \
			if (diff && (b_ | tid))                                                      \
				cp_async4(in_sm[st_] + 16 * kVec, src + b_ - 2); /* [x(i0-2), x(i0-1)] */ \
		}                                                                                    \
		if (need_m) {                                                                        \
			for (uint32_t v_ = 0; v_ < kVec; v_++)                                       \
				cp_async16(m_sm[st_] + 16 * v_, work + b_ + 8 * v_);                 \
		}                                                                                    \
		cp_async_commit();                                                                   \
	} while (0)

	AIRS_STAGE_TILE(0u);

	for (uint32_t t = 0; t < n_tiles; t++) {
		const uint32_t base = t * kTile, st = t & 1u;
		uint32_t w[kPairs], mw[kPairs], d[kPairs];

		PHASE_T(t0);
		/* next tile's copies go out before this tile is encoded */
		if (t + 1 < n_tiles) {
			AIRS_STAGE_TILE(t + 1);
			cp_async_wait<1>();
		} else {
			cp_async_wait<0>();
		}
		if (need_x) {
#pragma unroll
			for (uint32_t v = 0; v < kVec; v++) {
				const uint4 q = sh.in[st][tid * (kVec + 1) + v];
				w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
			}
		}
		if (need_m) {
#pragma unroll
			for (uint32_t v = 0; v < kVec; v++) {
				const uint4 q = sh.min[st][tid * kVec + v];
				mw[4 * v] = q.x; mw[4 * v + 1] = q.y; mw[4 * v + 2] = q.z; mw[4 * v + 3] = q.w;
			}
		}
		PHASE_T(t1);
		PHASE_ADD(0, t0, t1);

		/* packed residuals: ref preprocess.c:268-290,348-353,406-411 */
		if (pre == CMP_PREPROCESS_DIFF) {
			const uint32_t pw = (base | tid) ? sh.in[st][tid * (kVec + 1) + kVec].x : 0u; /* [x(i0-2), x(i0-1)] */
#pragma unroll
			for (int k = 0; k < (int)kPairs; k++)
				d[k] = __vsub2(w[k], __funnelshift_l(k ? w[k - 1] : pw, w[k], 16));
		} else if (pre == CMP_PREPROCESS_MODEL) {
#pragma unroll
			for (int k = 0; k < (int)kPairs; k++)
				d[k] = __vsub2(w[k], mw[k]);
		} else if (pre == CMP_PREPROCESS_IWT) {
			{
#pragma unroll
				for (uint32_t v = 0; v < kVec; v++) {
					const uint4 q = ((const uint4 *)(work + base))[v];
					d[4 * v] = q.x; d[4 * v + 1] = q.y; d[4 * v + 2] = q.z; d[4 * v + 3] = q.w;
				}
			}
		} else {
#pragma unroll
			for (int k = 0; k < (int)kPairs; k++)
				d[k] = w[k];
		}

		/* codewords; the two samples of a word are merged into one string of pl bits:
		 * (ph:pc) = code_lo << len_hi | code_hi */
		uint32_t pc[kPairs], ph[kPairs], pl[kPairs], tb = 0, mx = 0;
		if (enc == CMP_ENCODER_UNCOMPRESSED) {
#pragma unroll
			for (int k = 0; k < (int)kPairs; k++) {
				pc[k] = __byte_perm(d[k], 0, 0x1032); /* first sample in the upper half */
				ph[k] = 0;
				pl[k] = 32;
			}
			tb = 32 * kPairs;
			mx = 32;
		} else {
			uint32_t u[kPairs], chk = 0;
#pragma unroll
			for (int k = 0; k < (int)kPairs; k++) {
				asm("add.u16x2 %0, %1, %2;" : "=r"(u[k]) : "r"(d[k]), "r"(rp));
				chk |= u[k];
			}
			if (__all_sync(kFull, R >= 8u && (chk & notmask) == 0u)) {
#pragma unroll
				for (int k = 0; k < (int)kPairs; k++) {
					const uint32_t e0 = *(const uint32_t *)(lut + ((u[k] << 2) & imask));
					const uint32_t e1 = *(const uint32_t *)(lut + ((u[k] >> 14) & imask));
					const uint32_t l1 = e1 >> kLutLenShift, c0 = e0 & kLutCodeMask;
					pc[k] = __funnelshift_lc(0u, c0, l1) | (e1 & kLutCodeMask);
					ph[k] = __funnelshift_lc(c0, 0u, l1);
					pl[k] = (e0 >> kLutLenShift) + l1;
					tb += pl[k];
					mx = max(mx, pl[k]);
				}
			} else {
#pragma unroll
				for (int k = 0; k < (int)kPairs; k++) {
					encode_pair_compute(e, zigzag2(d[k]), pc[k], ph[k], pl[k]);
					tb += pl[k];
					mx = max(mx, pl[k]);
				}
			}
		}
		/* 0: every pair of the warp fits 32 bits, 1: 64 bits, 2: not even that (rare) */
		const uint32_t wide = __reduce_max_sync(kFull, mx > 64u ? 2u : (mx > 32u ? 1u : 0u));
		PHASE_T(t2);
		PHASE_ADD(1, t1, t2);

		uint32_t tile_bits, overlong;
		const uint32_t excl = block_scan(sh, tb, tile_bits, wide == 2u, &overlong);
		PHASE_T(t3);
		PHASE_ADD(2, t2, t3);
		if (size_only) {
			const uint32_t staged = c.sbits + tile_bits;
			c.gw0 += staged >> 5;
			c.sbits = staged & 31u;
			__syncthreads();
		} else if (overlong) {
			__syncthreads(); /* the scan's warp sums are reused by tile_generic */
			/* a pair longer than 64 bits (multi-escape pile-up): the slow way.  Nothing
			 * has been staged yet. */
			tile_generic(sh, o, a, c, base, false);
		} else {
			const uint32_t tile_end = c.gw0 * 32u + c.sbits + tile_bits - 8u * a; /* stream bits after this tile */
			/* two independent writers (pairs 0-3 and 4-7) double the instruction level
			 * parallelism of the packing chain */
			Packer pa, pb;
			const uint32_t pos = c.sbits + excl;
			packer_open(pa, sh.stg, pos);
			uint32_t half = 0;
#pragma unroll
			for (int k = 0; k < (int)kPairs / 2; k++)
				half += pl[k];
			packer_open(pb, sh.stg, pos + half);
			if (wide == 0u) {
#pragma unroll
				for (int k = 0; k < (int)kPairs / 2; k++) {
					packer_push(pa, pc[k], pl[k]);
					packer_push(pb, pc[k + kPairs / 2], pl[k + kPairs / 2]);
				}
			} else {
#pragma unroll
				for (int k = 0; k < (int)kPairs / 2; k++) {
					const uint32_t ha = pl[k] > 32u ? pl[k] - 32u : 0u;
					const uint32_t hb = pl[k + kPairs / 2] > 32u ? pl[k + kPairs / 2] - 32u : 0u;
					packer_push(pa, ph[k], ha);
					packer_push(pb, ph[k + kPairs / 2], hb);
					packer_push(pa, pc[k], pl[k] - ha);
					packer_push(pb, pc[k + kPairs / 2], pl[k + kPairs / 2] - hb);
				}
			}
			packer_close(pa);
			packer_close(pb);
			PHASE_T(t4);
			PHASE_ADD(3, t3, t4);

			/* model := samples, or model update (ref cmp.c:304-311) */
			if (model_mode) {
				if (tile_end < P.trip) {
					uint32_t nm[kPairs];
#pragma unroll
					for (int k = 0; k < (int)kPairs; k++) {
						if (model_mode == 1) {
							nm[k] = w[k];
						} else {
							uint32_t lo = airs_model_update(w[k] & 0xFFFFu, mw[k] & 0xFFFFu, P.rate, P.is_signed);
							uint32_t hi = airs_model_update(w[k] >> 16, mw[k] >> 16, P.rate, P.is_signed);
							nm[k] = lo | (hi << 16);
						}
					}
					uint4 *q = (uint4 *)(work + base);
#pragma unroll
					for (uint32_t v = 0; v < kVec; v++)
						q[v] = make_uint4(nm[4 * v], nm[4 * v + 1], nm[4 * v + 2], nm[4 * v + 3]);
				} else {
					/* the stream overflows its capacity inside this tile: per-sample gate */
					uint32_t cum = c.gw0 * 32u + c.sbits + excl - 8u * a;
					const uint32_t i0 = base + tid * kSpt;
					for (uint32_t i = i0; i < i0 + kSpt; i++) {
						uint32_t x = sample_at(P.src, P.dtype, i);
						uint32_t m = P.work[i];
						uint32_t cw, cl, rw, rl;
						encode_any(e, residual_at(P, i, x), cw, cl, rw, rl);
						cum += cl + rl;
						if (cum < P.trip)
							P.work[i] = (uint16_t)(model_mode == 1 ? x : airs_model_update(x, m, P.rate, P.is_signed));
					}
				}
			}
			PHASE_T(t5);
			__syncthreads();
			PHASE_T(t6);
			copy_out(sh, o, c, tile_bits);
			PHASE_T(t7);
			PHASE_ADD(4, t4, t5);
			PHASE_ADD(5, t5, t6);
			PHASE_ADD(6, t6, t7);
			PHASE_ADD(7, t0, t7);
		}
	}
#undef AIRS_STAGE_TILE
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

	const uint32_t n = P.n, pre = P.pre, model_mode = P.model_mode;
	const bool fast_ok = P.dtype != AIRS_DTYPE_I16_IN_I32 && ((uintptr_t)P.src & 15u) == 0 &&
			     (((uintptr_t)P.work & 15u) == 0 || (pre < CMP_PREPROCESS_IWT && !model_mode));
	const uint32_t n_fast = fast_ok ? n / kTile : 0u;
	if (n_fast)
		frame_fast(sh, o, a, c, n_fast, size_only);
	for (uint32_t base = n_fast * kTile; base < n; base += kTile)
		tile_generic(sh, o, a, c, base, size_only);
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

__global__ void __launch_bounds__(AIRS_THREADS, AIRS_CTAS_PER_SM) airs_encode_kernel(AirsLaunch b)
{
	__shared__ Shared sh;
	const uint32_t tid = threadIdx.x;

	for (uint32_t w = tid; w < kStgWords; w += kThreads)
		sh.stg[w] = 0;
	if (tid == 0)
		sh.ticket = atomicAdd(b.ticket, 1u);

	for (;;) {
		PHASE_T(j0);
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
		PHASE_T(j1);
		PHASE_ADD(8, j0, j1);
		if (!sh.plan.frame_err) { /* codeword tables of this job's encoders (uniform branch) */
			uint32_t r0 = build_lut(sh.plan.enc[0], sh.lut[0]);
			uint32_t r1 = sh.plan.sec_iter ? build_lut(sh.plan.enc[1], sh.lut[1]) : 0u;
			if (tid == 0) {
				sh.lut_range[0] = r0;
				sh.lut_range[1] = r1;
			}
		}
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
		PHASE_T(j2);
		PHASE_ADD(9, j1, j2);

		for (uint32_t f = 0; f < n_frames; f++) {
			PHASE_T(f0);
			if (tid == 0)
				plan_frame(sh, b, f);
			__syncthreads();
			PHASE_T(f1);
			PHASE_ADD(10, f0, f1);
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
			PHASE_T(f2);
			PHASE_ADD(11, f1, f2);
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

#ifdef AIRS_PHASE_CLOCKS
extern "C" void airs_phase_clocks(unsigned long long *out, int reset)
{
	cudaMemcpyFromSymbol(out, g_phase_clk, sizeof(g_phase_clk));
	if (reset) {
		unsigned long long z[16] = { 0 };
		cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z));
	}
}
#endif

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
