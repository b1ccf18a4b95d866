/*
 * airs_kernels.cu - the sm_100a compression kernels.
 *
 * airs_plan_kernel    one thread per job: cmp_initialise validation and every
 *                     constant that follows from the parameters (airs_plan.cuh).
 * airs_encode_kernel  persistent CTAs of 128 threads; a CTA takes one job (one
 *                     compression context, lib/cmp.h:129-137) at a time from a
 *                     ticket counter and pushes its frames through in order.
 *
 * Fast path (frame_fast), per 4096-sample tile.  A warp owns 1024 consecutive
 * samples as 128 pieces of 8; lane l holds pieces l, 32 + l, 64 + l, 96 + l
 * ("segments"), so every 128-bit load of a warp is fully coalesced:
 *   4 x LDG.128 (half of the next tile's loads in flight while this one is
 *   encoded) -> packed 16x2 biased residuals u = r + R (VIADD.16x2; none / diff
 *   / IWT coefficient / model) -> if every residual of the warp lies in [-R, R):
 *   one shared-memory load per PAIR of samples from a 4096-entry table that
 *   holds the merged codeword and length of both (build_pair_lut) -> pairs
 *   merged into quads, quads into one string of at most 64 bits per segment ->
 *   two shuffle scans over packed bit counts + REDUX over the 4 warp sums ->
 *   every string is shifted into place with three funnel shifts and OR-ed into
 *   the MSB-first staging words in shared memory (RED.OR) -> 128-bit
 *   byte-swapped coalesced stores of the tile before, which drains from a second
 *   staging area behind this tile's scan barrier: one block barrier per tile.
 * A warp with a residual outside the table range (escapes, wide data) computes
 * its codewords arithmetically instead, sample by sample, into the same staging
 * words.  Tiles that could overflow the staging area or the destination
 * capacity, ragged tails, unaligned frames and the i16-in-i32 container go
 * through tile_generic (rolled loops, same results).
 *
 * Reference being replaced: compress_engine and cmp_compress_generic
 * (lib/compress/cmp.c:213-393), preprocess.c:268-411, encoder.c:274-378,
 * bitstream_writer.h:124-227, header.c:24-67,137-163.  See DESIGN.md.
 *
 * airs_small_kernel     one warp per short single-frame job (see there).
 * airs_checksum_kernel  one thread per frame: the XXH32 trailers, behind the encoders.
 */
#include <cuda_runtime.h>

#include "airs_device.cuh"
#include "airs_fast.cuh"
#include "airs_launch.h"
#include "airs_plan.cuh"
#include "airs_private.h"

namespace {

constexpr uint32_t kThreads = AIRS_THREADS; /* 128 */
constexpr uint32_t kWarps = kThreads / 32;
constexpr uint32_t kSeg = 4;                /* fast path: segments (pieces of 8 samples) per thread and tile */
constexpr uint32_t kSegModel = 2;           /* ... of the model pass (old and new model words next to the samples) */
constexpr uint32_t kTilePieces = kThreads * kSeg;
constexpr uint32_t kTile = kTilePieces * 8; /* 4096 samples = 8 KiB of u16 */
constexpr uint32_t kGenSpt = 8;             /* generic path: samples per thread and tile */
constexpr uint32_t kGenTile = kThreads * kGenSpt;
/* staging area: one fast tile at <= 12 bits per sample (the table path needs at
 * most 8; a tile with more goes through the generic path), or one generic tile
 * at 48 bits per sample, plus the < 128 bits carried over from the tile before */
constexpr uint32_t kStgBits = kTile * 12 + 128;
constexpr uint32_t kStgWords = kStgBits / 32 + 4;
static_assert(kGenTile * 48 + 128 <= kStgBits, "a generic tile must fit the staging area");
constexpr uint32_t kLutR = 32;              /* pair table covers residuals in [-32, 32) */
constexpr uint32_t kLutStride = 2 * kLutR;
#ifndef AIRS_LUT_PITCH
#define AIRS_LUT_PITCH 65
#endif
/* entries between consecutive rows of the pair table: odd, so that the shared-memory bank of entry
 * [u_hi][u_lo] depends on both residuals (with a pitch of 64 it would be u_lo mod 32 alone, and lanes
 * with equal u_lo but different u_hi would queue up at one bank) */
constexpr uint32_t kLutPitch = AIRS_LUT_PITCH;
static_assert(kLutPitch >= kLutStride && kLutPitch < 256, "the pitch is a byte weight of the index IDP.2A");
constexpr uint32_t kLutLenShift = 26;       /* entry = pair length << 26 | pair codeword */
constexpr uint32_t kLutMaxLen = 13;         /* per sample, so that a pair's codeword fits 26 bits */
constexpr uint32_t kLutMinSamples = 1024;   /* frames shorter than this do not pay for a table build */
constexpr uint32_t kSmallMaxSamples = 32768; /* longest frame a single warp encodes (airs_small_kernel); measured crossover */
#ifndef AIRS_SEG_RUN
#define AIRS_SEG_RUN 4
#endif
constexpr uint32_t kSegRun = AIRS_SEG_RUN;  /* segments per thread and visit of model_run_fast() */
#ifndef AIRS_CTX_FRAMES
#define AIRS_CTX_FRAMES 64
#endif
constexpr uint32_t kCtxFrames = AIRS_CTX_FRAMES;         /* most frames of a run of model_run_fast(): their state must leave the SM a 196 KB carve-out and some L1 */
constexpr uint64_t kMask48 = 0xFFFFFFFFFFFFull;
constexpr uint32_t kFull = 0xFFFFFFFFu;
/* Development switches (make EXTRA=-D..., never in the shipped library):
 *   AIRS_EXP_NO_MODEL_IO  timing experiment of DESIGN.md section 4 - the frame-by-frame model pass without its model
 *                         loads and stores (wrong streams): the bound of any "model on chip" design with that loop order
 *   AIRS_CTX_DEBUG        counters of finished / abandoned runs of model_run_fast() in words 40-48 of the scratch header
 *                         (tools/perf_probe.py prints them with AIRS_PROBE_DEBUG=1) */
#ifdef AIRS_EXP_NO_MODEL_IO
constexpr bool kExpNoModelIo = true;
#else
constexpr bool kExpNoModelIo = false;
#endif

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
	uint32_t plut[kLutStride * kLutPitch]; /* pair table: length << 26 | merged codeword at [u_hi * kLutPitch + u_lo] */
	/* two staging areas (the fast path fills one while the other one drains): MSB-first
	 * 32-bit words of the stream being assembled, all zero when idle; the 4 pad words in
	 * front absorb the zeros that strings ending in word 0 or 1 OR below the area */
	alignas(16) uint32_t stg_mem[2][4 + kStgWords];
	uint2 slut[kLutStride];              /* single-sample table the pair table is built from */
	uint32_t wsum[2][kWarps];            /* warp totals of the tile scan, alternating between tiles */
	uint32_t plut_key[3];                /* encoder the pair table was built for: type, g, outlier */
	uint32_t plut_R;                     /* its usable half range: 32, 16, 8 or 0 (none) */
	uint32_t first_code[3];              /* DIFF: codeword (hi, lo, bits) of the frame's first sample, used by thread 0 */
	JobPlan plan;
	airs_job job;
	Pass pass;
	CtxState ctx;
	uint64_t offset;
	uint32_t ticket;
	/* context_fast(): where the stream of every frame of the context stands (bits from the start of its 16-byte aligned
	 * space) and its last, incomplete 16-byte group */
	uint32_t cx_bit[kCtxFrames];
	uint32_t cx_abort;
	alignas(16) uint4 cx_carry[kCtxFrames];
};

/* byte window of the destination a pass may write, in the 16-byte aligned
 * space that starts at dst - (dst & 15) */
struct OutWin {
	uint8_t *base;
	uint32_t lo, hi;
};

/* position of the stream under construction: word 0 of staging area `buf` is
 * word gw0 (a multiple of 4) of the aligned space and already holds sbits
 * (< 128) bits */
struct Cursor {
	uint32_t gw0, sbits, buf;
};

__device__ __forceinline__ uint32_t *stg_of(Shared &sh, uint32_t buf)
{
	return sh.stg_mem[buf] + 4;
}

__device__ __forceinline__ void cursor_advance(Cursor &c, uint32_t bits)
{
	const uint32_t staged = c.sbits + bits;
	c.gw0 += (staged >> 7) << 2;
	c.sbits = staged & 127u;
}

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
	const uint32_t v = __ldg((const uint16_t *)src + i);
	return (dtype & AIRS_DTYPE_BE) ? ((v << 8) | (v >> 8)) & 0xFFFFu : v;
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

/* one level of the transform at stride s over w[0..n), restricted to the positions lo <= i < hi
 * whose neighbours lie inside [lo, hi) or outside the frame; w is addressed as w[i - lo].  lo is a
 * multiple of 2s (tiles and their halos start at multiples of 32, s <= 8).  Stores truncate to
 * 16 bits, which is all the reference's casts do. */
__device__ __forceinline__ void iwt_level(int16_t *w, uint32_t s, uint32_t lo, uint32_t hi, uint32_t n)
{
	const uint32_t tid = threadIdx.x, s2 = 2u * s;
	int16_t *b = w - lo; /* b[i] is element i */

	for (uint32_t i = lo + s + s2 * tid; i < hi; i += s2 * kThreads) { /* details: odd multiples of s */
		if (i + s < hi)
			b[i] = (int16_t)(b[i] - (((int32_t)b[i - s] + b[i + s]) >> 1));
		else if (i + s >= n) /* the last one has no right neighbour */
			b[i] = (int16_t)(b[i] - b[i - s]);
	}
	__syncthreads();
	for (uint32_t i = lo + s2 * tid; i < hi; i += s2 * kThreads) { /* approximations: even multiples */
		const bool has_l = i >= s, has_r = i + s < n;
		if ((has_l && i - s < lo) || (has_r && i + s >= hi))
			continue; /* a neighbour outside the tile: this element belongs to the halo */
		int32_t t = 0;
		if (has_l && has_r)
			t = ((int32_t)b[i - s] + b[i + s]) >> 2;
		else if (has_r)
			t = (int32_t)b[i + s] >> 1;
		else if (has_l)
			t = (int32_t)b[i - s] >> 1;
		b[i] = (int16_t)(b[i] + t);
	}
	__syncthreads();
}

/* The transform of a whole frame into the work buffer, four levels at a time.  Stage 0 takes
 * the samples, stage 1 every 16th coefficient of the work buffer (the approximations stage 0
 * left), stage 2 every 256th, ..: to the decimated sequence y[k] = w[k * S] of length ceil(n / S)
 * levels 1-4 are what levels with the strides S, 2S, 4S, 8S are to w, edges included.  Each
 * stage runs in shared memory, tile by tile with a halo of 32 elements (a level-4 coefficient
 * depends on elements up to 30 positions away), so no lifting step ever waits for global
 * memory.  `buf32` is idle staging memory and is handed back zeroed. */
constexpr uint32_t kIwtTile = 4096, kIwtHalo = 32, kIwtStageLevels = 4;

__device__ __noinline__ void iwt_global(const Pass &P, uint32_t *buf32)
{
	int16_t *w = (int16_t *)P.work;
	int16_t *buf = (int16_t *)buf32;
	int16_t *keep = buf + kIwtTile + 2u * kIwtHalo; /* the last 32 untransformed elements of the tile before */
	const uint32_t n = P.n;
	const uint32_t tid = threadIdx.x;
	/* 16-byte accesses for stage 0: tiles and halos start at multiples of 32 samples */
	const bool vec = P.dtype != AIRS_DTYPE_I16_IN_I32 && !(P.dtype & AIRS_DTYPE_BE) && ((uintptr_t)P.src & 15u) == 0 && ((uintptr_t)w & 15u) == 0;

	for (uint32_t S = 1, stage = 0; stage == 0 || S < n; S <<= kIwtStageLevels, stage++) {
		const uint32_t m = (n + S - 1u) / S; /* elements of this stage */
		for (uint32_t t0 = 0; t0 < m; t0 += kIwtTile) {
			const uint32_t lo = t0 >= kIwtHalo ? t0 - kIwtHalo : 0u;
			const uint32_t hi = min(t0 + kIwtTile + kIwtHalo, m);
			const uint32_t t1 = min(t0 + kIwtTile, m);
			if (stage == 0 && vec) {
				const uint4 *s4 = reinterpret_cast<const uint4 *>(P.src) + lo / 8u;
				uint4 *b4 = reinterpret_cast<uint4 *>(buf);
				const uint32_t nv = (hi - lo) / 8u;
				for (uint32_t v0 = tid; v0 < nv; v0 += 2u * kThreads) { /* two loads in flight per thread */
					const uint4 q0 = __ldg(s4 + v0);
					const uint4 q1 = v0 + kThreads < nv ? __ldg(s4 + v0 + kThreads) : make_uint4(0, 0, 0, 0);
					b4[v0] = q0;
					if (v0 + kThreads < nv)
						b4[v0 + kThreads] = q1;
				}
				for (uint32_t i = lo + nv * 8u + tid; i < hi; i += kThreads)
					buf[i - lo] = (int16_t)sample_at(P.src, P.dtype, i);
			} else if (stage == 0) {
				for (uint32_t i = lo + tid; i < hi; i += kThreads)
					buf[i - lo] = (int16_t)sample_at(P.src, P.dtype, i);
			} else {
				/* in place: the elements in front of the tile have been transformed already; their
				 * old values wait in `keep` */
				for (uint32_t i = lo + tid; i < hi; i += kThreads)
					buf[i - lo] = i < t0 ? keep[i - lo] : w[(size_t)i * S];
				__syncthreads();
				if (tid < kIwtHalo && t1 == t0 + kIwtTile)
					keep[tid] = buf[t1 - kIwtHalo - lo + tid];
			}
			__syncthreads();
			for (uint32_t l = 0, ss = 1; l < kIwtStageLevels && ss < m; l++, ss <<= 1)
				iwt_level(buf, ss, lo, hi, m);
			if (stage == 0 && vec) {
				uint4 *w4 = reinterpret_cast<uint4 *>(w) + t0 / 8u;
				const uint4 *b4 = reinterpret_cast<const uint4 *>(buf + (t0 - lo));
				const uint32_t nv = (t1 - t0) / 8u;
				for (uint32_t v = tid; v < nv; v += kThreads)
					w4[v] = b4[v];
				for (uint32_t i = t0 + nv * 8u + tid; i < t1; i += kThreads)
					w[i] = buf[i - lo];
			} else {
				for (uint32_t i = t0 + tid; i < t1; i += kThreads)
					w[(size_t)i * S] = buf[i - lo];
			}
			__syncthreads();
		}
	}
	for (uint32_t i = tid; i < (kIwtTile + 3u * kIwtHalo) / 2u; i += kThreads)
		buf32[i] = 0;
	__syncthreads();
}

/* -------------------------------------------------------------------------
 * XXH32 of the big-endian samples of one frame (ref cmp_checksum,
 * header.c:137-163; xxHash 0.8.3 XXH32), by ONE thread: the hash is a serial
 * chain per stream, its four accumulators are the thread's instruction level
 * parallelism, and a batch has thousands of streams (airs_checksum_kernel).
 * ---------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t sample_pair_at(const uint8_t *src, uint32_t dtype, bool al4, uint32_t i)
{
	if (dtype == AIRS_DTYPE_I16_IN_I32) {
		const uint32_t *p = (const uint32_t *)src;
		return (__ldg(p + i) & 0xFFFFu) | (__ldg(p + i + 1) << 16);
	}
	uint32_t w;
	if (al4) {
		w = __ldg((const uint32_t *)((const uint16_t *)src + i));
	} else {
		const uint16_t *p = (const uint16_t *)src;
		w = (uint32_t)__ldg(p + i) | ((uint32_t)__ldg(p + i + 1) << 16);
	}
	return (dtype & AIRS_DTYPE_BE) ? airs_swap16x2(w) : w;
}

__device__ uint32_t stream_checksum(const uint8_t *src, uint32_t dtype, uint32_t n)
{
	const uint32_t nbytes = n * 2u, stripes = n / 8u;
	const uint32_t seed = AIRS_CHECKSUM_SEED;
	uint32_t v0 = seed + AIRS_XP1 + AIRS_XP2, v1 = seed + AIRS_XP2, v2 = seed, v3 = seed - AIRS_XP1;
	uint32_t s = 0;
	const bool be = (dtype & AIRS_DTYPE_BE) != 0u;

	if (dtype != AIRS_DTYPE_I16_IN_I32 && ((uintptr_t)src & 15u) == 0) {
		/* whole 128-byte lines: 8 x LDG.128 in flight, then 8 stripes of rounds */
		const uint4 *p = (const uint4 *)src;
		for (; s + 8 <= stripes; s += 8) {
			uint4 q[8];
#pragma unroll
			for (int k = 0; k < 8; k++)
				q[k] = __ldg(p + s + k);
#pragma unroll
			for (int k = 0; k < 8; k++) { /* (big-endian samples are the image the hash is taken of) */
				const uint4 e = be ? q[k] : make_uint4(airs_be_pair(q[k].x), airs_be_pair(q[k].y), airs_be_pair(q[k].z), airs_be_pair(q[k].w));
				v0 = airs_xxh_round(v0, e.x);
				v1 = airs_xxh_round(v1, e.y);
				v2 = airs_xxh_round(v2, e.z);
				v3 = airs_xxh_round(v3, e.w);
			}
		}
	}
	const bool al4 = ((uintptr_t)src & 3u) == 0;
	for (; s < stripes; s++) {
		v0 = airs_xxh_round(v0, airs_be_pair(sample_pair_at(src, dtype, al4, s * 8)));
		v1 = airs_xxh_round(v1, airs_be_pair(sample_pair_at(src, dtype, al4, s * 8 + 2)));
		v2 = airs_xxh_round(v2, airs_be_pair(sample_pair_at(src, dtype, al4, s * 8 + 4)));
		v3 = airs_xxh_round(v3, airs_be_pair(sample_pair_at(src, dtype, al4, s * 8 + 6)));
	}
	uint32_t h = nbytes >= 16 ? airs_rotl(v0, 1) + airs_rotl(v1, 7) + airs_rotl(v2, 12) + airs_rotl(v3, 18)
				  : seed + AIRS_XP5;
	h += nbytes;
	uint32_t i = stripes * 8;
	for (; i + 2 <= n; i += 2)
		h = airs_rotl(h + airs_be_pair(sample_pair_at(src, dtype, false, i)) * AIRS_XP3, 17) * AIRS_XP4;
	if (i < n) {
		const uint32_t sv = sample_at(src, dtype, i);
		h = airs_rotl(h + (sv >> 8) * AIRS_XP5, 11) * AIRS_XP1;
		h = airs_rotl(h + (sv & 0xFFu) * AIRS_XP5, 11) * AIRS_XP1;
	}
	h ^= h >> 15;
	h *= AIRS_XP2;
	h ^= h >> 13;
	h *= AIRS_XP3;
	h ^= h >> 16;
	return h;
}

/* -------------------------------------------------------------------------
 * building blocks shared by the fast and the generic tile
 * ---------------------------------------------------------------------- */

/* per-thread bit writer into the staging words: up to 32 bits per push.  Branch
 * free: the word store is a shared-memory reduction, so a push is a short
 * dependency chain (two funnel shifts, an OR, an add) whatever the data. */
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
		sh.wsum[0][warp] = incl;
	__syncthreads();
	const uint32_t ws = lane < kWarps ? sh.wsum[0][lane] : 0u;
	total = __reduce_add_sync(kFull, ws);
	const uint32_t wpre = __reduce_add_sync(kFull, lane < warp ? ws : 0u);
	return wpre + incl - tb;
}

/* Complete 16-byte groups of staged words leave as coalesced 128-bit stores and
 * the staging area is zeroed behind them; call after a barrier that follows the
 * staging.  `staged` = bits in area `buf` (word 0 = word gw0 of the aligned
 * space).  The trailing partial group moves to the front of area `carry_to`:
 * the same area (nobody else touches it until the next barrier), or the other
 * one, where it is OR-ed in because the next tile is being staged there at the
 * same time.  No barrier afterwards: the next staging into a drained area
 * happens behind the next scan barrier. */
__device__ __forceinline__ void copy_out(Shared &sh, const OutWin &o, uint32_t buf, uint32_t gw0, uint32_t staged,
					 uint32_t carry_to, uint4 *carry_slot = nullptr)
{
	const uint32_t tid = threadIdx.x;
	const uint32_t nvec = staged >> 7;
	uint4 *stg4 = reinterpret_cast<uint4 *>(stg_of(sh, buf));
	const uint32_t b0 = gw0 * 4u; /* streams are shorter than 2^24 bytes */

	if (b0 >= o.lo && b0 + 16u * nvec <= o.hi) { /* the usual case: all groups inside the window */
		uint4 *out = reinterpret_cast<uint4 *>(o.base + b0);
		for (uint32_t v = tid; v < nvec; v += kThreads) {
			const uint4 q = stg4[v];
			stg4[v] = make_uint4(0, 0, 0, 0);
			out[v] = make_uint4(airs_bswap32(q.x), airs_bswap32(q.y), airs_bswap32(q.z), airs_bswap32(q.w));
		}
	} else { /* an edge of the window: header in front, capacity or a neighbour stream behind */
		for (uint32_t v = tid; v < nvec; v += kThreads) {
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
						o.base[b + k] = s8[k ^ 3u]; /* stream byte k sits in the MSB-first word k / 4 */
			}
			stg4[v] = make_uint4(0, 0, 0, 0);
		}
	}
	if (tid == 0 && carry_slot) { /* (context_fast: the group waits in the frame's slot, the area is left all zero) */
		*carry_slot = stg4[nvec];
		stg4[nvec] = make_uint4(0, 0, 0, 0);
	} else if (tid == 0 && (nvec || carry_to != buf)) { /* thread 0 drained group 0 itself; group nvec is nobody else's */
		const uint4 carry = stg4[nvec];
		stg4[nvec] = make_uint4(0, 0, 0, 0);
		uint32_t *dst = stg_of(sh, carry_to);
		if (carry_to == buf) {
			*reinterpret_cast<uint4 *>(dst) = carry;
		} else {
			atomicOr(dst, carry.x);
			atomicOr(dst + 1, carry.y);
			atomicOr(dst + 2, carry.z);
			atomicOr(dst + 3, carry.w);
		}
	}
}

/* the synchronous form: drain into the stream, keep the rest in the same area */
__device__ __forceinline__ void copy_out_sync(Shared &sh, const OutWin &o, Cursor &c, uint32_t tile_bits)
{
	copy_out(sh, o, c.buf, c.gw0, c.sbits + tile_bits, c.buf);
	cursor_advance(c, tile_bits);
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

/* samples [base, min(base + kGenTile, end)) */
__device__ __forceinline__ void tile_generic_range(Shared &sh, const OutWin &o, uint32_t a, Cursor &c, uint32_t base,
						    uint32_t end, bool size_only)
{
	const Pass &P = sh.pass;
	const EncConst e = P.enc;
	const uint32_t tid = threadIdx.x;
	const uint32_t i0 = min(base + tid * kGenSpt, end);
	const uint32_t i1 = min(i0 + kGenSpt, end);
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
		cursor_advance(c, tile_bits);
		__syncthreads();
		return;
	}
	Packer pk;
	packer_open(pk, stg_of(sh, c.buf), c.sbits + excl);
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
	copy_out_sync(sh, o, c, tile_bits);
}

/* samples [s0, s1) through the generic path; the cursor travels by value so that the
 * caller's copy can stay in registers */
__device__ __noinline__ Cursor generic_span(Shared &sh, const OutWin o, uint32_t a, Cursor c, uint32_t s0, uint32_t s1,
					    bool size_only)
{
	for (uint32_t base = s0; base < s1; base += kGenTile)
		tile_generic_range(sh, o, a, c, base, s1, size_only);
	return c;
}

/* -------------------------------------------------------------------------
 * pair table.  Entry [u1 * 64 + u0], u = r + R, holds the codewords of two
 * consecutive residuals r0 (first in the stream), r1 merged into one string,
 * and its length (length << 26 | codeword).  Only residuals whose codeword
 * (escape part included) is at most 13 bits long qualify, so that a pair fits
 * the 26 codeword bits of an entry and a quad 64 bits; R shrinks (32, 16, 8)
 * until that holds, 0 = no table for this encoder.
 * All threads call it; two barriers.
 * ---------------------------------------------------------------------- */
template <class S>
__device__ __noinline__ void build_pair_lut(S &sh, const EncConst &e)
{
	const uint32_t tid = threadIdx.x;
	uint32_t bad = 0; /* bit i: a residual with |r| <= 8 << i does not qualify */

	if (tid < kLutStride) {
		const uint32_t r = (tid - kLutR) & 0xFFFFu;
		uint32_t cw, cl, rw, rl;
		if (e.type == CMP_ENCODER_GOLOMB_ZERO)
			airs_encode<CMP_ENCODER_GOLOMB_ZERO>(e, r, cw, cl, rw, rl);
		else
			airs_encode<CMP_ENCODER_GOLOMB_MULTI>(e, r, cw, cl, rw, rl);
		const uint32_t len = cl + rl;
		sh.slut[tid] = make_uint2((cw << rl) | rw, len);
		if (len > kLutMaxLen) {
			const uint32_t dist = tid >= kLutR ? tid - kLutR + 1u : kLutR - tid; /* r in [-R, R) <=> dist <= R */
			bad = dist <= 8u ? 7u : dist <= 16u ? 6u : 4u;
		}
	}
	bad = __syncthreads_or(bad);
	const uint32_t R = (bad & 1u) ? 0u : kLutR >> __popc(bad);
	if (R) {
		for (uint32_t idx = tid; idx < kLutStride * kLutStride; idx += kThreads) {
			const uint32_t u0 = idx % kLutStride, u1 = idx / kLutStride;
			if (u0 < 2u * R && u1 < 2u * R) {
				const uint2 e0 = sh.slut[u0 + kLutR - R], e1 = sh.slut[u1 + kLutR - R];
				sh.plut[u1 * kLutPitch + u0] = ((e0.y + e1.y) << kLutLenShift) | (e0.x << e1.y) | e1.x;
			}
		}
	}
	if (tid == 0) {
		sh.plut_key[0] = e.type;
		sh.plut_key[1] = e.g;
		sh.plut_key[2] = e.outlier;
		sh.plut_R = R;
	}
	__syncthreads();
}

/* -------------------------------------------------------------------------
 * fast path
 * ---------------------------------------------------------------------- */

/* packed 16x2 zig-zag: ref map_to_unsigned, encoder.c:274-286 */
__device__ __forceinline__ uint32_t zigzag2(uint32_t d)
{
	uint32_t sign; /* per 16-bit lane: 0xFFFF if negative */
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sign) : "r"(d), "r"(0u), "r"(0xBB99u));
	return ((d << 1) & 0xFFFEFFFEu) ^ sign;
}

/* pair table entry of the biased residual pair u (u0 | u1 << 16, both below 64): the entry index
 * u0 + 64 u1 is one IDP.2A (two 16-bit x 8-bit products), the address one LEA; lut_s is the
 * table's address in the shared window */
__device__ __forceinline__ uint32_t lut_pair(uint32_t lut_s, uint32_t u)
{
	uint32_t idx, ent;
	asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(idx) : "r"(u), "r"(1u | (kLutPitch << 8)), "r"(0u));
	asm("ld.shared.u32 %0, [%1];" : "=r"(ent) : "r"(lut_s + (idx << 2)));
	return ent;
}

/* OR a bit string of len <= 64 bits (hi:lo, right aligned) into the staging
 * words.  ne = -(bit position where the string starts), updated to the start of
 * the next one.  Three funnel shifts and three reductions whatever the length;
 * words the string does not reach receive zeros (stg_pad absorbs those below
 * word 0). */
__device__ __forceinline__ void put_unit(uint32_t *stg, int32_t &ne, uint32_t hi, uint32_t lo, uint32_t len)
{
	ne -= (int32_t)len;              /* -(end bit) */
	const uint32_t s = (uint32_t)ne; /* wrap-mode funnel shifts use s & 31: the bits free behind the string's last bit */
	/* word of the last bit: (end - 1) >> 5 with end - 1 = ~ne; as a byte offset ~(ne >> 3) & ~3 (SHF + LOP3) */
	uint32_t *p = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(stg) + (~(ne >> 3) & ~3));
	atomicOr(p, __funnelshift_l(0u, lo, s));
	atomicOr(p - 1, __funnelshift_l(lo, hi, s));
	atomicOr(p - 2, __funnelshift_l(hi, 0u, s));
}

/* packed model update of two samples (ref update_model_16, cmp.c:120-142):
 * (m * rate + x * (16 - rate)) >> 4 truncated to 16 bits, for 1 <= rate <= 15.
 * The pairs (m, x) of each lane go through one IDP.2A (two 16-bit x 8-bit
 * products summed) with the weights pre-multiplied by 16, so that the result
 * sits in bytes 1-2 of each sum and one PRMT packs both lanes.  wdp = 16 * rate |
 * 16 * (16 - rate) << 8; SIGNED: i16 containers (operands sign-extended). */
template <bool SIGNED>
__device__ __forceinline__ uint32_t model_update2(uint32_t x, uint32_t m, uint32_t wdp)
{
	const uint32_t lo = __byte_perm(m, x, 0x5410), hi = __byte_perm(m, x, 0x7632); /* (m, x) of the low / high lane */
	uint32_t tl, th;
	if (SIGNED) {
		asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(tl) : "r"(lo), "r"(wdp), "r"(0u));
		asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(th) : "r"(hi), "r"(wdp), "r"(0u));
	} else {
		asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(tl) : "r"(lo), "r"(wdp), "r"(0u));
		asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(th) : "r"(hi), "r"(wdp), "r"(0u));
	}
	return __byte_perm(tl, th, 0x6521);
}

/* one zig-zag mapped residual through the encoder the pass names at run time */
__device__ __forceinline__ void encode_mapped_rt(const EncConst &e, uint32_t m, uint32_t &cw, uint32_t &cl,
						 uint32_t &rw, uint32_t &rl)
{
	if (e.type == CMP_ENCODER_GOLOMB_ZERO)
		airs_encode_mapped<CMP_ENCODER_GOLOMB_ZERO>(e, m, cw, cl, rw, rl);
	else
		airs_encode_mapped<CMP_ENCODER_GOLOMB_MULTI>(e, m, cw, cl, rw, rl);
}

/* codeword (hi:lo, n <= 48 bits) of the first sample of a frame under DIFF preprocessing: its
 * residual is the sample itself (ref preprocess.c:284-290) */
__device__ __noinline__ void first_sample_code(const EncConst &e, uint32_t x0, uint32_t *out)
{
	uint32_t cw = 0, cl = 0, rw = 0, rl = 0;
	if (e.type != CMP_ENCODER_UNCOMPRESSED)
		encode_mapped_rt(e, airs_zigzag16(x0), cw, cl, rw, rl);
	out[0] = __funnelshift_lc(cw, 0u, rl);
	out[1] = __funnelshift_lc(0u, cw, rl) | rw;
	out[2] = cl + rl;
}

/* The arithmetic encoders: every code word is computed once.  slow_codes() takes the packed
 * residuals of the segments d[0..3], d[4..7], .. whose bit is set in `rows`, leaves their code
 * words in cw[] - per sample the low 32 bits, then the high 16 bits | length << 16 (a code word
 * has at most 48 bits) - and returns the bit counts of the segments (b01 = segment 0 | segment 1
 * << 16, b23 likewise); slow_put_codes() stages the 8 samples of one segment from there, the two
 * code words of a pair as one string where they fit 64 bits.  Rolled loops over local memory:
 * small code, no calls, so that the table path around them keeps its registers. */
template <int ENC>
__device__ __forceinline__ void slow_codes_t(const EncConst &e, const uint32_t *d, uint32_t n_words, uint32_t rows,
					     uint32_t *cw, uint32_t &b01, uint32_t &b23)
{
	b01 = 0;
	b23 = 0;
#pragma unroll 4
	for (uint32_t k = 0; k < n_words; k++) {
		if (!((rows >> (k >> 2)) & 1u)) /* segment k / 4 is not asked for */
			continue;
		const uint32_t z = zigzag2(d[k]);
		uint32_t n = 0;
#pragma unroll
		for (uint32_t h = 0; h < 2u; h++) {
			uint32_t c, cl, rw, rl;
			airs_encode_mapped<ENC>(e, h ? z >> 16 : z & 0xFFFFu, c, cl, rw, rl);
			if (ENC == CMP_ENCODER_GOLOMB_ZERO) { /* one string of at most 32 bits */
				cw[4u * k + 2u * h] = c;
				cw[4u * k + 2u * h + 1u] = cl << 16;
			} else { /* codeword then raw escape bits */
				cw[4u * k + 2u * h] = __funnelshift_lc(0u, c, rl) | rw;
				cw[4u * k + 2u * h + 1u] = __funnelshift_lc(c, 0u, rl) | ((cl + rl) << 16);
			}
			n += cl + rl;
		}
		n = (k & 4u) ? n << 16 : n;
		if (k < 8u)
			b01 += n;
		else
			b23 += n;
	}
}

__device__ __forceinline__ void slow_codes(const EncConst &e, const uint32_t *d, uint32_t n_words, uint32_t rows,
					   uint32_t *cw, uint32_t &b01, uint32_t &b23)
{
	if (e.type == CMP_ENCODER_GOLOMB_ZERO)
		slow_codes_t<CMP_ENCODER_GOLOMB_ZERO>(e, d, n_words, rows, cw, b01, b23);
	else
		slow_codes_t<CMP_ENCODER_GOLOMB_MULTI>(e, d, n_words, rows, cw, b01, b23);
}

__device__ __forceinline__ void slow_put_codes(const uint32_t *cw, uint32_t *stg, uint32_t start)
{
	int32_t ne = -(int32_t)start;
#pragma unroll
	for (uint32_t k = 0; k < 4u; k++) {
		const uint32_t lo0 = cw[4u * k], hn0 = cw[4u * k + 1u], lo1 = cw[4u * k + 2u], hn1 = cw[4u * k + 3u];
		const uint32_t n0 = hn0 >> 16, n1 = hn1 >> 16;
		if (n0 + n1 <= 64u) {
			const unsigned long long c0 = (unsigned long long)(hn0 & 0xFFFFu) << 32 | lo0;
			const unsigned long long c1 = (unsigned long long)(hn1 & 0xFFFFu) << 32 | lo1;
			const unsigned long long m = (c0 << n1) | c1;
			put_unit(stg, ne, (uint32_t)(m >> 32), (uint32_t)m, n0 + n1);
		} else {
			put_unit(stg, ne, hn0 & 0xFFFFu, lo0, n0);
			put_unit(stg, ne, hn1 & 0xFFFFu, lo1, n1);
		}
	}
}

/* One scan for the four segments of every thread: b01 = bits of segment 0 |
 * segment 1 << 16, b23 likewise.  Stream order inside a warp: segment 0 of all
 * lanes, then segment 1 of all lanes, ...  Returns the bits of the whole tile
 * and where this thread's segments start in the staging area (sbits bits are
 * staged already).  One block barrier; warps may call it from different
 * places.  The warp totals alternate between two slots (parity), so that no
 * second barrier is needed before the next tile's scan. */
template <int SEG>
__device__ __forceinline__ uint32_t tile_scan(Shared &sh, uint32_t parity, uint32_t lane, uint32_t warp, uint32_t b01,
					      uint32_t b23, uint32_t sbits, uint32_t (&pos)[4])
{
	uint32_t i01 = b01, i23 = b23;

	/* shuffle with its "source lane exists" predicate feeding the adds directly; two
	 * segments per thread need one chain only */
#define AIRS_SCAN_STEP(d_)                                                                           \
	do {                                                                                         \
		if (SEG > 2)                                                                         \
			asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t0, t1;\n\t"                       \
				     "shfl.sync.up.b32 t0|p, %0, " #d_ ", 0, 0xffffffff;\n\t"            \
				     "shfl.sync.up.b32 t1, %1, " #d_ ", 0, 0xffffffff;\n\t"              \
				     "@p add.u32 %0, %0, t0;\n\t@p add.u32 %1, %1, t1;\n\t}"             \
				     : "+r"(i01), "+r"(i23));                                        \
		else                                                                                 \
			asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t0;\n\t"                           \
				     "shfl.sync.up.b32 t0|p, %0, " #d_ ", 0, 0xffffffff;\n\t"            \
				     "@p add.u32 %0, %0, t0;\n\t}"                                      \
				     : "+r"(i01));                                                   \
	} while (0)
	AIRS_SCAN_STEP(1);
	AIRS_SCAN_STEP(2);
	AIRS_SCAN_STEP(4);
	AIRS_SCAN_STEP(8);
	AIRS_SCAN_STEP(16);
#undef AIRS_SCAN_STEP
	const uint32_t t01 = __shfl_sync(kFull, i01, 31), t23 = SEG > 2 ? __shfl_sync(kFull, i23, 31) : 0u;
	const uint32_t tot0 = t01 & 0xFFFFu, tot01 = tot0 + (t01 >> 16), tot012 = tot01 + (t23 & 0xFFFFu);
	if (lane == 31)
		sh.wsum[parity][warp] = tot012 + (t23 >> 16);
	__syncthreads();
	const uint32_t ws = lane < kWarps ? sh.wsum[parity][lane] : 0u;
	const uint32_t tile_bits = __reduce_add_sync(kFull, ws);
	const uint32_t base = sbits + __reduce_add_sync(kFull, lane < warp ? ws : 0u);
	const uint32_t e01 = i01 - b01, e23 = i23 - b23; /* exclusive, both halves */
	pos[0] = base + (e01 & 0xFFFFu);
	pos[1] = base + tot0 + (e01 >> 16);
	pos[2] = base + tot01 + (e23 & 0xFFFFu);
	pos[3] = base + tot012 + (e23 >> 16);
	return tile_bits;
}

/* global memory accesses with L2 eviction hints (createpolicy descriptors): samples pass
 * through once, models are read and rewritten once per frame and should stay in the L2 */
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
	uint64_t pol;
	asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}

__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
	uint64_t pol;
	asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}

__device__ __forceinline__ uint4 ld_stream(const uint4 *p, uint64_t pol)
{
	uint4 v;
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
		     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
	return v;
}

/* 8 samples held in the low halves of 8 consecutive 32-bit words (ref sample_read_i16,
 * sample_reader.h:63-72), packed like a 16-bit container's */
__device__ __forceinline__ uint4 ld_stream32(const uint4 *p, uint64_t pol)
{
	const uint4 a = ld_stream(p, pol), b = ld_stream(p + 1, pol);
	return make_uint4(__byte_perm(a.x, a.y, 0x5410), __byte_perm(a.z, a.w, 0x5410), __byte_perm(b.x, b.y, 0x5410),
			  __byte_perm(b.z, b.w, 0x5410));
}

__device__ __forceinline__ uint4 ld_keep(const uint4 *p, uint64_t pol)
{
	uint4 v;
	asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
		     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol) : "memory");
	return v;
}

__device__ __forceinline__ void st_keep(uint4 *p, const uint4 v, uint64_t pol)
{
	asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;"
		     :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

/* biased packed residuals u = r + R of one segment (4 words = 8 samples):
 * ref preprocess.c:268-290,348-353,406-411.  w: samples, m: work buffer words
 * (model or IWT coefficients), pw: the word in front of the segment (its upper
 * half is the previous sample), Rb = R per lane, B1 = R + 1 per lane. */
__device__ __forceinline__ void seg_residuals(uint32_t pre, const uint32_t (&w)[4], const uint32_t (&m)[4], uint32_t pw,
					      uint32_t Rb, uint32_t B1, uint32_t (&u)[4])
{
	if (pre == CMP_PREPROCESS_DIFF) {
		uint32_t nbp = __vadd2(~pw, B1); /* ~v + B1 = R - v per lane */
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t nb = __vadd2(~w[k], B1);
			u[k] = __vadd2(w[k], __byte_perm(nbp, nb, 0x5432));
			nbp = nb;
		}
	} else if (pre == CMP_PREPROCESS_MODEL) {
#pragma unroll
		for (int k = 0; k < 4; k++)
			u[k] = __vadd2(w[k], __vadd2(~m[k], B1));
	} else if (pre == CMP_PREPROCESS_IWT) {
#pragma unroll
		for (int k = 0; k < 4; k++)
			u[k] = __vadd2(m[k], Rb);
	} else {
#pragma unroll
		for (int k = 0; k < 4; k++)
			u[k] = __vadd2(w[k], Rb);
	}
}

/*
 * Tiles of pieces (8 samples each) of a frame whose source (and work buffer
 * when used) is 16-byte aligned, 16-bit container.  A tile is 512 pieces; a
 * warp owns 128 consecutive pieces, lane l holds pieces l, 32 + l, 64 + l and
 * 96 + l of them ("segments" 0-3), so every 128-bit load of a warp is fully
 * coalesced.
 *
 * PRE / MM / UNC / SZ >= 0 fix the preprocessing, the model mode, "uncompressed
 * encoder" and "size only" at compile time (the hot instantiations, inlined
 * into the kernel); -1 reads them from the pass at run time (the catch-all
 * instantiations, called).
 * PARTIAL = false: the full tiles t0 .. n_tiles-1 (tile t starts at piece
 * p0 + 512 t); the loads of the next tile are in flight while one is encoded.
 * PARTIAL = true: the single tile that starts at piece p0 and ends with the
 * frame (n_pieces); a warp then owns 32, 64, 96 or 128 pieces, whatever keeps
 * all warps busy.
 *
 * Arms per warp and tile: "table" - every residual hits the pair table and
 * the eight codewords of every segment fit 64 bits: one string per segment, 3
 * shared memory reductions each; "wide table" - all in the table but a segment
 * longer than 64 bits: one string per pair, looked up again behind the scan; "raw" - the uncompressed encoder, two 64-bit strings
 * per segment; "arithmetic" - everything else, sample by sample from reloaded
 * data (rolled loops).
 *
 * Returns n_tiles, or the index of a tile it left untouched because it has to
 * go through the generic path.  Inlined into the kernel and free of calls: a
 * called function only gets the registers its caller leaves over.
 */
template <int PRE, int MM, int UNC, int SZ, bool PARTIAL, int SEG>
__device__ __forceinline__ uint32_t frame_fast(Shared &sh, const OutWin &o, uint32_t a, Cursor &c_io, uint32_t t0,
					       uint32_t n_tiles, uint32_t p0, uint32_t n_pieces, bool size_only_rt)
{
	const Pass &P = sh.pass;
	const bool size_only = SZ >= 0 ? SZ != 0 : size_only_rt;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t pre = PRE >= 0 ? (uint32_t)PRE : P.pre;
	/* MM 2 / 3: model update with a rate of 1..15, zero- / sign-extended operands */
	const uint32_t mm = MM >= 0 ? (MM >= 2 ? 2u : (uint32_t)MM) : (size_only ? 0u : P.model_mode);
	const bool unc = UNC >= 0 ? UNC != 0 : P.enc.type == CMP_ENCODER_UNCOMPRESSED;
	/* samples in the low halves of 32-bit words (catch-all instantiations only) */
	const bool c32 = PRE < 0 && P.dtype == AIRS_DTYPE_I16_IN_I32;
	/* samples big-endian in memory (catch-all instantiations only) */
	const bool be = PRE < 0 && (P.dtype & AIRS_DTYPE_BE) != 0u;
	/* table of this pass: built for this encoder (encode_pass) */
	const bool have_lut = !unc && sh.plut_key[0] == P.enc.type && sh.plut_key[1] == P.enc.g &&
			      sh.plut_key[2] == P.enc.outlier;
	const uint32_t R = have_lut ? sh.plut_R : 0u;
	const uint32_t Rb = R * 0x00010001u;        /* + R per lane */
	const uint32_t B1 = (R + 1u) * 0x00010001u; /* ~v + B1 = R - v per lane */
	const uint32_t notmask = ~((2u * R - 1u) * 0x00010001u);
	const bool diff = pre == CMP_PREPROCESS_DIFF;
	const bool use_m = pre == CMP_PREPROCESS_MODEL || pre == CMP_PREPROCESS_IWT; /* residuals read the work buffer */
	const bool need_x = pre != CMP_PREPROCESS_IWT || mm;
	const bool need_m = use_m || mm == 2u;
	/* segments per warp in this tile */
	constexpr uint32_t kTP = kThreads * SEG; /* pieces per tile */
	const uint32_t nseg = PARTIAL ? min((uint32_t)SEG, (n_pieces - p0 + kThreads - 1u) / kThreads) : (uint32_t)SEG;
	const uint32_t pw0 = p0 + warp * 32u * nseg; /* first piece of the warp in tile 0 */
	const uint4 *src4 = reinterpret_cast<const uint4 *>(P.src);
	uint4 *work4 = reinterpret_cast<uint4 *>(P.work);
	const uint16_t *src16 = reinterpret_cast<const uint16_t *>(P.src);
	const char *lut = reinterpret_cast<const char *>(sh.plut);
	const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(sh.plut);
	const uint4 zero4 = make_uint4(0, 0, 0, 0);
	const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
	/* stream bits at which the reference's writer gives up, in staging coordinates */
	const uint32_t trip = mm ? P.trip + 8u * a : 0xFFFFFFFFu;
	/* c: where the next tile's bits go (area c.buf).  While pend is set, the tile before is
	 * still staged in the other area (pend_bits bits from word pend_gw0 on) and drains behind
	 * the next scan barrier: one barrier per tile. */
	Cursor c = c_io;
	bool pend = false;
	uint32_t pend_gw0 = 0, pend_bits = 0;

	/* samples / work words of segment j_ of the tile whose warp starts at piece pw_ (none beyond the frame) */
#define AIRS_SEG_VALID(pw_, j_) (!PARTIAL || ((j_) < nseg && (pw_) + 32u * (j_) + lane < n_pieces))
#define AIRS_LOAD_X(pw_, j_) ((need_x && AIRS_SEG_VALID(pw_, j_)) ? (c32 ? ld_stream32(src4 + 2u * ((pw_) + 32u * (j_) + lane), pol_stream) \
									 : be ? airs_swap16x8(ld_stream(src4 + (pw_) + 32u * (j_) + lane, pol_stream)) \
									      : ld_stream(src4 + (pw_) + 32u * (j_) + lane, pol_stream)) : zero4)
#ifdef AIRS_EXP_NO_MODEL_IO /* (timing experiment, wrong streams: the model never travels) */
#define AIRS_LOAD_M(pw_, j_) zero4
#else
#define AIRS_LOAD_M(pw_, j_) ((need_m && AIRS_SEG_VALID(pw_, j_)) ? ld_keep(work4 + (pw_) + 32u * (j_) + lane, pol_keep) : zero4)
#endif
	/* lane 0: the sample in front of the warp's first piece (previous warp or tile) */
#define AIRS_LOAD_PS(pw_) ((diff && lane == 0 && (pw_) != 0 && AIRS_SEG_VALID(pw_, 0)) ? \
	(c32 ? __ldg(reinterpret_cast<const uint32_t *>(src16) + 8u * (pw_) - 1u) & 0xFFFFu \
	     : be ? sample_at(P.src, P.dtype, 8u * (pw_) - 1u) : (uint32_t)__ldg(src16 + 8u * (pw_) - 1u)) : 0u)

	/* the whole next tile is loaded one tile ahead (the scheduler pulls the first consumers of
	 * all four segments to the top of the loop body, so a later load would be waited for) */
	/* work buffer words (model, IWT coefficients) are loaded ahead as well where a tile is
	 * small enough for the registers (the model pass) */
	constexpr bool kAheadM = SEG <= 2;
	uint4 nx[SEG], nm[SEG];
#pragma unroll
	for (int j = 0; j < SEG; j++) {
		nx[j] = AIRS_LOAD_X(pw0 + t0 * kTP, j);
		nm[j] = kAheadM ? AIRS_LOAD_M(pw0 + t0 * kTP, j) : zero4;
	}
	uint32_t nps = AIRS_LOAD_PS(pw0 + t0 * kTP);

	uint32_t t = t0;
	for (; t < n_tiles; t++) {
		const uint32_t pw = pw0 + t * kTP;
		bool v[SEG];
#pragma unroll
		for (int j = 0; j < SEG; j++)
			v[j] = AIRS_SEG_VALID(pw, j);
		/* this thread holds sample 0 of the frame and a table is there to be missed */
		const bool first = diff && R != 0u && pw == 0u && lane == 0u;

		uint32_t w[SEG][4], m[SEG][4];
#pragma unroll
		for (int j = 0; j < SEG; j++) {
			const uint4 mj = kAheadM ? nm[j] : AIRS_LOAD_M(pw, j);
			w[j][0] = nx[j].x; w[j][1] = nx[j].y; w[j][2] = nx[j].z; w[j][3] = nx[j].w;
			m[j][0] = mj.x; m[j][1] = mj.y; m[j][2] = mj.z; m[j][3] = mj.w;
			if (kExpNoModelIo && use_m) { /* (small residuals of about the real length) */
				m[j][0] = __vsub2(w[j][0], 0x00050002u); m[j][1] = __vsub2(w[j][1], 0x0001000Bu);
				m[j][2] = __vsub2(w[j][2], 0x00070003u); m[j][3] = __vsub2(w[j][3], 0x00020009u);
			}
		}
		const uint32_t ps = nps;
		/* the model words of the next tile come from the L2: requested a whole tile ahead */
		if (kAheadM && !PARTIAL && t + 1u < n_tiles) {
#pragma unroll
			for (int j = 0; j < SEG; j++)
				nm[j] = AIRS_LOAD_M(pw + kTP, j);
		}

		/* residuals of the four segments; the new model takes the place of the old one
		 * (ref cmp.c:304-311) */
		uint32_t u[SEG][4];
		uint32_t chk = 0;
		{
			const uint32_t src_lane = (lane - 1u) & 31u;
			uint32_t front0 = ps << 16; /* lane 0: the word in front of its segment j */
#pragma unroll
			for (int j = 0; j < SEG; j++) {
				uint32_t pw_word = 0;
				if (diff) {
					/* lanes 1-31 receive their left neighbour's last word; lane 0 receives lane
					 * 31's, which is in front of its segment j + 1 */
					const uint32_t up = __shfl_sync(kFull, w[j][3], src_lane);
					pw_word = lane ? up : front0;
					front0 = up;
				}
				seg_residuals(pre, w[j], m[j], pw_word, Rb, B1, u[j]);
				/* the first sample of a frame has no predecessor: its "difference" is the
				 * sample itself, far outside any table.  The thread that holds it encodes it
				 * on its own (first_sample_*) and lets residual 0 stand in here. */
				if (j == 0 && first)
					u[0][0] = (u[0][0] & 0xFFFF0000u) | (Rb & 0xFFFFu);
				if (mm == 2u) {
					/* rate 16 keeps the model as it is, rate 0 replaces it by the samples */
					const uint32_t rate = P.rate, wdp = (rate << 4) | ((16u - rate) << 12);
					if (MM == 2 || MM == 3) {
#pragma unroll
						for (int k = 0; k < 4; k++)
							m[j][k] = MM == 3 ? model_update2<true>(w[j][k], m[j][k], wdp)
									  : model_update2<false>(w[j][k], m[j][k], wdp);
					} else if (rate == 0u) {
#pragma unroll
						for (int k = 0; k < 4; k++)
							m[j][k] = w[j][k];
					} else if (rate < 16u) {
						if (P.is_signed) {
#pragma unroll
							for (int k = 0; k < 4; k++)
								m[j][k] = model_update2<true>(w[j][k], m[j][k], wdp);
						} else {
#pragma unroll
							for (int k = 0; k < 4; k++)
								m[j][k] = model_update2<false>(w[j][k], m[j][k], wdp);
						}
					}
				} else if (mm == 1u) {
#pragma unroll
					for (int k = 0; k < 4; k++)
						m[j][k] = w[j][k];
				}
#pragma unroll
				for (int k = 0; k < 4; k++) {
					if (PARTIAL) /* pieces beyond the frame: residual 0, so that they do not spoil the table check */
						u[j][k] = v[j] ? u[j][k] : Rb;
					chk |= u[j][k];
				}
			}
		}

		/* the samples are used up: their registers take the next tile's (requested here, not
		 * at the top of the tile, so that no second copy of a tile is alive) */
		if (!PARTIAL && t + 1u < n_tiles) {
#pragma unroll
			for (int j = 0; j < SEG; j++)
				nx[j] = AIRS_LOAD_X(pw + kTP, j);
			nps = AIRS_LOAD_PS(pw + kTP);
		}

		uint32_t tile_bits;
		uint32_t action; /* 0: bits staged, 1: size only, 2: tile handed back */
		uint32_t pos[4];
#define AIRS_AFTER_SCAN(put_)                                                                        \
	do {                                                                                         \
		/* a tile too big for the staging area, or one that crosses the point where the  \
		 * reference's writer gives up (the model then has to stay exact sample by       \
		 * sample, ref cmp.c:300-311), goes back untouched */                            \
		action = size_only ? 1u : ((c.sbits + tile_bits > kStgBits ||                        \
					    c.gw0 * 32u + c.sbits + tile_bits >= trip) ? 2u : 0u);   \
		if (action == 0u) {                                                                  \
			/* the scan barrier is behind us: the tile before is completely staged */   \
			if (pend)                                                                    \
				copy_out(sh, o, c.buf ^ 1u, pend_gw0, pend_bits, c.buf);             \
			if (mm && !kExpNoModelIo) {                                                  \
				_Pragma("unroll") for (int j = 0; j < SEG; j++)                \
					if (v[j])                                                    \
						st_keep(work4 + pw + 32u * j + lane, make_uint4(m[j][0], m[j][1], m[j][2], m[j][3]), pol_keep); \
			}                                                                            \
			uint32_t *stg = stg_of(sh, c.buf);                                           \
			put_                                                                         \
		}                                                                                    \
	} while (0)

		/* ---- table arm, first half: codewords of pairs, quads, segments */
		const bool hit = !unc && R != 0u && __all_sync(kFull, (chk & notmask) == 0u);
		bool table = hit;
		uint32_t sh_[SEG], sl_[SEG], sn_[SEG]; /* one string per segment: hi, lo, length */
		if (hit) {
			uint32_t qchk = 0;
#pragma unroll
			for (int j = 0; j < SEG; j++) {
				uint32_t pc[4], pl[4];
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const uint32_t ent = lut_pair(lut_s, u[j][k]);
					pc[k] = ent & ((1u << kLutLenShift) - 1u);
					pl[k] = ent >> kLutLenShift;
				}
				if (j == 0 && first) { /* drop the stand-in's codeword from the head of pair 0 */
					pl[0] -= sh.slut[kLutR].y;
					pc[0] &= (1u << pl[0]) - 1u;
				}
				/* the four pair strings appended one after the other: 64-bit shifts by < 32 */
				uint32_t lo = pc[0], hi = 0u, n = pl[0];
#pragma unroll
				for (int k = 1; k < 4; k++) {
					hi = __funnelshift_l(lo, hi, pl[k]);
					lo = (lo << pl[k]) | pc[k];
					n += pl[k];
				}
				qchk |= n + 63u; /* bit 7 set: a segment longer than 64 bits */
				sl_[j] = (PARTIAL && !v[j]) ? 0u : lo;
				sh_[j] = (PARTIAL && !v[j]) ? 0u : hi;
				sn_[j] = (PARTIAL && !v[j]) ? 0u : n;
			}
			table = __all_sync(kFull, (qchk & 128u) == 0u);
		}

		if (table) {
			/* bits of the frame's first sample, in front of this thread's segment 0 (thread 0
			 * wrote them itself, encode_pass) */
			const uint32_t n_first = first ? sh.first_code[2] : 0u;
			tile_bits = tile_scan<SEG>(sh, t & 1u, lane, warp, (sn_[0] + n_first) | (sn_[1] << 16),
					      SEG > 2 ? sn_[SEG - 2] | (sn_[SEG - 1] << 16) : 0u, c.sbits, pos);
			AIRS_AFTER_SCAN({
				if (first) {
					int32_t ne = -(int32_t)pos[0];
					put_unit(stg, ne, sh.first_code[0], sh.first_code[1], n_first);
					pos[0] += n_first;
				}
				_Pragma("unroll") for (int j = 0; j < SEG; j++) {
					int32_t ne = -(int32_t)pos[j];
					put_unit(stg, ne, sh_[j], sl_[j], sn_[j]);
				}
			});
		} else if (hit) {
			/* ---- wide table arm: every residual is in the table, but a segment is longer than
			 * 64 bits (more than 8 bits per sample).  The lengths are known; the biased
			 * residuals wait in local memory and are looked up again pair by pair. */
			uint32_t d[SEG * 4];
#pragma unroll
			for (int j = 0; j < SEG; j++)
#pragma unroll
				for (int k = 0; k < 4; k++)
					d[4 * j + k] = u[j][k];
			const uint32_t n_first = first ? sh.first_code[2] : 0u;
			tile_bits = tile_scan<SEG>(sh, t & 1u, lane, warp, (sn_[0] + n_first) | (sn_[1] << 16),
					      SEG > 2 ? sn_[SEG - 2] | (sn_[SEG - 1] << 16) : 0u, c.sbits, pos);
			AIRS_AFTER_SCAN({
				_Pragma("unroll 1") for (uint32_t j = 0; j < (uint32_t)SEG; j++) {
					if (PARTIAL && !(j < nseg && pw + 32u * j + lane < n_pieces))
						continue;
					int32_t ne = -(int32_t)(j == 0u ? pos[0] : j == 1u ? pos[1] : j == 2u ? pos[2] : pos[3]);
					if (j == 0u && first)
						put_unit(stg, ne, sh.first_code[0], sh.first_code[1], n_first);
					_Pragma("unroll 1") for (uint32_t k = 0; k < 4u; k++) {
						const uint32_t uu = d[4u * j + k];
						const uint32_t off = 4u * ((uu & 0xFFFFu) + kLutPitch * (uu >> 16));
						const uint32_t ent = *reinterpret_cast<const uint32_t *>(lut + off);
						uint32_t pc = ent & ((1u << kLutLenShift) - 1u);
						uint32_t pl = ent >> kLutLenShift;
						if (j == 0u && k == 0u && first) { /* drop the stand-in's codeword */
							pl -= sh.slut[kLutR].y;
							pc &= (1u << pl) - 1u;
						}
						put_unit(stg, ne, 0u, pc, pl);
					}
				}
			});
		} else if (unc) {
			/* ---- raw arm: 16 bits per sample, two 64-bit strings per segment */
			const uint32_t nb = 128u;
			tile_bits = tile_scan<SEG>(sh, t & 1u, lane, warp, (v[0] ? nb : 0u) | ((v[1] ? nb : 0u) << 16),
					      SEG > 2 ? (v[SEG - 2] ? nb : 0u) | ((v[SEG - 1] ? nb : 0u) << 16) : 0u, c.sbits, pos);
			AIRS_AFTER_SCAN({
				_Pragma("unroll") for (int j = 0; j < SEG; j++) {
					if (v[j]) {
						int32_t ne = -(int32_t)pos[j];
						/* first sample of a word in the upper half */
						put_unit(stg, ne, __byte_perm(u[j][0], 0u, 0x1032), __byte_perm(u[j][1], 0u, 0x1032), 64u);
						put_unit(stg, ne, __byte_perm(u[j][2], 0u, 0x1032), __byte_perm(u[j][3], 0u, 0x1032), 64u);
					}
				}
			});
		} else {
			/* ---- arithmetic arm, row by row: a "row" is segment j of all 32 lanes (256 samples).
			 * Rows whose residuals are all in the table (and whose segments fit 64 bits) still
			 * go through the table, one string per segment; the other rows are encoded sample by
			 * sample from plain residuals.  A few outliers in a tile thus cost their rows, not
			 * the whole warp tile.  Everything lives in local memory and rolled loops: this arm
			 * must not cost the table arm its registers. */
			uint32_t d[SEG * 4];    /* biased residuals; plain ones in the rows encoded arithmetically */
			uint32_t rstr[SEG * 3]; /* hi, lo, length of the strings of the rows on the table */
			uint32_t cwords[SEG * 16]; /* code words of the other rows, computed once (slow_codes) */
			uint32_t miss = 0;      /* bit j: row j is encoded arithmetically */
#pragma unroll
			for (int j = 0; j < SEG; j++)
#pragma unroll
				for (int k = 0; k < 4; k++)
					d[4 * j + k] = u[j][k];
			{
				const uint32_t negRb = ((0x10000u - R) & 0xFFFFu) * 0x00010001u;
				uint32_t b01 = 0, b23 = 0;
#pragma unroll 1
				for (uint32_t j = 0; j < (uint32_t)SEG; j++) {
					const bool valid = !PARTIAL || (j < nseg && pw + 32u * j + lane < n_pieces);
					uint32_t w4[4];
#pragma unroll
					for (int k = 0; k < 4; k++)
						w4[k] = d[4u * j + k];
					bool row_hit = !unc && R != 0u && !(diff && pw == 0u && j == 0u) &&
						       __all_sync(kFull, ((w4[0] | w4[1] | w4[2] | w4[3]) & notmask) == 0u);
					if (row_hit) {
						uint32_t lo = 0, hi = 0, n = 0;
#pragma unroll
						for (int k = 0; k < 4; k++) {
							const uint32_t off = 4u * ((w4[k] & 0xFFFFu) + kLutPitch * (w4[k] >> 16));
							const uint32_t ent = *reinterpret_cast<const uint32_t *>(lut + off);
							const uint32_t pl = ent >> kLutLenShift;
							hi = __funnelshift_l(lo, hi, pl);
							lo = (lo << pl) | (ent & ((1u << kLutLenShift) - 1u));
							n += pl;
						}
						row_hit = __all_sync(kFull, n <= 64u);
						rstr[3u * j] = valid ? hi : 0u;
						rstr[3u * j + 1u] = valid ? lo : 0u;
						rstr[3u * j + 2u] = valid ? n : 0u;
						if (row_hit) {
							if (j < 2u)
								b01 += (valid ? n : 0u) << (16u * (j & 1u));
							else
								b23 += (valid ? n : 0u) << (16u * (j & 1u));
						}
					}
					if (!row_hit) {
						miss |= 1u << j;
#pragma unroll
						for (int k = 0; k < 4; k++)
							d[4u * j + k] = __vadd2(w4[k], negRb);
					}
				}
				if (first) /* the stand-in goes, the sample comes back (its row is never on the table) */
					d[0] = (d[0] & 0xFFFF0000u) | sample_at(P.src, P.dtype, 0);
				uint32_t a01, a23;
				slow_codes(P.enc, d, 4u * SEG, miss, cwords, a01, a23);
				if (PARTIAL) {
					a01 = (v[0] ? a01 & 0xFFFFu : 0u) | (v[1] ? a01 & 0xFFFF0000u : 0u);
					if (SEG > 2)
						a23 = (v[SEG - 2] ? a23 & 0xFFFFu : 0u) | (v[SEG - 1] ? a23 & 0xFFFF0000u : 0u);
				}
				tile_bits = tile_scan<SEG>(sh, t & 1u, lane, warp, a01 + b01, a23 + b23, c.sbits, pos);
			}
			AIRS_AFTER_SCAN({
				_Pragma("unroll 1") for (uint32_t j = 0; j < (uint32_t)SEG; j++) {
					const uint32_t pj = j == 0u ? pos[0] : j == 1u ? pos[1] : j == 2u ? pos[2] : pos[3];
					if ((miss >> j) & 1u) {
						if (!PARTIAL || (j < nseg && pw + 32u * j + lane < n_pieces))
							slow_put_codes(cwords + 16u * j, stg, pj);
					} else {
						int32_t ne = -(int32_t)pj;
						put_unit(stg, ne, rstr[3u * j], rstr[3u * j + 1u], rstr[3u * j + 2u]);
					}
				}
			});
		}
#undef AIRS_AFTER_SCAN
		if (action == 1u) {
			cursor_advance(c, tile_bits);
			continue;
		}
		if (action == 2u) /* nothing of this tile has been staged or stored: it goes back to the caller
				   * and through the generic path */
			break;
		pend = true;
		pend_gw0 = c.gw0;
		pend_bits = c.sbits + tile_bits;
		cursor_advance(c, tile_bits);
		c.buf ^= 1u;
	}
#undef AIRS_SEG_VALID
#undef AIRS_LOAD_X
#undef AIRS_LOAD_M
#undef AIRS_LOAD_PS
	/* drain the tile still staged; what is left of it stays in its own area, which becomes
	 * the current one again */
	__syncthreads();
	if (pend) {
		c.buf ^= 1u;
		copy_out(sh, o, c.buf, pend_gw0, pend_bits, c.buf);
	}
	c_io = c;
	return t;
}

/* the catch-all instantiations: any preprocessing / model mode / encoder, sizing passes */
__device__ __noinline__ uint32_t frame_fast_full_rt(Shared &sh, const OutWin o, uint32_t a, Cursor &c, uint32_t t0,
						    uint32_t n_tiles, uint32_t n_pieces, bool size_only)
{
	return frame_fast<-1, -1, -1, -1, false, kSeg>(sh, o, a, c, t0, n_tiles, 0u, n_pieces, size_only);
}

/* ... with half-size tiles: the uncompressed encoder's 16 bits per sample must fit the staging area */
__device__ __noinline__ uint32_t frame_fast_half_rt(Shared &sh, const OutWin o, uint32_t a, Cursor &c, uint32_t t0,
						    uint32_t n_tiles, uint32_t n_pieces, bool size_only)
{
	return frame_fast<-1, -1, -1, -1, false, kSegModel>(sh, o, a, c, t0, n_tiles, 0u, n_pieces, size_only);
}

__device__ __noinline__ uint32_t frame_fast_tail_rt(Shared &sh, const OutWin o, uint32_t a, Cursor &c, uint32_t p0,
						    uint32_t n_pieces, bool size_only)
{
	return frame_fast<-1, -1, -1, -1, true, kSeg>(sh, o, a, c, 0u, 1u, p0, n_pieces, size_only);
}

/* all pieces of a frame: full tiles through the instantiation specialised for
 * this pass when there is one, else through the catch-all, like the partial
 * tile; a tile the fast path hands back goes through the generic path.  The
 * model pass works on half-size tiles (two segments per thread): it holds old
 * and new model words next to the samples. */
__device__ __forceinline__ void frame_fast_any(Shared &sh, const OutWin &o, uint32_t a, Cursor &c, uint32_t n_pieces,
					       bool size_only)
{
	const Pass &P = sh.pass;
	/* pass key: preprocessing * 4 + model mode, mode 2 -> 3 for sign-extended model updates;
	 * passes without a specialised instantiation get 99 */
	uint32_t key = P.pre * 4u + P.model_mode;
	if (P.model_mode == 2u)
		key = (P.rate >= 1u && P.rate <= 15u) ? key + P.is_signed : 99u;
	if (P.enc.type == CMP_ENCODER_UNCOMPRESSED || size_only || P.dtype == AIRS_DTYPE_I16_IN_I32 || (P.dtype & AIRS_DTYPE_BE))
		key = 99u;
	const bool model_pass = key == CMP_PREPROCESS_MODEL * 4u + 2u || key == CMP_PREPROCESS_MODEL * 4u + 3u;
	const bool half = model_pass || P.enc.type == CMP_ENCODER_UNCOMPRESSED;
	const uint32_t tp = half ? kThreads * kSegModel : kTilePieces; /* pieces per tile */
	const uint32_t n_full = n_pieces / tp;

	for (uint32_t t = 0; t < n_full;) {
		switch (key) {
#define AIRS_HOT(pre_, mm_, seg_)                                                                              \
	case (pre_) * 4u + (mm_):                                                                              \
		t = frame_fast<(pre_), (mm_), 0, 0, false, (seg_)>(sh, o, a, c, t, n_full, 0u, n_pieces, false); \
		break;
		AIRS_HOT(CMP_PREPROCESS_NONE, 0, kSeg)
		AIRS_HOT(CMP_PREPROCESS_DIFF, 0, kSeg)
		AIRS_HOT(CMP_PREPROCESS_DIFF, 1, kSeg)
		AIRS_HOT(CMP_PREPROCESS_MODEL, 2, kSegModel)
		AIRS_HOT(CMP_PREPROCESS_MODEL, 3, kSegModel)
#undef AIRS_HOT
		default:
			t = half ? frame_fast_half_rt(sh, o, a, c, t, n_full, n_pieces, size_only)
				 : frame_fast_full_rt(sh, o, a, c, t, n_full, n_pieces, size_only);
			break;
		}
		if (t < n_full) {
			c = generic_span(sh, o, a, c, t * tp * 8u, (t + 1u) * tp * 8u, size_only);
			t++;
		}
	}
	if (n_full * tp < n_pieces) {
		if (!frame_fast_tail_rt(sh, o, a, c, n_full * tp, n_pieces, size_only))
			c = generic_span(sh, o, a, c, n_full * tp * 8u, n_pieces * 8u, size_only);
	}
}

/*
 * A run of consecutive secondary MODEL passes of a context (ref cmp.c:228-262, 296-312) with the FRAMES IN THE INNER
 * LOOP: for every tile position (4096 samples) the CTA walks through all frames of the run, so that the model of the
 * position lives in registers from the first frame to the last (32 samples a thread; read from the work buffer once
 * and written back once per run) - the frame-by-frame order reads and rewrites the whole model through the L2 once
 * per frame, and with a context per CTA those are 58 MB in flight that the L2 does not hold (1.54 x the algorithmic
 * DRAM traffic).  The tiles of a frame are still visited in stream order, `run` visits apart: where its stream stands
 * and its last incomplete 16-byte group wait in shared memory (cx_bit, cx_carry).  One barrier per visit as in
 * frame_fast(); the samples of the next visit are in flight while one is encoded.
 * Frames f0 .. f0 + run - 1 of the job; the first one has sequence number seq0; all carry the context's identifier.
 * Warps whose residuals are not all in the pair table compute their code words arithmetically.  A stream that leaves
 * its slot (or comes near the point where the reference's writer gives up) - returns false: streams and model are
 * half done, the caller starts the context over frame by frame.
 */
/* the arithmetic arm of model_run_fast(), behind calls: the code words of the two segments of a thread from their
 * biased residuals d[0..7] into cw[0..31] (returns the segments' bit counts, packed), and from there into the staging
 * area */
__device__ __noinline__ uint32_t run_slow_codes(const EncConst &e, uint32_t *d, uint32_t R, uint32_t *cw, uint32_t &b23)
{
	const uint32_t negRb = ((0x10000u - R) & 0xFFFFu) * 0x00010001u;
	uint32_t b01;
#pragma unroll 1
	for (uint32_t k = 0; k < 4u * kSegRun; k++)
		d[k] = __vadd2(d[k], negRb);
	slow_codes(e, d, 4u * kSegRun, (1u << kSegRun) - 1u, cw, b01, b23);
	return b01;
}

__device__ __noinline__ void run_slow_put(const uint32_t *cw, uint32_t *stg, const uint32_t *pos)
{
#pragma unroll 1
	for (uint32_t j = 0; j < kSegRun; j++)
		slow_put_codes(cw + 16u * j, stg, pos[j]);
}

template <bool SIGNED>
__device__ __noinline__ bool model_run_fast(Shared &sh, const AirsLaunch &b, uint32_t f0, uint32_t run, uint32_t seq0)
{
	const JobPlan &pl = sh.plan;
	const airs_job &jb = sh.job;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	constexpr uint32_t SEG = kSegRun, kTP = kThreads * SEG;
	const uint32_t n_pos = pl.n / (8u * kTP);
	const uint8_t *src0 = b.src + jb.src_offset + (uint64_t)f0 * jb.src_frame_stride;
	uint8_t *dst0 = b.dst + jb.dst_offset + (uint64_t)f0 * jb.dst_frame_stride;
	uint4 *work4 = reinterpret_cast<uint4 *>(b.work + jb.work_offset);
	const uint64_t sstride = jb.src_frame_stride, dstride = jb.dst_frame_stride;
	const uint32_t hdr_len = CMP_HDR_SIZE + 6u, cap = pl.cap_eff;
	const uint32_t rate = pl.rate, wdp = (rate << 4) | ((16u - rate) << 12);
	const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(sh.plut);
	const uint64_t pol_stream = l2_policy_evict_first();
	const uint4 zero4 = make_uint4(0, 0, 0, 0);
	const EncConst &e = pl.enc[1];

	for (uint32_t i = tid; i < run; i += kThreads) {
		const uint32_t a = (uint32_t)((uintptr_t)(dst0 + i * dstride) & 15u);
		sh.cx_bit[i] = 8u * (a + hdr_len);
		sh.cx_carry[i] = zero4;
	}
	if (tid == 0)
		sh.cx_abort = 0;
	if (sh.plut_key[0] != e.type || sh.plut_key[1] != e.g || sh.plut_key[2] != e.outlier)
		build_pair_lut(sh, e);
	__syncthreads();
	const uint32_t R = sh.plut_R;
	if (R == 0u)
		return false; /* (nothing touched yet - but the caller does not tell) */
	const uint32_t Rb = R * 0x00010001u, B1 = (R + 1u) * 0x00010001u, notmask = ~((2u * R - 1u) * 0x00010001u);

	bool ok = true, pend = false;
	uint32_t it = 0, buf = 0, pend_i = 0, pend_gw0 = 0, pend_bits = 0;
	uint8_t *pend_d = dst0;

	/* samples of segment j_ of the visit (frame i_ of the run, first piece of the warp pw_) */
#define AIRS_CX_X(i_, pw_, j_) ld_stream(reinterpret_cast<const uint4 *>(src0 + (i_) * sstride) + (pw_) + 32u * (j_) + lane, pol_stream)
	uint4 nx[SEG];
#pragma unroll
	for (int j = 0; j < SEG; j++)
		nx[j] = AIRS_CX_X(0u, warp * 32u * SEG, j);

	for (uint32_t t = 0; t < n_pos && ok; t++) {
		const uint32_t pw = t * kTP + warp * 32u * SEG;
		uint32_t m[SEG][4];
#pragma unroll
		for (int j = 0; j < SEG; j++) {
			const uint4 mj = work4[pw + 32u * j + lane];
			m[j][0] = mj.x; m[j][1] = mj.y; m[j][2] = mj.z; m[j][3] = mj.w;
		}
		uint8_t *fd = dst0;        /* slot of the visit's frame */
		const uint8_t *xs = src0;  /* ... and its samples */
		for (uint32_t i = 0; i < run; i++, it++) {
			uint32_t w[SEG][4];
#pragma unroll
			for (int j = 0; j < SEG; j++) {
				w[j][0] = nx[j].x; w[j][1] = nx[j].y; w[j][2] = nx[j].z; w[j][3] = nx[j].w;
			}
			{ /* the next visit: the next frame at this position, or the first frame at the next one */
				const bool wrap = i + 1u == run;
				if (!wrap || t + 1u < n_pos) {
					const uint4 *nx4 = reinterpret_cast<const uint4 *>(wrap ? src0 : xs + sstride) + (wrap ? pw + kTP : pw) + lane;
#pragma unroll
					for (int j = 0; j < SEG; j++)
						nx[j] = ld_stream(nx4 + 32u * j, pol_stream);
				}
			}
			uint32_t u[SEG][4], chk = 0;
#pragma unroll
			for (int j = 0; j < SEG; j++) {
				/* (per 16-bit half and modulo 2^16: one 32-bit subtraction per word is NOT the same - samples on either
				 * side of the 0 / 65535 boundary give a small residual with a borrow that nothing takes back) */
				seg_residuals(CMP_PREPROCESS_MODEL, w[j], m[j], 0u, Rb, B1, u[j]);
#pragma unroll
				for (int k = 0; k < 4; k++) /* (rates 1 .. 15: model_run_ok()) */
					m[j][k] = model_update2<SIGNED>(w[j][k], m[j][k], wdp);
#pragma unroll
				for (int k = 0; k < 4; k++)
					chk |= u[j][k];
			}
			/* the table arm of frame_fast(), or nothing */
			uint32_t sh_[SEG], sl_[SEG], sn_[SEG], qchk = 128u;
			const bool hit = __all_sync(kFull, (chk & notmask) == 0u);
			if (hit) {
				qchk = 0;
#pragma unroll
				for (int j = 0; j < SEG; j++) {
					uint32_t pc[4], pl_[4];
#pragma unroll
					for (int k = 0; k < 4; k++) {
						const uint32_t ent = lut_pair(lut_s, u[j][k]);
						pc[k] = ent & ((1u << kLutLenShift) - 1u);
						pl_[k] = ent >> kLutLenShift;
					}
					uint32_t lo = pc[0], hi = 0u, nb = pl_[0];
#pragma unroll
					for (int k = 1; k < 4; k++) {
						hi = __funnelshift_l(lo, hi, pl_[k]);
						lo = (lo << pl_[k]) | pc[k];
						nb += pl_[k];
					}
					qchk |= nb + 63u; /* bit 7 set: a segment longer than 64 bits */
					sl_[j] = lo; sh_[j] = hi; sn_[j] = nb;
				}
			} else {
#pragma unroll
				for (int j = 0; j < SEG; j++)
					sl_[j] = sh_[j] = sn_[j] = 0;
			}
			/* a warp with a residual outside the table or a segment longer than 64 bits: all its code words
			 * arithmetically, from plain residuals (the arithmetic arm of frame_fast(), for the whole warp) */
			const bool slow = !(hit && __all_sync(kFull, (qchk & 128u) == 0u));
			uint32_t cwords[SEG * 16]; /* (local memory, touched by the two calls below only) */
			uint32_t b01 = sn_[0] | (sn_[1] << 16), b23 = SEG > 2 ? sn_[SEG - 2] | (sn_[SEG - 1] << 16) : 0u;
			if (slow) {
				uint32_t d[SEG * 4];
#pragma unroll
				for (int j = 0; j < SEG; j++)
#pragma unroll
					for (int k = 0; k < 4; k++)
						d[4 * j + k] = u[j][k];
				b01 = run_slow_codes(e, d, R, cwords, b23);
			}
			const uint32_t bit = sh.cx_bit[i];
			const uint32_t gw0 = (bit >> 7) << 2, sbits = bit & 127u;
			uint32_t pos[4];
			const uint32_t tile_bits = tile_scan<SEG>(sh, it & 1u, lane, warp, b01, b23, sbits, pos);
			/* behind the scan barrier: everything of the visit before is staged; every thread takes the same turns */
			if (pend) {
				const uint32_t pa = (uint32_t)((uintptr_t)pend_d & 15u);
				OutWin po;
				po.base = pend_d - pa;
				po.lo = pa + hdr_len;
				po.hi = pa + cap;
				copy_out(sh, po, buf ^ 1u, pend_gw0, pend_bits, buf ^ 1u, &sh.cx_carry[pend_i]);
				pend = false;
			}
			if (sbits + tile_bits > kStgBits) { /* (no stream can leave its slot: model_run_ok()) */
				ok = false; /* a tile beyond the staging area (most of its samples escape): the frame-by-frame path */
#ifdef AIRS_CTX_DEBUG
				if (tid == 0 && atomicAdd(&b.ticket[41], 1u) == 0u) {
					b.ticket[42] = t; b.ticket[43] = i; b.ticket[44] = 2u; b.ticket[45] = R; b.ticket[46] = bit; b.ticket[47] = tile_bits;
				}
#endif
				break;
			}
			uint32_t *stg = stg_of(sh, buf);
			if (tid == 0) { /* the frame's incomplete group from its tile before */
				const uint4 cg = sh.cx_carry[i];
				atomicOr(stg, cg.x);
				atomicOr(stg + 1, cg.y);
				atomicOr(stg + 2, cg.z);
				atomicOr(stg + 3, cg.w);
				sh.cx_bit[i] = bit + tile_bits;
			}
			if (slow) {
				run_slow_put(cwords, stg, pos);
			} else {
#pragma unroll
				for (int j = 0; j < SEG; j++) {
					int32_t ne = -(int32_t)pos[j];
					put_unit(stg, ne, sh_[j], sl_[j], sn_[j]);
				}
			}
			pend = true;
			pend_i = i;
			pend_d = fd;
			pend_gw0 = gw0;
			pend_bits = sbits + tile_bits;
			buf ^= 1u;
			fd += dstride;
			xs += sstride;
		}
		if (ok) { /* the model of this position behind the run's last frame (ref cmp.c:304-311) */
#pragma unroll
			for (int j = 0; j < SEG; j++)
				work4[pw + 32u * j + lane] = make_uint4(m[j][0], m[j][1], m[j][2], m[j][3]);
		}
	}
#undef AIRS_CX_X
	__syncthreads();
	if (ok && pend) {
		const uint32_t pa = (uint32_t)((uintptr_t)pend_d & 15u);
		OutWin po;
		po.base = pend_d - pa;
		po.lo = pa + hdr_len;
		po.hi = pa + cap;
		copy_out(sh, po, buf ^ 1u, pend_gw0, pend_bits, buf ^ 1u, &sh.cx_carry[pend_i]);
	}
	__syncthreads();
	/* sizes; a stream that does not fit sends the context to the frame-by-frame path */
	const uint32_t checksum = (pl.flags & AIRS_PF_CHECKSUM) ? 1u : 0u;
	if (ok) {
		for (uint32_t i = tid; i < run; i += kThreads) {
			const uint32_t a = (uint32_t)((uintptr_t)(dst0 + i * dstride) & 15u);
			const uint32_t size = ((sh.cx_bit[i] - 8u * a + 7u) >> 3) + 4u * checksum;
			if (size > cap || size > CMP_HDR_MAX_COMPRESSED_SIZE)
				sh.cx_abort = 1;
		}
	}
	__syncthreads();
#ifdef AIRS_CTX_DEBUG
	if (tid == 0)
		atomicAdd(&b.ticket[(!ok || sh.cx_abort) ? 48 : 40], 1u);
#endif
	if (!ok || sh.cx_abort) {
		for (uint32_t wd = tid; wd < 2u * (4u + kStgWords); wd += kThreads)
			(&sh.stg_mem[0][0])[wd] = 0;
		__syncthreads();
		return false;
	}
	/* last bytes (zero padded, ref bitstream_writer.h:205-227), headers (ref cmp.c:265-279, 329-334), results */
	for (uint32_t i = tid; i < run; i += kThreads) {
		uint8_t *fd = dst0 + i * dstride;
		const uint32_t a = (uint32_t)((uintptr_t)fd & 15u);
		const uint32_t bit = sh.cx_bit[i], size = ((bit - 8u * a + 7u) >> 3) + 4u * checksum;
		const uint32_t cg[4] = {sh.cx_carry[i].x, sh.cx_carry[i].y, sh.cx_carry[i].z, sh.cx_carry[i].w};
		uint8_t *gp = fd - a + (size_t)(bit >> 7) * 16u;
		const uint32_t nb = ((bit & 127u) + 7u) >> 3;
#pragma unroll 1
		for (uint32_t k = 0; k < nb; k++)
			gp[k] = (uint8_t)((k < 4u ? cg[0] : k < 8u ? cg[1] : k < 12u ? cg[2] : cg[3]) >> (24u - 8u * (k & 3u)));
		Pass P;
		P.enc = e;
		P.pre = CMP_PREPROCESS_MODEL;
		P.n = pl.n;
		P.identifier = sh.ctx.identifier;
		P.seq = seq0 + i;
		P.checksum = checksum;
		P.rate = rate;
#pragma unroll 1
		for (uint32_t k = 0; k < hdr_len; k++)
			fd[k] = (uint8_t)header_byte(P, k, size);
		b.results[jb.first_result + f0 + i] = size;
	}
	return true;
}

/* what model_run_fast() asks of a job (uniform over the CTA) */
__device__ __forceinline__ bool model_run_ok(const Shared &sh, const AirsLaunch &b)
{
	const JobPlan &pl = sh.plan;
	const airs_job &jb = sh.job;
	return b.layout == AIRS_LAYOUT_SLOTS && !b.ctx_io && b.dst && b.work && b.src && (pl.flags & AIRS_PF_VALID) &&
	       (pl.flags & AIRS_PF_MODEL) && !(pl.flags & AIRS_PF_BE) && !pl.frame_err && !pl.orig_err && !pl.pre_err[1] &&
	       !pl.model_err && pl.pre[1] == CMP_PREPROCESS_MODEL && pl.enc[1].type != CMP_ENCODER_UNCOMPRESSED &&
	       pl.sec_iter >= 2u && jb.dtype != AIRS_DTYPE_I16_IN_I32 && jb.n_frames >= 3u && pl.rate >= 1u && pl.rate <= 15u &&
	       pl.n >= 8u * kThreads * kSegRun && pl.n % (8u * kThreads * kSegRun) == 0 && pl.cap_eff >= CMP_HDR_SIZE + 6u &&
	       ((uintptr_t)(b.src + jb.src_offset) & 15u) == 0 && (jb.src_frame_stride & 15u) == 0 &&
	       ((uintptr_t)(b.work + jb.work_offset) & 15u) == 0 && ((uintptr_t)(b.dst + jb.dst_offset) & 7u) == 0 &&
	       (jb.dst_frame_stride & 7u) == 0 &&
	       /* no frame can fail or fall back (slots of cmp_compress_bound() bytes: 48 bits a sample, ref cmp.c:59-74): a run
		* writes the streams of all its frames side by side, and a frame that fails changes the passes - and the streams -
		* of every frame behind it (ref cmp.c:228-262), in whose slots bytes behind the final stream would then stay */
	       !(pl.flags & AIRS_PF_FALLBACK_OK) && !jb.params.uncompressed_fallback_enabled &&
	       (uint64_t)jb.dst_capacity >= CMP_HDR_SIZE + 6u + 4u + 6ull * pl.n;
}

/* one pass over one frame; returns the stream size or an error (uniform over the CTA).
 * ref compress_engine, cmp.c:213-338 */
__device__ __forceinline__ uint32_t encode_pass(Shared &sh, bool size_only, bool suppress)
{
	const Pass &P = sh.pass;
	const uint32_t tid = threadIdx.x;

	if (P.err)
		return P.err;

	const uint32_t a = (uint32_t)((uintptr_t)P.dst & 15u);
	OutWin o;
	o.base = P.dst - a;
	o.lo = a + P.hdr_len;
	o.hi = suppress ? o.lo : a + P.cap_eff; /* suppress: run for the model side effects only */
	Cursor c;
	c.gw0 = ((8u * (a + P.hdr_len)) >> 7) << 2;
	c.sbits = (8u * (a + P.hdr_len)) & 127u;
	c.buf = 0;

	if (P.pre == CMP_PREPROCESS_IWT)
		iwt_global(P, &sh.stg_mem[0][0]);

	const uint32_t n = P.n, pre = P.pre, model_mode = P.model_mode;
	const bool fast_ok = ((uintptr_t)P.src & 15u) == 0 &&
			     (((uintptr_t)P.work & 15u) == 0 || (pre < CMP_PREPROCESS_IWT && !model_mode));
	const uint32_t n_pieces = fast_ok ? n / 8u : 0u;
	if (n_pieces) {
		/* pair table of this pass's encoder, kept across frames and jobs while the encoder stays */
		const EncConst &e = P.enc;
		if (e.type != CMP_ENCODER_UNCOMPRESSED && n >= kLutMinSamples &&
		    (sh.plut_key[0] != e.type || sh.plut_key[1] != e.g || sh.plut_key[2] != e.outlier))
			build_pair_lut(sh, e);
		if (tid == 0 && pre == CMP_PREPROCESS_DIFF)
			first_sample_code(e, sample_at(P.src, P.dtype, 0), sh.first_code);
		frame_fast_any(sh, o, a, c, n_pieces, size_only);
	}
	if (n_pieces * 8u < n)
		c = generic_span(sh, o, a, c, n_pieces * 8u, n, size_only);
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

	/* (the 4 checksum bytes behind the stream are filled in by airs_checksum_kernel) */
	if (tid == 32) { /* last partial group, zero padded (ref bitstream_writer.h:205-227) */
		const uint32_t nb = (c.sbits + 7u) >> 3;
		uint32_t *stg = stg_of(sh, c.buf);
		for (uint32_t k = 0; k < nb; k++) {
			const uint64_t b = (uint64_t)c.gw0 * 4 + k;
			if (b >= o.lo && b < o.hi)
				o.base[b] = (uint8_t)(stg[k >> 2] >> (24 - 8 * (k & 3)));
		}
		stg[0] = stg[1] = stg[2] = stg[3] = 0;
	}
	__syncthreads();

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

/* two-phase CONCAT: is this launch the phase that runs? (see airs_launch.h) */
__device__ __forceinline__ bool gate_closed(const AirsLaunch &b)
{
	return b.ticket[AIRS_TICKET_INVALID] != 0u || (b.gate && (*b.gate != 0u) != (b.gate_want != 0u));
}

/* Job descriptors come in through shared memory: the 32 jobs of a warp are 3840 consecutive bytes, fetched with
 * coalesced 16-byte loads; a thread that reads its own 120-byte descriptor field by field from global memory
 * waits for one dependent load after the other.  Plans and fast-job records stay in registers and leave with
 * 16-byte stores. */
struct PlanWarp {
	alignas(16) uint32_t job[32 * sizeof(airs_job) / 4];
};

__global__ void __launch_bounds__(128) airs_plan_kernel(AirsLaunch b)
{
	__shared__ PlanWarp psh[4];
	PlanWarp &pw = psh[threadIdx.x >> 5];
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31u, j0 = j - lane;

	if (gate_closed(b))
		return;
	const uint32_t in_warp = j0 < b.n_jobs ? min(32u, b.n_jobs - j0) : 0u; /* jobs of this warp */
	{
		constexpr uint32_t kVec = sizeof(airs_job) / 4u; /* a descriptor is 30 words; 32 of them are 240 x 16 bytes */
		const uint32_t *src = reinterpret_cast<const uint32_t *>(b.jobs + j0);
		if (((uintptr_t)b.jobs & 15u) == 0 && in_warp == 32u) {
			for (uint32_t v = lane; v < 32u * kVec / 4u; v += 32u)
				reinterpret_cast<uint4 *>(pw.job)[v] = __ldg(reinterpret_cast<const uint4 *>(src) + v);
		} else {
			for (uint32_t w = lane; w < in_warp * kVec; w += 32u)
				pw.job[w] = __ldg(src + w);
		}
		__syncwarp();
	}
	const bool have = j < b.n_jobs;
	const airs_job &job = *reinterpret_cast<const airs_job *>(pw.job + (have ? lane : 0u) * (sizeof(airs_job) / 4u));
	/* The contract of include/airs_cuda.h: every frame has a result index below n_results; in the CONCAT layout the
	 * jobs list the frames 0 .. n_results - 1 in order (the look-back scan waits for every index in front of a frame:
	 * an index nobody covers would make it wait forever).  A batch that breaks it is not run: the kernels behind this
	 * one return at once and the checksum kernel turns every result into CMP_ERR_GENERIC. */
	if (have) {
		bool bad = (uint64_t)job.first_result + job.n_frames > b.n_results;
		if (b.ordered) {
			uint32_t want = 0;
			if (j != 0u) {
				const airs_job &prev = lane ? *reinterpret_cast<const airs_job *>(pw.job + (lane - 1u) * (sizeof(airs_job) / 4u)) : b.jobs[j - 1u];
				want = prev.first_result + prev.n_frames;
			}
			bad = bad || job.first_result != want || (j + 1u == b.n_jobs && job.first_result + job.n_frames != b.n_results);
		}
		if (bad)
			atomicExch(&b.ticket[AIRS_TICKET_INVALID], 1u);
	}
	/* the job of every frame (airs_checksum_kernel, the CONCAT copy): short jobs by their own
	 * thread, long ones by the whole warp */
	{
		const uint32_t first = have ? job.first_result : 0u;
		const uint32_t nf = have ? job.n_frames : 0u;
		const bool wide = nf > 16u;
		if (!wide)
			for (uint32_t f = 0; f < nf && (uint64_t)first + f < b.n_results; f++)
				b.result_job[first + f] = j;
		for (uint32_t todo = __ballot_sync(kFull, wide); todo; todo &= todo - 1u) {
			const int l = __ffs((int)todo) - 1;
			const uint32_t fl = __shfl_sync(kFull, first, l), nl = __shfl_sync(kFull, nf, l);
			const uint32_t jl = __shfl_sync(kFull, j, l);
			for (uint32_t f = lane; f < nl && (uint64_t)fl + f < b.n_results; f += 32u)
				b.result_job[fl + f] = jl;
		}
	}
	if (in_warp == 0u)
		return;
	JobPlan pl;
	airs_make_plan(pl, job, b.src, b.work, b.be_batch != 0u);
	/* Short single-frame jobs without model, with a Golomb encoder, none / diff preprocessing,
	 * an aligned 16-bit source and nothing that could fail before the encoding go to
	 * airs_fast_kernel (one warp per job); everything else to airs_encode_kernel.  Jobs keep their
	 * order inside a warp of this kernel, warps take their places in the lists as they come, except
	 * in the CONCAT layout and on the host-shim path, where every job is "big" and the list is the
	 * identity (the look-back scan needs the frames to start in result order). */
	const bool listed = b.layout == AIRS_LAYOUT_SLOTS && !b.ctx_io;
	/* what the fast kernels ask of a job */
	/* Single frames under the wavelet transform: airs_iwt_kernel (all CTAs over the tiles of all such frames) leaves the
	 * coefficients in the job's work buffer, exactly where the reference keeps them (preprocess.c:337-353), and the warp
	 * encoders code them like samples without preprocessing - the coarse coefficients are far outside any code word
	 * table, which the arithmetic code words do not mind.  Not with the uncompressed fallback (the fast kernel would
	 * store the coefficients raw) and not beyond the frame length whose last levels fit one CTA. */
	const bool iwt = have && pl.pre[0] == CMP_PREPROCESS_IWT && b.work && job.n_frames == 1 && !(pl.flags & (AIRS_PF_FALLBACK_OK | AIRS_PF_BE)) &&
			 pl.n <= AIRS_IWT_MAX_SAMPLES && ((uintptr_t)(b.work + job.work_offset) & 15u) == 0;
	const bool common = have && listed && (pl.flags & AIRS_PF_VALID) && !pl.frame_err && !pl.orig_err && !pl.pre_err[0] &&
			    (pl.pre[0] == CMP_PREPROCESS_NONE || pl.pre[0] == CMP_PREPROCESS_DIFF || iwt) &&
			    pl.enc[0].type != CMP_ENCODER_UNCOMPRESSED && job.dtype != AIRS_DTYPE_I16_IN_I32 && b.dst &&
			    ((uintptr_t)(b.dst + job.dst_offset) & 7u) == 0 && ((uintptr_t)(b.src + job.src_offset) & 15u) == 0 &&
			    pl.cap_eff >= (CMP_HDR_SIZE + 6u) && pl.enc[0].g <= AIRS_FAST_MAX_G;
	const bool quick = common && !(pl.flags & AIRS_PF_MODEL) && job.n_frames == 1;
	const bool small = quick && pl.n <= kSmallMaxSamples; /* one warp per job: airs_fast_kernel */
	/* Contexts of several frames (with or without model) in a batch of few jobs: their frames go through
	 * airs_tile_kernel tile by tile, frame by frame (a CTA per context would leave most of the device idle).
	 * Only contexts in which no frame can fail or fall back - slots of cmp_compress_bound() bytes (48 bits a
	 * sample, ref cmp.c:59-74), no uncompressed fallback: a frame that fails changes the passes of all frames
	 * behind it (ref cmp.c:228-262), whose tiles are already on their way */
	const bool secondary_ok = pl.sec_iter == 0u ||
				  (!pl.pre_err[1] && pl.enc[1].type != CMP_ENCODER_UNCOMPRESSED && pl.enc[1].g <= AIRS_FAST_MAX_G &&
				   (pl.pre[1] == CMP_PREPROCESS_NONE || pl.pre[1] == CMP_PREPROCESS_DIFF || pl.pre[1] == CMP_PREPROCESS_MODEL));
	const bool never_fails = !(pl.flags & AIRS_PF_FALLBACK_OK) && !job.params.uncompressed_fallback_enabled &&
				 (uint64_t)job.dst_capacity >= CMP_HDR_SIZE + 6u + 4u + 6ull * pl.n;
	const bool frames_tiled = common && !iwt && job.n_frames > 1u && b.n_jobs < b.tile_below_jobs && secondary_ok && !pl.model_err && never_fails &&
				  pl.n >= 2u * AIRS_TILE_SAMPLES && (pl.n & 7u) == 0 && (job.src_frame_stride & 15u) == 0 &&
				  (job.dst_frame_stride & 7u) == 0 &&
				  (!(pl.flags & AIRS_PF_MODEL) || (b.work && ((uintptr_t)(b.work + job.work_offset) & 15u) == 0));
	/* Tiles over all warps (airs_tile_kernel, arithmetic code words) for long frames when the batch has too few jobs to
	 * give every resident CTA of airs_encode_kernel one, or when that kernel's code word table cannot cover the
	 * residuals anyway (no preprocessing: the samples themselves are coded; tiny g: long code words) */
	const bool tiled = frames_tiled ||
			   (quick && !small && (b.n_jobs < b.tile_below_jobs || pl.pre[0] == CMP_PREPROCESS_NONE || pl.enc[0].g < 4u || iwt));
	const bool iwt_job = iwt && (small || tiled);
	const uint32_t my_iwt_tiles = iwt_job ? (pl.n + AIRS_IWT_TILE - 1u) / AIRS_IWT_TILE : 0u;
	/* Single frames under the UNCOMPRESSED encoder whose slot holds the stream: a copy with a byte swap, airs_raw_kernel */
	const bool raw = have && listed && (pl.flags & AIRS_PF_VALID) && !pl.frame_err && !pl.orig_err && !pl.pre_err[0] &&
			 !(pl.flags & AIRS_PF_MODEL) && job.n_frames == 1 && pl.enc[0].type == CMP_ENCODER_UNCOMPRESSED &&
			 (pl.pre[0] == CMP_PREPROCESS_NONE || pl.pre[0] == CMP_PREPROCESS_DIFF) &&
			 (job.dtype == AIRS_DTYPE_I16 || job.dtype == AIRS_DTYPE_U16) && b.dst &&
			 ((uintptr_t)(b.dst + job.dst_offset) & 7u) == 0 && ((uintptr_t)(b.src + job.src_offset) & 15u) == 0 &&
			 (uint64_t)pl.cap_eff >= (pl.pre[0] == CMP_PREPROCESS_DIFF ? CMP_HDR_SIZE + 6u : CMP_HDR_SIZE) + 2ull * pl.n +
							 ((pl.flags & AIRS_PF_CHECKSUM) ? 4u : 0u);
	const uint32_t frame_tiles = (pl.n + AIRS_TILE_SAMPLES - 1u) / AIRS_TILE_SAMPLES;
	const uint32_t my_tiles = tiled ? frame_tiles * job.n_frames : 0u;
	if (small)
		pl.flags |= AIRS_PF_SMALL;
	if (have && b.init_results && !b.ctx_io)
		b.init_results[j] = pl.init_result;
	/* one atomic per warp and counter: the lanes' places follow from the ballots */
	const uint32_t below = (1u << lane) - 1u;
	const uint32_t m_cs = __ballot_sync(kFull, have && (pl.flags & AIRS_PF_CHECKSUM));
	const uint32_t m_small = __ballot_sync(kFull, small);
	const uint32_t m_tiled = __ballot_sync(kFull, tiled);
	const uint32_t m_raw = __ballot_sync(kFull, raw);
	const uint32_t m_big = __ballot_sync(kFull, have && listed && !small && !tiled && !raw);
	uint32_t tiles_incl = my_tiles; /* tiles of the lanes up to this one */
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(kFull, tiles_incl, d);
		if (lane >= (uint32_t)d)
			tiles_incl += t;
	}
	const uint32_t warp_tiles = __shfl_sync(kFull, tiles_incl, 31);
	uint32_t base_small = 0, base_big = 0, base_tslot = 0, base_tile = 0, base_raw = 0, base_raw_chunk = 0;
	/* chunks of the raw jobs of the lanes up to this one */
	const uint32_t raw_hdr = pl.pre[0] == CMP_PREPROCESS_DIFF ? CMP_HDR_SIZE + 6u : CMP_HDR_SIZE;
	const uint32_t my_raw_chunks = raw ? ((raw_hdr + 2u * pl.n) / 8u + AIRS_RAW_CHUNK - 1u) / AIRS_RAW_CHUNK : 0u;
	uint32_t raw_incl = raw ? max(my_raw_chunks, 1u) : 0u;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(kFull, raw_incl, d);
		if (lane >= (uint32_t)d)
			raw_incl += t;
	}
	const uint32_t warp_raw_chunks = __shfl_sync(kFull, raw_incl, 31);
	const uint32_t m_iwt = __ballot_sync(kFull, iwt_job);
	uint32_t iwt_incl = my_iwt_tiles, base_iwt = 0, base_iwt_tile = 0;
	if (m_iwt) {
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(kFull, iwt_incl, d);
			if (lane >= (uint32_t)d)
				iwt_incl += t;
		}
		if (lane == 31) { /* records and tiles from ONE counter: the order of the records is the order of their tiles */
			const unsigned long long got = atomicAdd(reinterpret_cast<unsigned long long *>(b.ticket + AIRS_TICKET_IWT),
								 ((unsigned long long)__popc(m_iwt) << 40) | iwt_incl);
			base_iwt = (uint32_t)(got >> 40);
			base_iwt_tile = (uint32_t)(got & ((1ull << 40) - 1u));
		}
		base_iwt = __shfl_sync(kFull, base_iwt, 31);
		base_iwt_tile = __shfl_sync(kFull, base_iwt_tile, 31);
	}
	if (m_tiled) { /* do all tile jobs have the same shape?  (airs_tile_kernel deals their tiles frame by frame then) */
		const uint32_t fr = tiled ? job.n_frames : 0u;
		const uint32_t tmin = __reduce_min_sync(kFull, tiled ? frame_tiles : 0xFFFFFFFFu), tmax = __reduce_max_sync(kFull, tiled ? frame_tiles : 0u);
		const uint32_t fmin = __reduce_min_sync(kFull, tiled ? fr : 0xFFFFFFFFu), fmax = __reduce_max_sync(kFull, fr);
		if (lane == 0) {
			atomicMax(&b.ticket[AIRS_TICKET_TILE_SHAPE], ~tmin);
			atomicMax(&b.ticket[AIRS_TICKET_TILE_SHAPE + 1], tmax);
			atomicMax(&b.ticket[AIRS_TICKET_TILE_SHAPE + 2], ~fmin);
			atomicMax(&b.ticket[AIRS_TICKET_TILE_SHAPE + 3], fmax);
		}
	}
	if (lane == 0) {
		if (m_cs)
			atomicAdd(&b.ticket[4], (uint32_t)__popc(m_cs));
		if (m_small)
			base_small = atomicAdd(&b.ticket[3], (uint32_t)__popc(m_small));
		if (m_big)
			base_big = atomicAdd(&b.ticket[2], (uint32_t)__popc(m_big));
		if (m_raw) { /* slots and chunks from ONE counter: the order of the records is the order of their chunks */
			const unsigned long long got = atomicAdd(reinterpret_cast<unsigned long long *>(b.ticket + AIRS_TICKET_RAW),
								 ((unsigned long long)__popc(m_raw) << 40) | warp_raw_chunks);
			base_raw = (uint32_t)(got >> 40);
			base_raw_chunk = (uint32_t)(got & ((1ull << 40) - 1u));
		}
		if (m_tiled) { /* slots and tiles from ONE counter: the order of the slots is the order of their tiles */
			const unsigned long long got = atomicAdd(reinterpret_cast<unsigned long long *>(b.ticket + 10),
								 ((unsigned long long)__popc(m_tiled) << 40) | warp_tiles);
			base_tslot = (uint32_t)(got >> 40);
			base_tile = (uint32_t)(got & ((1ull << 40) - 1u));
		}
	}
	base_small = __shfl_sync(kFull, base_small, 0);
	base_big = __shfl_sync(kFull, base_big, 0);
	base_tslot = __shfl_sync(kFull, base_tslot, 0);
	base_tile = __shfl_sync(kFull, base_tile, 0);
	base_raw = __shfl_sync(kFull, base_raw, 0);
	base_raw_chunk = __shfl_sync(kFull, base_raw_chunk, 0);
	if (have) {
		FastJob *recs = reinterpret_cast<FastJob *>(b.fast_jobs);
		if (!listed) {
			b.big_list[j] = j;
			if (j == 0)
				b.ticket[2] = b.n_jobs;
		} else if (small) {
			const uint32_t slot = base_small + (uint32_t)__popc(m_small & below);
			b.small_list[slot] = j;
			FastJob fj;
			airs_fill_fast_job(fj, job, pl, b.src, b.dst, j, 0u, 0u, iwt ? b.work + job.work_offset : nullptr);
			recs[slot] = fj;
		} else if (raw) { /* (FastJob records in the front of the tile extensions) */
			FastJob fj;
			const uint32_t chunks = max(my_raw_chunks, 1u);
			airs_fill_fast_job(fj, job, pl, b.src, b.dst, j, base_raw_chunk + raw_incl - chunks, chunks);
			reinterpret_cast<FastJob *>(b.tile_ext)[base_raw + (uint32_t)__popc(m_raw & below)] = fj;
		} else if (tiled) { /* the records of the long jobs fill the array from its end */
			const uint32_t slot = base_tslot + (uint32_t)__popc(m_tiled & below);
			FastJob fj;
			airs_fill_fast_job(fj, job, pl, b.src, b.dst, j, base_tile + tiles_incl - my_tiles, my_tiles,
					   iwt ? b.work + job.work_offset : nullptr);
			recs[b.n_jobs - 1u - slot] = fj;
			TileExt tx;
			airs_fill_tile_ext(tx, job, pl, b.work, frame_tiles);
			reinterpret_cast<TileExt *>(b.tile_ext)[b.n_jobs - 1u - slot] = tx;
		} else {
			b.big_list[base_big + (uint32_t)__popc(m_big & below)] = j;
		}
		if (iwt_job) {
			IwtRec ir;
			ir.src = (uint64_t)(uintptr_t)(b.src + job.src_offset);
			ir.work = (uint64_t)(uintptr_t)(b.work + job.work_offset);
			ir.n = pl.n;
			ir.tile_base = base_iwt_tile + iwt_incl - my_iwt_tiles;
			ir.n_tiles = my_iwt_tiles;
			ir.job = j;
			reinterpret_cast<IwtRec *>(b.iwt_recs)[base_iwt + (uint32_t)__popc(m_iwt & below)] = ir;
		}
		/* the plan of a short job is read again only by the checksum kernels and by airs_encode_kernel when the job is
		 * handed back to it: a million 4 KiB chunks without either save a third of this kernel's memory traffic */
		if (!small || (pl.flags & (AIRS_PF_CHECKSUM | AIRS_PF_FALLBACK_OK)))
			b.plans[j] = pl;
		else
			b.plans[j].flags = pl.flags; /* (the checksum kernels look at the flags of every frame's job: no stale ones) */
	}
}

__global__ void __launch_bounds__(AIRS_THREADS, AIRS_CTAS_PER_SM) airs_encode_kernel(AirsLaunch b)
{
	__shared__ Shared sh;
	const uint32_t tid = threadIdx.x;

	if (gate_closed(b) || b.ticket[2] == 0u) /* not this phase's launch, or no job for this kernel */
		return;

	for (uint32_t w = tid; w < 2u * (4u + kStgWords); w += kThreads)
		(&sh.stg_mem[0][0])[w] = 0;
	if (tid == 0) {
		sh.plut_key[0] = 0xFFFFFFFFu; /* no table yet */
		sh.plut_R = 0;
		sh.ticket = atomicAdd(b.ticket, 1u);
	}

	for (;;) {
		__syncthreads();
		if (sh.ticket >= b.ticket[2]) /* entries of big_list, written by airs_plan_kernel */
			break;
		const uint32_t job = b.big_list[sh.ticket];
		/* plan (32 words) and job descriptor (30 words) into shared memory, coalesced */
		if (tid < 32)
			((uint32_t *)&sh.plan)[tid] = ((const uint32_t *)&b.plans[job])[tid];
		else if (tid < 62)
			((uint32_t *)&sh.job)[tid - 32] = ((const uint32_t *)&b.jobs[job])[tid - 32];
		__syncthreads();
		if (tid == 0) {
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

		const bool concat = b.layout == AIRS_LAYOUT_CONCAT;
		if (tid == 0 && n_frames == 0)
			sh.ticket = atomicAdd(b.ticket, 1u);
		/* runs of secondary MODEL passes: all their frames tile position by tile position, the model in registers */
		const bool runs_ok = model_run_ok(sh, b);
		bool runs_off = false; /* a run gave up: the context is redone frame by frame */
		for (uint32_t f = 0; f < n_frames; f++) {
			if (tid == 0) {
				/* the next job is drawn while the last frame of this one is encoded: early
				 * enough to hide the atomic, late enough that a CTA does not sit on a second
				 * long job while other CTAs run dry */
				if (f + 1 == n_frames)
					sh.ticket = atomicAdd(b.ticket, 1u);
				plan_frame(sh, b, f);
			}
			__syncthreads();
			/* One call site for encode_pass (it is inlined).  SLOTS: stage 0 encodes, stage 1 is
			 * the raw retry of the fallback (ref cmp.c:380-392).  CONCAT: stage 0 sizes the
			 * stream (exact, no output), the scan turns sizes into offsets, stage 1 writes;
			 * a frame that fails contributes no bytes. */
			uint32_t r = 0, stage = 0;
			bool fits = true;
			for (;;) {
				const uint32_t rp = encode_pass(sh, concat && stage == 0, concat && stage == 1 && !fits);
				const bool fall_back = stage == 0 && (sh.plan.flags & AIRS_PF_FALLBACK_OK) &&
						       rp == AIRS_ERR(DST_TOO_SMALL);
				if (fall_back) { /* the frame is stored raw as a fresh primary pass */
					__syncthreads();
					if (tid == 0) {
						ctx_reset(sh.ctx);
						plan_pass(sh, true, !concat);
					}
					__syncthreads();
				}
				if (!concat) {
					r = rp;
					if (!fall_back)
						break;
					stage = 1;
					continue;
				}
				if (stage == 1) {
					if (!airs_failed(r))
						r = fits ? rp : AIRS_ERR(DST_TOO_SMALL);
					break;
				}
				r = fall_back ? (sh.pass.err ? sh.pass.err : sh.plan.raw_size) : rp;
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
				fits = !airs_failed(r) && sh.offset + r <= b.dst_size;
				if (sh.pass.err)
					break;
				stage = 1;
			}
			__syncthreads();
			if (tid == 0) {
				if (!airs_failed(r))
					sh.ctx.seq = (sh.ctx.seq + 1u) & 0xFFu;
				b.results[first + f] = r;
			}
			if (runs_ok && !runs_off && !airs_failed(r) && f + 1u < n_frames) {
				__syncthreads();
				const uint32_t seq = sh.ctx.seq;
				const uint32_t run = (seq >= 1u && seq <= sh.plan.sec_iter) ? min(min(n_frames - (f + 1u), sh.plan.sec_iter - seq + 1u), kCtxFrames) : 0u;
				if (run >= 2u) {
					const bool done = (sh.plan.flags & AIRS_PF_SIGNED) ? model_run_fast<true>(sh, b, f + 1u, run, seq)
											    : model_run_fast<false>(sh, b, f + 1u, run, seq);
					__syncthreads();
					if (done) {
						f += run;
						if (tid == 0) {
							sh.ctx.seq = (seq + run) & 0xFFu;
							if (f + 1u == n_frames) /* (the run ended the job: its last frame did not draw the next one) */
								sh.ticket = atomicAdd(b.ticket, 1u);
						}
					} else { /* from the first frame again: its primary pass sets the model anew */
						runs_off = true;
						if (tid == 0) {
							sh.ctx.counter = sh.job.identifier_base;
							ctx_reset(sh.ctx);
						}
						f = 0xFFFFFFFFu;
					}
				}
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

/* The XXH32 trailer of every successfully encoded stream whose job asked for a checksum (ref
 * cmp.c:314-319: zero padded to a byte, then 4 bytes big endian).  Runs behind the encode
 * kernels; a batch without checksums costs one launch of early exits.  One thread per frame: for
 * batches of many frames (few, long frames: airs_checksum_warp_kernel below). */
__global__ void __launch_bounds__(128) airs_checksum_kernel(AirsLaunch b)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;

	if (b.ticket[AIRS_TICKET_INVALID]) { /* a job table that breaks the contract: nothing has run */
		if (k < b.n_results)
			b.results[k] = AIRS_ERR(GENERIC);
		return;
	}
	if (b.ticket[4] == 0 || gate_closed(b))
		return;
	bool todo = false;
	const uint8_t *src = nullptr;
	uint8_t *stream = nullptr;
	uint32_t r = 0, n = 0, dtype = AIRS_DTYPE_U16;
	if (k < b.n_results) {
		const uint32_t j = b.result_job[k];
		if (j < b.n_jobs) {
			const JobPlan &pl = b.plans[j];
			r = b.results[k];
			if ((pl.flags & AIRS_PF_CHECKSUM) && !airs_failed(r) && r >= 4u) {
				const airs_job &job = b.jobs[j];
				const uint32_t f = k - job.first_result;
				src = b.src + job.src_offset + (uint64_t)f * job.src_frame_stride;
				stream = b.layout == AIRS_LAYOUT_CONCAT ? b.dst + b.out_offsets[k]
									 : b.dst + job.dst_offset + (uint64_t)f * job.dst_frame_stride;
				n = pl.n;
				dtype = job.dtype;
				todo = true;
			}
		}
	}
	if (!todo)
		return;
	const uint32_t h = stream_checksum(src, dtype, n);
	{
		stream[r - 4u] = (uint8_t)(h >> 24);
		stream[r - 3u] = (uint8_t)(h >> 16);
		stream[r - 2u] = (uint8_t)(h >> 8);
		stream[r - 1u] = (uint8_t)h;
	}
}

/* -------------------------------------------------------------------------
 * Few, long frames: FOUR LANES PER FRAME, eight frames per warp.  The XXH32 of a stream is one serial chain per
 * accumulator (a round is IMAD, SHF, IMAD: ~13 cycles a 16-byte stripe, 0.9 ms for a 2 MiB frame), so what a
 * frame needs is its bytes waiting in shared memory when the chain asks for them: the four lanes of a group keep
 * kCsRing - 1 blocks of kCsBlock bytes in flight with 16-byte asynchronous copies (cp.async, one commit group per
 * block) and run one accumulator each over the block that has landed (conflict-free inside a group: its lanes
 * read the four words of a stripe).  Nothing in the loop waits for global memory once the ring is full.  A CTA is
 * one warp; frames are dealt to CTAs first and to the groups of a CTA second, so that a batch of few frames
 * spreads over all SMs.  (Bulk copies - cp.async.bulk, one per block and mbarrier - were measured first: 1 KiB
 * copies ran at 1.4 TB/s over the device, a fixed cost of ~200 cycles per copy and SM.)
 * ---------------------------------------------------------------------- */
constexpr uint32_t kCsBlock = 2048; /* bytes per ring slot: 128 stripes = 0.9 us of chain */
constexpr uint32_t kCsRing = 4;
constexpr uint32_t kCsGroups = 8;   /* frames a warp works on at the same time */
constexpr uint32_t kCsSmem = kCsGroups * kCsRing * kCsBlock;

/* this lane's quarter of a block: 16-byte pieces acc, acc + 4, ... (a group's copy instruction covers 64 contiguous bytes) */
__device__ __forceinline__ void cs_request(uint32_t smem_slot, const uint8_t *src, uint32_t bytes, uint32_t acc)
{
	for (uint32_t o = 16u * acc; o < bytes; o += 64u)
		asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_slot + o), "l"(src + o) : "memory");
}

__global__ void __launch_bounds__(32) airs_checksum_warp_kernel(AirsLaunch b)
{
	extern __shared__ __align__(128) uint8_t cs_ring[]; /* [group][slot][kCsBlock] */
	const uint32_t lane = threadIdx.x, grp = lane >> 2, acc = lane & 3u;

	if (b.ticket[AIRS_TICKET_INVALID]) { /* a job table that breaks the contract: nothing has run */
		for (uint32_t k = blockIdx.x * 32u + lane; k < b.n_results; k += gridDim.x * 32u)
			b.results[k] = AIRS_ERR(GENERIC);
		return;
	}
	if (b.ticket[4] == 0 || gate_closed(b))
		return;
	const uint8_t *ring = cs_ring + grp * kCsRing * kCsBlock;
	const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(ring);
	const uint32_t seed = AIRS_CHECKSUM_SEED;

	/* the loops are the same for all groups of the warp (a group without a frame, or with a shorter one, idles):
	 * every barrier in them is a plain warp barrier */
	for (uint32_t base = blockIdx.x; base < b.n_results; base += kCsGroups * gridDim.x) {
		const uint32_t k = base + grp * gridDim.x;
		bool todo = false;
		const uint8_t *src = nullptr;
		uint8_t *stream = nullptr;
		uint32_t r = 0, n = 0, dtype = AIRS_DTYPE_U16;
		if (k < b.n_results) {
			const uint32_t j = b.result_job[k];
			if (j < b.n_jobs) {
				const JobPlan &pl = b.plans[j];
				r = b.results[k];
				if ((pl.flags & AIRS_PF_CHECKSUM) && !airs_failed(r) && r >= 4u) {
					const airs_job &job = b.jobs[j];
					const uint32_t f = k - job.first_result;
					src = b.src + job.src_offset + (uint64_t)f * job.src_frame_stride;
					stream = b.layout == AIRS_LAYOUT_CONCAT ? b.dst + b.out_offsets[k]
										 : b.dst + job.dst_offset + (uint64_t)f * job.dst_frame_stride;
					n = pl.n;
					dtype = job.dtype;
					todo = true;
				}
			}
		}
		const bool staged = todo && dtype != AIRS_DTYPE_I16_IN_I32 && ((uintptr_t)src & 15u) == 0;
		const bool be = (dtype & AIRS_DTYPE_BE) != 0u; /* the samples are the big-endian image already */
		const uint32_t nbytes = 2u * n, full = staged ? nbytes & ~15u : 0u; /* whole stripes */
		const uint32_t n_blocks = (full + kCsBlock - 1u) / kCsBlock;
		const uint32_t max_blocks = __reduce_max_sync(kFull, n_blocks);
		uint32_t v = acc == 0 ? seed + AIRS_XP1 + AIRS_XP2 : acc == 1 ? seed + AIRS_XP2 : acc == 2 ? seed : seed - AIRS_XP1;

		/* one commit group per block, empty ones where a group has nothing to fetch: block blk has landed when at
		 * most kCsRing - 1 groups were committed behind it */
		for (uint32_t q = 0; q + 1u < kCsRing; q++) {
			if (q < n_blocks)
				cs_request(ring0 + q * kCsBlock, src + (size_t)q * kCsBlock, min(kCsBlock, full - q * kCsBlock), acc);
			asm volatile("cp.async.commit_group;" ::: "memory");
		}
		for (uint32_t blk = 0; blk < max_blocks; blk++) {
			{ /* the slot read one iteration ago takes the block kCsRing - 1 ahead */
				const uint32_t nb = blk + kCsRing - 1u;
				if (nb < n_blocks)
					cs_request(ring0 + (nb % kCsRing) * kCsBlock, src + (size_t)nb * kCsBlock, min(kCsBlock, full - nb * kCsBlock), acc);
				asm volatile("cp.async.commit_group;" ::: "memory");
			}
			asm volatile("cp.async.wait_group %0;" ::"n"(kCsRing - 1u) : "memory");
			__syncwarp(); /* the quarters of the other lanes of the group are there as well */
			if (blk < n_blocks) {
				const uint32_t nst = min(kCsBlock, full - blk * kCsBlock) / 16u;
				const uint32_t *w = reinterpret_cast<const uint32_t *>(ring + (blk % kCsRing) * kCsBlock) + acc;
				uint32_t st = 0;
				for (; st + 16u <= nst; st += 16u) { /* sixteen words at once: their latency is paid once per sixteen rounds */
					uint32_t q[16];
#pragma unroll
					for (int i = 0; i < 16; i++)
						q[i] = w[4u * (st + i)];
#pragma unroll
					for (int i = 0; i < 16; i++)
						v = airs_xxh_round(v, be ? q[i] : airs_be_pair(q[i]));
				}
				for (; st < nst; st++)
					v = airs_xxh_round(v, be ? w[4u * st] : airs_be_pair(w[4u * st]));
			}
			__syncwarp(); /* the slot has been read: it may be filled again */
		}
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		const uint32_t l0 = 4u * grp;
		const uint32_t v0 = __shfl_sync(kFull, v, l0), v1 = __shfl_sync(kFull, v, l0 + 1u);
		const uint32_t v2 = __shfl_sync(kFull, v, l0 + 2u), v3 = __shfl_sync(kFull, v, l0 + 3u);
		if (!todo || acc != 0)
			continue;
		uint32_t h;
		if (!staged) { /* no 16-byte copies from this source: plain loads */
			h = stream_checksum(src, dtype, n);
		} else {
			h = nbytes >= 16u ? airs_rotl(v0, 1) + airs_rotl(v1, 7) + airs_rotl(v2, 12) + airs_rotl(v3, 18) : seed + AIRS_XP5;
			h += nbytes;
			uint32_t i = full / 2u; /* what is left of the frame behind its last whole stripe */
			for (; i + 2 <= n; i += 2)
				h = airs_rotl(h + airs_be_pair(sample_pair_at(src, dtype, true, i)) * AIRS_XP3, 17) * AIRS_XP4;
			if (i < n) {
				const uint32_t sv = sample_at(src, dtype, i);
				h = airs_rotl(h + (sv >> 8) * AIRS_XP5, 11) * AIRS_XP1;
				h = airs_rotl(h + (sv & 0xFFu) * AIRS_XP5, 11) * AIRS_XP1;
			}
			h ^= h >> 15;
			h *= AIRS_XP2;
			h ^= h >> 13;
			h *= AIRS_XP3;
			h ^= h >> 16;
		}
		stream[r - 4u] = (uint8_t)(h >> 24);
		stream[r - 3u] = (uint8_t)(h >> 16);
		stream[r - 2u] = (uint8_t)(h >> 8);
		stream[r - 1u] = (uint8_t)h;
	}
}

extern "C" cudaError_t airs_launch_plan(const AirsLaunch *b, cudaStream_t stream)
{
	airs_plan_kernel<<<(b->n_jobs + 127) / 128, 128, 0, stream>>>(*b);
	return cudaGetLastError();
}

/* resident CTAs per SM of the encode kernel on the current device */
extern "C" cudaError_t airs_encode_ctas_per_sm(int *out)
{
	/* shared memory for AIRS_CTAS_PER_SM CTAs (1 KiB per CTA is reserved by the system), the
	 * rest of the 228 KiB stays L1 */
	cudaFuncAttributes fa;
	cudaError_t e = cudaFuncGetAttributes(&fa, airs_encode_kernel);
	if (e != cudaSuccess)
		return e;
	const size_t need = (size_t)AIRS_CTAS_PER_SM * (fa.sharedSizeBytes + 1024);
	int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
	e = cudaFuncSetAttribute(airs_encode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
	if (e != cudaSuccess)
		return e;
	return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, airs_encode_kernel, AIRS_THREADS, 0);
}

extern "C" cudaError_t airs_launch_encode(const AirsLaunch *b, unsigned int grid, cudaStream_t stream)
{
	airs_encode_kernel<<<grid, AIRS_THREADS, 0, stream>>>(*b);
	return cudaGetLastError();
}

extern "C" cudaError_t airs_launch_checksum(const AirsLaunch *b, cudaStream_t stream)
{
	if (b->n_results == 0)
		return cudaSuccess;
	if (b->n_results >= 32768u) { /* enough frames to fill the device with one thread each */
		airs_checksum_kernel<<<(b->n_results + 127) / 128, 128, 0, stream>>>(*b);
	} else { /* four lanes per frame, at most the warps the device holds at once (3 one-warp CTAs of 64 KiB per SM) */
		static thread_local int dev_cached = -1, max_grid = 0;
		int dev;
		cudaError_t e = cudaGetDevice(&dev);
		if (e != cudaSuccess)
			return e;
		if (dev != dev_cached) {
			int sms = 0;
			if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
				return e;
			if ((e = cudaFuncSetAttribute(airs_checksum_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCsSmem)) != cudaSuccess)
				return e;
			max_grid = 3 * sms;
			dev_cached = dev;
		}
		airs_checksum_warp_kernel<<<b->n_results < (unsigned int)max_grid ? b->n_results : (unsigned int)max_grid, 32, kCsSmem, stream>>>(*b);
	}
	return cudaGetLastError();
}
