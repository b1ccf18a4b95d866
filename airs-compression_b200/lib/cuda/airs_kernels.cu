/*
 * airs_kernels.cu - the sm_100a compression kernel.
 *
 * One CTA owns one job (= one compression context, lib/cmp.h:129-137) at a
 * time, taken from a device-side ticket counter, and pushes the job's frames
 * through in order; the context state machine of cmp_compress_generic /
 * compress_engine (ref lib/compress/cmp.c:213-393) runs on thread 0, the
 * per-sample work of the hot loop (cmp.c:296-312) on all 256 threads:
 *
 *   128-bit loads of 8 samples per thread -> residual (none / diff / IWT
 *   coefficient / model) -> zig-zag -> Golomb / escape codeword and length, all
 *   in registers -> warp-shuffle + block exclusive scan over bit lengths ->
 *   each thread shifts its codewords into place with funnel shifts and writes
 *   32-bit words of the MSB-first stream into shared memory (only the first
 *   and last word of a thread can be shared with a neighbour: shared-memory
 *   atomicOr) -> the staged words leave as coalesced, byte-swapped 32-bit
 *   stores.  Header bytes, the zero padding and the XXH32 trailer are written
 *   last, when the size is known.
 *
 * See DESIGN.md for the data layout and the roofline of this kernel.
 */
#include <cuda_runtime.h>

#include "airs_device.cuh"
#include "airs_launch.h"
#include "airs_private.h"

namespace {

constexpr uint32_t kThreads = AIRS_THREADS;
constexpr uint32_t kSpt = 8;                  /* samples per thread and tile */
constexpr uint32_t kTile = kThreads * kSpt;   /* 2048 samples = 4 KiB of u16 */
constexpr uint32_t kWarps = kThreads / 32;
constexpr uint32_t kStgWords = kTile * 48 / 32 + 16;
constexpr uint64_t kMask48 = 0xFFFFFFFFFFFFull;

/* context state between frames: struct cmp_context (ref cmp.h:129-137) */
struct JobState {
	airs_job job;
	uint64_t identifier;
	uint64_t counter; /* the timestamp source of this job */
	uint32_t valid;
	uint32_t seq;
	uint32_t model_size;
	uint32_t model_needed;
	uint8_t *work; /* this context's work buffer or NULL */
};

/* everything one pass over one frame needs, written by thread 0 */
struct PassPlan {
	EncConst enc;
	const uint8_t *src;
	uint8_t *dst;
	uint16_t *work;
	uint32_t pre;
	uint32_t n;
	uint32_t dtype;
	uint32_t hdr_len;
	uint32_t cap_eff;    /* bytes this pass may write */
	uint32_t trip;       /* cumulative bit count at which the reference's writer gives up */
	uint32_t model_mode; /* 0: none, 1: model := samples, 2: model update */
	uint32_t rate;
	uint32_t checksum;
	uint32_t fallback_ok;
	uint32_t align_check; /* SLOTS layout: dst must be 8-byte aligned like the reference's */
	uint32_t err;
	uint8_t hdr[24];
};

/* -------------------------------------------------------------------------
 * thread-0 logic
 * ---------------------------------------------------------------------- */

__device__ bool model_needed(const cmp_params &p) /* ref cmp.c:145-149 */
{
	return p.secondary_preprocessing == CMP_PREPROCESS_MODEL && p.secondary_iterations != 0;
}

__device__ uint32_t pre_work_size(uint32_t pre, uint32_t src_size, uint32_t &out)
{
	switch (pre) { /* ref preprocess.c:233,301-304,364-367 */
	case CMP_PREPROCESS_NONE:
	case CMP_PREPROCESS_DIFF:
		out = 0;
		return 0;
	case CMP_PREPROCESS_IWT:
	case CMP_PREPROCESS_MODEL:
		out = (src_size + 1u) & ~1u;
		return 0;
	default:
		return AIRS_ERR(PARAMS_INVALID);
	}
}

__device__ uint32_t work_buf_size(const cmp_params &p, uint32_t src_size) /* ref cmp.c:77-103 */
{
	uint32_t a = 0, b = 0, r;

	if (p.primary_preprocessing == CMP_PREPROCESS_MODEL)
		return AIRS_ERR(PARAMS_INVALID);
	r = pre_work_size(p.primary_preprocessing, src_size, a);
	if (airs_failed(r))
		return r;
	if (p.secondary_iterations) {
		r = pre_work_size(p.secondary_preprocessing, src_size, b);
		if (airs_failed(r))
			return r;
	}
	return a > b ? a : b;
}

__device__ void ctx_reset(JobState &s) /* ref cmp_reset, cmp.c:452-465 */
{
	s.seq = 0;
	s.identifier = s.counter++ & kMask48;
	s.model_size = 0;
}

/* cmp_initialise (ref cmp.c:152-209) */
__device__ __noinline__ uint32_t job_begin(JobState &s, const airs_job &j, uint8_t *work_base)
{
	const cmp_params &p = j.params;
	uint32_t r;

	s.job = j;
	s.valid = 0;
	s.seq = 0;
	s.identifier = 0;
	s.model_size = 0;
	s.counter = j.identifier_base;
	s.model_needed = model_needed(p);
	s.work = (work_base && j.work_size) ? work_base + j.work_offset : nullptr;

	if (airs_failed(j.work_size))
		return AIRS_ERR(GENERIC);
	if (p.secondary_iterations >= 256u)
		return AIRS_ERR(PARAMS_INVALID);
	r = airs_encoder_check(p.primary_encoder_type, p.primary_encoder_param, p.primary_encoder_outlier);
	if (airs_failed(r))
		return r;
	if (p.secondary_iterations) {
		r = airs_encoder_check(p.secondary_encoder_type, p.secondary_encoder_param,
				       p.secondary_encoder_outlier);
		if (airs_failed(r))
			return r;
	}
	if (s.model_needed && p.model_rate > 16u)
		return AIRS_ERR(PARAMS_INVALID);
	r = work_buf_size(p, 2);
	if (airs_failed(r))
		return r;
	if (r > 0) {
		if (!s.work)
			return AIRS_ERR(WORK_BUF_NULL);
		if (j.work_size == 0)
			return AIRS_ERR(WORK_BUF_TOO_SMALL);
		if ((uintptr_t)s.work & 1u)
			return AIRS_ERR(WORK_BUF_UNALIGNED);
	}
	s.valid = 1;
	ctx_reset(s);
	return 0;
}

/* continue a context the host shim keeps in the caller's struct cmp_context */
__device__ __noinline__ void job_resume(JobState &s, const airs_job &j, uint8_t *work_base,
					const airs_ctx_state &st)
{
	s.job = j;
	s.valid = st.valid;
	s.seq = st.seq;
	s.identifier = st.identifier;
	s.model_size = st.model_size;
	s.counter = st.counter;
	s.model_needed = model_needed(j.params);
	s.work = (work_base && j.work_size) ? work_base + j.work_offset : nullptr;
}

/* header bytes (ref cmp_hdr_serialize, header.c:24-67; fields cmp.c:265-279) */
__device__ uint32_t build_header(uint8_t *h, uint32_t n, uint64_t id, uint32_t seq, uint32_t pre,
				 const EncConst &e, uint32_t checksum, uint32_t rate)
{
	uint32_t orig = n * 2u;

	for (int k = 0; k < 24; k++)
		h[k] = 0;
	h[0] = (uint8_t)(0x80u | (CMP_VERSION_NUMBER >> 8));
	h[1] = (uint8_t)(CMP_VERSION_NUMBER & 0xFF);
	h[5] = (uint8_t)(orig >> 16);
	h[6] = (uint8_t)(orig >> 8);
	h[7] = (uint8_t)orig;
	for (int k = 0; k < 6; k++)
		h[8 + k] = (uint8_t)(id >> (8 * (5 - k)));
	h[14] = (uint8_t)seq;
	h[15] = (uint8_t)((pre << 4) | ((checksum ? 1u : 0u) << 3) | e.type);
	if (pre == CMP_PREPROCESS_NONE && e.type == CMP_ENCODER_UNCOMPRESSED)
		return CMP_HDR_SIZE;
	if (pre == CMP_PREPROCESS_MODEL)
		h[16] = (uint8_t)rate;
	if (e.type != CMP_ENCODER_UNCOMPRESSED) {
		h[17] = (uint8_t)(e.g >> 8);
		h[18] = (uint8_t)e.g;
		h[19] = (uint8_t)(e.outlier >> 16);
		h[20] = (uint8_t)(e.outlier >> 8);
		h[21] = (uint8_t)e.outlier;
	}
	return CMP_HDR_SIZE + 6u;
}

/* pass selection and every check that precedes the sample loop (ref cmp.c:228-294) */
__device__ __noinline__ void plan_pass(JobState &s, PassPlan &P, bool forced_raw)
{
	const cmp_params &p = s.job.params;
	uint32_t n = P.n, packed = n * 2u;
	uint32_t pre, type, g, user;

	P.err = 0;
	if (s.seq == 0 || s.seq > p.secondary_iterations) {
		ctx_reset(s);
		pre = forced_raw ? (uint32_t)CMP_PREPROCESS_NONE : (uint32_t)p.primary_preprocessing;
		type = forced_raw ? (uint32_t)CMP_ENCODER_UNCOMPRESSED : (uint32_t)p.primary_encoder_type;
		g = p.primary_encoder_param;
		user = p.primary_encoder_outlier;
		s.model_size = packed;
	} else {
		pre = p.secondary_preprocessing;
		type = p.secondary_encoder_type;
		g = p.secondary_encoder_param;
		user = p.secondary_encoder_outlier;
		if (s.model_needed && packed != s.model_size) {
			P.err = AIRS_ERR(SRC_SIZE_MISMATCH);
			return;
		}
	}
	P.pre = pre;
	P.work = (uint16_t *)s.work;
	P.model_mode = 0;
	if (s.model_needed) {
		if (s.job.work_size < packed) {
			P.err = AIRS_ERR(WORK_BUF_TOO_SMALL);
			return;
		}
		P.model_mode = s.seq == 0 ? 1u : 2u;
	}
	if (!P.dst) { /* ref bitstream_writer.h:65-68 */
		P.err = AIRS_ERR(DST_NULL);
		return;
	}
	if (P.align_check && ((uintptr_t)P.dst & 7u)) {
		P.err = AIRS_ERR(DST_UNALIGNED);
		return;
	}
	airs_enc_const(&P.enc, type, g, user);
	if (type != CMP_ENCODER_UNCOMPRESSED && P.enc.outlier == 0) {
		P.err = AIRS_ERR(PARAMS_INVALID);
		return;
	}
	if ((uint64_t)n * 2u > CMP_HDR_MAX_ORIGINAL_SIZE) {
		P.err = AIRS_ERR(HDR_ORIGINAL_TOO_LARGE);
		return;
	}
	P.rate = p.model_rate;
	P.checksum = p.checksum_enabled ? 1u : 0u;
	P.hdr_len = build_header(P.hdr, n, s.identifier, s.seq, pre, P.enc, P.checksum, p.model_rate);
	if (P.hdr_len > P.cap_eff) {
		P.err = AIRS_ERR(DST_TOO_SMALL);
		return;
	}
	if (pre == CMP_PREPROCESS_IWT || pre == CMP_PREPROCESS_MODEL) { /* ref preprocess.c:321-335,382-393 */
		if (!s.work)
			P.err = AIRS_ERR(WORK_BUF_NULL);
		else if (s.job.work_size < ((packed + 1u) & ~1u))
			P.err = AIRS_ERR(WORK_BUF_TOO_SMALL);
		else if ((uintptr_t)s.work & 1u)
			P.err = AIRS_ERR(WORK_BUF_UNALIGNED);
		if (P.err)
			return;
	}
	uint64_t trip = 64ull * ((uint64_t)P.cap_eff / 8 + 1);
	P.trip = trip > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)trip;
}

/* the checks of cmp_compress_* and cmp_compress_generic (ref cmp.c:342-364,396-435) */
__device__ __noinline__ void plan_frame(JobState &s, PassPlan &P, const AirsLaunch &b, uint32_t frame)
{
	const airs_job &j = s.job;
	uint32_t stride = j.dtype == AIRS_DTYPE_I16_IN_I32 ? 4u : 2u;

	P.err = 0;
	P.fallback_ok = 0;
	P.align_check = b.layout == AIRS_LAYOUT_SLOTS;
	P.src = b.src ? b.src + j.src_offset + (uint64_t)frame * j.src_frame_stride : nullptr;
	P.dst = b.dst ? b.dst + j.dst_offset + (uint64_t)frame * j.dst_frame_stride : nullptr;
	P.dtype = j.dtype;
	if (!P.src) {
		P.err = AIRS_ERR(SRC_NULL);
		return;
	}
	if (j.src_size == 0 || j.dtype > AIRS_DTYPE_U16 || j.src_size % stride) {
		P.err = AIRS_ERR(SRC_SIZE_WRONG);
		return;
	}
	if (!s.valid) {
		P.err = AIRS_ERR(CONTEXT_INVALID);
		return;
	}
	if (airs_failed(j.dst_capacity)) {
		P.err = AIRS_ERR(GENERIC);
		return;
	}
	P.n = j.src_size / stride;
	uint32_t raw_size = CMP_HDR_SIZE + P.n * 2u + (j.params.checksum_enabled ? 4u : 0u);
	P.fallback_ok = j.params.uncompressed_fallback_enabled && j.dst_capacity >= raw_size;
	P.cap_eff = P.fallback_ok ? raw_size : j.dst_capacity;
	plan_pass(s, P, false);
}

/* -------------------------------------------------------------------------
 * sample access
 * ---------------------------------------------------------------------- */

__device__ __forceinline__ uint32_t sample_at(const uint8_t *src, uint32_t dtype, uint32_t i)
{
	/* ref sample_read_i16, sample_reader.h:63-72 */
	if (dtype == AIRS_DTYPE_I16_IN_I32)
		return __ldg((const uint32_t *)src + i) & 0xFFFFu;
	return __ldg((const uint16_t *)src + i);
}

__device__ __forceinline__ void unpack8(const uint4 &v, uint32_t x[8])
{
	x[0] = v.x & 0xFFFFu;
	x[1] = v.x >> 16;
	x[2] = v.y & 0xFFFFu;
	x[3] = v.y >> 16;
	x[4] = v.z & 0xFFFFu;
	x[5] = v.z >> 16;
	x[6] = v.w & 0xFFFFu;
	x[7] = v.w >> 16;
}

/* 8 consecutive samples starting at i0 (zeros behind the end of the frame) */
__device__ __forceinline__ void load_samples(const uint8_t *src, uint32_t dtype, bool vec, uint32_t i0,
					     uint32_t n, uint32_t x[8])
{
	if (vec && i0 + 8 <= n) {
		if (dtype == AIRS_DTYPE_I16_IN_I32) {
			const uint4 *p = (const uint4 *)(src + (uint64_t)i0 * 4);
			uint4 a = __ldg(p), c = __ldg(p + 1);
			x[0] = a.x & 0xFFFFu;
			x[1] = a.y & 0xFFFFu;
			x[2] = a.z & 0xFFFFu;
			x[3] = a.w & 0xFFFFu;
			x[4] = c.x & 0xFFFFu;
			x[5] = c.y & 0xFFFFu;
			x[6] = c.z & 0xFFFFu;
			x[7] = c.w & 0xFFFFu;
		} else {
			unpack8(__ldg((const uint4 *)(src + (uint64_t)i0 * 2)), x);
		}
	} else {
#pragma unroll
		for (int j = 0; j < 8; j++)
			x[j] = i0 + j < n ? sample_at(src, dtype, i0 + j) : 0u;
	}
}

/* 8 consecutive 16-bit values of the work buffer (model or IWT coefficients) */
__device__ __forceinline__ void load_work(const uint16_t *w, bool vec, uint32_t i0, uint32_t n, uint32_t x[8])
{
	if (vec && i0 + 8 <= n) {
		unpack8(*(const uint4 *)(w + i0), x);
	} else {
#pragma unroll
		for (int j = 0; j < 8; j++)
			x[j] = i0 + j < n ? (uint32_t)w[i0 + j] : 0u;
	}
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

__device__ __noinline__ void iwt_global(const PassPlan &P)
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
__device__ __forceinline__ uint32_t pair_at(const PassPlan &P, bool al4, uint32_t i)
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

__device__ __noinline__ uint32_t frame_checksum(const PassPlan &P)
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
	uint32_t v1 = __shfl_sync(0xFFFFFFFFu, v, 1);
	uint32_t v2 = __shfl_sync(0xFFFFFFFFu, v, 2);
	uint32_t v3 = __shfl_sync(0xFFFFFFFFu, v, 3);
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
	return __shfl_sync(0xFFFFFFFFu, h, 0);
}

/* -------------------------------------------------------------------------
 * the sample loop
 * ---------------------------------------------------------------------- */

struct Shared {
	uint32_t stg[kStgWords]; /* MSB-first 32-bit words of the stream being assembled */
	uint32_t wsum[kWarps];
	JobState js;
	PassPlan plan;
	uint32_t job;
	uint32_t checksum;
	uint64_t offset; /* CONCAT: where the current frame's stream starts */
};

/* byte-granular window of the destination a pass may write: [lo, hi) in the
 * 4-byte-aligned address space that starts at dst - (dst & 3) */
struct OutWin {
	uint8_t *base4;
	uint32_t lo, hi;
};

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

template <int ENC, int PRE>
__device__ __noinline__ void encode_tiles(Shared &sh, const OutWin &o, uint32_t a, uint32_t &gw0, uint32_t &sbits,
					  bool size_only)
{
	const PassPlan &P = sh.plan;
	const EncConst e = P.enc;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t n = P.n, dtype = P.dtype, model_mode = P.model_mode;
	const bool src_vec = ((uintptr_t)P.src & 15u) == 0;
	const bool work_vec = ((uintptr_t)P.work & 15u) == 0;
	const bool is_signed = dtype != AIRS_DTYPE_U16;
	uint32_t *stg = sh.stg;

	for (uint32_t base = 0; base < n; base += kTile) {
		const uint32_t i0 = base + tid * kSpt;
		uint32_t x[8], m[8], r[8];

		if (PRE != CMP_PREPROCESS_IWT || model_mode)
			load_samples(P.src, dtype, src_vec, i0, n, x);
		if (PRE == CMP_PREPROCESS_MODEL || model_mode == 2)
			load_work(P.work, work_vec, i0, n, m);

		/* residuals: ref preprocess.c:268-290,348-353,406-411 */
		if (PRE == CMP_PREPROCESS_NONE) {
#pragma unroll
			for (int j = 0; j < 8; j++)
				r[j] = x[j];
		} else if (PRE == CMP_PREPROCESS_DIFF) {
			uint32_t prev = (i0 > 0 && i0 < n) ? sample_at(P.src, dtype, i0 - 1) : 0u;
			r[0] = x[0] - prev;
#pragma unroll
			for (int j = 1; j < 8; j++)
				r[j] = x[j] - x[j - 1];
		} else if (PRE == CMP_PREPROCESS_IWT) {
			load_work(P.work, work_vec, i0, n, r);
		} else {
#pragma unroll
			for (int j = 0; j < 8; j++)
				r[j] = x[j] - m[j];
		}

		/* codewords and lengths */
		uint32_t cw[8], cl[8], rw[8], rl[8], tb = 0;
#pragma unroll
		for (int j = 0; j < 8; j++) {
			airs_encode<ENC>(e, r[j], cw[j], cl[j], rw[j], rl[j]);
			if (i0 + j >= n) {
				cl[j] = 0;
				rl[j] = 0;
				cw[j] = 0;
				rw[j] = 0;
			}
			tb += cl[j] + rl[j];
		}

		/* exclusive scan of the per-thread bit counts over the CTA */
		uint32_t incl = tb;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
			if (lane >= (uint32_t)d)
				incl += t;
		}
		if (lane == 31)
			sh.wsum[warp] = incl;
		__syncthreads();
		uint32_t ws = lane < kWarps ? sh.wsum[lane] : 0u;
		uint32_t tile_bits = __reduce_add_sync(0xFFFFFFFFu, ws);
		uint32_t wpre = __reduce_add_sync(0xFFFFFFFFu, lane < warp ? ws : 0u);
		uint32_t excl = wpre + incl - tb;

		if (size_only) { /* CONCAT sizing pass: only the bit count is wanted */
			const uint32_t staged0 = sbits + tile_bits;
			gw0 += staged0 >> 5;
			sbits = staged0 & 31u;
			__syncthreads();
			continue;
		}

		/* bit packing: this thread's codewords start at staging bit `pos` */
		uint32_t pos = sbits + excl;
		uint32_t wp = pos >> 5, fill = pos & 31u, lo = 0;
		const uint32_t wfirst = wp;
		/* bits of the stream before this thread's first sample, for the model gate */
		uint32_t cum = gw0 * 32u + pos - 8u * a;
		uint32_t upd = 0; /* bit j: sample j still updates the model */

#pragma unroll
		for (int j = 0; j < 8; j++) {
			/* push (cw, cl) */
			{
				uint32_t len = cl[j];
				uint32_t hi = __funnelshift_lc(lo, 0u, len);
				lo = __funnelshift_lc(0u, lo, len) | cw[j];
				fill += len;
				if (fill >= 32u) {
					uint32_t word = __funnelshift_r(lo, hi, fill);
					if (wp == wfirst)
						atomicOr(&stg[wp], word);
					else
						stg[wp] = word;
					wp++;
					fill -= 32u;
				}
			}
			if (ENC == CMP_ENCODER_GOLOMB_MULTI) { /* raw part of an escape */
				uint32_t len = rl[j];
				uint32_t hi = __funnelshift_lc(lo, 0u, len);
				lo = __funnelshift_lc(0u, lo, len) | rw[j];
				fill += len;
				if (fill >= 32u) {
					uint32_t word = __funnelshift_r(lo, hi, fill);
					if (wp == wfirst)
						atomicOr(&stg[wp], word);
					else
						stg[wp] = word;
					wp++;
					fill -= 32u;
				}
			}
			cum += cl[j] + rl[j];
			if (cum < P.trip) /* ref cmp.c:300-302: the writer has not given up yet */
				upd |= 1u << j;
		}
		if (fill && tb)
			atomicOr(&stg[wp], lo << (32u - fill));

		/* model := samples, or model update (ref cmp.c:304-311) */
		if (model_mode) {
			uint32_t nm[8];
#pragma unroll
			for (int j = 0; j < 8; j++)
				nm[j] = model_mode == 1 ? x[j] : airs_model_update(x[j], m[j], P.rate, is_signed);
			if (work_vec && i0 + 8 <= n && upd == 0xFFu) {
				uint4 v;
				v.x = nm[0] | (nm[1] << 16);
				v.y = nm[2] | (nm[3] << 16);
				v.z = nm[4] | (nm[5] << 16);
				v.w = nm[6] | (nm[7] << 16);
				*(uint4 *)(P.work + i0) = v;
			} else {
#pragma unroll
				for (int j = 0; j < 8; j++)
					if (i0 + j < n && (upd >> j & 1u))
						P.work[i0 + j] = (uint16_t)nm[j];
			}
		}
		__syncthreads();

		/* staged full words leave as coalesced stores; the staging area is zeroed behind */
		const uint32_t staged = sbits + tile_bits;
		const uint32_t wfull = staged >> 5;
		for (uint32_t w = tid; w < wfull; w += kThreads) {
			uint32_t v = stg[w];
			stg[w] = 0;
			store_word(o, gw0 + w, v);
		}
		__syncthreads();
		if (tid == 0 && wfull) { /* the trailing partial word becomes word 0 of the next tile */
			uint32_t c = stg[wfull];
			stg[wfull] = 0;
			stg[0] = c;
		}
		gw0 += wfull;
		sbits = staged & 31u;
	}
	__syncthreads();
}

/* one pass over one frame; returns the stream size or an error (uniform over the CTA).
 * ref compress_engine, cmp.c:213-338 */
__device__ uint32_t encode_pass(Shared &sh, bool size_only, bool suppress)
{
	const PassPlan &P = sh.plan;
	const uint32_t tid = threadIdx.x;

	if (P.err)
		return P.err;

	const uint32_t a = (uint32_t)((uintptr_t)P.dst & 3u);
	OutWin o;
	o.base4 = P.dst - a;
	o.lo = a + P.hdr_len;
	o.hi = suppress ? o.lo : a + P.cap_eff; /* suppress: run for the model side effects only */
	/* position of stg[0] in 32-bit words of the aligned space, and bits already in it */
	uint32_t gw0 = (8u * (a + P.hdr_len)) >> 5;
	uint32_t sbits = (8u * (a + P.hdr_len)) & 31u;

	if (P.pre == CMP_PREPROCESS_IWT)
		iwt_global(P);
	if (size_only) {
		const uint32_t fb = gw0 * 32u + sbits - 8u * a;
		(void)fb;
	}

#define AIRS_DISPATCH_PRE(ENC)                                                              \
	switch (P.pre) {                                                                    \
	case CMP_PREPROCESS_NONE:  encode_tiles<ENC, CMP_PREPROCESS_NONE>(sh, o, a, gw0, sbits, size_only); break;  \
	case CMP_PREPROCESS_DIFF:  encode_tiles<ENC, CMP_PREPROCESS_DIFF>(sh, o, a, gw0, sbits, size_only); break;  \
	case CMP_PREPROCESS_IWT:   encode_tiles<ENC, CMP_PREPROCESS_IWT>(sh, o, a, gw0, sbits, size_only); break;   \
	default:                   encode_tiles<ENC, CMP_PREPROCESS_MODEL>(sh, o, a, gw0, sbits, size_only); break; \
	}
	switch (P.enc.type) {
	case CMP_ENCODER_UNCOMPRESSED: AIRS_DISPATCH_PRE(CMP_ENCODER_UNCOMPRESSED) break;
	case CMP_ENCODER_GOLOMB_ZERO:  AIRS_DISPATCH_PRE(CMP_ENCODER_GOLOMB_ZERO) break;
	default:                       AIRS_DISPATCH_PRE(CMP_ENCODER_GOLOMB_MULTI) break;
	}
#undef AIRS_DISPATCH_PRE

	if (size_only) { /* the size is all the CONCAT look-back needs */
		const uint32_t bits = gw0 * 32u + sbits - 8u * a;
		const uint32_t sz = ((bits + 7u) >> 3) + (P.checksum ? 4u : 0u);
		if (sz > P.cap_eff)
			return AIRS_ERR(DST_TOO_SMALL);
		if (sz > CMP_HDR_MAX_COMPRESSED_SIZE)
			return AIRS_ERR(HDR_CMP_SIZE_TOO_LARGE);
		return sz;
	}

	/* checksum of the samples while the tail is flushed */
	if (P.checksum && !suppress && tid < 32) {
		uint32_t h = frame_checksum(P);
		if (tid == 0)
			sh.checksum = h;
	}

	const uint32_t frame_bits = gw0 * 32u + sbits - 8u * a;
	const uint32_t payload_end = (frame_bits + 7u) >> 3; /* header + code bytes */
	const uint32_t size = payload_end + (P.checksum ? 4u : 0u);

	if (tid == 32 || (kThreads <= 32 && tid == 0)) { /* last partial word, zero padded (ref bitstream_writer.h:205-227) */
		uint32_t v = sh.stg[0];
		uint32_t nb = (sbits + 7u) >> 3;
		for (uint32_t k = 0; k < nb; k++) {
			uint64_t b = (uint64_t)gw0 * 4 + k;
			if (b >= o.lo && b < o.hi)
				o.base4[b] = (uint8_t)(v >> (24 - 8 * k));
		}
		sh.stg[0] = 0;
	}
	__syncthreads();

	uint32_t result;
	if (size > P.cap_eff)
		result = AIRS_ERR(DST_TOO_SMALL);
	else if (size > CMP_HDR_MAX_COMPRESSED_SIZE)
		result = AIRS_ERR(HDR_CMP_SIZE_TOO_LARGE);
	else
		result = size;

	if (P.checksum && !suppress && tid < 4) { /* trailer, big endian (ref cmp.c:314-319) */
		uint64_t b = (uint64_t)a + payload_end + tid;
		if (b < o.hi)
			o.base4[b] = (uint8_t)(sh.checksum >> (24 - 8 * tid));
	}
	if (!airs_failed(result) && !suppress && tid < P.hdr_len) { /* header with the final size (ref cmp.c:329-334) */
		uint8_t v = P.hdr[tid];
		if (tid >= 2 && tid <= 4)
			v = (uint8_t)(size >> (8 * (4 - tid)));
		P.dst[tid] = v;
	}
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
		} while (__any_sync(0xFFFFFFFFu, (v >> 62) == 0));
		const uint32_t incl = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
		if (incl) {
			const int first = __ffs((int)incl) - 1; /* nearest predecessor holding a prefix */
			excl += __reduce_add_sync(0xFFFFFFFFu, (int)lane < first ? (uint32_t)(v & kValue) : 0u);
			const uint64_t pv = v & kValue;
			const uint32_t plo = __shfl_sync(0xFFFFFFFFu, (uint32_t)pv, first);
			const uint32_t phi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(pv >> 32), first);
			excl += ((uint64_t)phi << 32) | plo;
			break;
		}
		excl += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)(v & kValue));
	}
	if (lane == 0)
		st[k] = (2ull << 62) | (excl + my_size);
	return excl;
}

} /* namespace */

__global__ void __launch_bounds__(AIRS_THREADS, 3) airs_encode_kernel(AirsLaunch b)
{
	__shared__ Shared sh;
	const uint32_t tid = threadIdx.x;

	for (uint32_t w = tid; w < kStgWords; w += kThreads)
		sh.stg[w] = 0;

	for (;;) {
		__syncthreads();
		if (tid == 0)
			sh.job = atomicAdd(b.ticket, 1u);
		__syncthreads();
		const uint32_t job = sh.job;
		if (job >= b.n_jobs)
			break;
		if (tid == 0) {
			if (b.ctx_io) {
				job_resume(sh.js, b.jobs[job], b.work, b.ctx_io[job]);
			} else {
				uint32_t r = job_begin(sh.js, b.jobs[job], b.work);
				if (b.init_results)
					b.init_results[job] = r;
			}
		}
		__syncthreads();
		const uint32_t n_frames = sh.js.job.n_frames;
		const uint32_t first = sh.js.job.first_result;

		for (uint32_t f = 0; f < n_frames; f++) {
			if (tid == 0)
				plan_frame(sh.js, sh.plan, b, f);
			__syncthreads();
			uint32_t r;
			if (b.layout == AIRS_LAYOUT_CONCAT) {
				/* size first (exact, no output), then the offset from the scan, then one
				 * pass that writes; a frame that fails contributes no bytes */
				r = encode_pass(sh, true, false);
				if (sh.plan.fallback_ok && r == AIRS_ERR(DST_TOO_SMALL)) {
					__syncthreads();
					if (tid == 0) {
						ctx_reset(sh.js);
						plan_pass(sh.js, sh.plan, true);
					}
					__syncthreads();
					r = sh.plan.err ? sh.plan.err : sh.plan.cap_eff;
				}
				const uint32_t k = first + f;
				if (tid < 32) {
					uint64_t off = lookback_offset(b.lookback, k, airs_failed(r) ? 0u : r);
					if (tid == 0) {
						sh.offset = off;
						b.out_offsets[k] = off;
						if (k + 1 == b.n_results)
							b.out_offsets[k + 1] = off + (airs_failed(r) ? 0u : r);
						sh.plan.dst = b.dst + off;
					}
				}
				__syncthreads();
				const bool fits = !airs_failed(r) && sh.offset + r <= b.dst_size;
				if (!sh.plan.err) {
					uint32_t r2 = encode_pass(sh, false, !fits);
					if (!airs_failed(r))
						r = fits ? r2 : AIRS_ERR(DST_TOO_SMALL);
				}
			} else {
				r = encode_pass(sh, false, false);
				if (sh.plan.fallback_ok && r == AIRS_ERR(DST_TOO_SMALL)) {
					/* store the frame raw as a fresh primary pass (ref cmp.c:380-392) */
					__syncthreads();
					if (tid == 0) {
						ctx_reset(sh.js);
						plan_pass(sh.js, sh.plan, true);
					}
					__syncthreads();
					r = encode_pass(sh, false, false);
				}
			}
			__syncthreads();
			if (tid == 0) {
				if (!airs_failed(r))
					sh.js.seq = (sh.js.seq + 1u) & 0xFFu;
				b.results[first + f] = r;
			}
		}
		if (tid == 0 && b.ctx_io) {
			airs_ctx_state &st = b.ctx_io[job];
			st.identifier = sh.js.identifier;
			st.counter = sh.js.counter;
			st.seq = sh.js.seq;
			st.model_size = sh.js.model_size;
		}
	}
}

extern "C" cudaError_t airs_launch_encode(const AirsLaunch *b, unsigned int grid, cudaStream_t stream)
{
	airs_encode_kernel<<<grid, AIRS_THREADS, 0, stream>>>(*b);
	return cudaGetLastError();
}
