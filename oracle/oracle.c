/*
 * oracle.c - CPU restatement of the AIRSPACE (airs-compression v0.6.0)
 * compression path, written from the format specification in SURVEY.md App. A
 * and from reading the reference.  Whole-array, two-step style (residuals
 * first, then codes through a byte-wise bit sink); no vtables, no streaming
 * 64-bit cache.  Citations are into /root/reference/.
 *
 * TEST INFRASTRUCTURE ONLY - see oracle.h.  Parity status: PINNED (oracle.h).
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define E(name) ((uint32_t)0 - (uint32_t)CMP_ERR_##name)

static int failed(uint32_t r)
{
	return r > (uint32_t)0 - (uint32_t)CMP_ERR_MAX_CODE; /* err_private.h:44-47 */
}

/* ------------------------------------------------------------------ XXH32 */
/* xxHash 0.8.3 (subprojects/xxhash.wrap:2), 32-bit variant, from its spec. */

#define XP1 0x9E3779B1u
#define XP2 0x85EBCA77u
#define XP3 0xC2B2AE3Du
#define XP4 0x27D4EB2Fu
#define XP5 0x165667B1u

static uint32_t rotl(uint32_t v, unsigned int r)
{
	return (v << r) | (v >> (32u - r));
}

static uint32_t le32(const uint8_t *p)
{
	return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}

uint32_t oracle_xxh32(const void *data, uint32_t len, uint32_t seed)
{
	const uint8_t *p = data;
	uint32_t left = len;
	uint32_t h;

	if (len >= 16) {
		uint32_t lane[4];
		int k;

		lane[0] = seed + XP1 + XP2;
		lane[1] = seed + XP2;
		lane[2] = seed;
		lane[3] = seed - XP1;
		for (; left >= 16; left -= 16, p += 16)
			for (k = 0; k < 4; k++)
				lane[k] = rotl(lane[k] + le32(p + 4 * k) * XP2, 13) * XP1;
		h = rotl(lane[0], 1) + rotl(lane[1], 7) + rotl(lane[2], 12) + rotl(lane[3], 18);
	} else {
		h = seed + XP5;
	}
	h += len;
	for (; left >= 4; left -= 4, p += 4)
		h = rotl(h + le32(p) * XP3, 17) * XP4;
	for (; left; left--, p++)
		h = rotl(h + (uint32_t)*p * XP5, 11) * XP1;
	h ^= h >> 15;
	h *= XP2;
	h ^= h >> 13;
	h *= XP3;
	h ^= h >> 16;
	return h;
}

/* ---------------------------------------------------------------- samples */

static uint16_t sample(const void *src, uint32_t i, uint32_t dtype)
{
	/* sample_reader.h:63-72: an i32 container contributes its low 16 bits */
	if (dtype == AIRS_DTYPE_I16_IN_I32)
		return (uint16_t)(((const uint32_t *)src)[i] & 0xFFFFu);
	return ((const uint16_t *)src)[i];
}

/* header.c:137-163 with header_private.h:46: hash of the big-endian samples */
uint32_t oracle_checksum(const void *src, uint32_t n, uint32_t dtype)
{
	uint8_t *be = malloc(n ? (size_t)n * 2 : 1);
	uint32_t i, h;

	for (i = 0; i < n; i++) {
		uint16_t v = sample(src, i, dtype);

		be[2 * i] = (uint8_t)(v >> 8);
		be[2 * i + 1] = (uint8_t)v;
	}
	h = oracle_xxh32(be, n * 2, 419764627u);
	free(be);
	return h;
}

/* ---------------------------------------------------------------- encoder */

static uint32_t floor_log2(uint32_t v) /* v > 0 */
{
	uint32_t l = 0;

	while (v >>= 1)
		l++;
	return l;
}

/* encoder.c:63-110, 154-182, 185-224 */
uint32_t oracle_outlier(uint32_t encoder_type, uint32_t g, uint32_t user_outlier)
{
	uint64_t cutoff, limit, o;
	uint32_t L;

	if (g < 1 || g > 65535)
		return 0;
	L = floor_log2(g);
	cutoff = (2ull << L) - g;
	/* first value whose Golomb codeword would exceed 32 bits */
	limit = cutoff + (uint64_t)(31 - L) * g;
	if (encoder_type == CMP_ENCODER_GOLOMB_MULTI) {
		if (limit <= 8) /* eight escape symbols must stay codable */
			return 0;
		limit -= 8;
		o = user_outlier;
	} else {
		o = cutoff + 16ull * g - 1; /* escape as soon as it is not longer */
	}
	return (uint32_t)(o < limit ? o : limit);
}

static uint32_t encoder_check(uint32_t type, uint32_t g, uint32_t user_outlier)
{
	if (type == CMP_ENCODER_UNCOMPRESSED)
		return 0;
	if (type != CMP_ENCODER_GOLOMB_ZERO && type != CMP_ENCODER_GOLOMB_MULTI)
		return E(PARAMS_INVALID);
	if (oracle_outlier(type, g, user_outlier) == 0)
		return E(PARAMS_INVALID);
	return 0;
}

/* encoder.c:303-324: returns the length, code in *cw (<= 32 bits) */
static uint32_t golomb(uint32_t v, uint32_t g, uint32_t *cw)
{
	uint32_t L = floor_log2(g);
	uint32_t cutoff = (2u << L) - g;
	uint32_t q, r;

	if (v < cutoff) {
		*cw = v;
		return L + 1;
	}
	q = (v - cutoff) / g;
	r = (v - cutoff) % g;
	*cw = (q >= 32 ? 0 : ((1u << q) - 1) << ((L + 2) & 31)) + 2 * cutoff + r;
	return L + 2 + q;
}

/* encoder.c:274-286, 327-378 */
uint32_t oracle_encode_residual(uint32_t encoder_type, uint32_t g, uint32_t outlier, int16_t r,
				uint64_t *code)
{
	uint16_t ur = (uint16_t)r;
	uint16_t m = (uint16_t)((uint16_t)(ur << 1) ^ (uint16_t)(0 - (ur >> 15))); /* zig-zag */
	uint32_t cw, len;

	switch (encoder_type) {
	case CMP_ENCODER_UNCOMPRESSED:
		*code = ur;
		return 16;
	case CMP_ENCODER_GOLOMB_ZERO:
		if (m < outlier) {
			len = golomb((uint32_t)m + 1, g, &cw);
			*code = cw;
			return len;
		}
		/* codeword 0 (floor_log2(g)+1 zero bits), then the mapped value raw */
		*code = m;
		return floor_log2(g) + 1 + 16;
	default: { /* CMP_ENCODER_GOLOMB_MULTI */
		uint32_t d, level, raw_bits;

		if (m < outlier) {
			len = golomb(m, g, &cw);
			*code = cw;
			return len;
		}
		d = m - outlier;
		level = d < 4 ? 0 : floor_log2(d) / 2;
		raw_bits = 2 * (level + 1);
		len = golomb(outlier + level, g, &cw);
		*code = ((uint64_t)cw << raw_bits) | d;
		return len + raw_bits;
	}
	}
}

/* ------------------------------------------------------------ preprocessing */

static int16_t w16(int32_t v)
{
	return (int16_t)(uint16_t)(uint32_t)v; /* every IWT result wraps to 16 bit */
}

/* preprocess.c:140-177: one lifting level over the positions k*s */
static void iwt_level(int16_t *v, uint32_t n, uint32_t s)
{
	uint32_t i;

	/* details (odd multiples of s) from the untouched even neighbours */
	for (i = s; i < n; i += 2 * s) {
		if (i + s < n)
			v[i] = w16(v[i] - w16(((int32_t)v[i - s] + v[i + s]) >> 1));
		else
			v[i] = w16(v[i] - v[i - s]);
	}
	/* approximations (even multiples of s) from the new details */
	for (i = 0; i < n; i += 2 * s) {
		int has_l = i >= s, has_r = i + s < n;

		if (has_l && has_r)
			v[i] = w16(v[i] + w16(((int32_t)v[i - s] + v[i + s]) >> 2));
		else if (has_r)
			v[i] = w16(v[i] + w16((int32_t)v[i + s] >> 1));
		else if (has_l)
			v[i] = w16(v[i] + w16((int32_t)v[i - s] >> 1));
	}
}

/* preprocess.c:190-221 */
void oracle_iwt(int16_t *v, uint32_t n)
{
	uint64_t s;

	for (s = 1; s < n; s <<= 1)
		iwt_level(v, n, (uint32_t)s);
}

/* cmp.c:120-142 */
uint16_t oracle_model_update(uint16_t data, uint16_t model, uint32_t rate, uint32_t dtype)
{
	int32_t d = dtype == AIRS_DTYPE_U16 ? (int32_t)data : (int32_t)(int16_t)data;
	int32_t m = dtype == AIRS_DTYPE_U16 ? (int32_t)model : (int32_t)(int16_t)model;
	int32_t mix = m * (int32_t)rate + d * (16 - (int32_t)rate);

	return (uint16_t)(uint32_t)(mix >> 4); /* arithmetic shift (gcc) */
}

/* ------------------------------------------------------------------ bit sink */
/*
 * Models lib/common/bitstream_writer.h:124-227 by its observable behaviour:
 * bits go out MSB first; a write that completes 64-bit word number
 * floor(cap/8) (0-based) raises a sticky DST_TOO_SMALL and is dropped, like
 * every write after it; the final flush needs ceil(bits/8) <= cap.
 */
struct sink {
	uint8_t *dst;
	uint32_t cap;
	uint64_t nbits;
	uint64_t trip;  /* first cumulative bit count that overflows */
	uint64_t acc;
	uint32_t fill;  /* pending bits in acc, < 8 between calls */
	int overflow;
};

static void sink_open(struct sink *s, uint8_t *dst, uint32_t cap)
{
	memset(s, 0, sizeof(*s));
	s->dst = dst;
	s->cap = cap;
	s->trip = 64ull * ((uint64_t)cap / 8 + 1);
}

static void sink_put(struct sink *s, uint32_t bits, uint32_t len) /* len <= 32 */
{
	if (s->overflow || len == 0)
		return;
	if (s->nbits + len >= s->trip) {
		s->overflow = 1;
		return;
	}
	s->acc = (s->acc << len) | bits;
	s->fill += len;
	while (s->fill >= 8) {
		uint64_t at = (s->nbits + len - s->fill) >> 3;

		if (at < s->cap)
			s->dst[at] = (uint8_t)(s->acc >> (s->fill - 8));
		s->fill -= 8;
	}
	s->nbits += len;
}

static void sink_align(struct sink *s) /* bitstream_writer.h:186-192 */
{
	if (s->nbits & 7)
		sink_put(s, 0, 8 - (uint32_t)(s->nbits & 7));
}

/* total bytes or DST_TOO_SMALL; pads the last byte with zeros (:205-227) */
static uint32_t sink_close(struct sink *s)
{
	uint64_t bytes = (s->nbits + 7) >> 3;

	if (s->overflow || bytes > s->cap)
		return E(DST_TOO_SMALL);
	if (s->fill)
		s->dst[bytes - 1] = (uint8_t)(s->acc << (8 - s->fill));
	return (uint32_t)bytes;
}

/* ------------------------------------------------------------------ context */

static int model_needed(const struct cmp_params *p) /* cmp.c:145-149 */
{
	return p->secondary_preprocessing == CMP_PREPROCESS_MODEL && p->secondary_iterations != 0;
}

uint32_t oracle_compress_bound(uint32_t packed_size) /* cmp.c:59-74, encoder.c:381-386 */
{
	uint64_t n, bound;

	if (packed_size > CMP_HDR_MAX_ORIGINAL_SIZE)
		return E(HDR_ORIGINAL_TOO_LARGE);
	n = ((uint64_t)packed_size * 8 + 15) / 16;
	bound = 22 + 4 + (n * 48 + 7) / 8;
	if (bound > CMP_HDR_MAX_COMPRESSED_SIZE)
		return E(HDR_CMP_SIZE_TOO_LARGE);
	return (uint32_t)bound;
}

static uint32_t pre_work_size(uint32_t pre, uint32_t src_size, uint32_t *out)
{
	switch (pre) {
	case CMP_PREPROCESS_NONE:
	case CMP_PREPROCESS_DIFF:
		*out = 0;
		return 0;
	case CMP_PREPROCESS_IWT:
	case CMP_PREPROCESS_MODEL:
		*out = (src_size + 1u) & ~1u; /* preprocess.h ROUND_UP_TO_NEXT_2 */
		return 0;
	default:
		return E(PARAMS_INVALID);
	}
}

uint32_t oracle_work_buf_size(const struct cmp_params *p, uint32_t src_size) /* cmp.c:77-103 */
{
	uint32_t a, b = 0, r;

	if (!p)
		return E(GENERIC);
	if (p->primary_preprocessing == CMP_PREPROCESS_MODEL)
		return E(PARAMS_INVALID);
	r = pre_work_size(p->primary_preprocessing, src_size, &a);
	if (failed(r))
		return r;
	if (p->secondary_iterations) {
		r = pre_work_size(p->secondary_preprocessing, src_size, &b);
		if (failed(r))
			return r;
	}
	return a > b ? a : b;
}

uint32_t oracle_reset(struct oracle_ctx *c) /* cmp.c:452-465, 438-449 */
{
	if (!c)
		return E(GENERIC);
	if (!c->valid)
		return E(CONTEXT_INVALID);
	c->seq = 0;
	c->identifier = c->counter++ & 0xFFFFFFFFFFFFull;
	c->model_size = 0;
	return 0;
}

uint32_t oracle_init(struct oracle_ctx *c, const struct cmp_params *p, void *work, uint32_t work_size,
		     uint64_t identifier_base) /* cmp.c:152-209 */
{
	uint32_t r, need;

	if (!c)
		return E(GENERIC);
	memset(c, 0, sizeof(*c));
	c->counter = identifier_base;
	if (!p)
		return E(GENERIC);
	if (failed(work_size))
		return E(GENERIC);
	if (p->secondary_iterations >= 256)
		return E(PARAMS_INVALID);
	r = encoder_check(p->primary_encoder_type, p->primary_encoder_param, p->primary_encoder_outlier);
	if (failed(r))
		return r;
	if (p->secondary_iterations) {
		r = encoder_check(p->secondary_encoder_type, p->secondary_encoder_param,
				  p->secondary_encoder_outlier);
		if (failed(r))
			return r;
	}
	if (model_needed(p) && p->model_rate > 16)
		return E(PARAMS_INVALID);
	need = oracle_work_buf_size(p, 2);
	if (failed(need))
		return need;
	if (need > 0) {
		if (!work)
			return E(WORK_BUF_NULL);
		if (work_size == 0)
			return E(WORK_BUF_TOO_SMALL);
		if ((uintptr_t)work & 1)
			return E(WORK_BUF_UNALIGNED);
	}
	c->params = *p;
	c->work = work;
	c->work_size = work_size;
	c->valid = 1;
	return oracle_reset(c);
}

/* ------------------------------------------------------------------- engine */

/* header.c:24-67 and cmp.c:265-279; returns the header length */
static uint32_t make_header(uint8_t h[22], uint32_t n, uint64_t identifier, uint8_t seq, uint32_t pre,
			    uint32_t enc, int checksum, uint32_t model_rate, uint32_t g,
			    uint32_t outlier)
{
	uint32_t orig = n * 2;
	int k;

	memset(h, 0, 22);
	h[0] = (uint8_t)(0x80 | (CMP_VERSION_NUMBER >> 8));
	h[1] = (uint8_t)CMP_VERSION_NUMBER;
	/* h[2..4]: compressed size, filled in at the end */
	h[5] = (uint8_t)(orig >> 16);
	h[6] = (uint8_t)(orig >> 8);
	h[7] = (uint8_t)orig;
	for (k = 0; k < 6; k++)
		h[8 + k] = (uint8_t)(identifier >> (8 * (5 - k)));
	h[14] = seq;
	h[15] = (uint8_t)((pre << 4) | ((checksum ? 1u : 0u) << 3) | enc);
	if (pre == CMP_PREPROCESS_NONE && enc == CMP_ENCODER_UNCOMPRESSED)
		return 16;
	if (pre == CMP_PREPROCESS_MODEL)
		h[16] = (uint8_t)model_rate;
	if (enc != CMP_ENCODER_UNCOMPRESSED) {
		h[17] = (uint8_t)(g >> 8);
		h[18] = (uint8_t)g;
		h[19] = (uint8_t)(outlier >> 16);
		h[20] = (uint8_t)(outlier >> 8);
		h[21] = (uint8_t)outlier;
	}
	return 22;
}

/* cmp.c:213-338 */
static uint32_t engine(struct oracle_ctx *c, uint8_t *dst, uint32_t cap, const void *src, uint32_t n,
		       uint32_t dtype)
{
	const struct cmp_params *p = &c->params;
	uint32_t pre, enc, g, user_outlier, outlier = 0;
	uint16_t *model = NULL;
	const int16_t *coef = NULL;
	struct sink s;
	uint8_t hdr[22];
	uint32_t hdr_len, i, size, k;
	uint64_t packed = (uint64_t)n * 2;

	if (c->seq == 0 || c->seq > p->secondary_iterations) { /* :228-236 */
		uint32_t r = oracle_reset(c);

		if (failed(r))
			return r;
		pre = p->primary_preprocessing;
		enc = p->primary_encoder_type;
		g = p->primary_encoder_param;
		user_outlier = p->primary_encoder_outlier;
		c->model_size = (uint32_t)packed;
	} else { /* :237-248 */
		pre = p->secondary_preprocessing;
		enc = p->secondary_encoder_type;
		g = p->secondary_encoder_param;
		user_outlier = p->secondary_encoder_outlier;
		if (model_needed(p) && (uint32_t)packed != c->model_size)
			return E(SRC_SIZE_MISMATCH);
	}
	if (model_needed(p)) { /* :250-254 */
		if (c->work_size < (uint32_t)packed)
			return E(WORK_BUF_TOO_SMALL);
		model = c->work;
	}
	if (!dst) /* bitstream_writer.h:65-68 */
		return E(DST_NULL);
	if ((uintptr_t)dst & 7)
		return E(DST_UNALIGNED);
	if (enc != CMP_ENCODER_UNCOMPRESSED) { /* encoder.c:185-224 */
		if (enc != CMP_ENCODER_GOLOMB_ZERO && enc != CMP_ENCODER_GOLOMB_MULTI)
			return E(PARAMS_INVALID);
		outlier = oracle_outlier(enc, g, user_outlier);
		if (!outlier)
			return E(PARAMS_INVALID);
	}
	if (packed > CMP_HDR_MAX_ORIGINAL_SIZE) /* header.c:34 */
		return E(HDR_ORIGINAL_TOO_LARGE);

	sink_open(&s, dst, cap);
	hdr_len = make_header(hdr, n, c->identifier, c->seq, pre, enc, p->checksum_enabled,
			      p->model_rate, g, outlier);
	if (hdr_len > cap) /* the placeholder header does not fit: header.c:62 */
		return E(DST_TOO_SMALL);
	for (k = 0; k < hdr_len; k++)
		sink_put(&s, hdr[k], 8);

	switch (pre) { /* preprocess.c:250-393 */
	case CMP_PREPROCESS_NONE:
	case CMP_PREPROCESS_DIFF:
		break;
	case CMP_PREPROCESS_IWT:
	case CMP_PREPROCESS_MODEL:
		if (!c->work)
			return E(WORK_BUF_NULL);
		if (c->work_size < (((uint32_t)packed + 1u) & ~1u))
			return E(WORK_BUF_TOO_SMALL);
		if ((uintptr_t)c->work & 1)
			return E(WORK_BUF_UNALIGNED);
		if (pre == CMP_PREPROCESS_IWT) {
			for (i = 0; i < n; i++)
				c->work[i] = sample(src, i, dtype);
			oracle_iwt((int16_t *)c->work, n);
			coef = (const int16_t *)c->work;
		}
		break;
	default:
		return E(PARAMS_INVALID);
	}

	for (i = 0; i < n; i++) { /* cmp.c:296-312 */
		uint16_t x = sample(src, i, dtype);
		uint16_t r;
		uint64_t code;
		uint32_t len;

		switch (pre) {
		case CMP_PREPROCESS_DIFF:
			r = i ? (uint16_t)(x - sample(src, i - 1, dtype)) : x;
			break;
		case CMP_PREPROCESS_IWT:
			r = (uint16_t)coef[i];
			break;
		case CMP_PREPROCESS_MODEL:
			r = (uint16_t)(x - c->work[i]);
			break;
		default:
			r = x;
		}
		len = oracle_encode_residual(enc, g, outlier, (int16_t)r, &code);
		if (len > 32) {
			sink_put(&s, (uint32_t)(code >> 16), len - 16);
			sink_put(&s, (uint32_t)(code & 0xFFFF), 16);
		} else {
			sink_put(&s, (uint32_t)code, len);
		}
		if (s.overflow)
			break; /* :300-302: sample i does not update the model */
		if (model)
			model[i] = c->seq == 0 ? x : oracle_model_update(x, model[i], p->model_rate, dtype);
	}

	if (p->checksum_enabled) { /* :314-319 */
		uint32_t h = oracle_checksum(src, n, dtype);

		sink_align(&s);
		sink_put(&s, h, 32);
	}
	size = sink_close(&s);
	if (failed(size))
		return size;
	if (size > CMP_HDR_MAX_COMPRESSED_SIZE) /* second serialise, header.c:31 */
		return E(HDR_CMP_SIZE_TOO_LARGE);
	dst[2] = (uint8_t)(size >> 16);
	dst[3] = (uint8_t)(size >> 8);
	dst[4] = (uint8_t)size;
	c->seq++;
	return size;
}

/* cmp.c:342-435 (with the container checks of sample_reader.h:19-51) */
uint32_t oracle_compress(struct oracle_ctx *c, void *dst, uint32_t cap, const void *src,
			 uint32_t src_size, uint32_t dtype)
{
	uint32_t stride = dtype == AIRS_DTYPE_I16_IN_I32 ? 4 : 2;
	uint32_t n, raw_size, r;
	uint32_t keep_pre, keep_enc;

	if (!src)
		return E(SRC_NULL);
	if (src_size == 0 || dtype > AIRS_DTYPE_U16 || src_size % stride)
		return E(SRC_SIZE_WRONG);
	n = src_size / stride;
	if (!c)
		return E(GENERIC);
	if (!c->valid)
		return E(CONTEXT_INVALID);
	if (failed(cap))
		return E(GENERIC);

	raw_size = 16 + n * 2 + (c->params.checksum_enabled ? 4 : 0);
	if (!c->params.uncompressed_fallback_enabled || cap < raw_size)
		return engine(c, dst, cap, src, n, dtype);

	r = engine(c, dst, raw_size, src, n, dtype);
	if (r != E(DST_TOO_SMALL))
		return r;
	/* :380-392: store it raw, as a fresh primary pass */
	r = oracle_reset(c);
	if (failed(r))
		return r;
	keep_pre = c->params.primary_preprocessing;
	keep_enc = c->params.primary_encoder_type;
	c->params.primary_preprocessing = CMP_PREPROCESS_NONE;
	c->params.primary_encoder_type = CMP_ENCODER_UNCOMPRESSED;
	r = engine(c, dst, raw_size, src, n, dtype);
	c->params.primary_preprocessing = keep_pre;
	c->params.primary_encoder_type = keep_enc;
	return r;
}

/* -------------------------------------------------------------------- batch */

int oracle_run_jobs(const void *src, void *dst, void *work, const struct airs_job *jobs,
		    uint32_t job_begin, uint32_t job_end, uint32_t layout, uint32_t *results,
		    uint32_t *init_results, uint64_t *out_offsets)
{
	uint64_t cursor = 0;
	uint32_t j, f;
	/* CONCAT streams need not start 8-byte aligned: encode into a bounce buffer */
	uint8_t *bounce = NULL;
	uint32_t bounce_cap = 0;

	for (j = job_begin; j < job_end; j++) {
		const struct airs_job *job = &jobs[j];
		struct oracle_ctx ctx;
		uint32_t r;

		r = oracle_init(&ctx, &job->params,
				work && job->work_size ? (uint8_t *)work + job->work_offset : NULL,
				job->work_size, job->identifier_base);
		if (init_results)
			init_results[j] = r;
		for (f = 0; f < job->n_frames; f++) {
			const uint8_t *s = (const uint8_t *)src + job->src_offset + f * job->src_frame_stride;
			uint32_t k = job->first_result + f;

			if (layout == AIRS_LAYOUT_SLOTS) {
				uint8_t *d = (uint8_t *)dst + job->dst_offset + f * job->dst_frame_stride;

				results[k] = oracle_compress(&ctx, d, job->dst_capacity, s, job->src_size,
							     job->dtype);
			} else {
				uint32_t want = job->dst_capacity;

				if (failed(want))
					want = 0;
				if (want > bounce_cap || !bounce) {
					free(bounce);
					bounce_cap = want > 64 ? want : 64;
					if (posix_memalign((void **)&bounce, 8, bounce_cap))
						return -1;
				}
				r = oracle_compress(&ctx, bounce, job->dst_capacity, s, job->src_size,
						    job->dtype);
				results[k] = r;
				out_offsets[k] = cursor;
				if (!failed(r)) {
					memcpy((uint8_t *)dst + cursor, bounce, r);
					cursor += r;
				}
			}
		}
	}
	if (layout == AIRS_LAYOUT_CONCAT && job_end > job_begin) {
		const struct airs_job *last = &jobs[job_end - 1];

		out_offsets[last->first_result + last->n_frames] = cursor;
	}
	free(bounce);
	return 0;
}

/* the same loop, streams hashed in scratch memory of the calling thread instead of kept (hash_jobs.h) */
#include "hash_jobs.h"
AIRS_DEFINE_HASH_JOBS(oracle_hash_jobs, oracle_run_jobs)
