/*
 * cmp_shim.c - the cmp.h API (include/cmp.h) on top of the CUDA backend.
 *
 * Plain C host code, the "lib/cuda backend next to lib/compress" of the
 * reference tree: it keeps the reference's context struct, validation order and
 * error codes (ref lib/compress/cmp.c:53-209,396-472) and hands every
 * cmp_compress_*() call to the device through airs_cuda_compress_resume().
 * Nothing here encodes a sample: without a CUDA device the compress calls fail.
 *
 * Identifier bookkeeping (ref cmp.c:27-50,438-449): the timestamp callback is
 * host code, so the device numbers the identifiers it draws 0,1,..; afterwards
 * the shim calls the callback exactly that many times and patches the last
 * value into bytes 8..13 of the stream, which is the value the reference's
 * header would carry.
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../../include/cmp.h"
#include "../../../include/cmp_errors.h"
#include "../cuda/airs_private.h"

#define CMP_MAGIC 34021395u /* ref cmp.c:23 */
#define ERR(name) ((uint32_t)0 - (uint32_t)CMP_ERR_##name)

static int failed(uint32_t r)
{
	return r > (uint32_t)0 - (uint32_t)CMP_ERR_MAX_CODE;
}

/* ---- identifiers ------------------------------------------------------- */

static void counter_timestamp(uint32_t *coarse, uint16_t *fine)
{
	static uint64_t count; /* ref cmp.c:27-34: post-incremented, starts at 0 */

	*coarse = (uint32_t)(count >> 16);
	*fine = (uint16_t)count;
	count++;
}

static void (*timestamp_source)(uint32_t *, uint16_t *) = counter_timestamp;

void cmp_set_timestamp_func(void (*get_current_timestamp_func)(uint32_t *coarse, uint16_t *fine))
{
	timestamp_source = get_current_timestamp_func ? get_current_timestamp_func : counter_timestamp;
}

static uint64_t draw_identifier(void)
{
	uint32_t coarse = 0;
	uint16_t fine = 0;

	timestamp_source(&coarse, &fine);
	return ((uint64_t)coarse << 16) | fine;
}

/* ---- helpers without device work --------------------------------------- */

unsigned int cmp_is_error(uint32_t code)
{
	return failed(code);
}

uint32_t cmp_compress_bound(uint32_t packed_size)
{
	uint64_t samples, bound;

	if (packed_size > CMP_HDR_MAX_ORIGINAL_SIZE)
		return ERR(HDR_ORIGINAL_TOO_LARGE);
	samples = ((uint64_t)packed_size * 8 + 15) / 16;
	bound = (CMP_HDR_SIZE + 6) + CMP_CHECKSUM_SIZE + (samples * 48 + 7) / 8;
	if (bound > CMP_HDR_MAX_COMPRESSED_SIZE)
		return ERR(HDR_CMP_SIZE_TOO_LARGE);
	return (uint32_t)bound;
}

static uint32_t work_size_of(uint32_t preprocessing, uint32_t src_size, uint32_t *out)
{
	switch (preprocessing) {
	case CMP_PREPROCESS_NONE:
	case CMP_PREPROCESS_DIFF:
		*out = 0;
		return 0;
	case CMP_PREPROCESS_IWT:
	case CMP_PREPROCESS_MODEL:
		*out = (src_size + 1u) & ~1u;
		return 0;
	default:
		return ERR(PARAMS_INVALID);
	}
}

uint32_t cmp_cal_work_buf_size(const struct cmp_params *params, uint32_t src_size)
{
	uint32_t first = 0, second = 0, r;

	if (!params)
		return ERR(GENERIC);
	if (params->primary_preprocessing == CMP_PREPROCESS_MODEL)
		return ERR(PARAMS_INVALID);
	r = work_size_of(params->primary_preprocessing, src_size, &first);
	if (failed(r))
		return r;
	if (params->secondary_iterations) {
		r = work_size_of(params->secondary_preprocessing, src_size, &second);
		if (failed(r))
			return r;
	}
	return first > second ? first : second;
}

static unsigned int floor_log2(uint32_t v)
{
	unsigned int l = 0;

	while (v >>= 1)
		l++;
	return l;
}

/* parameter check of one encoder (ref encoder.c:185-233) */
static uint32_t encoder_check(uint32_t type, uint32_t g, uint32_t user_outlier)
{
	uint64_t cutoff, limit, outlier;
	unsigned int L;

	if (type == CMP_ENCODER_UNCOMPRESSED)
		return 0;
	if (type != CMP_ENCODER_GOLOMB_ZERO && type != CMP_ENCODER_GOLOMB_MULTI)
		return ERR(PARAMS_INVALID);
	if (g < 1 || g > 65535)
		return ERR(PARAMS_INVALID);
	L = floor_log2(g);
	cutoff = (2ull << L) - g;
	limit = cutoff + (uint64_t)(31 - L) * g;
	if (type == CMP_ENCODER_GOLOMB_MULTI) {
		if (limit <= 8)
			return ERR(PARAMS_INVALID);
		limit -= 8;
		outlier = user_outlier;
	} else {
		outlier = cutoff + 16ull * g - 1;
	}
	if (outlier > limit)
		outlier = limit;
	return outlier ? 0 : ERR(PARAMS_INVALID);
}

static int model_needed(const struct cmp_params *p)
{
	return p->secondary_preprocessing == CMP_PREPROCESS_MODEL && p->secondary_iterations != 0;
}

/* ---- context life cycle ------------------------------------------------ */

void cmp_deinitialise(struct cmp_context *ctx)
{
	if (ctx)
		memset(ctx, 0, sizeof(*ctx));
}

uint32_t cmp_reset(struct cmp_context *ctx)
{
	if (!ctx)
		return ERR(GENERIC);
	if (ctx->magic != CMP_MAGIC)
		return ERR(CONTEXT_INVALID);
	ctx->sequence_number = 0;
	ctx->identifier = draw_identifier();
	ctx->model_size = 0;
	return 0;
}

uint32_t cmp_initialise(struct cmp_context *ctx, const struct cmp_params *params, void *work_buf,
			uint32_t work_buf_size)
{
	uint32_t r, need;

	if (!ctx)
		return ERR(GENERIC);
	cmp_deinitialise(ctx);
	if (!params)
		return ERR(GENERIC);
	if (failed(work_buf_size))
		return ERR(GENERIC);
	if (params->secondary_iterations >= 256)
		return ERR(PARAMS_INVALID);
	r = encoder_check(params->primary_encoder_type, params->primary_encoder_param,
			  params->primary_encoder_outlier);
	if (failed(r))
		return r;
	if (params->secondary_iterations) {
		r = encoder_check(params->secondary_encoder_type, params->secondary_encoder_param,
				  params->secondary_encoder_outlier);
		if (failed(r))
			return r;
	}
	if (model_needed(params) && params->model_rate > 16)
		return ERR(PARAMS_INVALID);
	need = cmp_cal_work_buf_size(params, 2);
	if (failed(need))
		return need;
	if (need > 0) {
		if (!work_buf)
			return ERR(WORK_BUF_NULL);
		if (work_buf_size == 0)
			return ERR(WORK_BUF_TOO_SMALL);
		if ((uintptr_t)work_buf & 1u)
			return ERR(WORK_BUF_UNALIGNED);
	}
	ctx->params = *params;
	ctx->work_buf = work_buf;
	ctx->work_buf_size = work_buf_size;
	ctx->magic = CMP_MAGIC;
	return cmp_reset(ctx);
}

/* ---- compression: everything below the checks runs on the GPU ----------- */

static uint32_t compress_any(struct cmp_context *ctx, void *dst, uint32_t dst_capacity, const void *src,
			     uint32_t src_size, uint32_t dtype)
{
	struct airs_job job;
	struct airs_ctx_state st;
	uint32_t stride = dtype == AIRS_DTYPE_I16_IN_I32 ? 4u : 2u;
	uint32_t result = 0;
	uint64_t id = 0, k;
	int rc;

	/* container checks first (ref sample_reader.h:19-51), then the context (ref cmp.c:350-357) */
	if (!src)
		return ERR(SRC_NULL);
	if (src_size == 0 || src_size % stride)
		return ERR(SRC_SIZE_WRONG);
	if (!ctx)
		return ERR(GENERIC);
	if (ctx->magic != CMP_MAGIC)
		return ERR(CONTEXT_INVALID);
	if (failed(dst_capacity))
		return ERR(GENERIC);

	memset(&job, 0, sizeof(job));
	job.params = ctx->params;
	job.src_size = src_size;
	job.dst_capacity = dst_capacity;
	job.dst_offset = (uint64_t)((uintptr_t)dst & 7u);
	job.work_size = ctx->work_buf ? ctx->work_buf_size : 0;
	job.n_frames = 1;
	job.dtype = dtype;
	memset(&st, 0, sizeof(st));
	st.identifier = ctx->identifier;
	st.valid = 1;
	st.seq = ctx->sequence_number;
	st.model_size = ctx->model_size;

	rc = airs_cuda_compress_resume(&job, &st, src, dst, ctx->work_buf, model_needed(&ctx->params), &result);
	if (rc != AIRS_OK) {
		fprintf(stderr, "libcmp_b200: %s (this library has no CPU path)\n", airs_cuda_last_error());
		return ERR(GENERIC);
	}

	/* replay the identifier draws of the device state machine on the host callback */
	for (k = 0; k < st.counter; k++)
		id = draw_identifier();
	if (st.counter) {
		ctx->identifier = id;
		if (!failed(result)) {
			uint8_t be[6];
			int b;

			for (b = 0; b < 6; b++)
				be[b] = (uint8_t)(id >> (8 * (5 - b)));
			if (airs_cuda_patch_bytes(dst, be, CMP_HDR_OFFSET_IDENTIFIER, 6) != AIRS_OK)
				return ERR(GENERIC);
		}
	}
	ctx->sequence_number = (uint8_t)st.seq;
	ctx->model_size = st.model_size;
	return result;
}

uint32_t cmp_compress_u16(struct cmp_context *ctx, void *dst, uint32_t dst_capacity, const uint16_t *src,
			  uint32_t src_size)
{
	return compress_any(ctx, dst, dst_capacity, src, src_size, AIRS_DTYPE_U16);
}

uint32_t cmp_compress_i16(struct cmp_context *ctx, void *dst, uint32_t dst_capacity, const int16_t *src,
			  uint32_t src_size)
{
	return compress_any(ctx, dst, dst_capacity, src, src_size, AIRS_DTYPE_I16);
}

uint32_t cmp_compress_i16_in_i32(struct cmp_context *ctx, void *dst, uint32_t dst_capacity,
				 const int32_t *src, uint32_t src_size)
{
	return compress_any(ctx, dst, dst_capacity, src, src_size, AIRS_DTYPE_I16_IN_I32);
}

/* -------------------------------------------------------------------------
 * error code helpers of include/cmp_errors.h.  Same codes and, so that log
 * output stays comparable, the same message texts as the reference
 * (lib/common/cmp_errors.c:17-90), kept in a table.
 * ---------------------------------------------------------------------- */
enum cmp_error cmp_get_error_code(uint32_t code)
{
	if (code <= (uint32_t)0 - (uint32_t)CMP_ERR_MAX_CODE)
		return CMP_ERR_NO_ERROR;
	return (enum cmp_error)((uint32_t)0 - code);
}

static const struct {
	enum cmp_error code;
	const char *text;
} messages[] = {
	{ CMP_ERR_NO_ERROR, "No error detected" },
	{ CMP_ERR_GENERIC, "Error (generic)" },
	{ CMP_ERR_PARAMS_INVALID, "Invalid compression parameters" },
	{ CMP_ERR_DST_TOO_SMALL, "Destination buffer is too small to hold the content" },
	{ CMP_ERR_DST_NULL, "Destination buffer pointer is NULL" },
	{ CMP_ERR_DST_UNALIGNED, "Destination buffer pointer is unaligned" },
	{ CMP_ERR_SRC_SIZE_WRONG, "Source buffer size is invalid" },
	{ CMP_ERR_SRC_NULL, "Source buffer pointer is NULL" },
	{ CMP_ERR_SRC_SIZE_MISMATCH,
	  "Source data size changed using model preprocessing; not allowed until reset" },
	{ CMP_ERR_WORK_BUF_TOO_SMALL, "Work buffer is too small" },
	{ CMP_ERR_WORK_BUF_NULL, "Work buffer is NULL but required" },
	{ CMP_ERR_WORK_BUF_UNALIGNED, "Work buffer is unaligned" },
	{ CMP_ERR_HDR_CMP_SIZE_TOO_LARGE, "Compressed size exceeds header field limit" },
	{ CMP_ERR_HDR_ORIGINAL_TOO_LARGE, "Original size exceeds header field limit" },
	{ CMP_ERR_CONTEXT_INVALID, "Compression context uninitialised or corrupted" },
	{ CMP_ERR_INT_HDR, "Internal header processing error" },
	{ CMP_ERR_INT_ENCODER, "Internal data encoder error" },
	{ CMP_ERR_INT_BITSTREAM, "Internal bitstream writer error" },
};

const char *cmp_get_error_string(enum cmp_error code)
{
	unsigned int i;

	for (i = 0; i < sizeof(messages) / sizeof(messages[0]); i++)
		if (messages[i].code == code)
			return messages[i].text;
	return "Unspecified error code";
}

const char *cmp_get_error_message(uint32_t code)
{
	return cmp_get_error_string(cmp_get_error_code(code));
}
