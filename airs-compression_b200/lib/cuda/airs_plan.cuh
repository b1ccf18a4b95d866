/*
 * airs_plan.cuh - per-job plan: everything cmp_initialise (ref lib/compress/cmp.c:152-209)
 * and the checks at the head of compress_engine (cmp.c:228-294) decide from the
 * parameters alone, computed once per job by one thread of airs_plan_kernel so
 * that the encode kernel never runs serial per-job code.
 */
#ifndef AIRS_PLAN_CUH
#define AIRS_PLAN_CUH

#include "airs_device.cuh"

/* flags */
#define AIRS_PF_VALID        1u  /* cmp_initialise succeeded */
#define AIRS_PF_MODEL        2u  /* model_is_needed (cmp.c:145-149) */
#define AIRS_PF_CHECKSUM     4u
#define AIRS_PF_FALLBACK_OK  8u  /* fallback enabled and capacity >= raw size (cmp.c:363) */
#define AIRS_PF_SIGNED      16u  /* i16 containers: model update sign-extends (cmp.c:132-142) */
#define AIRS_PF_SMALL       32u  /* one short frame without model: a warp of airs_small_kernel encodes it */
#define AIRS_PF_BE          64u  /* AIRS_DTYPE_BE: samples big-endian in memory */

/* 128 bytes, read by the encode kernel with one coalesced 32-lane load */
struct alignas(16) JobPlan {
	EncConst enc[2];      /* [0] primary, [1] secondary encoder constants            (56 B) */
	uint32_t pre[2];      /* preprocessing of primary / secondary passes                    */
	uint32_t pre_err[2];  /* work-buffer error of that preprocessing or 0 (pre.c:321-393)   */
	uint32_t init_result; /* what cmp_initialise returns                                    */
	uint32_t flags;
	uint32_t n;           /* samples per frame                                              */
	uint32_t frame_err;   /* error every frame of this job returns before any state change  */
	uint32_t raw_size;    /* 16 + 2n (+4)                                                   */
	uint32_t cap_eff;     /* bytes a pass may write: raw_size when fallback can happen      */
	uint32_t trip;        /* cumulative bits at which the reference's writer gives up       */
	uint32_t sec_iter;
	uint32_t rate;
	uint32_t model_err;   /* WORK_BUF_TOO_SMALL of the model buffer (cmp.c:250-254) or 0    */
	uint32_t orig_err;    /* HDR_ORIGINAL_TOO_LARGE (header.c:34) or 0                      */
	uint32_t pad[3];
};
static_assert(sizeof(JobPlan) == 128, "JobPlan is read as 32 words");

#ifdef __CUDACC__

__device__ inline uint32_t airs_pre_work_size(uint32_t pre, uint32_t src_size, uint32_t &out)
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

__device__ inline uint32_t airs_work_buf_size(const cmp_params &p, uint32_t src_size) /* ref cmp.c:77-103 */
{
	uint32_t a = 0, b = 0, r;

	if (p.primary_preprocessing == CMP_PREPROCESS_MODEL)
		return AIRS_ERR(PARAMS_INVALID);
	r = airs_pre_work_size(p.primary_preprocessing, src_size, a);
	if (airs_failed(r))
		return r;
	if (p.secondary_iterations) {
		r = airs_pre_work_size(p.secondary_preprocessing, src_size, b);
		if (airs_failed(r))
			return r;
	}
	return a > b ? a : b;
}

/* cmp_initialise (ref cmp.c:152-209) without the final cmp_reset */
__device__ inline uint32_t airs_validate(const airs_job &j, const uint8_t *work)
{
	const cmp_params &p = j.params;
	uint32_t r;

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
	if (p.secondary_preprocessing == CMP_PREPROCESS_MODEL && p.secondary_iterations != 0 && p.model_rate > 16u)
		return AIRS_ERR(PARAMS_INVALID);
	r = airs_work_buf_size(p, 2);
	if (airs_failed(r))
		return r;
	if (r > 0) {
		if (!work)
			return AIRS_ERR(WORK_BUF_NULL);
		if (j.work_size == 0)
			return AIRS_ERR(WORK_BUF_TOO_SMALL);
		if ((uintptr_t)work & 1u)
			return AIRS_ERR(WORK_BUF_UNALIGNED);
	}
	return 0;
}

__device__ inline uint32_t airs_pre_err(uint32_t pre, const uint8_t *work, uint32_t work_size, uint32_t packed)
{
	if (pre != CMP_PREPROCESS_IWT && pre != CMP_PREPROCESS_MODEL)
		return 0; /* ref preprocess.c:321-335,382-393 */
	if (!work)
		return AIRS_ERR(WORK_BUF_NULL);
	if (work_size < ((packed + 1u) & ~1u))
		return AIRS_ERR(WORK_BUF_TOO_SMALL);
	if ((uintptr_t)work & 1u)
		return AIRS_ERR(WORK_BUF_UNALIGNED);
	return 0;
}

__device__ inline void airs_make_plan(JobPlan &pl, const airs_job &j, const uint8_t *src_base, uint8_t *work_base, bool be_batch)
{
	const cmp_params &p = j.params;
	const uint8_t *work = (work_base && j.work_size) ? work_base + j.work_offset : nullptr;
	const uint32_t stride = j.dtype == AIRS_DTYPE_I16_IN_I32 ? 4u : 2u;
	const bool dtype_ok = j.dtype <= AIRS_DTYPE_U16 || (be_batch && (j.dtype == AIRS_DTYPE_I16_BE || j.dtype == AIRS_DTYPE_U16_BE));

	memset(&pl, 0, sizeof(pl));
	pl.init_result = airs_validate(j, work);
	if (!airs_failed(pl.init_result))
		pl.flags |= AIRS_PF_VALID;
	/* the checks of cmp_compress_* / cmp_compress_generic that come before the engine
	 * (ref sample_reader.h:19-51, cmp.c:350-357), in their order */
	if (!src_base)
		pl.frame_err = AIRS_ERR(SRC_NULL);
	else if (j.src_size == 0 || !dtype_ok || j.src_size % stride)
		pl.frame_err = AIRS_ERR(SRC_SIZE_WRONG);
	else if (!(pl.flags & AIRS_PF_VALID))
		pl.frame_err = AIRS_ERR(CONTEXT_INVALID);
	else if (airs_failed(j.dst_capacity))
		pl.frame_err = AIRS_ERR(GENERIC);
	if (pl.frame_err)
		return;

	const uint32_t n = j.src_size / stride, packed = n * 2u;
	pl.n = n;
	airs_enc_const(&pl.enc[0], p.primary_encoder_type, p.primary_encoder_param, p.primary_encoder_outlier);
	pl.pre[0] = p.primary_preprocessing;
	pl.pre_err[0] = airs_pre_err(pl.pre[0], work, j.work_size, packed);
	if (p.secondary_iterations) {
		airs_enc_const(&pl.enc[1], p.secondary_encoder_type, p.secondary_encoder_param,
			       p.secondary_encoder_outlier);
		pl.pre[1] = p.secondary_preprocessing;
		pl.pre_err[1] = airs_pre_err(pl.pre[1], work, j.work_size, packed);
	}
	if (p.secondary_preprocessing == CMP_PREPROCESS_MODEL && p.secondary_iterations != 0) {
		pl.flags |= AIRS_PF_MODEL;
		if (j.work_size < packed)
			pl.model_err = AIRS_ERR(WORK_BUF_TOO_SMALL);
	}
	if (p.checksum_enabled)
		pl.flags |= AIRS_PF_CHECKSUM;
	if ((j.dtype & 3u) != AIRS_DTYPE_U16)
		pl.flags |= AIRS_PF_SIGNED;
	if (j.dtype & AIRS_DTYPE_BE)
		pl.flags |= AIRS_PF_BE;
	pl.raw_size = CMP_HDR_SIZE + packed + (p.checksum_enabled ? 4u : 0u);
	if (p.uncompressed_fallback_enabled && j.dst_capacity >= pl.raw_size)
		pl.flags |= AIRS_PF_FALLBACK_OK;
	pl.cap_eff = (pl.flags & AIRS_PF_FALLBACK_OK) ? pl.raw_size : j.dst_capacity;
	uint64_t trip = 64ull * ((uint64_t)pl.cap_eff / 8 + 1);
	pl.trip = trip > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)trip;
	pl.sec_iter = p.secondary_iterations;
	pl.rate = p.model_rate;
	if ((uint64_t)n * 2u > CMP_HDR_MAX_ORIGINAL_SIZE)
		pl.orig_err = AIRS_ERR(HDR_ORIGINAL_TOO_LARGE);
}

#endif /* __CUDACC__ */
#endif /* AIRS_PLAN_CUH */
