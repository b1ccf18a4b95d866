/*
 * ref_driver.c - batch loop of include/airs_cuda.h over the UNMODIFIED
 * reference library.  Compiled together with the reference's own sources
 * (taken where they lie under /root/reference, never copied) into
 * oracle/_ref/libcmp_ref.so by oracle/Makefile.
 *
 * TEST INFRASTRUCTURE ONLY: the arbiter for the oracle restatement and for the
 * CUDA path, and the "reference" CPU baseline of bench.py.
 */
#include <stdlib.h>
#include <string.h>

#include "cmp.h"        /* the reference's lib/cmp.h (found first via -I) */
#include "cmp_errors.h"
#include "../include/airs_cuda.h" /* struct airs_job; its cmp.h include is guarded out */
#include "hash_jobs.h"

/* one counter per thread so that jobs can be spread over threads */
static __thread uint64_t tls_counter;

static void counter_timestamp(uint32_t *coarse, uint16_t *fine)
{
	*coarse = (uint32_t)(tls_counter >> 16);
	*fine = (uint16_t)tls_counter;
	tls_counter++;
}

void ref_driver_install(void)
{
	cmp_set_timestamp_func(counter_timestamp);
}

void ref_driver_set_counter(uint64_t v)
{
	tls_counter = v;
}

uint64_t ref_driver_get_counter(void)
{
	return tls_counter;
}

static uint32_t call(struct cmp_context *ctx, void *d, uint32_t cap, const void *s, uint32_t size,
		     uint32_t dtype)
{
	switch (dtype) {
	case AIRS_DTYPE_I16:
		return cmp_compress_i16(ctx, d, cap, s, size);
	case AIRS_DTYPE_I16_IN_I32:
		return cmp_compress_i16_in_i32(ctx, d, cap, s, size);
	case AIRS_DTYPE_U16:
		return cmp_compress_u16(ctx, d, cap, s, size);
	default:
		return (uint32_t)0 - (uint32_t)CMP_ERR_SRC_SIZE_WRONG;
	}
}

int ref_run_jobs(const void *src, void *dst, void *work, const struct airs_job *jobs,
		 uint32_t job_begin, uint32_t job_end, uint32_t layout, uint32_t *results,
		 uint32_t *init_results, uint64_t *out_offsets)
{
	uint64_t cursor = 0;
	uint8_t *bounce = NULL;
	uint32_t bounce_cap = 0;
	uint32_t j, f;

	cmp_set_timestamp_func(counter_timestamp);
	for (j = job_begin; j < job_end; j++) {
		const struct airs_job *job = &jobs[j];
		struct cmp_context ctx;
		uint32_t r;

		tls_counter = job->identifier_base;
		r = cmp_initialise(&ctx, &job->params,
				   work && job->work_size ? (uint8_t *)work + job->work_offset : NULL,
				   job->work_size);
		if (init_results)
			init_results[j] = r;
		for (f = 0; f < job->n_frames; f++) {
			const uint8_t *s = (const uint8_t *)src + job->src_offset + f * job->src_frame_stride;
			uint32_t k = job->first_result + f;

			if (layout == AIRS_LAYOUT_SLOTS) {
				uint8_t *d = (uint8_t *)dst + job->dst_offset + f * job->dst_frame_stride;

				results[k] = call(&ctx, d, job->dst_capacity, s, job->src_size, job->dtype);
			} else {
				uint32_t want = cmp_is_error(job->dst_capacity) ? 0 : job->dst_capacity;

				if (want > bounce_cap || !bounce) {
					free(bounce);
					bounce_cap = want > 64 ? want : 64;
					if (posix_memalign((void **)&bounce, 8, bounce_cap))
						return -1;
				}
				r = call(&ctx, bounce, job->dst_capacity, s, job->src_size, job->dtype);
				results[k] = r;
				out_offsets[k] = cursor;
				if (!cmp_is_error(r)) {
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
AIRS_DEFINE_HASH_JOBS(ref_hash_jobs, ref_run_jobs)
