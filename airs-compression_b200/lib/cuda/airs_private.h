/*
 * airs_private.h - interface between the host shim (lib/host/cmp_shim.c, plain
 * C) and the CUDA side (lib/cuda/airs_cuda_api.cu).  Not installed.
 */
#ifndef AIRS_PRIVATE_H
#define AIRS_PRIVATE_H

#include <stdint.h>

#include "../../../include/airs_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* the mutable part of struct cmp_context (ref cmp.h:129-137) as the device sees it */
struct airs_ctx_state {
	uint64_t identifier;
	uint64_t counter;   /* in: 0; out: number of identifiers drawn by the device state machine */
	uint32_t valid;
	uint32_t seq;
	uint32_t model_size;
	uint32_t reserved;
};

/*
 * One cmp_compress_*() call on an already initialised context.
 *
 * job: n_frames must be 1; src_offset/dst_offset/work_offset are ignored except
 *      that (dst_offset & 7) is added to the device-side destination so that the
 *      device sees the caller's alignment.
 * src / dst / work: host or device pointers (cudaPointerGetAttributes decides);
 *      host memory is staged through cached device buffers.  With
 *      work_is_state != 0 the work buffer is copied in before and out after the
 *      call (it holds the model).
 * Returns AIRS_OK or AIRS_E_*; *result is the value cmp_compress_*() returns.
 * Identifiers drawn by the device are numbered 0,1,.. (state->counter says how
 * many); the caller replaces bytes 8..13 of the stream with the real one.
 */
int airs_cuda_compress_resume(const struct airs_job *job, struct airs_ctx_state *state, const void *src,
			      void *dst, void *work, int work_is_state, uint32_t *result);

/* overwrite the 6 identifier bytes of a stream that may live on the device */
int airs_cuda_patch_bytes(void *dst, const uint8_t *bytes, uint32_t offset, uint32_t count);

#ifdef __cplusplus
}
#endif

#endif
