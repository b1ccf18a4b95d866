/*
 * airs_cuda_decode.h - C-ABI of the batched AIRSPACE stream decoder (sm_100a).
 *
 * The reference has no decoder (programs/airspacecli.c:421-423 refuses -d); the
 * only consumer of a stream it ships is cmp_hdr_deserialize (lib/common/header.c:
 * 89-134), used by its tests to look at headers.  This interface is therefore
 * additive.  Its semantics are DEFINED by the
 * reference's encoder: decoding the streams the reference produces for a
 * sequence of cmp_compress_*() calls on one context returns the samples that
 * were passed in, for every parameter set cmp_initialise accepts
 * (lib/compress/cmp.c:213-393, preprocess.c:268-411, encoder.c:303-378, header
 * layout lib/cmp_header.h:26-58).
 *
 * Plain C: pointers, sizes, fixed-width integers.
 */
#ifndef AIRS_CUDA_DECODE_H
#define AIRS_CUDA_DECODE_H

#include <stddef.h>
#include <stdint.h>

#include "airs_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* result codes the decoder adds to enum cmp_error (all below CMP_ERR_MAX_CODE,
 * so cmp_is_error() recognises them); header trouble is CMP_ERR_INT_HDR, a
 * short destination CMP_ERR_DST_TOO_SMALL, a stream that does not fit its slot
 * CMP_ERR_SRC_SIZE_WRONG */
#define AIRS_DEC_ERR_CHECKSUM 110 /* the XXH32 trailer does not match the decoded samples */
#define AIRS_DEC_ERR_NO_MODEL 111 /* a MODEL frame without a decodable first frame in front of it */
#define AIRS_DEC_ERR_CORRUPT  112 /* the code words do not end where the header says the stream ends */

/*
 * One compressed sequence: the streams of one compression context, in the
 * order they were produced (a model frame needs its predecessors).  Independent
 * chunks are jobs with n_frames == 1.  56 bytes.
 */
struct airs_dec_job {
	uint64_t src_offset;       /* stream of frame 0, bytes from the stream base */
	uint64_t src_frame_stride; /* bytes between consecutive stream slots; 0: the streams lie back to back */
	uint64_t dst_offset;       /* samples of frame 0, bytes from the sample base; a multiple of the container size */
	uint64_t dst_frame_stride; /* bytes between the sample arrays of consecutive frames */
	uint32_t src_size;         /* bytes available per stream slot (back to back: for the whole sequence) */
	uint32_t dst_capacity;     /* bytes available per decoded frame */
	uint32_t n_frames;
	uint32_t dtype;            /* AIRS_DTYPE_*: container of the decoded samples.  The stream does not record
				    * it; it decides between sign and zero extension in the model update
				    * (ref cmp.c:132-142) and the width of a sample in dst */
	uint32_t first_result;     /* index of frame 0 in results[] / info[] */
	uint32_t reserved;         /* must be 0 */
};

/* the header of a stream, field by field (ref struct cmp_hdr, lib/common/header_private.h:58-76,
 * as filled in by cmp_hdr_deserialize, header.c:89-134).  32 bytes. */
struct airs_frame_info {
	uint64_t identifier;
	uint32_t compressed_size;
	uint32_t original_size;
	uint32_t encoder_outlier;
	uint16_t version;          /* version_flag << 15 | version_id */
	uint16_t encoder_param;
	uint8_t sequence_number;
	uint8_t preprocessing;
	uint8_t checksum_enabled;
	uint8_t encoder_type;
	uint8_t model_rate;
	uint8_t header_size;       /* 16 or 22; 0 if the header could not be read */
	uint8_t reserved[2];
};

/* Device-resident batch: every pointer is a device pointer. */
struct airs_dec_batch {
	const void *src;                 /* stream base */
	void *dst;                       /* sample base */
	const struct airs_dec_job *jobs; /* n_jobs descriptors */
	uint32_t *results;               /* n_results: bytes written for the frame or (uint32_t)-error */
	struct airs_frame_info *info;    /* n_results parsed headers, or NULL */
	void *scratch;                   /* airs_cuda_decode_scratch_size() bytes, 16-byte aligned */
	uint32_t n_jobs;
	uint32_t n_results;              /* total number of frames */
};

/* Bytes of device scratch a decode batch needs (one record per frame). */
size_t airs_cuda_decode_scratch_size(uint32_t n_jobs, uint32_t n_results);

/*
 * Decode a device-resident batch on `stream` (a cudaStream_t passed as void *,
 * NULL = default stream).  Asynchronous.  Samples are written in the job's
 * container: 2 bytes each for AIRS_DTYPE_U16 / AIRS_DTYPE_I16, a sign-extended
 * 32-bit word for AIRS_DTYPE_I16_IN_I32 (the encoder only ever saw the low
 * halves, ref sample_reader.h:63-72).  Returns AIRS_OK or AIRS_E_*; there is no
 * CPU fallback.
 */
int airs_cuda_decompress_batch(const struct airs_dec_batch *batch, void *stream);

#ifdef __cplusplus
}
#endif

#endif /* AIRS_CUDA_DECODE_H */
