/*
 * cmp.h - AIRSPACE compression API, served by the B200 (sm_100a) backend.
 *
 * This is the drop-in boundary: same structs, enums, macros and the same
 * eleven functions as the reference's lib/cmp.h:64-137,154-344 (AIRSPACE
 * v0.6.0), so a program written against the reference library links against
 * libcmp_b200.so unchanged.  Differences that a caller can observe:
 *
 *  - src / dst / work_buf may be HOST pointers (staged over PCIe by the host
 *    shim) or DEVICE pointers (used in place, no copy);
 *    the shim asks cudaPointerGetAttributes() which one it got.
 *  - there is no CPU fallback: without a usable CUDA device every compress
 *    call fails with CMP_ERR_GENERIC after printing the reason to stderr.
 *
 * For many independent chunks use the batched driver in airs_cuda.h instead of
 * calling cmp_compress_*() in a loop.
 */
#ifndef CMP_H
#define CMP_H

#include <stdint.h>

#include "cmp_header.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CMP_QUOTE(str)            #str
#define CMP_EXPAND_AND_QUOTE(str) CMP_QUOTE(str)

/* format version written into every header (ref: cmp.h:39-53) */
#define CMP_VERSION_MAJOR   0
#define CMP_VERSION_MINOR   6
#define CMP_VERSION_RELEASE 0
#define CMP_VERSION_NUMBER \
	(CMP_VERSION_MAJOR * 100 * 100 + CMP_VERSION_MINOR * 100 + CMP_VERSION_RELEASE)
#define CMP_VERSION_STRING \
	CMP_EXPAND_AND_QUOTE(CMP_VERSION_MAJOR.CMP_VERSION_MINOR.CMP_VERSION_RELEASE)

/* ref: cmp.h:64-71 */
enum cmp_preprocessing {
	CMP_PREPROCESS_NONE,  /* samples are coded as they are */
	CMP_PREPROCESS_DIFF,  /* x[i] - x[i-1] (mod 2^16), x[0] kept */
	CMP_PREPROCESS_IWT,   /* multi-level 5/3 integer wavelet transform */
	CMP_PREPROCESS_MODEL  /* x[i] - model[i]; secondary passes only */
};

/* ref: cmp.h:78-82 */
enum cmp_encoder_type {
	CMP_ENCODER_UNCOMPRESSED, /* 16 raw bits per sample */
	CMP_ENCODER_GOLOMB_ZERO,  /* Golomb code, codeword 0 escapes to a raw sample */
	CMP_ENCODER_GOLOMB_MULTI  /* Golomb code, several escape symbols + raw offset */
};

/* ref: cmp.h:94-116 (44 bytes, field order is ABI) */
struct cmp_params {
	enum cmp_preprocessing primary_preprocessing;
	enum cmp_encoder_type primary_encoder_type;
	uint32_t primary_encoder_param;   /* Golomb parameter, 1..65535 */
	uint32_t primary_encoder_outlier; /* GOLOMB_MULTI only */

	uint32_t secondary_iterations;    /* passes after a primary one, 0..255 */
	enum cmp_preprocessing secondary_preprocessing;
	enum cmp_encoder_type secondary_encoder_type;
	uint32_t secondary_encoder_param;
	uint32_t secondary_encoder_outlier;
	uint32_t model_rate;              /* 0..16, weight of the old model in 1/16 */

	uint8_t checksum_enabled;              /* append XXH32 of the samples */
	uint8_t uncompressed_fallback_enabled; /* store raw if that is smaller */
};

/* ref: cmp.h:129-137 (80 bytes; opaque to callers, layout is ABI) */
struct cmp_context {
	uint32_t magic;
	struct cmp_params params;
	void *work_buf;
	uint32_t work_buf_size;
	uint32_t model_size;
	uint64_t identifier;
	uint8_t sequence_number;
};

/* Install the source of the 48-bit identifiers (coarse<<16 | fine); NULL
 * restores the built-in counter.                          ref: cmp.h:154 */
void cmp_set_timestamp_func(void (*get_current_timestamp_func)(uint32_t *coarse, uint16_t *fine));

/* non-zero when a result is an error                      ref: cmp.h:166 */
unsigned int cmp_is_error(uint32_t code);

/* worst-case stream size for packed_size bytes of samples ref: cmp.h:184 */
uint32_t cmp_compress_bound(uint32_t packed_size);

/* stream size of the uncompressed representation          ref: cmp.h:212-215 */
#define CMP_UNCOMPRESSED_BOUND(packed_size)                                                  \
	((packed_size) <= (CMP_HDR_MAX_COMPRESSED_SIZE - CMP_HDR_SIZE - CMP_CHECKSUM_SIZE) ? \
		 (CMP_HDR_SIZE + (packed_size) + CMP_CHECKSUM_SIZE) :                        \
		 SIZE_MAX)

/* bytes of working buffer the parameter set needs         ref: cmp.h:234 */
uint32_t cmp_cal_work_buf_size(const struct cmp_params *params, uint32_t src_size);

/* validate params, bind the work buffer, start a fresh identifier
 *                                                         ref: cmp.h:261 */
uint32_t cmp_initialise(struct cmp_context *ctx, const struct cmp_params *params, void *work_buf,
			uint32_t work_buf_size);

/* compress one buffer; dst must be 8-byte aligned; returns the stream size
 *                                                         ref: cmp.h:284,299,309 */
uint32_t cmp_compress_i16(struct cmp_context *ctx, void *dst, uint32_t dst_capacity,
			  const int16_t *src, uint32_t src_size);
uint32_t cmp_compress_i16_in_i32(struct cmp_context *ctx, void *dst, uint32_t dst_capacity,
				 const int32_t *src, uint32_t src_size);
uint32_t cmp_compress_u16(struct cmp_context *ctx, void *dst, uint32_t dst_capacity,
			  const uint16_t *src, uint32_t src_size);

/* next call is a primary pass with a new identifier       ref: cmp.h:328 */
uint32_t cmp_reset(struct cmp_context *ctx);

/* forget the context (zeroes it; frees nothing of the caller's)
 *                                                         ref: cmp.h:344 */
void cmp_deinitialise(struct cmp_context *ctx);

#ifdef __cplusplus
}
#endif

#endif /* CMP_H */
