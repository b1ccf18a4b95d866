/*
 * cmp_errors.h - error codes of the AIRSPACE compression API (B200 backend).
 *
 * Interface-compatible with the reference's lib/cmp_errors.h:28-105: same enum
 * names and values, same three helper functions.  A result of any cmp_*
 * function is either a size (or 0) or (uint32_t)-code; test it with
 * cmp_is_error().
 */
#ifndef CMP_ERRORS_H
#define CMP_ERRORS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum cmp_error {
	CMP_ERR_NO_ERROR = 0,

	CMP_ERR_GENERIC = 1,
	CMP_ERR_PARAMS_INVALID = 10,

	CMP_ERR_DST_TOO_SMALL = 30,
	CMP_ERR_DST_NULL = 31,
	CMP_ERR_DST_UNALIGNED = 32,

	CMP_ERR_SRC_SIZE_WRONG = 40,
	CMP_ERR_SRC_NULL = 41,
	CMP_ERR_SRC_SIZE_MISMATCH = 42,

	CMP_ERR_WORK_BUF_TOO_SMALL = 50,
	CMP_ERR_WORK_BUF_NULL = 51,
	CMP_ERR_WORK_BUF_UNALIGNED = 52,

	CMP_ERR_HDR_CMP_SIZE_TOO_LARGE = 60,
	CMP_ERR_HDR_ORIGINAL_TOO_LARGE = 61,

	CMP_ERR_CONTEXT_INVALID = 70,

	CMP_ERR_INT_HDR = 100,
	CMP_ERR_INT_ENCODER = 101,
	CMP_ERR_INT_BITSTREAM = 102,

	CMP_ERR_MAX_CODE = 128 /* upper limit marker, never returned */
};

/* result -> error code (CMP_ERR_NO_ERROR for a size)   ref: cmp_errors.h:74 */
enum cmp_error cmp_get_error_code(uint32_t code);
/* result -> static description string                  ref: cmp_errors.h:90 */
const char *cmp_get_error_message(uint32_t code);
/* error code -> static description string              ref: cmp_errors.h:105 */
const char *cmp_get_error_string(enum cmp_error code);

#ifdef __cplusplus
}
#endif

#endif /* CMP_ERRORS_H */
