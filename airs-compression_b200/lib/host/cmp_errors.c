/*
 * cmp_errors.c - error code helpers of the cmp.h API (include/cmp_errors.h).
 * Same codes and, so that log output stays comparable, the same message
 * strings as the reference's lib/common/cmp_errors.c:17-90.
 */
#include <stdint.h>

#include "../../../include/cmp_errors.h"

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
