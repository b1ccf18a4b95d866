/*
 * cmp_header.h - layout constants of the compressed-stream header.
 *
 * Same macro names and values as the reference's lib/cmp_header.h:19-62 (they are part of the API).
 * A stream is: 16-byte base header [+ 6-byte extension] | code bits
 * (MSB first) | zero padding to a byte | [4-byte big-endian XXH32].
 *
 *  byte 0-1   1-bit version flag (=1) + 15-bit version id
 *  byte 2-4   compressed size (whole stream, bytes)
 *  byte 5-7   original size (packed 16-bit samples, bytes)
 *  byte 8-13  identifier (48 bit)
 *  byte 14    sequence number
 *  byte 15    preprocessing<<4 | checksum_enabled<<3 | encoder_type
 *  -- extension, present unless (NONE, UNCOMPRESSED) --
 *  byte 16    model rate
 *  byte 17-18 encoder parameter (Golomb parameter)
 *  byte 19-21 encoder outlier (as derived by the encoder)
 */
#ifndef CMP_HEADER_H
#define CMP_HEADER_H

/* field by field, in stream order: where it starts (byte) and how wide it is (bits) */
#define CMP_HDR_OFFSET_VERSION 0
#define CMP_HDR_BITS_VERSION_FLAG 1
#define CMP_HDR_BITS_VERSION_ID 15
#define CMP_HDR_BITS_VERSION (CMP_HDR_BITS_VERSION_FLAG + CMP_HDR_BITS_VERSION_ID) /* 16 */

#define CMP_HDR_OFFSET_COMPRESSED_SIZE 2
#define CMP_HDR_BITS_COMPRESSED_SIZE 24
#define CMP_HDR_MAX_COMPRESSED_SIZE ((1ULL << CMP_HDR_BITS_COMPRESSED_SIZE) - 1) /* 16 MiB - 1 */

#define CMP_HDR_OFFSET_ORIGINAL_SIZE 5
#define CMP_HDR_BITS_ORIGINAL_SIZE 24
#define CMP_HDR_MAX_ORIGINAL_SIZE ((1ULL << CMP_HDR_BITS_ORIGINAL_SIZE) - 1)

#define CMP_HDR_OFFSET_IDENTIFIER 8
#define CMP_HDR_BITS_IDENTIFIER 48

#define CMP_HDR_OFFSET_SEQUENCE_NUMBER 14
#define CMP_HDR_BITS_SEQUENCE_NUMBER 8

#define CMP_HDR_OFFSET_METHOD 15
#define CMP_HDR_BITS_METHOD_PREPROCESSING 4    /* bits 7-4 of the method byte */
#define CMP_HDR_BITS_METHOD_CHECKSUM_ENABLED 1 /* bit 3 */
#define CMP_HDR_BITS_METHOD_ENCODER_TYPE 3     /* bits 2-0 */
#define CMP_HDR_BITS_METHOD \
	(CMP_HDR_BITS_METHOD_PREPROCESSING + CMP_HDR_BITS_METHOD_CHECKSUM_ENABLED + CMP_HDR_BITS_METHOD_ENCODER_TYPE)

/* the base header in bytes: 16 */
#define CMP_HDR_SIZE \
	((CMP_HDR_BITS_VERSION + CMP_HDR_BITS_COMPRESSED_SIZE + CMP_HDR_BITS_ORIGINAL_SIZE + \
	  CMP_HDR_BITS_IDENTIFIER + CMP_HDR_BITS_SEQUENCE_NUMBER + CMP_HDR_BITS_METHOD) / 8)

/* the trailer */
#define CMP_CHECKSUM_SIZE sizeof(uint32_t)

#endif /* CMP_HEADER_H */
