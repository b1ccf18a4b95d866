/*
 * airs_device.cuh - device-side building blocks of the sm_100a compression path.
 *
 * Every function states which piece of the reference (AIRSPACE v0.6.0,
 * /root/reference) it stands in for.  Written for Blackwell: 32-bit integer
 * pipes only (no tensor work exists on this path), funnel shifts (SHF) for all
 * bit placement, PRMT for byte order, REDUX for small warp reductions.
 */
#ifndef AIRS_DEVICE_CUH
#define AIRS_DEVICE_CUH

#include <stdint.h>

#include "../../../include/airs_cuda.h"
#include "../../../include/cmp_errors.h"

#define AIRS_ERR(name) ((uint32_t)0 - (uint32_t)CMP_ERR_##name)

__host__ __device__ __forceinline__ bool airs_failed(uint32_t r)
{
	return r > (uint32_t)0 - (uint32_t)CMP_ERR_MAX_CODE; /* ref err_private.h:44-47 */
}

/* ------------------------------------------------------------------------
 * Encoder constants of one pass, derived once per pass by one thread.
 * Stands in for struct cmp_encoder + cmp_encoder_init (ref encoder.h:40-47,
 * encoder.c:185-224) and for the per-call constants of golomb_encode
 * (encoder.c:303-324).
 *
 * Golomb code of v with parameter g, L = floor(log2 g), cutoff = 2^(L+1) - g,
 * restated without the group-0 branch: with q' = floor((v + g - cutoff) / g)
 *      length = L + 1 + q'
 *      code   = (2^(L+1) << q') - 2^(L+1) + v - q' * g        (mod 2^32)
 * (q' = 0 reproduces group 0, q' = q + 1 the other groups.)
 * The division is a multiply-high by magic = floor((2^(32+L) - 1) / g) applied
 * to the dividend + 1 followed by a shift by L; exact while dividend * g <
 * 2^(32+L), i.e. for every dividend below 2^31 (tests/test_host_logic.py
 * checks all g and all reachable dividends).
 * ------------------------------------------------------------------------ */
struct EncConst {
	uint32_t type;    /* enum cmp_encoder_type */
	uint32_t g;
	uint32_t L;       /* floor(log2 g) */
	uint32_t outlier; /* derived outlier (goes into the header too) */
	uint32_t magic;   /* floor((2^(32+L) - 1) / g) */
	uint32_t bias;    /* g - cutoff + 1: dividend + 1 = v + bias */
	uint32_t two_l1;  /* 2^(L+1) */
};

__host__ __device__ __forceinline__ uint32_t airs_floor_log2(uint32_t v)
{
#ifdef __CUDA_ARCH__
	return 31u - (uint32_t)__clz((int)v);
#else
	uint32_t l = 0;
	while (v >>= 1)
		l++;
	return l;
#endif
}

/* derived outlier, 0 = invalid (ref encoder.c:63-110, 154-182, 205-216) */
__host__ __device__ inline uint32_t airs_derive_outlier(uint32_t type, uint32_t g, uint32_t user)
{
	if (g < 1u || g > 65535u)
		return 0;
	uint32_t L = airs_floor_log2(g);
	uint64_t cutoff = (2ull << L) - g;
	uint64_t limit = cutoff + (uint64_t)(31u - L) * g; /* first value with a >32-bit codeword */
	uint64_t o;
	if (type == CMP_ENCODER_GOLOMB_MULTI) {
		if (limit <= 8)
			return 0;
		limit -= 8; /* room for the 8 escape symbols */
		o = user;
	} else {
		o = cutoff + 16ull * g - 1;
	}
	return (uint32_t)(o < limit ? o : limit);
}

/* cmp_encoder_params_check (ref encoder.c:227-233) */
__host__ __device__ inline uint32_t airs_encoder_check(uint32_t type, uint32_t g, uint32_t user)
{
	if (type == CMP_ENCODER_UNCOMPRESSED)
		return 0;
	if (type != CMP_ENCODER_GOLOMB_ZERO && type != CMP_ENCODER_GOLOMB_MULTI)
		return AIRS_ERR(PARAMS_INVALID);
	return airs_derive_outlier(type, g, user) ? 0 : AIRS_ERR(PARAMS_INVALID);
}

__host__ __device__ inline void airs_enc_const(EncConst *e, uint32_t type, uint32_t g, uint32_t user)
{
	e->type = type;
	e->g = g;
	e->L = 0;
	e->outlier = 0;
	e->magic = 0;
	e->bias = 0;
	e->two_l1 = 0;
	if (type == CMP_ENCODER_UNCOMPRESSED)
		return;
	e->L = airs_floor_log2(g);
	e->outlier = airs_derive_outlier(type, g, user);
	e->two_l1 = 2u << e->L;
	e->bias = g - (e->two_l1 - g) + 1u;
	{ /* floor((2^(32+L) - 1) / g) by two 32-bit divisions: the dividend is (2^L - 1) 2^32 + (2^32 - 1) and
	   * 2^L - 1 < g < 2^16, so both partial dividends stay below g 2^16 */
		const uint32_t t1 = (((1u << e->L) - 1u) << 16) | 0xFFFFu;
		const uint32_t q1 = t1 / g, r1 = t1 - q1 * g;
		e->magic = (q1 << 16) + (((r1 << 16) | 0xFFFFu) / g);
	}
}

#ifdef __CUDACC__

/* zig-zag of a 16-bit residual held in the low half of r (upper half ignored):
 * ref map_to_unsigned, encoder.c:274-286 */
__device__ __forceinline__ uint32_t airs_zigzag16(uint32_t r)
{
	uint32_t sign; /* PRMT in sign-replicate mode: byte 1's msb over all four bytes */
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sign) : "r"(r), "r"(0u), "r"(0x9999u));
	return ((r << 1) ^ sign) & 0xFFFFu;
}

/* Golomb codeword of v (ref golomb_encode, encoder.c:303-324), branch free */
__device__ __forceinline__ void airs_golomb(const EncConst &e, uint32_t v, uint32_t &code, uint32_t &len)
{
	uint32_t q = __umulhi(v + e.bias, e.magic) >> e.L;
	code = (e.two_l1 << q) - e.two_l1 + v - q * e.g;
	len = e.L + 1u + q;
}

/*
 * One zig-zag mapped residual m -> up to two bit strings: (cw, cwlen) then (raw, rawlen).
 * ref cmp_encoder_encode_s16, encoder.c:335-376.  ENC is a compile-time copy of e.type.
 */
template <int ENC>
__device__ __forceinline__ void airs_encode_mapped(const EncConst &e, uint32_t m, uint32_t &cw, uint32_t &cwlen,
						   uint32_t &raw, uint32_t &rawlen)
{
	if (ENC == CMP_ENCODER_GOLOMB_ZERO) {
		uint32_t c, l;
		airs_golomb(e, m + 1u, c, l);
		bool esc = m >= e.outlier;
		cw = esc ? m : c; /* escape: codeword 0 (L+1 zero bits) then m raw in 16 bits */
		cwlen = esc ? e.L + 17u : l;
		raw = 0;
		rawlen = 0;
	} else {
		bool esc = m >= e.outlier;
		uint32_t d = m - e.outlier;
		uint32_t level = d < 4u ? 0u : (31u - (uint32_t)__clz((int)d)) >> 1;
		airs_golomb(e, esc ? e.outlier + level : m, cw, cwlen);
		raw = esc ? d : 0u;
		rawlen = esc ? 2u * level + 2u : 0u;
	}
}

/* one residual (low 16 bits of r16): ref cmp_encoder_encode_s16, encoder.c:327-378 */
template <int ENC>
__device__ __forceinline__ void airs_encode(const EncConst &e, uint32_t r16, uint32_t &cw, uint32_t &cwlen,
					    uint32_t &raw, uint32_t &rawlen)
{
	if (ENC == CMP_ENCODER_UNCOMPRESSED) {
		cw = r16 & 0xFFFFu;
		cwlen = 16;
		raw = 0;
		rawlen = 0;
	} else {
		airs_encode_mapped<ENC>(e, airs_zigzag16(r16), cw, cwlen, raw, rawlen);
	}
}

/* both samples of a word from big-endian to host order (AIRS_DTYPE_BE) */
__device__ __forceinline__ uint32_t airs_swap16x2(uint32_t w)
{
	return __byte_perm(w, 0, 0x2301);
}

__device__ __forceinline__ uint4 airs_swap16x8(uint4 v)
{
	return make_uint4(airs_swap16x2(v.x), airs_swap16x2(v.y), airs_swap16x2(v.z), airs_swap16x2(v.w));
}

/* one model update (ref update_model, cmp.c:120-142); SIGNED for i16 containers */
__device__ __forceinline__ uint32_t airs_model_update(uint32_t x, uint32_t m, uint32_t rate, bool is_signed)
{
	int32_t xv = is_signed ? (int32_t)(int16_t)x : (int32_t)x;
	int32_t mv = is_signed ? (int32_t)(int16_t)m : (int32_t)m;
	return (uint32_t)((mv * (int32_t)rate + xv * (16 - (int32_t)rate)) >> 4) & 0xFFFFu;
}

/* XXH32 pieces (xxHash 0.8.3 spec; ref header.c:137-163 uses XXH32_reset/update/digest) */
#define AIRS_XP1 0x9E3779B1u
#define AIRS_XP2 0x85EBCA77u
#define AIRS_XP3 0xC2B2AE3Du
#define AIRS_XP4 0x27D4EB2Fu
#define AIRS_XP5 0x165667B1u
#define AIRS_CHECKSUM_SEED 419764627u /* ref header_private.h:46 */

__device__ __forceinline__ uint32_t airs_rotl(uint32_t v, int r)
{
	return __funnelshift_l(v, v, r);
}

__device__ __forceinline__ uint32_t airs_xxh_round(uint32_t acc, uint32_t word)
{
	return airs_rotl(acc + word * AIRS_XP2, 13) * AIRS_XP1;
}

/* two native little-endian samples (s0 | s1 << 16) -> the LE32 lane word of their big-endian bytes */
__device__ __forceinline__ uint32_t airs_be_pair(uint32_t w)
{
	return __byte_perm(w, 0, 0x2301);
}

__device__ __forceinline__ uint32_t airs_bswap32(uint32_t w)
{
	return __byte_perm(w, 0, 0x0123);
}

#endif /* __CUDACC__ */
#endif /* AIRS_DEVICE_CUH */
