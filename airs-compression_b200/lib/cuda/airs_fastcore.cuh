/*
 * airs_fastcore.cuh - the warp-level encoder shared by airs_fast_kernel (one warp per short job, airs_fast.cu)
 * and airs_tile_kernel (the tiles of long frames spread over CTAs, airs_tile.cu).
 *
 * A "unit" is 512 consecutive samples of one warp: lane l holds the pieces 2 l and 2 l + 1 of 8 samples ("rows"
 * 0 and 1), 16 consecutive samples.  The strings of one staging instruction then lie a lane's whole bit count
 * apart (3 to 6 words for most data) instead of one piece's (1.5 words: neighbouring lanes in the same word,
 * 2.8 wavefronts per shared-memory reduction in ncu) - the price is two 128-bit loads per lane that each
 * touch only half of every sector (the other half is the other load's: L1 hits).  Per unit:
 *   packed 16x2 residuals (none / diff, VIADD.16x2 + PRMT) -> packed zig-zag -> per sample the Golomb code
 *   word ARITHMETICALLY: quotient by one multiply-high (airs_fast.cuh), code word by one shift, one
 *   multiply-add and one three-input add, escapes by selects (ref cmp_encoder_encode_s16, encoder.c:327-378;
 *   golomb_encode, encoder.c:303-324) -> GOLOMB_ZERO: the code words of a pair of samples merged into one
 *   string of at most 62 bits; GOLOMB_MULTI: one string of at most 48 bits per sample -> one shuffle scan
 *   over the packed bit counts of both rows -> every string shifted into place by three funnel shifts and
 *   OR-ed into MSB-first staging words in shared memory (reductions; ref bitstream_add_bits32,
 *   bitstream_writer.h:124-158).
 * Any data goes through this one path: escapes cost nothing extra.  The kernels are bound by instruction
 * issue and by the integer ALU pipe (LOP3 / SHF / PRMT / IADD3 / SEL share one pipe of 64 lanes per clock and
 * SM, tools/micro/pipes.cu), so the code is written to need few instructions per sample.
 */
#ifndef AIRS_FASTCORE_CUH
#define AIRS_FASTCORE_CUH

#include "airs_device.cuh"
#include "airs_fast.cuh"

namespace fastcore {

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kRows = 2;               /* pieces per lane and unit */
constexpr uint32_t kUnitPieces = 32 * kRows;
constexpr uint32_t kUnit = 8 * kUnitPieces; /* 512 samples */
constexpr uint32_t kUnitMaxBits = kUnit * 48;

/* the piece of row j of a lane, counted inside the unit */
__device__ __forceinline__ uint32_t unit_piece(uint32_t lane, uint32_t j)
{
	return kRows * lane + j;
}

/* AIRS_BOUNDS_CHECK builds: the shared-memory window of the warp's staging words, and a counter of the
 * strings that would have left it (airs_fast_bounds_violations() reads and clears it) */
#ifdef AIRS_BOUNDS_CHECK
static __device__ unsigned int airs_bounds_violations; /* one per translation unit */
struct Dbg {
	uint32_t lo, hi;
};
#else
struct Dbg {
};
#endif

/* per-job encoder constants, uniform over the warp (registers) */
struct FK {
	uint32_t neg_g;   /* -g */
	uint32_t M;       /* airs_fast_magic(g) */
	uint32_t bias_q;  /* quotient q' = umulhi(value + bias_q, M), see airs_golomb() in airs_device.cuh */
	uint32_t two_l1;  /* 2^(L+1) */
	uint32_t c0;      /* code word = (two_l1 << q') + q' * -g + value + c0 */
	uint32_t L1;      /* L + 1: length = L1 + q' */
	uint32_t outlier;
	uint32_t esc_len; /* GOLOMB_ZERO: L + 17 */
};

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t s) /* 0 for s >= 32 */
{
	uint32_t r;
	asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
	return r;
}

/* Golomb code word of `value` (ref golomb_encode, encoder.c:303-324) without the group-0 branch */
__device__ __forceinline__ void golomb_k(const FK &k, uint32_t value, uint32_t &code, uint32_t &len)
{
	const uint32_t q = __umulhi(value + k.bias_q, k.M);
	code = shl_clamp(k.two_l1, q) + (q * k.neg_g + value) + k.c0;
	len = q + k.L1;
}

/* one zig-zag mapped residual m -> one string (hi:lo, n bits); ref cmp_encoder_encode_s16, encoder.c:327-378 */
template <bool MULTI>
__device__ __forceinline__ void encode_k(const FK &k, uint32_t m, uint32_t &hi, uint32_t &lo, uint32_t &n)
{
	if (!MULTI) {
		uint32_t c, l;
		golomb_k(k, m, c, l); /* value m + 1: the + 1 lives in bias_q and c0 */
		const bool esc = m >= k.outlier;
		lo = esc ? m : c; /* escape: L + 1 zero bits, then m in 16 bits */
		n = esc ? k.esc_len : l;
		hi = 0;
	} else {
		/* escape: Golomb code word of outlier + level, then d = m - outlier in 2 level + 2 bits, level =
		 * d < 4 ? 0 : floor(log2 d) / 2 (encoder.c:356-372).  level <= d, so min(m, outlier + level) is
		 * the value to encode whether the sample escapes or not (d wraps to a huge number if not) */
		const bool esc = m >= k.outlier;
		const uint32_t d = m - k.outlier;
		uint32_t msb;
		asm("bfind.u32 %0, %1;" : "=r"(msb) : "r"(d | 1u));
		const uint32_t level = msb >> 1;
		uint32_t c, l;
		golomb_k(k, min(m, k.outlier + level), c, l);
		const uint32_t rl = esc ? 2u * level + 2u : 0u; /* raw bits */
		const uint32_t pw = 1u << rl;
		lo = c * pw + (esc ? d : 0u);
		hi = __umulhi(c, pw);
		n = l + rl;
	}
}

/* OR a bit string of len <= 64 bits (hi:lo, right aligned) into the staging words.  ne = -(absolute
 * bit address where the string starts: 8 * shared-memory byte address + bit), updated to the start of
 * the next string.  Three funnel shifts and three reductions whatever the length. */
__device__ __forceinline__ void put(const Dbg &dbg, int32_t &ne, uint32_t hi, uint32_t lo, uint32_t len)
{
	ne -= (int32_t)len;
	const uint32_t s = (uint32_t)ne;               /* wrap-mode shifts use s & 31 = the bits free behind the string */
	const uint32_t a = ~(uint32_t)(ne >> 3) & ~3u; /* shared-memory address of the word holding its last bit */
#ifdef AIRS_BOUNDS_CHECK
	/* development builds (compute-sanitizer is not available on the GPU pool): every word a string touches
	 * must lie inside the staging words of the warp, pad included */
	if (a - 8u < dbg.lo || a >= dbg.hi) {
		atomicAdd(&airs_bounds_violations, 1u);
		return;
	}
#endif
	asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(__funnelshift_l(0u, lo, s)) : "memory");
	asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a - 4u), "r"(__funnelshift_l(lo, hi, s)) : "memory");
	asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a - 8u), "r"(__funnelshift_l(hi, 0u, s)) : "memory");
}

/* the constants of one encoder (g, L = floor(log2 g), derived outlier, magic = airs_fast_magic(g)) */
__device__ __forceinline__ FK make_fk(bool multi, uint32_t g, uint32_t L, uint32_t outlier, uint32_t magic)
{
	FK k;
	const uint32_t cutoff = (2u << L) - g;
	k.M = magic;
	k.neg_g = 0u - g;
	k.two_l1 = 2u << L;
	k.L1 = L + 1u;
	k.outlier = outlier;
	k.esc_len = L + 17u;
	/* dividend of the quotient: value + g - cutoff; GOLOMB_ZERO codes m + 1 (encoder.c:344) */
	k.bias_q = g - cutoff + (multi ? 0u : 1u) + (g == 1u ? 1u : 0u);
	k.c0 = (multi ? 0u : 1u) - k.two_l1;
	return k;
}

/* the incomplete last piece of a frame: cnt (1..7) samples from sample `first` on, the rest 0.  Called by
 * one lane once per job: kept out of line */
static __device__ __noinline__ uint4 load_partial_piece(const uint16_t *s16, uint32_t first, uint32_t cnt, bool be = false)
{
	uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll 1
	for (uint32_t i = 0; i < cnt; i++) {
		uint32_t x = __ldg(s16 + first + i);
		if (be)
			x = ((x << 8) | (x >> 8)) & 0xFFFFu;
		const uint32_t v = x << (16u * (i & 1u));
		if (i < 2u)
			w0 |= v;
		else if (i < 4u)
			w1 |= v;
		else if (i < 6u)
			w2 |= v;
		else
			w3 |= v;
	}
	return make_uint4(w0, w1, w2, w3);
}

/* the whole piece p of a frame (16-byte aligned 16-bit samples, `be`: big-endian in memory: AIRS_DTYPE_BE) */
template <bool BE>
__device__ __forceinline__ uint4 load_piece(const uint4 *src4, uint32_t p, bool be)
{
	const uint4 v = __ldg(src4 + p);
	return (BE && be) ? airs_swap16x8(v) : v;
}

/* the strings of one unit, in registers: GOLOMB_ZERO one string per pair of samples (hi, lo, bits);
 * GOLOMB_MULTI one string of at most 48 bits per sample (lo, and hi | bits << 16) */
template <bool MULTI>
struct UnitStrings {
	static constexpr int kStr = MULTI ? 8 : 4;
	uint32_t a[kRows][MULTI ? 8 : 4], lo[kRows][MULTI ? 8 : 4], n[kRows][MULTI ? 1 : 4];
};

/* preprocessing of a unit: what the residual of a sample is taken against */
constexpr int kPreNone = 0, kPreDiff = 1, kPreModel = 2;

/* model of a secondary pass (ref cmp.c:120-142,304-311): rate 16 keeps the model, rate 0 replaces it by the samples,
 * anything between is (rate * model + (16 - rate) * sample) / 16 on two lanes at once (IDP.2A: 16-bit x 8-bit
 * products, weights 16 rate and 16 (16 - rate), the quotient in bytes 1-2 of each sum) */
struct ModelK {
	uint32_t rate, wdp;
	bool is_signed; /* i16 containers: the average is taken of sign-extended values */
};

__device__ __forceinline__ ModelK make_model_k(uint32_t rate, bool is_signed)
{
	ModelK mk;
	mk.rate = rate;
	mk.wdp = (rate << 4) | ((16u - rate) << 12);
	mk.is_signed = is_signed;
	return mk;
}

template <bool SIGNED>
__device__ __forceinline__ uint32_t model_update_pair(uint32_t x, uint32_t m, uint32_t wdp)
{
	const uint32_t lo = __byte_perm(m, x, 0x5410), hi = __byte_perm(m, x, 0x7632); /* (m, x) of the low / high lane */
	uint32_t tl, th;
	if (SIGNED) {
		asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(tl) : "r"(lo), "r"(wdp), "r"(0u));
		asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(th) : "r"(hi), "r"(wdp), "r"(0u));
	} else {
		asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(tl) : "r"(lo), "r"(wdp), "r"(0u));
		asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(th) : "r"(hi), "r"(wdp), "r"(0u));
	}
	return __byte_perm(tl, th, 0x6521);
}

/*
 * One unit, first half: the two pieces x[0], x[1] of this lane (rows 0 and 1) -> strings.
 * PRE = kPreDiff: front is lane 0's word in front of its row-0 piece (upper half = the sample before the unit,
 * 0 at the start of a frame: the first "difference" is the sample itself, ref preprocess.c:284-290); updated
 * for the next unit of the same warp.  PRE = kPreModel: m[0], m[1] are the model words of the two pieces; the
 * residual is taken against them and they are replaced by the updated model (ref cmp.c:304-311).
 * nv[j]: valid samples of the lane's piece in row j (RAGGED units only).
 * Returns the bits of the lane's rows: row 0 | row 1 << 16.
 */
template <bool MULTI, int PRE, bool RAGGED>
__device__ __forceinline__ uint32_t unit_codes_pre(const FK &k, const uint4 (&x)[kRows], uint4 (&m)[kRows], const ModelK &mk, uint32_t &front,
						   const uint32_t (&nv)[kRows], uint32_t lane, UnitStrings<MULTI> &s)
{
	uint32_t row_bits[kRows];

#pragma unroll
	for (uint32_t j = 0; j < kRows; j++) {
		const uint32_t w[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
		uint32_t z[4];
		if (PRE == kPreDiff) {
			/* t = ~r = ~w + predecessor, per 16-bit half; zig-zag of r = ((t << 1) | 1) ^ sign(t)
			 * (ref map_to_unsigned, encoder.c:274-286) */
			uint32_t prev;
			if (j == 0) { /* the last word of the lane in front */
				const uint32_t up = __shfl_sync(kFull, x[kRows - 1].w, (lane - 1u) & 31u);
				prev = lane ? up : front;
				front = up; /* lane 0: lane 31's last word, in front of the next unit */
			} else {
				prev = x[j - 1].w;
			}
#pragma unroll
			for (int i = 0; i < 4; i++) {
				const uint32_t t = __vadd2(~w[i], __byte_perm(prev, w[i], 0x5432));
				uint32_t sign;
				asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sign) : "r"(t), "r"(0u), "r"(0xBB99u));
				z[i] = ((t << 1) | 0x00010001u) ^ sign;
				prev = w[i];
			}
		} else if (PRE == kPreModel) {
			uint32_t mw[4] = {m[j].x, m[j].y, m[j].z, m[j].w};
#pragma unroll
			for (int i = 0; i < 4; i++) {
				const uint32_t t = __vadd2(~w[i], mw[i]); /* ~(sample - model) */
				uint32_t sign;
				asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sign) : "r"(t), "r"(0u), "r"(0xBB99u));
				z[i] = ((t << 1) | 0x00010001u) ^ sign;
				if (mk.rate == 0u)
					mw[i] = w[i];
				else if (mk.rate < 16u)
					mw[i] = mk.is_signed ? model_update_pair<true>(w[i], mw[i], mk.wdp) : model_update_pair<false>(w[i], mw[i], mk.wdp);
			}
			m[j] = make_uint4(mw[0], mw[1], mw[2], mw[3]);
		} else {
#pragma unroll
			for (int i = 0; i < 4; i++) {
				uint32_t sign;
				asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sign) : "r"(w[i]), "r"(0u), "r"(0xBB99u));
				z[i] = ((w[i] << 1) & 0xFFFEFFFEu) ^ sign;
			}
		}
		uint32_t bits = 0;
#pragma unroll
		for (uint32_t i = 0; i < 4; i++) {
			uint32_t h0, l0, n0, h1, l1, n1;
			encode_k<MULTI>(k, z[i] & 0xFFFFu, h0, l0, n0);
			encode_k<MULTI>(k, z[i] >> 16, h1, l1, n1);
			if (RAGGED) {
				if (2u * i >= nv[j])
					h0 = l0 = n0 = 0;
				if (2u * i + 1u >= nv[j])
					h1 = l1 = n1 = 0;
			}
			if (!MULTI) {
				s.a[j][i] = __funnelshift_l(l0, 0u, n1);
				s.lo[j][i] = (l0 << n1) | l1;
				s.n[j][i] = n0 + n1;
			} else {
				s.a[j][2 * i] = h0 | (n0 << 16);
				s.lo[j][2 * i] = l0;
				s.a[j][2 * i + 1] = h1 | (n1 << 16);
				s.lo[j][2 * i + 1] = l1;
			}
			bits += n0 + n1;
		}
		row_bits[j] = bits;
	}
	return row_bits[0] | (row_bits[1] << 16);
}

/* ... without a model */
template <bool MULTI, bool DIFF, bool RAGGED>
__device__ __forceinline__ uint32_t unit_codes(const FK &k, const uint4 (&x)[kRows], uint32_t &front, const uint32_t (&nv)[kRows],
					       uint32_t lane, UnitStrings<MULTI> &s)
{
	uint4 none[kRows];
	const ModelK mk = {0u, 0u, false};
	return unit_codes_pre<MULTI, DIFF ? kPreDiff : kPreNone, RAGGED>(k, x, none, mk, front, nv, lane, s);
}

/* the bits of a lane's rows together */
__device__ __forceinline__ uint32_t lane_bits(uint32_t packed)
{
	return (packed & 0xFFFFu) + (packed >> 16);
}

/* inclusive warp scan of the lanes' bit counts (stream order is lane by lane).  The shuffle's "source lane
 * exists" predicate feeds the adds directly. */
__device__ __forceinline__ uint32_t unit_scan(uint32_t b)
{
	uint32_t incl = b;
#define AIRS_SCAN_STEP(d_)                                                                    \
	asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t0;\n\t"                                  \
		     "shfl.sync.up.b32 t0|p, %0, " #d_ ", 0, 0xffffffff;\n\t"                     \
		     "@p add.u32 %0, %0, t0;\n\t}"                                               \
		     : "+r"(incl))
	AIRS_SCAN_STEP(1);
	AIRS_SCAN_STEP(2);
	AIRS_SCAN_STEP(4);
	AIRS_SCAN_STEP(8);
	AIRS_SCAN_STEP(16);
#undef AIRS_SCAN_STEP
	return incl;
}

/* One unit, second half: the strings of row j go to the absolute bit address pos[j] (see put()) */
template <bool MULTI>
__device__ __forceinline__ void unit_put(const Dbg &dbg, const UnitStrings<MULTI> &s, const uint32_t (&pos)[kRows])
{
#pragma unroll
	for (uint32_t j = 0; j < kRows; j++) {
		int32_t ne = -(int32_t)pos[j];
#pragma unroll
		for (int i = 0; i < UnitStrings<MULTI>::kStr; i++) {
			if (!MULTI)
				put(dbg, ne, s.a[j][i], s.lo[j][i], s.n[j][i]);
			else
				put(dbg, ne, s.a[j][i] & 0xFFFFu, s.lo[j][i], s.a[j][i] >> 16);
		}
	}
}

/* a whole unit of a warp that owns its staging words: strings, scan, staging at abs_bit; returns the unit's bits */
template <bool MULTI, int PRE, bool RAGGED>
__device__ __forceinline__ uint32_t encode_unit_pre(const Dbg &dbg, const FK &k, const uint4 (&x)[kRows], uint4 (&m)[kRows], const ModelK &mk,
						    uint32_t &front, const uint32_t (&nv)[kRows], uint32_t lane, uint32_t abs_bit)
{
	UnitStrings<MULTI> s;
	const uint32_t b = unit_codes_pre<MULTI, PRE, RAGGED>(k, x, m, mk, front, nv, lane, s);
	const uint32_t mine = lane_bits(b);
	const uint32_t incl = unit_scan(mine);
	const uint32_t tot = __shfl_sync(kFull, incl, 31), at = abs_bit + incl - mine;
	const uint32_t pos[kRows] = {at, at + (b & 0xFFFFu)};
	unit_put<MULTI>(dbg, s, pos);
	return tot;
}

template <bool MULTI, bool DIFF, bool RAGGED>
__device__ __forceinline__ uint32_t encode_unit(const Dbg &dbg, const FK &k, const uint4 (&x)[kRows], uint32_t &front,
						const uint32_t (&nv)[kRows], uint32_t lane, uint32_t abs_bit)
{
	UnitStrings<MULTI> s;
	const uint32_t b = unit_codes<MULTI, DIFF, RAGGED>(k, x, front, nv, lane, s);
	const uint32_t mine = lane_bits(b);
	const uint32_t incl = unit_scan(mine);
	const uint32_t tot = __shfl_sync(kFull, incl, 31), at = abs_bit + incl - mine;
	const uint32_t pos[kRows] = {at, at + (b & 0xFFFFu)};
	unit_put<MULTI>(dbg, s, pos);
	return tot;
}

} /* namespace fastcore */

#endif /* AIRS_FASTCORE_CUH */
