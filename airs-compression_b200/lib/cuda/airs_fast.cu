/*
 * airs_fast.cu - airs_fast_kernel: one WARP per short single-frame job (a chunk of up to
 * 32768 samples without model: BASELINE config 3 and the small-chunk cuts of config 4).
 *
 * The kernel is bound by the integer ALU pipe (LOP3 / SHF / PRMT / IADD3 / SEL share one pipe
 * of 64 lanes per clock and SM, tools/micro/pipes.cu), so everything here is written to need
 * few of those instructions per sample, and to be warp-synchronous: a warp owns its job, its
 * staging words and its position in the stream; there is no block barrier and no code that
 * runs on one thread while the others wait.
 *
 * Per unit of 512 samples ("rows" 0 and 1: lane l holds pieces l and 32 + l of 8 samples, so
 * both 128-bit loads of the warp are coalesced; the next unit's loads are in flight while one
 * is encoded):
 *   packed 16x2 residuals (none / diff, VIADD.16x2 + PRMT) -> packed zig-zag -> per sample
 *   the Golomb code word ARITHMETICALLY: quotient by one multiply-high (airs_fast.cuh), code
 *   word by one shift, one multiply-add and one three-input add, escape by two selects (ref
 *   cmp_encoder_encode_s16, encoder.c:327-378; golomb_encode, encoder.c:303-324) -> the code
 *   words of a pair of samples merged into one string of at most 64 bits -> one shuffle scan
 *   over the packed bit counts of both rows -> every string shifted into place by three funnel
 *   shifts and OR-ed into the warp's MSB-first staging words (shared-memory reductions; ref
 *   bitstream_add_bits32, bitstream_writer.h:124-158) -> complete 16-byte groups leave as
 *   byte-swapped coalesced 128-bit stores.
 * Any data goes through this one path - escapes cost nothing extra, which is what BASELINE
 * config 3 (escape heavy, mixed parameters) asks for.  A batch whose jobs mostly share one
 * encoder additionally gets the pair-table arm (airs_fast_table.cuh).
 *
 * Header (ref cmp_hdr_serialize, header.c:24-67) and the uncompressed fallback (cmp.c:363-392)
 * are handled by the same warp; XXH32 trailers by airs_checksum_kernel behind it.
 */
#include <cuda_runtime.h>

#include "airs_device.cuh"
#include "airs_fast.cuh"
#include "airs_launch.h"

namespace {

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kFWarps = AIRS_FAST_THREADS / 32;
constexpr uint32_t kRows = 2;                        /* pieces per lane and unit */
constexpr uint32_t kUnitPieces = 32 * kRows;
constexpr uint32_t kUnit = 8 * kUnitPieces;          /* 512 samples */
/* The staging words are drained when more than one unit at 48 bits per sample is staged (for most data: every
 * few units, so that the fixed cost of a drain is shared); the next unit then still fits, and so do the 22
 * header bytes behind up to 8 bytes of alignment in front of the first one */
constexpr uint32_t kUnitMaxBits = kUnit * 48;
constexpr uint32_t kStgWords = 2 * kUnitMaxBits / 32 + 8;

/* AIRS_BOUNDS_CHECK builds: the shared-memory window of the warp's staging words, and a counter of the
 * strings that would have left it (airs_fast_bounds_violations() reads and clears it) */
#ifdef AIRS_BOUNDS_CHECK
__device__ unsigned int airs_bounds_violations;
struct Dbg {
	uint32_t lo, hi;
};
#else
struct Dbg {
};
#endif

/* the staging words of one warp: MSB-first 32-bit words of the stream under construction, all
 * zero when idle; the pad in front absorbs the zeros that strings ending in word 0 or 1 OR
 * below the area */
struct FastWarp {
	alignas(16) uint32_t pad[4];
	uint32_t stg[kStgWords];
};

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

/* where the stream of the job stands: word 0 of the staging area is 32-bit word gw0 (a multiple of 4)
 * of the 16-byte aligned space that starts at base = dst - (dst & 15) and holds sbits bits (< 128 behind a
 * drain); [lo, hi) is the byte window of that space the job may write */
struct Out {
	uint8_t *base;
	uint32_t lo, hi;
	uint32_t gw0, sbits;
};

/* complete 16-byte groups of the staging area leave as coalesced stores; the partial group moves to the front */
__device__ __forceinline__ void drain(FastWarp &ws, Out &o, uint32_t lane)
{
	const uint32_t staged = o.sbits, nvec = staged >> 7, b0 = o.gw0 * 4u;
	uint4 *stg4 = reinterpret_cast<uint4 *>(ws.stg);

	__syncwarp();
	if (b0 >= o.lo && b0 + 16u * nvec <= o.hi) {
		uint4 *out = reinterpret_cast<uint4 *>(o.base + b0);
		for (uint32_t v = lane; v < nvec; v += 32u) {
			const uint4 q = stg4[v];
			stg4[v] = make_uint4(0, 0, 0, 0);
			out[v] = make_uint4(airs_bswap32(q.x), airs_bswap32(q.y), airs_bswap32(q.z), airs_bswap32(q.w));
		}
	} else { /* an edge of the window: the stream starts at dst % 16 == 8, or the capacity ends here */
		for (uint32_t v = lane; v < nvec; v += 32u) {
			const uint32_t b = b0 + 16u * v;
			const uint4 q = stg4[v];
			if (b >= o.lo && b + 16u <= o.hi) {
				*reinterpret_cast<uint4 *>(o.base + b) =
					make_uint4(airs_bswap32(q.x), airs_bswap32(q.y), airs_bswap32(q.z), airs_bswap32(q.w));
			} else {
				const uint8_t *s8 = reinterpret_cast<const uint8_t *>(stg4 + v);
#pragma unroll 1
				for (uint32_t k = 0; k < 16u; k++)
					if (b + k >= o.lo && b + k < o.hi)
						o.base[b + k] = s8[k ^ 3u];
			}
			stg4[v] = make_uint4(0, 0, 0, 0);
		}
	}
	__syncwarp();
	if (lane == 0 && nvec) {
		const uint4 carry = stg4[nvec];
		stg4[nvec] = make_uint4(0, 0, 0, 0);
		stg4[0] = carry;
	}
	__syncwarp();
	o.gw0 += nvec * 4u;
	o.sbits = staged & 127u;
}

/* the incomplete last piece of a frame: cnt (1..7) samples from sample `first` on, the rest 0.  Called by
 * one lane once per job: kept out of line */
__device__ __noinline__ uint4 load_partial_piece(const uint16_t *s16, uint32_t first, uint32_t cnt)
{
	uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll 1
	for (uint32_t i = 0; i < cnt; i++) {
		const uint32_t v = (uint32_t)__ldg(s16 + first + i) << (16u * (i & 1u));
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

/*
 * One unit: the two pieces x[0], x[1] of this lane (rows 0 and 1) -> strings -> staging words.
 * front: lane 0's word in front of its row-0 piece (upper half = the sample before the unit, 0 at the
 * start of a frame: the first "difference" is the sample itself, ref preprocess.c:284-290); updated for
 * the next unit.  nv[j]: valid samples of the lane's piece in row j (RAGGED units only).
 * Returns the bits of the unit.
 */
template <bool MULTI, bool DIFF, bool RAGGED>
__device__ __forceinline__ uint32_t encode_unit(const Dbg &dbg, const FK &k, const uint4 (&x)[kRows], uint32_t &front, const uint32_t (&nv)[kRows],
						uint32_t lane, uint32_t abs_bit)
{
	/* GOLOMB_ZERO: one string per pair of samples (two code words of at most 31 bits): hi, lo, bits.
	 * GOLOMB_MULTI: one string of at most 48 bits per sample: lo, and hi | bits << 16 */
	constexpr int kStr = MULTI ? 8 : 4;
	uint32_t s_a[kRows][kStr], s_lo[kRows][kStr], s_n[kRows][MULTI ? 1 : 4], row_bits[kRows];

#pragma unroll
	for (uint32_t j = 0; j < kRows; j++) {
		const uint32_t w[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
		uint32_t z[4];
		if (DIFF) {
			/* t = ~r = ~w + predecessor, per 16-bit half; zig-zag of r = ((t << 1) | 1) ^ sign(t)
			 * (ref map_to_unsigned, encoder.c:274-286) */
			const uint32_t up = __shfl_sync(kFull, w[3], (lane - 1u) & 31u);
			uint32_t prev = lane ? up : front;
			front = up; /* lane 0: lane 31's last word, in front of lane 0's piece of the next row */
#pragma unroll
			for (int i = 0; i < 4; i++) {
				const uint32_t t = __vadd2(~w[i], __byte_perm(prev, w[i], 0x5432));
				uint32_t sign;
				asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sign) : "r"(t), "r"(0u), "r"(0xBB99u));
				z[i] = ((t << 1) | 0x00010001u) ^ sign;
				prev = w[i];
			}
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
				s_a[j][i] = __funnelshift_l(l0, 0u, n1);
				s_lo[j][i] = (l0 << n1) | l1;
				s_n[j][i] = n0 + n1;
			} else {
				s_a[j][2 * i] = h0 | (n0 << 16);
				s_lo[j][2 * i] = l0;
				s_a[j][2 * i + 1] = h1 | (n1 << 16);
				s_lo[j][2 * i + 1] = l1;
			}
			bits += n0 + n1;
		}
		row_bits[j] = bits;
	}

	/* one scan for both rows: stream order is row 0 of all lanes, then row 1 of all lanes.  The
	 * shuffle's "source lane exists" predicate feeds the adds directly. */
	const uint32_t b = row_bits[0] | (row_bits[1] << 16);
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
	const uint32_t tot = __shfl_sync(kFull, incl, 31), excl = incl - b;
	const uint32_t tot0 = tot & 0xFFFFu;
	const uint32_t pos[kRows] = {abs_bit + (excl & 0xFFFFu), abs_bit + tot0 + (excl >> 16)};

#pragma unroll
	for (uint32_t j = 0; j < kRows; j++) {
		int32_t ne = -(int32_t)pos[j];
#pragma unroll
		for (int i = 0; i < kStr; i++) {
			if (!MULTI)
				put(dbg, ne, s_a[j][i], s_lo[j][i], s_n[j][i]);
			else
				put(dbg, ne, s_a[j][i] & 0xFFFFu, s_lo[j][i], s_a[j][i] >> 16);
		}
	}
	return tot0 + (tot >> 16);
}

/* the pass of one job through its units */
template <bool MULTI, bool DIFF>
__device__ __forceinline__ void encode_units(const Dbg &dbg, const FK &k, FastWarp &ws, Out &o, const uint8_t *src, uint32_t n, uint32_t lane)
{
	const uint32_t n_whole = n / 8u;         /* complete pieces */
	const uint32_t n_full_units = n / kUnit; /* units whose 64 pieces are all complete */
	const uint32_t stg_bit = 8u * (uint32_t)__cvta_generic_to_shared(ws.stg);
	const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
	const uint4 zero4 = make_uint4(0, 0, 0, 0);
	const uint32_t full_nv[kRows] = {8u, 8u};
	uint32_t front = 0;
	uint4 x[kRows], nx[kRows];

	/* the complete pieces of the next unit are requested while one unit is encoded */
#pragma unroll
	for (uint32_t j = 0; j < kRows; j++)
		nx[j] = 32u * j + lane < n_whole ? __ldg(src4 + 32u * j + lane) : zero4;
	uint32_t u = 0;
	for (; u < n_full_units; u++) {
#pragma unroll
		for (uint32_t j = 0; j < kRows; j++)
			x[j] = nx[j];
		const uint32_t p1 = (u + 1u) * kUnitPieces + lane;
#pragma unroll
		for (uint32_t j = 0; j < kRows; j++)
			nx[j] = p1 + 32u * j < n_whole ? __ldg(src4 + p1 + 32u * j) : zero4;
		o.sbits += encode_unit<MULTI, DIFF, false>(dbg, k, x, front, full_nv, lane, stg_bit + o.sbits);
		if (o.sbits > kUnitMaxBits)
			drain(ws, o, lane);
	}
	if (u * kUnit < n) { /* the ragged last unit */
		uint32_t nv[kRows];
#pragma unroll
		for (uint32_t j = 0; j < kRows; j++) {
			const uint32_t p = u * kUnitPieces + 32u * j + lane;
			nv[j] = 8u * p >= n ? 0u : min(8u, n - 8u * p);
			if (nv[j] != 0u && nv[j] != 8u)
				nx[j] = load_partial_piece(reinterpret_cast<const uint16_t *>(src), 8u * p, nv[j]);
		}
		o.sbits += encode_unit<MULTI, DIFF, true>(dbg, k, nx, front, nv, lane, stg_bit + o.sbits);
	}
	drain(ws, o, lane);
}

} /* namespace */

__global__ void __launch_bounds__(AIRS_FAST_THREADS, AIRS_FAST_CTAS_PER_SM) airs_fast_kernel(AirsLaunch b)
{
	__shared__ FastWarp wsh[kFWarps];
	const uint32_t lane = threadIdx.x & 31u;
	FastWarp &ws = wsh[threadIdx.x >> 5];

	if (b.gate && (*b.gate != 0u) != (b.gate_want != 0u)) /* two-phase CONCAT: not the phase that runs */
		return;
	const uint32_t n_fast = b.ticket[3]; /* entries of fast_jobs, written by airs_plan_kernel */
	if (n_fast == 0u)
		return;
	for (uint32_t w = lane; w < 4u + kStgWords; w += 32u)
		ws.pad[w] = 0;
	__syncwarp();
	Dbg dbg;
#ifdef AIRS_BOUNDS_CHECK
	dbg.lo = (uint32_t)__cvta_generic_to_shared(&ws);
	dbg.hi = dbg.lo + (uint32_t)sizeof(FastWarp);
#endif

	/* tickets are drawn two jobs ahead and the record of the next job one job ahead, so that neither
	 * the atomic's round trip nor the record's load is waited for */
	const FastJob *recs = reinterpret_cast<const FastJob *>(b.fast_jobs);
	uint32_t t = 0, t1 = 0, t2 = 0;
	if (lane == 0) {
		t = atomicAdd(&b.ticket[1], 1u);
		t1 = atomicAdd(&b.ticket[1], 1u);
	}
	t = __shfl_sync(kFull, t, 0);
	t1 = __shfl_sync(kFull, t1, 0);
	uint32_t rec = (t < n_fast && lane < 16u) ? reinterpret_cast<const uint32_t *>(recs + t)[lane] : 0u;

	while (t < n_fast) {
		if (lane == 0)
			t2 = atomicAdd(&b.ticket[1], 1u);
		const uint32_t nrec = (t1 < n_fast && lane < 16u) ? __ldg(reinterpret_cast<const uint32_t *>(recs + t1) + lane) : 0u;

#define AIRS_REC(i) __shfl_sync(kFull, rec, (i))
		const uint8_t *src = reinterpret_cast<const uint8_t *>((uintptr_t)(AIRS_REC(0) | (uint64_t)AIRS_REC(1) << 32));
		uint8_t *dst = reinterpret_cast<uint8_t *>((uintptr_t)(AIRS_REC(2) | (uint64_t)AIRS_REC(3) << 32));
		const uint32_t id_lo = AIRS_REC(4), id_hi = AIRS_REC(5);
		const uint32_t n = AIRS_REC(6), cap_eff = AIRS_REC(7), flags = AIRS_REC(8), first_result = AIRS_REC(9);
		const uint32_t g = AIRS_REC(10), outlier = AIRS_REC(11), L = AIRS_REC(13);
		FK k;
		k.M = AIRS_REC(12);
#undef AIRS_REC
		const bool multi = (flags & AIRS_FJ_MULTI) != 0u;
		const uint32_t cutoff = (2u << L) - g;
		k.neg_g = 0u - g;
		k.two_l1 = 2u << L;
		k.L1 = L + 1u;
		k.outlier = outlier;
		k.esc_len = L + 17u;
		/* dividend of the quotient: value + g - cutoff; GOLOMB_ZERO codes m + 1 (encoder.c:344) */
		k.bias_q = g - cutoff + (multi ? 0u : 1u) + (g == 1u ? 1u : 0u);
		k.c0 = (multi ? 0u : 1u) - k.two_l1;

		const uint32_t a = (uint32_t)((uintptr_t)dst & 15u);
		const uint32_t checksum = (flags & AIRS_FJ_CHECKSUM) ? 1u : 0u;
		Out o;
		o.base = dst - a;
		o.lo = a;
		o.hi = a + cap_eff;
		o.gw0 = 0;
		o.sbits = 8u * a;

		/* the 22 header bytes travel through the staging words in front of the code words (size field zero,
		 * patched at the end), as five and a half big-endian words: ref cmp_hdr_serialize, header.c:24-67;
		 * fields cmp.c:265-279.  The area is all zero here, a is 0 or 8: plain stores */
		if (lane == 0) {
			uint32_t *h = ws.stg + a / 4u;
			const uint32_t pre = (flags & AIRS_FJ_PRE_DIFF) ? CMP_PREPROCESS_DIFF : CMP_PREPROCESS_NONE;
			const uint32_t enc = multi ? CMP_ENCODER_GOLOMB_MULTI : CMP_ENCODER_GOLOMB_ZERO;
			h[0] = (0x8000u | CMP_VERSION_NUMBER) << 16;
			h[1] = (2u * n) & 0xFFFFFFu;
			h[2] = (id_hi << 16) | (id_lo >> 16);
			h[3] = (id_lo << 16) | (pre << 4) | (checksum << 3) | enc;
			h[4] = (g << 8) | ((outlier >> 16) & 0xFFu);
			h[5] = outlier << 16;
		}
		__syncwarp();
		o.sbits += 8u * (CMP_HDR_SIZE + 6u);

		if (multi) {
			if (flags & AIRS_FJ_PRE_DIFF)
				encode_units<true, true>(dbg, k, ws, o, src, n, lane);
			else
				encode_units<true, false>(dbg, k, ws, o, src, n, lane);
		} else {
			if (flags & AIRS_FJ_PRE_DIFF)
				encode_units<false, true>(dbg, k, ws, o, src, n, lane);
			else
				encode_units<false, false>(dbg, k, ws, o, src, n, lane);
		}

		const uint32_t frame_bits = o.gw0 * 32u + o.sbits - 8u * a;
		const uint32_t payload_end = (frame_bits + 7u) >> 3; /* header + code bytes, zero padded (bitstream_writer.h:205-227) */
		const uint32_t size = payload_end + 4u * checksum;
		uint32_t result = size > cap_eff ? AIRS_ERR(DST_TOO_SMALL) : size;

		{ /* the last partial group: one byte per lane */
			const uint32_t nb = (o.sbits + 7u) >> 3, pos = o.gw0 * 4u + lane;
			if (lane < nb && pos >= o.lo && pos < o.hi)
				o.base[pos] = (uint8_t)(ws.stg[lane >> 2] >> (24u - 8u * (lane & 3u)));
			__syncwarp();
			if (lane < 4u)
				ws.stg[lane] = 0;
			__syncwarp();
		}
		if (!airs_failed(result)) {
			if (lane < 3u) /* the size field of the header that went out with the stream */
				dst[2u + lane] = (uint8_t)(size >> (16u - 8u * lane));
		} else if (flags & AIRS_FJ_FALLBACK_OK) {
			/* stored raw as a fresh primary pass, two more identifiers drawn (ref cmp.c:380-392): 16-byte
			 * header of a NONE + UNCOMPRESSED stream, the samples big endian (dst is 8-byte aligned) */
			const uint32_t raw_size = CMP_HDR_SIZE + 2u * n + 4u * checksum;
			const uint64_t id = ((((uint64_t)id_hi << 32) | id_lo) + 2u) & 0xFFFFFFFFFFFFull;
			uint32_t *out = reinterpret_cast<uint32_t *>(dst + CMP_HDR_SIZE);
			const uint32_t *in = reinterpret_cast<const uint32_t *>(src);
			for (uint32_t i = lane; i < n / 2u; i += 32u)
				out[i] = airs_be_pair(__ldg(in + i));
			if (lane == 0 && (n & 1u)) {
				const uint32_t xs = __ldg(reinterpret_cast<const uint16_t *>(src) + n - 1u);
				dst[CMP_HDR_SIZE + 2u * (n - 1u)] = (uint8_t)(xs >> 8);
				dst[CMP_HDR_SIZE + 2u * (n - 1u) + 1u] = (uint8_t)xs;
			}
			if (lane < CMP_HDR_SIZE) {
				uint32_t v;
				switch (lane) {
				case 0: v = 0x80u | (CMP_VERSION_NUMBER >> 8); break;
				case 1: v = CMP_VERSION_NUMBER & 0xFFu; break;
				case 2: v = raw_size >> 16; break;
				case 3: v = raw_size >> 8; break;
				case 4: v = raw_size; break;
				case 5: v = (2u * n) >> 16; break;
				case 6: v = (2u * n) >> 8; break;
				case 7: v = 2u * n; break;
				case 14: v = 0; break;
				case 15: v = checksum << 3; break;
				default: v = (uint32_t)(id >> (8u * (13u - lane))); break;
				}
				dst[lane] = (uint8_t)v;
			}
			result = raw_size;
		}
		if (lane == 0)
			b.results[first_result] = result;

		t = t1;
		rec = nrec;
		t1 = __shfl_sync(kFull, t2, 0);
	}
}

extern "C" cudaError_t airs_launch_fast(const AirsLaunch *b, unsigned int grid, cudaStream_t stream)
{
	airs_fast_kernel<<<grid, AIRS_FAST_THREADS, 0, stream>>>(*b);
	return cudaGetLastError();
}

/* AIRS_BOUNDS_CHECK builds: strings that would have been staged outside their warp's words since the last call
 * (-1: not such a build) */
extern "C" int airs_fast_bounds_violations(void)
{
#ifdef AIRS_BOUNDS_CHECK
	unsigned int v = 0, zero = 0;
	cudaMemcpyFromSymbol(&v, airs_bounds_violations, sizeof(v));
	cudaMemcpyToSymbol(airs_bounds_violations, &zero, sizeof(zero));
	return (int)v;
#else
	return -1;
#endif
}

/* resident CTAs of airs_fast_kernel on the current device */
extern "C" cudaError_t airs_fast_resident_ctas(int *out)
{
	int dev = 0, sms = 0, per_sm = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e == cudaSuccess)
		e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (e == cudaSuccess) {
		/* shared memory for the CTAs the kernel is compiled for (1 KiB per CTA is reserved by the system); the
		 * rest of the 228 KiB stays L1 */
		const size_t need = (size_t)AIRS_FAST_CTAS_PER_SM * (sizeof(FastWarp) * kFWarps + 1024);
		const int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
		e = cudaFuncSetAttribute(airs_fast_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
	}
	if (e == cudaSuccess)
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, airs_fast_kernel, AIRS_FAST_THREADS, 0);
	*out = sms * per_sm;
	return e;
}
