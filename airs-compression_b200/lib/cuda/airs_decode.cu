/*
 * airs_decode.cu - batched decoder of AIRSPACE streams for sm_100a
 * (include/airs_cuda_decode.h).
 *
 * The inverse of the path in airs_kernels.cu, i.e. of the reference's
 * compress_engine (lib/compress/cmp.c:213-338): header (lib/common/header.c:
 * 24-67,89-134), code words (lib/compress/encoder.c:303-378), zig-zag
 * (encoder.c:274-286), preprocessing (lib/compress/preprocess.c:268-411),
 * model update (cmp.c:120-142), checksum (header.c:137-163).
 *
 * What is parallel in a stream of variable-length codes is decided by the
 * format: code word k starts where code word k-1 ends, so one stream is one
 * serial chain - but residuals do not depend on the model, so the frames of a
 * context decode independently of each other.  Five kernels, all in stream
 * order:
 *
 *   dec_index_kernel   one thread per job: where every frame's stream starts
 *                      (streams that lie back to back are found by walking
 *                      their headers)
 *   dec_stream_kernel  one thread per frame: header, code words -> residuals,
 *                      running sum for DIFF frames, 16-byte stores
 *   dec_iwt_kernel     one CTA per IWT frame: the lifting steps undone level
 *                      by level, coarsest first
 *   dec_model_kernel   threads over sample positions, loop over the frames of
 *                      a job: x = model + residual, model update; the model
 *                      lives in a register and never touches memory
 *   dec_verify_kernel  one thread per frame: XXH32 of the decoded samples
 *                      against the trailer; writes results[]
 *
 * First version of this row: parity (round trips against the reference's
 * streams) before speed.  A batch of many frames keeps the device busy; a
 * single long stream is decoded by a single thread.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "airs_device.cuh"
#include "../../../include/airs_cuda_decode.h"

/* airs_cuda_api.cu */
extern "C" int airs_internal_fail(int code, const char *text);
extern "C" void airs_internal_set_launches(int n);
extern "C" int airs_internal_check_device(void);

namespace {

constexpr uint32_t kVersionWord = 0x8000u | 600u; /* version flag and id the encoder writes (ref cmp.c:265-279) */
constexpr uint32_t kHdr = 16u, kExt = 6u;         /* ref cmp_header.h:46-58 */

#define DEC_ERR(code) ((uint32_t)0 - (uint32_t)(code))

/* one record per frame, written by dec_index_kernel and dec_stream_kernel */
struct DecFrame {
	uint64_t off;     /* start of the stream, bytes from the stream base */
	uint32_t avail;   /* bytes the stream may occupy */
	uint32_t n;       /* samples */
	uint32_t err;     /* 0 or (uint32_t)-code */
	uint32_t trailer; /* checksum found behind the code words */
	uint32_t job;
	uint8_t pre, seq, rate, cks;
};
static_assert(sizeof(DecFrame) == 32, "one record per frame in the scratch area");

struct DecLaunch {
	const uint8_t *src;
	uint8_t *dst;
	const airs_dec_job *jobs;
	uint32_t *results;
	airs_frame_info *info;
	DecFrame *frames;
	uint32_t n_jobs, n_results, split;
};

__device__ __forceinline__ uint32_t be16(const uint8_t *p)
{
	return (uint32_t)p[0] << 8 | p[1];
}

__device__ __forceinline__ uint32_t be24(const uint8_t *p)
{
	return (uint32_t)p[0] << 16 | (uint32_t)p[1] << 8 | p[2];
}

__device__ __forceinline__ uint32_t container_size(uint32_t dtype)
{
	return dtype == AIRS_DTYPE_I16_IN_I32 ? 4u : 2u;
}

/* ------------------------------------------------------------------------- */

__global__ void __launch_bounds__(128) dec_index_kernel(DecLaunch b)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= b.n_jobs)
		return;
	const airs_dec_job job = b.jobs[j];
	uint64_t off = job.src_offset, left = job.src_size;
	bool lost = false; /* back to back: a header in front could not be trusted, the rest cannot be found */

	for (uint32_t f = 0; f < job.n_frames; f++) {
		const uint64_t k = (uint64_t)job.first_result + f;
		if (k >= b.n_results)
			break;
		DecFrame r;
		memset(&r, 0, sizeof(r));
		r.job = j;
		if (job.src_frame_stride) {
			r.off = job.src_offset + (uint64_t)f * job.src_frame_stride;
			r.avail = job.src_size;
		} else {
			r.off = off;
			r.avail = (uint32_t)left;
			if (lost) {
				r.err = DEC_ERR(CMP_ERR_INT_HDR);
			} else if (left >= kHdr && b.src) {
				const uint32_t csize = be24(b.src + off + CMP_HDR_OFFSET_COMPRESSED_SIZE);
				if (csize >= kHdr && csize <= left) {
					r.avail = csize;
					off += csize;
					left -= csize;
				} else {
					lost = true; /* this frame reports the trouble itself */
				}
			} else {
				lost = true;
			}
		}
		b.frames[k] = r;
	}
}

/* ------------------------------------------------------------------------- */

/* MSB-first reader over a stream at any byte address: 32-bit words of the
 * 4-byte aligned space around it, at least 33 valid bits in `buf` between calls */
struct Reader {
	const uint32_t *wb;
	uint32_t limit; /* aligned-space byte index behind the stream: nothing is read from there on */
	uint32_t next;
	uint32_t cnt;
	uint32_t used;
	uint64_t buf;

	__device__ __forceinline__ uint32_t fetch(uint32_t i) const
	{
		const uint32_t byte = 4u * i;
		if (byte + 4u <= limit)
			return airs_bswap32(__ldg(wb + i));
		uint32_t v = 0;
		const uint8_t *p = reinterpret_cast<const uint8_t *>(wb);
		for (uint32_t k = 0; k < 4u; k++)
			if (byte + k < limit)
				v |= (uint32_t)p[byte + k] << (24u - 8u * k);
		return v;
	}
	__device__ __forceinline__ void open(const uint8_t *stream, uint32_t size, uint32_t start_byte)
	{
		open_bits(stream, size, 8u * start_byte);
	}
	__device__ __forceinline__ void open_bits(const uint8_t *stream, uint32_t size, uint32_t start_bit)
	{
		const uint32_t lead = (uint32_t)((uintptr_t)stream & 3u);
		wb = reinterpret_cast<const uint32_t *>(stream - lead);
		limit = lead + size;
		const uint32_t bit = 8u * lead + start_bit, i = bit >> 5, sk = bit & 31u;
		buf = ((uint64_t)fetch(i) << 32 | fetch(i + 1u)) << sk;
		cnt = 64u - sk;
		next = i + 2u;
		used = 0;
	}
	__device__ __forceinline__ uint32_t peek() const
	{
		return (uint32_t)(buf >> 32);
	}
	__device__ __forceinline__ void skip(uint32_t n) /* n <= 32 */
	{
		buf <<= n;
		cnt -= n;
		used += n;
		if (cnt <= 32u) {
			buf |= (uint64_t)fetch(next++) << (32u - cnt);
			cnt += 32u;
		}
	}
};

/* constants of the code of one stream (ref golomb_encode, encoder.c:303-324: L = floor(log2 g),
 * cutoff = 2^(L+1) - g) */
struct DecConst {
	uint32_t g, L, cutoff, outlier;
	uint32_t two_cutoff; /* (L+2)-bit windows below this are (L+1)-bit code words of the values below cutoff */
	uint32_t top;        /* windows from 2^(L+1) + cutoff on start with a 1 of the unary part */
	uint32_t sh;         /* 30 - L: a 32-bit window >> sh leaves L + 2 bits */
};

/* value of the Golomb code word at the head of the reader.  A code word is q ones, then the
 * (L+2)-bit field 2 cutoff + r (r < g); the values below cutoff are their own (L+1)-bit code
 * words.  The field itself starts with at most one 1 (2 cutoff + g <= 3 * 2^L), so q is the
 * number of leading ones or one less. */
__device__ __forceinline__ uint32_t golomb_value(Reader &rd, const DecConst &dc, bool &bad)
{
	const uint32_t w = rd.peek();
	const uint32_t ones = (uint32_t)__clz((int)~w);
	uint32_t q = ones ? ones - 1u : 0u;
	uint32_t field = (w << q) >> dc.sh;
	if (field >= dc.top) {
		q++;
		field = (q < 32u ? w << q : 0u) >> dc.sh;
	}
	uint32_t v, len;
	if (q == 0u && field < dc.two_cutoff) {
		v = field >> 1;
		len = dc.L + 1u;
	} else {
		v = dc.cutoff + q * dc.g + field - dc.two_cutoff;
		len = q + dc.L + 2u;
	}
	if (len > 32u) { /* the encoder never writes more than 32 bits (ref encoder.c:154-182) */
		bad = true;
		len = 32u;
	}
	rd.skip(len);
	return v;
}

/* one residual (16 bits): ref cmp_encoder_encode_s16, encoder.c:327-378, backwards */
template <int ENC>
__device__ __forceinline__ uint32_t decode_one(Reader &rd, const DecConst &dc, bool &bad)
{
	uint32_t m;
	if (ENC == CMP_ENCODER_UNCOMPRESSED) {
		m = rd.peek() >> 16;
		rd.skip(16u);
		return m; /* raw residual, no zig-zag */
	}
	const uint32_t v = golomb_value(rd, dc, bad);
	if (ENC == CMP_ENCODER_GOLOMB_ZERO) {
		if (v == 0u) { /* escape: the mapped value in 16 raw bits */
			m = rd.peek() >> 16;
			rd.skip(16u);
		} else {
			m = v - 1u;
		}
	} else {
		if (v >= dc.outlier) { /* escape symbol outlier + level: 2 level + 2 raw bits of m - outlier */
			uint32_t nb = 2u * (v - dc.outlier) + 2u;
			if (nb > 16u) {
				bad = true;
				nb = 16u;
			}
			m = dc.outlier + (rd.peek() >> (32u - nb));
			rd.skip(nb);
		} else {
			m = v;
		}
	}
	if (m > 0xFFFFu)
		bad = true;
	return ((m >> 1) ^ (0u - (m & 1u))) & 0xFFFFu; /* ref map_to_unsigned, encoder.c:274-286 */
}

__device__ __forceinline__ uint32_t sext16(uint32_t v)
{
	return (uint32_t)(int32_t)(int16_t)v;
}

/* n residuals of one stream into `out`; DIFF frames leave as samples (ref diff_process,
 * preprocess.c:284-290: the first residual is the sample itself) */
template <int ENC>
__device__ void decode_samples(Reader &rd, const DecConst &dc, uint32_t n, bool diff, uint8_t *out, bool c32,
			       bool &bad, uint32_t prev = 0)
{
	const bool vec = ((uintptr_t)out & 15u) == 0;
	uint32_t i = 0;

	for (; i + 8u <= n; i += 8u) {
		uint32_t pk[4];
#pragma unroll
		for (int k = 0; k < 8; k++) {
			uint32_t r = decode_one<ENC>(rd, dc, bad);
			if (diff) {
				prev = (prev + r) & 0xFFFFu;
				r = prev;
			}
			if (k & 1)
				pk[k >> 1] |= r << 16;
			else
				pk[k >> 1] = r;
		}
		if (c32) {
			uint32_t *o = reinterpret_cast<uint32_t *>(out) + i;
			if (vec) {
				reinterpret_cast<uint4 *>(o)[0] = make_uint4(sext16(pk[0]), sext16(pk[0] >> 16),
									     sext16(pk[1]), sext16(pk[1] >> 16));
				reinterpret_cast<uint4 *>(o)[1] = make_uint4(sext16(pk[2]), sext16(pk[2] >> 16),
									     sext16(pk[3]), sext16(pk[3] >> 16));
			} else {
#pragma unroll
				for (int k = 0; k < 8; k++)
					o[k] = sext16(pk[k >> 1] >> (16 * (k & 1)));
			}
		} else {
			uint16_t *o = reinterpret_cast<uint16_t *>(out) + i;
			if (vec) {
				*reinterpret_cast<uint4 *>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
			} else {
#pragma unroll
				for (int k = 0; k < 8; k++)
					o[k] = (uint16_t)(pk[k >> 1] >> (16 * (k & 1)));
			}
		}
	}
	for (; i < n; i++) {
		uint32_t r = decode_one<ENC>(rd, dc, bad);
		if (diff) {
			prev = (prev + r) & 0xFFFFu;
			r = prev;
		}
		if (c32)
			reinterpret_cast<uint32_t *>(out)[i] = sext16(r);
		else
			reinterpret_cast<uint16_t *>(out)[i] = (uint16_t)r;
	}
}

/* what both stream kernels need of a frame: header read and checked, code constants */
struct StreamPlan {
	const uint8_t *s;
	uint8_t *out;
	uint32_t err, hdr_size, csize, n, pre, enc, tail;
	bool c32, diff;
	DecConst dc;
};

/* header: ref cmp_hdr_deserialize, header.c:89-134; then what the encoder can have written (ref
 * cmp.c:265-279, cmp_initialise cmp.c:152-209) */
__device__ void plan_stream(const DecLaunch &b, uint32_t k, DecFrame &fr, airs_frame_info &inf, StreamPlan &sp)
{
	memset(&inf, 0, sizeof(inf));
	memset(&sp, 0, sizeof(sp));
	if (fr.job >= b.n_jobs) { /* a result index no job claims */
		fr.err = DEC_ERR(CMP_ERR_GENERIC);
		fr.job = 0;
	}
	const airs_dec_job job = b.jobs[fr.job];
	const uint8_t *s = b.src + fr.off;
	uint32_t err = fr.err;
	uint32_t hdr_size = 0, csize = 0, n = 0, pre = 0, enc = 0;
	DecConst &dc = sp.dc;

	if (!err && (!b.src || !b.dst))
		err = DEC_ERR(!b.src ? CMP_ERR_SRC_NULL : CMP_ERR_DST_NULL);
	if (!err && job.dtype > AIRS_DTYPE_U16)
		err = DEC_ERR(CMP_ERR_PARAMS_INVALID);
	if (!err && fr.avail < kHdr)
		err = DEC_ERR(CMP_ERR_INT_HDR);
	if (!err) {
		const uint32_t method = s[CMP_HDR_OFFSET_METHOD];
		inf.version = (uint16_t)be16(s + CMP_HDR_OFFSET_VERSION);
		inf.compressed_size = csize = be24(s + CMP_HDR_OFFSET_COMPRESSED_SIZE);
		inf.original_size = be24(s + CMP_HDR_OFFSET_ORIGINAL_SIZE);
		inf.identifier = (uint64_t)be24(s + CMP_HDR_OFFSET_IDENTIFIER) << 24 | be24(s + CMP_HDR_OFFSET_IDENTIFIER + 3);
		inf.sequence_number = s[CMP_HDR_OFFSET_SEQUENCE_NUMBER];
		inf.preprocessing = pre = (method >> 4) & 0xFu;
		inf.checksum_enabled = (method >> 3) & 1u;
		inf.encoder_type = enc = method & 7u;
		hdr_size = kHdr;
		if (pre != CMP_PREPROCESS_NONE || enc != CMP_ENCODER_UNCOMPRESSED) {
			if (fr.avail < kHdr + kExt) {
				memset(&inf, 0, sizeof(inf));
				err = DEC_ERR(CMP_ERR_INT_HDR);
			} else {
				inf.model_rate = s[kHdr];
				inf.encoder_param = (uint16_t)be16(s + kHdr + 1);
				inf.encoder_outlier = be24(s + kHdr + 3);
				hdr_size = kHdr + kExt;
			}
		}
		inf.header_size = err ? 0 : (uint8_t)hdr_size;
	}
	const uint32_t tail = inf.checksum_enabled ? 4u : 0u;
	if (!err) {
		n = inf.original_size / 2u;
		if (inf.version != kVersionWord || (inf.original_size & 1u) || n == 0u || pre > CMP_PREPROCESS_MODEL ||
		    enc > CMP_ENCODER_GOLOMB_MULTI || csize < hdr_size + tail ||
		    (pre == CMP_PREPROCESS_MODEL && (inf.model_rate > 16u || inf.sequence_number == 0u)))
			err = DEC_ERR(CMP_ERR_INT_HDR);
		else if (csize > fr.avail)
			err = DEC_ERR(CMP_ERR_SRC_SIZE_WRONG);
		else if ((uint64_t)n * container_size(job.dtype) > job.dst_capacity)
			err = DEC_ERR(CMP_ERR_DST_TOO_SMALL);
		else if ((job.dst_offset | job.dst_frame_stride) & (container_size(job.dtype) - 1u))
			err = DEC_ERR(CMP_ERR_DST_UNALIGNED);
		if (!err && enc != CMP_ENCODER_UNCOMPRESSED) {
			dc.g = inf.encoder_param;
			dc.outlier = inf.encoder_outlier;
			/* the header carries the outlier the encoder derived (ref encoder.c:185-224) */
			if (dc.g == 0u || dc.outlier == 0u ||
			    airs_derive_outlier(enc, dc.g, enc == CMP_ENCODER_GOLOMB_MULTI ? dc.outlier : 0u) != dc.outlier) {
				err = DEC_ERR(CMP_ERR_INT_HDR);
			} else {
				dc.L = airs_floor_log2(dc.g);
				dc.cutoff = (2u << dc.L) - dc.g;
				dc.two_cutoff = 2u * dc.cutoff;
				dc.top = (2u << dc.L) + dc.cutoff;
				dc.sh = 30u - dc.L;
			}
		}
	}
	if (!err && tail)
		fr.trailer = (uint32_t)be16(s + csize - 4u) << 16 | be16(s + csize - 2u);
	sp.s = s;
	sp.out = b.dst ? b.dst + job.dst_offset + (uint64_t)(k - job.first_result) * job.dst_frame_stride : nullptr;
	sp.err = err;
	sp.hdr_size = hdr_size;
	sp.csize = csize;
	sp.n = n;
	sp.pre = pre;
	sp.enc = enc;
	sp.tail = tail;
	sp.c32 = job.dtype == AIRS_DTYPE_I16_IN_I32;
	sp.diff = pre == CMP_PREPROCESS_DIFF;
	fr.n = n;
	fr.pre = (uint8_t)pre;
	fr.seq = inf.sequence_number;
	fr.rate = inf.model_rate;
	fr.cks = inf.checksum_enabled;
}

/* one thread per frame: the kernel for batches of many frames */
__global__ void __launch_bounds__(64) dec_stream_kernel(DecLaunch b)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= b.n_results)
		return;
	DecFrame fr = b.frames[k];
	airs_frame_info inf;
	StreamPlan sp;
	plan_stream(b, k, fr, inf, sp);
	uint32_t err = sp.err;
	if (!err) {
		bool bad = false;
		Reader rd;
		rd.open(sp.s, sp.csize - sp.tail, sp.hdr_size);
		if (sp.enc == CMP_ENCODER_UNCOMPRESSED)
			decode_samples<CMP_ENCODER_UNCOMPRESSED>(rd, sp.dc, sp.n, sp.diff, sp.out, sp.c32, bad);
		else if (sp.enc == CMP_ENCODER_GOLOMB_ZERO)
			decode_samples<CMP_ENCODER_GOLOMB_ZERO>(rd, sp.dc, sp.n, sp.diff, sp.out, sp.c32, bad);
		else
			decode_samples<CMP_ENCODER_GOLOMB_MULTI>(rd, sp.dc, sp.n, sp.diff, sp.out, sp.c32, bad);
		/* the code words are padded to a byte, then comes the trailer (ref cmp.c:314-332) */
		if (bad || sp.hdr_size + (rd.used + 7u) / 8u + sp.tail != sp.csize)
			err = DEC_ERR(AIRS_DEC_ERR_CORRUPT);
	}
	fr.err = err;
	b.frames[k] = fr;
	if (b.info)
		b.info[k] = inf;
}

/*
 * One WARP per frame: the kernel for batches of few frames, where a thread per frame leaves the
 * device empty.  Code words are self-delimiting, and a decoder that starts inside a code word
 * falls into step with the true sequence of code word boundaries after a few symbols.  The
 * stream is taken in windows of 32 x 128 bits: lane l decodes from bit 128 l of the window on
 * (lane 0 from the exact start) until it passes the end of its 128 bits, and remembers where
 * it ended.  Then every lane whose predecessor ended somewhere else than where it started
 * decodes again from there, until nothing changes any more - lane l is final after l + 1
 * rounds at the latest, in practice after two.  The symbols wait in shared memory, one row
 * per lane; the warp writes them out side by side (running sum for DIFF frames by a warp scan).
 */
constexpr uint32_t kSubBits = 128;             /* bits per lane and window: at most 128 symbols */
constexpr uint32_t kSymStride = kSubBits + 2u; /* halfwords per row: rows start in different banks */
constexpr uint32_t kWarpsPerCta = 4;

template <int ENC>
__device__ __forceinline__ uint32_t decode_run(const StreamPlan &sp, uint32_t start, uint32_t stop, uint32_t data_end,
					       uint32_t max_sym, uint16_t *row, uint32_t &cnt, bool &bad)
{
	cnt = 0;
	bad = false;
	if (start >= stop || start >= data_end)
		return start;
	Reader rd;
	rd.open_bits(sp.s, sp.csize - sp.tail, 8u * sp.hdr_size + start);
	uint32_t pos = start;
	while (pos < stop && pos < data_end && cnt < max_sym) {
		const uint32_t r = decode_one<ENC>(rd, sp.dc, bad);
		if (row)
			row[cnt] = (uint16_t)r;
		cnt++;
		pos = start + rd.used;
	}
	return pos;
}

template <int ENC>
__device__ uint32_t decode_stream_warp(const StreamPlan &sp, uint16_t (*sym)[kSymStride])
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t data_end = 8u * (sp.csize - sp.tail - sp.hdr_size); /* code word bits, padding included */
	const uint32_t n = sp.n;
	uint32_t wstart = 0; /* bits from the first code word to the start of the window: exact */
	uint32_t done = 0;   /* symbols written */
	uint32_t prev = 0;   /* DIFF: the sample in front */
	uint32_t end_bits = 0;
	uint32_t slow_windows = 0; /* windows in a row that took many rounds */
	bool bad = false;

	while (done < n) {
		if (wstart >= data_end) { /* the stream ends before the samples do */
			bad = true;
			break;
		}
		const uint32_t stop = wstart + (lane + 1u) * kSubBits;
		uint32_t start = wstart + lane * kSubBits, endp = 0, cnt = 0;
		bool lbad = false, dirty = true;
		uint32_t round = 0;
		for (; round < 33u; round++) {
			if (dirty)
				endp = decode_run<ENC>(sp, start, stop, data_end, kSubBits, sym[lane], cnt, lbad);
			const uint32_t pe = __shfl_up_sync(0xFFFFFFFFu, endp, 1);
			const uint32_t want = lane == 0u ? wstart : pe;
			dirty = want != start;
			start = want;
			if (!__any_sync(0xFFFFFFFFu, dirty))
				break;
		}
		/* symbols of the window: exclusive prefix of the counts */
		uint32_t inc = cnt;
		for (uint32_t d = 1; d < 32u; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
			if (lane >= d)
				inc += t;
		}
		const uint32_t excl = inc - cnt, total = __shfl_sync(0xFFFFFFFFu, inc, 31);
		const uint32_t take = min(total, n - done);
		if (lbad && excl < take)
			bad = true;
		__syncwarp();
		for (uint32_t i0 = 0; i0 < take; i0 += 32u) {
			const uint32_t i = i0 + lane;
			/* the row of symbol i: the last one that starts at or in front of it */
			uint32_t row = 0;
			for (uint32_t step = 16u; step; step >>= 1) {
				const uint32_t cand = row + step;
				const uint32_t ex = __shfl_sync(0xFFFFFFFFu, excl, cand & 31u);
				if (cand < 32u && ex <= i)
					row = cand;
			}
			const uint32_t ex_row = __shfl_sync(0xFFFFFFFFu, excl, row);
			uint32_t v = i < take ? sym[row][i - ex_row] : 0u;
			if (sp.diff) { /* ref diff_process, preprocess.c:284-290, undone by a running sum */
				for (uint32_t d = 1; d < 32u; d <<= 1) {
					const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
					if (lane >= d)
						v += t;
				}
				v = (v + prev) & 0xFFFFu;
				prev = __shfl_sync(0xFFFFFFFFu, v, 31);
			}
			if (i < take) {
				if (sp.c32)
					reinterpret_cast<uint32_t *>(sp.out)[done + i] = sext16(v);
				else
					reinterpret_cast<uint16_t *>(sp.out)[done + i] = (uint16_t)v;
			}
		}
		__syncwarp();
		done += take;
		if (done == n) { /* where the last symbol ends: the lane that holds it knows, or counts again */
			uint32_t row = 0;
			for (uint32_t step = 16u; step; step >>= 1) {
				const uint32_t cand = row + step;
				const uint32_t ex = __shfl_sync(0xFFFFFFFFu, excl, cand & 31u);
				if (cand < 32u && ex <= take - 1u)
					row = cand;
			}
			uint32_t e = endp;
			if (lane == row && take - excl != cnt) {
				uint32_t c2;
				bool b2;
				e = decode_run<ENC>(sp, start, stop, data_end, take - excl, nullptr, c2, b2);
			}
			end_bits = __shfl_sync(0xFFFFFFFFu, e, row);
		}
		wstart = __shfl_sync(0xFFFFFFFFu, endp, 31);
		slow_windows = round > 6u ? slow_windows + 1u : 0u;
		if (slow_windows >= 3u && done < n) {
			/* three windows in a row: the lanes keep correcting each other - code words of (nearly) fixed length, or raw
			 * escape bits everywhere, never fall into step.  The rest of the stream is one
			 * lane's work, straight to the destination. */
			if (lane == 0u) {
				Reader rd;
				rd.open_bits(sp.s, sp.csize - sp.tail, 8u * sp.hdr_size + wstart);
				decode_samples<ENC>(rd, sp.dc, n - done, sp.diff, sp.out + (size_t)done * (sp.c32 ? 4u : 2u), sp.c32,
						    bad, prev);
				end_bits = wstart + rd.used;
			}
			end_bits = __shfl_sync(0xFFFFFFFFu, end_bits, 0);
			break;
		}
	}
	if (__any_sync(0xFFFFFFFFu, bad) || sp.hdr_size + (end_bits + 7u) / 8u + sp.tail != sp.csize)
		return DEC_ERR(AIRS_DEC_ERR_CORRUPT);
	return 0;
}

__global__ void __launch_bounds__(32 * kWarpsPerCta) dec_stream_warp_kernel(DecLaunch b)
{
	__shared__ uint16_t sym[kWarpsPerCta][32][kSymStride];
	const uint32_t k = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
	if (k >= b.n_results)
		return;
	DecFrame fr = b.frames[k];
	airs_frame_info inf;
	StreamPlan sp;
	plan_stream(b, k, fr, inf, sp); /* every lane: the loads are broadcasts */
	uint32_t err = sp.err;
	if (!err) {
		uint16_t (*rows)[kSymStride] = sym[threadIdx.x >> 5];
		if (sp.enc == CMP_ENCODER_UNCOMPRESSED)
			err = decode_stream_warp<CMP_ENCODER_UNCOMPRESSED>(sp, rows);
		else if (sp.enc == CMP_ENCODER_GOLOMB_ZERO)
			err = decode_stream_warp<CMP_ENCODER_GOLOMB_ZERO>(sp, rows);
		else
			err = decode_stream_warp<CMP_ENCODER_GOLOMB_MULTI>(sp, rows);
	}
	if ((threadIdx.x & 31u) == 0) {
		fr.err = err;
		b.frames[k] = fr;
		if (b.info)
			b.info[k] = inf;
	}
}

/* ------------------------------------------------------------------------- */

__device__ __forceinline__ int32_t ld_s16(const uint8_t *p, bool c32, uint32_t i)
{
	return c32 ? (int32_t)(int16_t)reinterpret_cast<const uint32_t *>(p)[i]
		   : (int32_t)reinterpret_cast<const int16_t *>(p)[i];
}

__device__ __forceinline__ void st_s16(uint8_t *p, bool c32, uint32_t i, int32_t v)
{
	const int16_t t = (int16_t)(uint16_t)(uint32_t)v;
	if (c32)
		reinterpret_cast<int32_t *>(p)[i] = t;
	else
		reinterpret_cast<int16_t *>(p)[i] = t;
}

__device__ __forceinline__ int32_t w16(int32_t v)
{
	return (int16_t)(uint16_t)(uint32_t)v;
}

/* IWT frames hold coefficients in place (ref iwt_multi_level_decomposition_i16,
 * preprocess.c:190-221: strides 1, 2, 4, .. < n; per level first the details at the odd
 * multiples of the stride, then the approximations at the even ones, preprocess.c:140-177).
 * Undone in the opposite order: coarsest level first, approximations before details. */
__global__ void __launch_bounds__(256) dec_iwt_kernel(DecLaunch b)
{
	const uint32_t k = blockIdx.x;
	const DecFrame fr = b.frames[k];
	if (fr.err || fr.pre != CMP_PREPROCESS_IWT || fr.n < 2u)
		return;
	const airs_dec_job job = b.jobs[fr.job];
	uint8_t *w = b.dst + job.dst_offset + (uint64_t)(k - job.first_result) * job.dst_frame_stride;
	const bool c32 = job.dtype == AIRS_DTYPE_I16_IN_I32;
	const uint32_t n = fr.n;

	for (uint32_t s = 1u << (31 - __clz((int)(n - 1u))); s >= 1u; s >>= 1) {
		for (uint32_t i = 2u * s * threadIdx.x; i < n; i += 2u * s * blockDim.x) {
			const bool has_l = i >= s, has_r = i + s < n;
			int32_t t = 0;
			if (has_l && has_r)
				t = w16((ld_s16(w, c32, i - s) + ld_s16(w, c32, i + s)) >> 2);
			else if (has_r)
				t = w16(ld_s16(w, c32, i + s) >> 1);
			else if (has_l)
				t = w16(ld_s16(w, c32, i - s) >> 1);
			if (has_l || has_r)
				st_s16(w, c32, i, ld_s16(w, c32, i) - t);
		}
		__syncthreads();
		for (uint32_t i = s + 2u * s * threadIdx.x; i < n; i += 2u * s * blockDim.x) {
			const int32_t t = i + s < n ? w16((ld_s16(w, c32, i - s) + ld_s16(w, c32, i + s)) >> 1)
						    : ld_s16(w, c32, i - s);
			st_s16(w, c32, i, ld_s16(w, c32, i) + t);
		}
		__syncthreads();
	}
}

/* ------------------------------------------------------------------------- */

/* MODEL frames: x = model + residual (ref model_process, preprocess.c:406-411), then the model
 * update (ref cmp.c:120-142,304-311): after a frame with sequence number 0 the model is that
 * frame, after a MODEL frame it is the weighted mean.  Sample positions are independent of
 * each other, so a thread carries the model of its position in a register through all frames
 * of the job.  b.split CTAs share a job. */
__global__ void __launch_bounds__(256) dec_model_kernel(DecLaunch b)
{
	const uint32_t j = blockIdx.x / b.split, part = blockIdx.x % b.split;
	const airs_dec_job job = b.jobs[j];
	const bool c32 = job.dtype == AIRS_DTYPE_I16_IN_I32, is_signed = job.dtype != AIRS_DTYPE_U16;
	__shared__ uint32_t s_any, s_max;

	if ((uint64_t)job.first_result + job.n_frames > b.n_results)
		return;
	const DecFrame *frames = b.frames + job.first_result;
	if (threadIdx.x == 0) {
		s_any = 0;
		s_max = 0;
	}
	__syncthreads();
	for (uint32_t f = threadIdx.x; f < job.n_frames; f += blockDim.x) {
		if (frames[f].pre == CMP_PREPROCESS_MODEL && !frames[f].err) {
			s_any = 1;
			atomicMax(&s_max, frames[f].n);
		}
	}
	__syncthreads();
	if (!s_any)
		return;
	const uint32_t n_max = s_max;

	for (uint32_t i0 = part * blockDim.x; i0 < n_max; i0 += b.split * blockDim.x) {
		const uint32_t i = i0 + threadIdx.x;
		const bool marker = i == 0; /* the thread that reports frames without a model */
		uint32_t m = 0, n_m = 0;
		bool have = false;

		for (uint32_t f = 0; f < job.n_frames; f++) {
			const DecFrame fr = frames[f];
			uint8_t *x = b.dst + job.dst_offset + (uint64_t)f * job.dst_frame_stride;
			if (fr.err) { /* what follows up to the next first frame cannot be rebuilt */
				have = false;
				continue;
			}
			if (fr.seq == 0u) {
				have = true;
				n_m = fr.n;
				if (i < fr.n)
					m = (uint32_t)ld_s16(x, c32, i) & 0xFFFFu;
				continue;
			}
			if (fr.pre != CMP_PREPROCESS_MODEL)
				continue;
			if (!have || fr.n != n_m) {
				have = false;
				if (marker)
					b.frames[job.first_result + f].err = DEC_ERR(AIRS_DEC_ERR_NO_MODEL);
				continue;
			}
			if (i < fr.n) {
				const uint32_t v = (m + (uint32_t)ld_s16(x, c32, i)) & 0xFFFFu;
				st_s16(x, c32, i, (int32_t)v);
				m = airs_model_update(v, m, fr.rate, is_signed);
			}
		}
	}
}

/* ------------------------------------------------------------------------- */

/* XXH32 over the big-endian samples (ref cmp_checksum, header.c:137-163; xxHash 0.8.3) */
__device__ uint32_t samples_xxh32(const uint8_t *p, bool c32, uint32_t n)
{
	const uint32_t seed = AIRS_CHECKSUM_SEED, nbytes = 2u * n;
	uint32_t v0 = seed + AIRS_XP1 + AIRS_XP2, v1 = seed + AIRS_XP2, v2 = seed, v3 = seed - AIRS_XP1;
	uint32_t i = 0;
#define DEC_PAIR(i_) airs_be_pair(((uint32_t)ld_s16(p, c32, (i_)) & 0xFFFFu) | ((uint32_t)ld_s16(p, c32, (i_) + 1u) << 16))
	for (; i + 8u <= n; i += 8u) {
		v0 = airs_xxh_round(v0, DEC_PAIR(i));
		v1 = airs_xxh_round(v1, DEC_PAIR(i + 2u));
		v2 = airs_xxh_round(v2, DEC_PAIR(i + 4u));
		v3 = airs_xxh_round(v3, DEC_PAIR(i + 6u));
	}
	uint32_t h = nbytes >= 16u ? airs_rotl(v0, 1) + airs_rotl(v1, 7) + airs_rotl(v2, 12) + airs_rotl(v3, 18)
				   : seed + AIRS_XP5;
	h += nbytes;
	for (; i + 2u <= n; i += 2u)
		h = airs_rotl(h + DEC_PAIR(i) * AIRS_XP3, 17) * AIRS_XP4;
#undef DEC_PAIR
	if (i < n) {
		const uint32_t sv = (uint32_t)ld_s16(p, c32, i) & 0xFFFFu;
		h = airs_rotl(h + (sv >> 8) * AIRS_XP5, 11) * AIRS_XP1;
		h = airs_rotl(h + (sv & 0xFFu) * AIRS_XP5, 11) * AIRS_XP1;
	}
	h ^= h >> 15;
	h *= AIRS_XP2;
	h ^= h >> 13;
	h *= AIRS_XP3;
	h ^= h >> 16;
	return h;
}

__global__ void __launch_bounds__(64) dec_verify_kernel(DecLaunch b)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= b.n_results)
		return;
	const DecFrame fr = b.frames[k];
	uint32_t r = fr.err;
	if (!r) {
		const airs_dec_job job = b.jobs[fr.job];
		const bool c32 = job.dtype == AIRS_DTYPE_I16_IN_I32;
		const uint8_t *x = b.dst + job.dst_offset + (uint64_t)(k - job.first_result) * job.dst_frame_stride;
		r = fr.n * container_size(job.dtype);
		if (fr.cks && samples_xxh32(x, c32, fr.n) != fr.trailer)
			r = DEC_ERR(AIRS_DEC_ERR_CHECKSUM);
	}
	b.results[k] = r;
}

} /* namespace */

extern "C" size_t airs_cuda_decode_scratch_size(uint32_t n_jobs, uint32_t n_results)
{
	(void)n_jobs;
	return sizeof(DecFrame) * ((size_t)n_results + 1);
}

extern "C" int airs_cuda_decompress_batch(const struct airs_dec_batch *bt, void *stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	char msg[256];

	airs_internal_set_launches(0);
	if (!bt || !bt->jobs || !bt->results || !bt->scratch)
		return airs_internal_fail(AIRS_E_ARGUMENT, "batch, jobs, results and scratch must be non-NULL");
	if ((uintptr_t)bt->scratch & 15u)
		return airs_internal_fail(AIRS_E_ARGUMENT, "scratch must be 16-byte aligned");
	if (bt->n_jobs == 0 || bt->n_results == 0)
		return AIRS_OK;
	int rc = airs_internal_check_device();
	if (rc != AIRS_OK)
		return rc;

	DecLaunch l;
	memset(&l, 0, sizeof(l));
	l.src = (const uint8_t *)bt->src;
	l.dst = (uint8_t *)bt->dst;
	l.jobs = bt->jobs;
	l.results = bt->results;
	l.info = bt->info;
	l.frames = (DecFrame *)bt->scratch;
	l.n_jobs = bt->n_jobs;
	l.n_results = bt->n_results;
	/* CTAs per job of the model kernel: enough to fill the device when jobs are few */
	int dev = 0, sms = 148;
	if (cudaGetDevice(&dev) == cudaSuccess)
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const uint32_t want = 4u * (uint32_t)sms;
	l.split = bt->n_jobs >= want ? 1u : (want + bt->n_jobs - 1u) / bt->n_jobs;
	if (l.split > 128u)
		l.split = 128u;

	cudaError_t e;
	/* frames no job claims keep job = 0xFFFFFFFF and report an error */
	e = cudaMemsetAsync(l.frames, 0xFF, sizeof(DecFrame) * (size_t)bt->n_results, stream);
	if (e == cudaSuccess) {
		dec_index_kernel<<<(bt->n_jobs + 127u) / 128u, 128, 0, stream>>>(l);
		/* a warp per frame until a thread per frame fills the device */
		if (bt->n_results < 64u * 1024u)
			dec_stream_warp_kernel<<<(bt->n_results + kWarpsPerCta - 1u) / kWarpsPerCta, 32 * kWarpsPerCta, 0, stream>>>(l);
		else
			dec_stream_kernel<<<(bt->n_results + 63u) / 64u, 64, 0, stream>>>(l);
		dec_iwt_kernel<<<bt->n_results, 256, 0, stream>>>(l);
		dec_model_kernel<<<bt->n_jobs * l.split, 256, 0, stream>>>(l);
		dec_verify_kernel<<<(bt->n_results + 63u) / 64u, 64, 0, stream>>>(l);
		e = cudaGetLastError();
	}
	if (e != cudaSuccess) {
		snprintf(msg, sizeof(msg), "decode launch: %s", cudaGetErrorString(e));
		return airs_internal_fail(AIRS_E_CUDA, msg);
	}
	airs_internal_set_launches(5);
	return AIRS_OK;
}
