/*
 * airs_cuda.h - C-ABI of the B200 (sm_100a) backend of the AIRSPACE compressor.
 *
 * Plain C: pointers, sizes, fixed-width integers; no C++ or torch types.  This
 * is what the host shim (lib/host/cmp_shim.c, the cmp.h API) calls, and what a
 * foreign-function binding (ctypes, cgo, JNI ...) would bind.  Every entry point
 * below names the reference interface it stands in for.
 *
 * The batched driver has no counterpart function in the reference; its
 * semantics are DEFINED by the reference's API: running a batch is
 * byte-identical to this loop over the reference library (lib/cmp.h:154-344),
 * with a timestamp callback that returns a counter:
 *
 *   for each job j:
 *       counter = j.identifier_base              (48-bit counter: coarse<<16|fine)
 *       init_results[j] = cmp_initialise(&ctx, &j.params, work + j.work_offset, j.work_size)
 *       for f in 0 .. j.n_frames-1:
 *           results[j.first_result + f] =
 *               cmp_compress_<dtype>(&ctx, dst_f, j.dst_capacity,
 *                                    src + j.src_offset + f*j.src_frame_stride, j.src_size)
 *   every timestamp request returns counter++ (cmp.c:27-34 is the same counter,
 *   shared between contexts; here each job owns one so that jobs are independent).
 *
 * dst_f is dst + j.dst_offset + f*j.dst_frame_stride in the SLOTS layout.  In
 * the CONCAT layout the streams of all frames are written back to back, in
 * result-index order, starting at dst (what `airspace -c --stdout f1 f2 ...`
 * produces, programs/airspacecli.c:131-202): out_offsets[k] is where stream k
 * starts, out_offsets[n_results] the total; a frame that failed contributes 0
 * bytes; dst_capacity still bounds each stream.
 *
 * Contract of the job table: the frames of job j have the result indices
 * first_result .. first_result + n_frames - 1, all below n_results.  In the
 * CONCAT layout the jobs list the frames 0 .. n_results - 1 in order (job 0
 * starts at 0, every job where the one before ended, the last one ends at
 * n_results) - the order of the concatenation is the order of the table.  A
 * table that breaks this is refused on the device: nothing is encoded and
 * every results[k] is (uint32_t)-CMP_ERR_GENERIC.
 *
 * Speed, not results, depends on two things a caller chooses.  dst_capacity:
 * with cmp_compress_bound(src_size) bytes per stream (lib/cmp.h) and no
 * uncompressed fallback no frame of a context can fail, and only then are the
 * frames of a context encoded out of order (runs of secondary MODEL passes with
 * the model in registers; contexts of few-job batches tile by tile).  work:
 * a batch whose jobs need no work buffer passes NULL and skips the launches of
 * the wavelet transform kernels.  Work buffers hold afterwards what the
 * reference leaves in work_buf: the model, or the IWT coefficients.
 */
#ifndef AIRS_CUDA_H
#define AIRS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#include "cmp.h"

#ifdef __cplusplus
extern "C" {
#endif

/* sample containers; values equal the reference's private enum cmp_type
 * (lib/common/sample_reader.h:9) */
#define AIRS_DTYPE_I16        0u
#define AIRS_DTYPE_I16_IN_I32 1u
#define AIRS_DTYPE_U16        2u
/* ... | AIRS_DTYPE_BE: the 16-bit samples lie BIG-ENDIAN in memory - a sample file as the reference's front end
 * reads it (programs/file.c:337-358 loads it and swaps every sample to host order before it compresses).  The
 * kernels swap while they load; the streams are those of the swapped samples in the plain container.  Not with
 * AIRS_DTYPE_I16_IN_I32. */
#define AIRS_DTYPE_BE         4u
#define AIRS_DTYPE_I16_BE     (AIRS_DTYPE_I16 | AIRS_DTYPE_BE)
#define AIRS_DTYPE_U16_BE     (AIRS_DTYPE_U16 | AIRS_DTYPE_BE)

#define AIRS_LAYOUT_SLOTS  0u /* every frame has its own dst slot */
#define AIRS_LAYOUT_CONCAT 1u /* streams laid out back to back by a device-wide scan */

/* flags of a batch */
#define AIRS_BATCH_BIG_ENDIAN 1u /* the batch may hold jobs with AIRS_DTYPE_BE containers (the kernel variants that
				    know them run; without the flag such a job fails like an unknown container) */

/* result value of a frame that was not attempted because its job failed to
 * initialise: the reference returns CMP_ERR_CONTEXT_INVALID there (cmp.c:353) */

/*
 * One compression context and the frames pushed through it, in order.
 * Independent chunks are jobs with n_frames == 1.  120 bytes.
 */
struct airs_job {
	uint64_t src_offset;       /* frame 0, bytes from the source base */
	uint64_t src_frame_stride; /* bytes between consecutive frames */
	uint64_t dst_offset;       /* SLOTS: slot of frame 0, bytes from the dst base */
	uint64_t dst_frame_stride; /* SLOTS: bytes between consecutive slots */
	uint64_t work_offset;      /* this context's work buffer, bytes from the work base */
	uint64_t identifier_base;  /* timestamp counter when the job starts */
	uint32_t src_size;         /* bytes per frame, in the source container type */
	uint32_t dst_capacity;     /* bytes available per stream */
	uint32_t work_size;        /* bytes of work buffer (0: none) */
	uint32_t n_frames;
	uint32_t dtype;            /* AIRS_DTYPE_* */
	uint32_t first_result;     /* index of frame 0 in results[] / out_offsets[] */
	struct cmp_params params;  /* as passed to cmp_initialise() */
	uint32_t reserved;         /* must be 0 */
};

/* Device-resident batch: every pointer is a device pointer. */
struct airs_batch {
	const void *src;             /* source base */
	void *dst;                   /* destination base, 16-byte aligned */
	void *work;                  /* work-buffer base (models, IWT scratch) or NULL */
	const struct airs_job *jobs; /* n_jobs descriptors */
	uint32_t *results;           /* n_results: stream size or (uint32_t)-cmp_error */
	uint32_t *init_results;      /* n_jobs: result of cmp_initialise, or NULL */
	uint64_t *out_offsets;       /* CONCAT: n_results + 1 entries; else NULL */
	void *scratch;               /* airs_cuda_batch_scratch_size() bytes */
	uint64_t dst_size;           /* bytes available behind dst */
	uint32_t n_jobs;
	uint32_t n_results;          /* total number of frames */
	uint32_t layout;             /* AIRS_LAYOUT_* */
	uint32_t flags;              /* AIRS_BATCH_* */
	void *tmp;                   /* CONCAT: temporary device memory, 16-byte aligned, or NULL (see below) */
	uint64_t tmp_size;           /* bytes behind tmp */
};

/*
 * CONCAT with temporary memory: the batch is encoded into temporary slots at the speed of the
 * SLOTS layout, one scan turns sizes into offsets and a copy lays the streams out back to back.
 * tmp must hold airs_cuda_concat_tmp_size(S) bytes, S = sum over the jobs of
 * n_frames * dst_capacity.  Without tmp (or with too little of it, or when the streams do not
 * all fit dst_size) the single-phase path runs: every frame is sized, then encoded at its final
 * place - same bytes, no extra memory, about twice the work and contexts one after the other.
 */
size_t airs_cuda_concat_tmp_size(uint64_t sum_of_capacities, uint32_t n_results);

/* 0 on success, else a negative value; airs_cuda_last_error() explains. */
#define AIRS_OK              0
#define AIRS_E_NO_DEVICE    -1 /* no CUDA device / driver: there is NO CPU fallback */
#define AIRS_E_CUDA         -2 /* a CUDA runtime call failed */
#define AIRS_E_ARGUMENT     -3
#define AIRS_E_NOMEM        -4

/* Number of usable sm_100 devices (0: none).  No reference counterpart. */
int airs_cuda_device_count(void);

/* Number of jobs (compression contexts) the current device works on at the same
 * time: one per resident CTA of the encode kernel.  A batch keeps the device busy
 * when it holds at least this many jobs of similar length (0: no usable device).
 * No reference counterpart. */
int airs_cuda_concurrent_jobs(void);

/* Text of the last failure on this thread. */
const char *airs_cuda_last_error(void);

/* Bytes of device scratch a batch needs (look-back state, ticket counters). */
size_t airs_cuda_batch_scratch_size(uint32_t n_jobs, uint32_t n_results);

/*
 * Compress a device-resident batch on `stream` (a cudaStream_t passed as
 * void *, NULL = default stream).  Asynchronous: results are valid once the
 * stream has been synchronised.  Replaces, for a whole batch, the call chain
 * cmp_initialise -> cmp_compress_{u16,i16,i16_in_i32} (lib/compress/cmp.c:152-209,
 * 396-435) including compress_engine (cmp.c:213-338), the preprocessing methods
 * (lib/compress/preprocess.c:268-411), the encoder (lib/compress/encoder.c:185-378),
 * the bitstream writer (lib/common/bitstream_writer.h:124-227), the header
 * (lib/common/header.c:24-67) and the checksum (header.c:137-163).
 */
int airs_cuda_compress_batch(const struct airs_batch *batch, void *stream);

/*
 * Verification aid: the 64-bit hash of include/airs_stream_hash.h over every stream a batch
 * produced (0 for a frame that failed), hashes[n_results] in device memory.  Call after
 * airs_cuda_compress_batch() with the same descriptor (the scratch memory still holds the
 * frame-to-job map), on the same stream.  The CPU side of the comparison hashes the reference's
 * output the same way (oracle/hash_jobs.h), so whole workloads are compared through 12 bytes per
 * stream.  No reference counterpart.
 */
int airs_cuda_hash_streams(const struct airs_batch *batch, uint64_t *hashes, void *stream);

/* ... and the same hash over n byte ranges base + offsets[k], sizes[k] bytes (a size that is an error code
 * hashes to 0): streams that were moved, e.g. gathered from other GPUs.  All pointers are device pointers. */
int airs_cuda_hash_ranges(const void *base, const uint64_t *offsets, const uint32_t *sizes, uint32_t n,
			  uint64_t *hashes, void *stream);

/* Number of kernels the last airs_cuda_compress_batch() on this thread launched. */
int airs_cuda_last_launch_count(void);

/*
 * Host-buffer batch: src/dst/jobs/results/out_offsets are HOST pointers (pinned
 * or pageable; work may be NULL, models then live in device memory owned by the
 * call).  Copies the inputs to the device, runs airs_cuda_compress_batch,
 * copies streams and results back, synchronises.  This is the end-to-end path
 * a host caller of the reference library is switched to.  SLOTS: the whole dst
 * range is copied back, so slot bytes behind a stream are unspecified; CONCAT:
 * exactly out_offsets[n_results] bytes are copied back.
 */
struct airs_host_batch {
	const void *src;
	uint64_t src_size;           /* bytes behind src */
	void *dst;
	uint64_t dst_size;
	void *work;                  /* host models in/out, or NULL */
	uint64_t work_size;
	const struct airs_job *jobs;
	uint32_t *results;
	uint32_t *init_results;      /* may be NULL */
	uint64_t *out_offsets;       /* CONCAT only */
	uint32_t n_jobs;
	uint32_t n_results;
	uint32_t layout;
	uint32_t flags;              /* AIRS_BATCH_* */
};
int airs_cuda_compress_batch_host(const struct airs_host_batch *batch);

/*
 * Residual statistics per job, for choosing encoder parameters (the reference leaves them to
 * the user, lib/cmp.h:64-112).  For the first frame of every job, under the job's primary
 * preprocessing (NONE: the samples; anything else: first differences, ref
 * lib/compress/preprocess.c:268-290): sum, maximum and a log2 histogram of the zig-zag mapped
 * residuals (ref map_to_unsigned, lib/compress/encoder.c:274-286).  Bin 0 counts the value 0, bin b
 * the values 2^(b-1) .. 2^b - 1.  src, jobs and stats are device pointers; asynchronous on `stream`.
 */
struct airs_stats {
	uint64_t sum_mapped;
	uint32_t n_samples;
	uint32_t max_mapped;
	uint32_t log2_hist[17];
	uint32_t reserved;
};
int airs_cuda_residual_stats(const void *src, const struct airs_job *jobs, uint32_t n_jobs,
			     struct airs_stats *stats, void *stream);

/* Golomb parameter g for a mean mapped residual sum / n: the code is shortest near g = mean * ln 2.
 * Plain integer arithmetic (ln 2 = 45426 / 65536), 1 <= g <= 65535.  Host function. */
uint32_t airs_cuda_golomb_param_for_mean(uint64_t sum_mapped, uint32_t n_samples);

/*
 * Parameter search.  The reference leaves primary_encoder_param / primary_encoder_outlier to the user (what it
 * derives itself is the zero-escape outlier and the upper bound of the multi-escape one, lib/compress/encoder.c:
 * 154-182, 205-216).  airs_cuda_candidate_bits() computes, for the first frame of every job under the job's primary
 * preprocessing (NONE: the samples; anything else: first differences - exact for NONE and DIFF), the EXACT number of
 * code bits the frame takes with each of n_cand candidate encoders (at most AIRS_MAX_CANDIDATES), in one pass over the
 * samples: bits[j * n_cand + c].  Compressed with candidate c the frame is CMP_HDR_SIZE + 6 + ceil(bits / 8) (+ 4 with
 * checksum) bytes long.  A candidate cmp_initialise() would refuse has bits = UINT64_MAX; GOLOMB_ZERO ignores
 * `outlier`.  src, jobs, cand and bits are device pointers; asynchronous on `stream`.
 */
#define AIRS_MAX_CANDIDATES 32u
struct airs_candidate {
	uint32_t encoder_type; /* enum cmp_encoder_type */
	uint32_t g;            /* encoder_param */
	uint32_t outlier;      /* encoder_outlier */
	uint32_t reserved;
};
int airs_cuda_candidate_bits(const void *src, const struct airs_job *jobs, uint32_t n_jobs,
			     const struct airs_candidate *cand, uint32_t n_cand, uint64_t *bits, void *stream);

/* Candidates around what the statistics of a job suggest, for airs_cuda_candidate_bits(): Golomb parameters from
 * half to twice g = mean * ln 2, and for GOLOMB_MULTI outliers of 4, 8 and 16 g (where valid).  Fills at most
 * `max` entries of `out` and returns their number.  Host function. */
uint32_t airs_cuda_param_candidates(const struct airs_stats *stats, uint32_t encoder_type, struct airs_candidate *out,
				    uint32_t max);

/* Release cached device staging buffers of this thread (optional). */
void airs_cuda_release_cache(void);

#ifdef __cplusplus
}
#endif

#endif /* AIRS_CUDA_H */
