/* airs_launch.h - what the C-ABI layer (airs_cuda_api.cu) hands to the kernels. */
#ifndef AIRS_LAUNCH_H
#define AIRS_LAUNCH_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/airs_cuda.h"

struct airs_ctx_state;
struct JobPlan;

#ifndef AIRS_THREADS
#define AIRS_THREADS 128
#endif
#ifndef AIRS_CTAS_PER_SM
#define AIRS_CTAS_PER_SM 5 /* resident CTAs per SM the encode kernel is compiled for: 96 registers a thread.  (Six CTAs of 80
			    * registers were the choice until the runs of model passes came: with four segments per thread and
			    * visit they need the registers - config 2 batched 0.464 -> 0.504 - and the other workloads of this
			    * kernel stay within 1 - 4 % either way: measured) */
#endif

#ifndef AIRS_FAST_THREADS
#define AIRS_FAST_THREADS 128
#endif
#ifndef AIRS_FAST_CTAS_PER_SM
#define AIRS_FAST_CTAS_PER_SM 6 /* resident CTAs per SM airs_fast_kernel is compiled for */
#endif

#ifndef AIRS_TILE_THREADS
#define AIRS_TILE_THREADS 96 /* three warps with two staging areas of a whole tile each: 37 KB a CTA */
#endif
#ifndef AIRS_TILE_CTAS_PER_SM
#define AIRS_TILE_CTAS_PER_SM 5 /* 128 registers; with six (96 registers) the inlined warp encoders spill: measured slower */
#endif
#define AIRS_TICKET_INVALID 6u
#define AIRS_TICKET_RAW 8u /* 64-bit: jobs of airs_raw_kernel << 40 | their chunks */
#define AIRS_RAW_CHUNK 4096u /* 8-byte groups of stream per chunk */
#define AIRS_TICKET_IWT 14u /* 64-bit: jobs whose transform airs_iwt_kernel computes << 40 | their tiles */
#define AIRS_IWT_TILE 8192u /* samples of a tile of airs_iwt_kernel */
#define AIRS_IWT_MAX_SAMPLES (1u << 21) /* longest frame of that kernel: what is left behind its eight levels fits one CTA */
#define AIRS_TICKET_TILE_SHAPE 20u /* four words: ~min and max of the tile jobs' tiles per frame, ~min and max of their frames */
#define AIRS_TILE_RING 65536u /* tile descriptors kept (a power of two, far more than tiles are in flight) */
/* 64-bit words of the rings: tile descriptors, tile tails, blocks of 32 tiles, blocks of 1024 tiles */
#define AIRS_TILE_RING_BYTES (8u * (2u * AIRS_TILE_RING + AIRS_TILE_RING / 32u + AIRS_TILE_RING / 1024u))

struct AirsLaunch {
	const uint8_t *src;
	uint8_t *dst;
	uint8_t *work;
	const struct airs_job *jobs;
	uint32_t *results;
	uint32_t *init_results;
	uint64_t *out_offsets;
	uint32_t *ticket;      /* zeroed before the launch: [0] next entry of big_list, [1] of small_list,
				  [2] entries in big_list, [3] entries in small_list,
				  [4] jobs with checksum, [5] gate of the two-phase CONCAT path,
				  [6] set by airs_plan_kernel when the job table breaks the contract (AIRS_TICKET_INVALID),
				  [10..11] 64-bit: jobs of airs_tile_kernel << 40 | their tiles, [12] next tile */
	uint32_t *big_list;    /* job indices for airs_encode_kernel, filled by airs_plan_kernel */
	uint32_t *small_list;  /* job indices of airs_fast_kernel's records, in their order (SLOTS layout only) */
	void *fast_jobs;       /* n_jobs 64-byte records (FastJob, airs_fast.cuh): the short jobs of airs_fast_kernel from
				  the front in the order of small_list, the long jobs of airs_tile_kernel from the back */
	void *tile_ext;        /* n_jobs 64-byte records: TileExt of the tile jobs from the back (like their FastJob records), FastJob
				  records of the jobs of airs_raw_kernel from the front */
	void *iwt_recs;        /* n_jobs 32-byte records (IwtRec, airs_fast.cuh): the frames airs_iwt_kernel transforms into their
				  work buffers in front of the warp encoders */
	uint64_t *tile_ring;   /* AIRS_TILE_RING_BYTES, zeroed before the launch: tile descriptors, tile tails, block sums */
	uint32_t *result_job;  /* n_results entries: the job a frame belongs to (airs_checksum_kernel) */
	struct JobPlan *plans; /* n_jobs plans written by airs_plan_kernel */
	uint64_t *lookback;    /* CONCAT: one status word per job, zeroed before the launch */
	struct airs_ctx_state *ctx_io; /* host-shim path: context state in/out per job, else NULL */
	uint64_t dst_size;
	uint32_t n_jobs;
	uint32_t n_results;
	uint32_t layout;
	/* two-phase CONCAT (airs_cuda_api.cu): the kernels of a phase return at once unless
	 * (*gate != 0) == gate_want; NULL: no gate */
	const uint32_t *gate;
	uint32_t gate_want;
	uint32_t tile_below_jobs; /* batches with fewer jobs send every long single-frame job to airs_tile_kernel */
	uint32_t be_batch;        /* AIRS_BATCH_BIG_ENDIAN: jobs may have AIRS_DTYPE_BE containers */
	uint32_t ordered;         /* the caller's layout is CONCAT: the jobs must list the frames 0 .. n_results - 1 in order */
};

/* what the two-phase CONCAT path adds: temporary slots, the job table rewritten onto them */
struct AirsConcat {
	const struct airs_job *jobs;  /* the caller's jobs */
	struct airs_job *slot_jobs;   /* n_jobs copies whose dst_offset / dst_frame_stride address the temporary slots */
	const uint32_t *results;
	const uint32_t *result_job;   /* the job of every frame (airs_plan_kernel) */
	uint64_t *out_offsets;
	uint8_t *tmp;
	uint8_t *dst;
	uint32_t *flag;               /* set when the path has to be abandoned: tmp or dst too small; flag[1] is the
					 batch's AIRS_TICKET_INVALID word */
	uint64_t *sums;               /* airs_concat_scratch_bytes(): tile sums of the scans ... */
	uint32_t *big_list;           /* ... the frames whose streams a whole CTA copies ... */
	uint32_t *n_big;              /* ... and their number */
	uint64_t tmp_size;
	uint64_t dst_size;
	uint32_t n_jobs;
	uint32_t n_results;
	/* a slice of a larger batch (pipelined host batches): the jobs' first_result counts from result_base (results,
	 * result_job and out_offsets start at the slice's first frame) and the slice's streams start at *base, the
	 * total the slice before left in its out_offsets[n_results]; 0 / NULL: a whole batch */
	uint32_t result_base;
	const uint64_t *base;
};

#ifdef __cplusplus
extern "C" {
#endif
cudaError_t airs_launch_plan(const struct AirsLaunch *b, cudaStream_t stream);
cudaError_t airs_launch_encode(const struct AirsLaunch *b, unsigned int grid, cudaStream_t stream);
cudaError_t airs_encode_ctas_per_sm(int *out);
cudaError_t airs_launch_fast(const struct AirsLaunch *b, unsigned int grid, cudaStream_t stream);
cudaError_t airs_fast_resident_ctas(int *out);
cudaError_t airs_launch_tile(const struct AirsLaunch *b, cudaStream_t stream);
cudaError_t airs_launch_iwt(const struct AirsLaunch *b, cudaStream_t stream);
cudaError_t airs_launch_raw(const struct AirsLaunch *b, unsigned int grid, cudaStream_t stream);
cudaError_t airs_launch_checksum(const struct AirsLaunch *b, cudaStream_t stream);
cudaError_t airs_launch_hash(const struct AirsLaunch *b, uint64_t *hashes, cudaStream_t stream);
cudaError_t airs_launch_hash_ranges(const uint8_t *base, const uint64_t *offsets, const uint32_t *sizes, uint32_t n,
				    uint64_t *hashes, cudaStream_t stream);
size_t airs_concat_scratch_bytes(uint32_t n_jobs, uint32_t n_results);
cudaError_t airs_launch_concat_slots(const struct AirsConcat *c, cudaStream_t stream);
cudaError_t airs_launch_concat_gather(const struct AirsConcat *c, unsigned int grid, cudaStream_t stream);
#ifdef __cplusplus
}
#endif

#endif
