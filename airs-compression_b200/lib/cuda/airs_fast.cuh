/*
 * airs_fast.cuh - what airs_plan_kernel hands to airs_fast_kernel (airs_fast.cu): one
 * 64-byte record per short single-frame job, in ticket order, holding everything the
 * warp that encodes the job needs (no second look at the 120-byte job descriptor or
 * the 128-byte plan).
 */
#ifndef AIRS_FAST_CUH
#define AIRS_FAST_CUH

#include "airs_plan.cuh"

/* flags of a FastJob */
#define AIRS_FJ_PRE_DIFF    1u  /* DIFF preprocessing (else NONE) */
#define AIRS_FJ_MULTI       2u  /* GOLOMB_MULTI (else GOLOMB_ZERO) */
#define AIRS_FJ_CHECKSUM    4u
#define AIRS_FJ_FALLBACK_OK 8u
#define AIRS_FJ_BE          16u /* samples big-endian in memory (AIRS_DTYPE_BE) */
#define AIRS_FJ_IWT         32u /* src is the job's work buffer, holding the IWT coefficients airs_iwt_kernel left there;
				   they are coded like samples without preprocessing, the header says IWT */

/* samples of a tile of airs_tile_kernel */
#define AIRS_TILE_SAMPLES 1024u /* two units of 512 samples */

/* largest Golomb parameter of the multiply-high division used by the fast kernels (below) */
#define AIRS_FAST_MAX_G 32767u

struct alignas(16) FastJob {
	uint64_t src;          /* device address of the frame's samples (16-byte aligned, 16-bit container) */
	uint64_t dst;          /* device address of its slot (8-byte aligned) */
	uint64_t identifier;   /* of the primary pass: identifier_base + 1 (ref cmp.c:208,229) */
	uint32_t n;            /* samples */
	uint32_t cap_eff;      /* bytes the pass may write */
	uint32_t flags;        /* AIRS_FJ_* | floor(log2 g) << 8 */
	uint32_t first_result;
	uint32_t g;
	uint32_t outlier;      /* derived outlier (header field, encoder.c:185-224) */
	uint32_t magic;        /* airs_fast_magic(g) */
	uint32_t job;          /* index of the job in the batch */
	uint32_t tile_base;    /* airs_tile_kernel: global id of the job's first tile ... */
	uint32_t n_tiles;      /* ... and the number of its tiles */
};
static_assert(sizeof(FastJob) == 64, "FastJob is read as 16 words");

/* a frame whose wavelet transform (ref preprocess.c:140-221) airs_iwt_kernel computes: tiles of AIRS_IWT_TILE samples,
 * numbered over all such frames of the batch */
struct alignas(16) IwtRec {
	uint64_t src;       /* device address of the samples (16-bit container, host order, 4-byte aligned) */
	uint64_t work;      /* device address of the work buffer the coefficients go to */
	uint32_t n;         /* samples */
	uint32_t tile_base; /* global id of the frame's first tile ... */
	uint32_t n_tiles;   /* ... and the number of its tiles */
	uint32_t job;
};
static_assert(sizeof(IwtRec) == 32, "IwtRec is read as two 16-byte words");

/* flags2 of a TileExt: the secondary passes of a context whose frames go through airs_tile_kernel */
#define AIRS_TX_PRE2_DIFF  1u
#define AIRS_TX_MULTI2     2u
#define AIRS_TX_PRE2_MODEL 4u
#define AIRS_TX_MODEL      8u  /* the context keeps a model (AIRS_PF_MODEL): primary passes store the samples as model */
#define AIRS_TX_SIGNED    16u  /* i16 container: the model update sign-extends */
/* ... | floor(log2 g2) << 8 | model_rate << 16 */

/* What a tile job has beyond its FastJob: the 16 words behind it in a tile's record (lanes 16-31).  A context
 * of n_frames frames has n_frames * tiles_per_frame tiles, frame by frame. */
struct alignas(16) TileExt {
	uint64_t src_frame_stride;
	uint64_t dst_frame_stride;
	uint64_t model;          /* device address of the context's model (16-byte aligned), 0: none */
	uint32_t n_frames;
	uint32_t tiles_per_frame;
	uint32_t g2, outlier2, magic2; /* secondary encoder */
	uint32_t flags2;         /* AIRS_TX_* */
	uint32_t sec_iter;       /* secondary passes behind every primary one */
	uint32_t pad[3];
};
static_assert(sizeof(TileExt) == 64, "TileExt is read as 16 words");

/*
 * floor(x / g) for 0 <= x <= 65536 + g and 1 <= g <= AIRS_FAST_MAX_G as one multiply-high:
 *   g >= 2: M = ceil(2^32 / g), q = umulhi(x, M).  With e = M g - 2^32 (0 <= e < g) the result is
 *           exact while x e < 2^32, and (65536 + g) g < 2^32 for g < 2^15.
 *   g == 1: M = 2^32 - 1 and the caller adds 1 to x: umulhi(x + 1, 2^32 - 1) = x.
 * tests/test_abi.py checks every g against every dividend at and next to the multiples of g.
 */
__host__ __device__ inline uint32_t airs_fast_magic(uint32_t g)
{
	return g <= 1u ? 0xFFFFFFFFu : 0xFFFFFFFFu / g + 1u; /* ceil(2^32 / g) */
}

#ifdef __CUDACC__
__device__ inline void airs_fill_fast_job(FastJob &f, const airs_job &job, const JobPlan &pl, const uint8_t *src_base,
					  uint8_t *dst_base, uint32_t job_index, uint32_t tile_base, uint32_t n_tiles,
					  const uint8_t *iwt_work = nullptr)
{
	/* (iwt_work: the coefficients in the job's work buffer take the place of the samples) */
	f.src = iwt_work ? (uint64_t)(uintptr_t)iwt_work : (uint64_t)(uintptr_t)(src_base + job.src_offset);
	f.dst = (uint64_t)(uintptr_t)(dst_base + job.dst_offset);
	f.identifier = (job.identifier_base + 1u) & 0xFFFFFFFFFFFFull;
	f.n = pl.n;
	f.cap_eff = pl.cap_eff;
	f.flags = (pl.pre[0] == CMP_PREPROCESS_DIFF ? AIRS_FJ_PRE_DIFF : 0u) |
		  (pl.enc[0].type == CMP_ENCODER_GOLOMB_MULTI ? AIRS_FJ_MULTI : 0u) |
		  ((pl.flags & AIRS_PF_CHECKSUM) ? AIRS_FJ_CHECKSUM : 0u) |
		  ((pl.flags & AIRS_PF_FALLBACK_OK) ? AIRS_FJ_FALLBACK_OK : 0u) | ((pl.flags & AIRS_PF_BE) ? AIRS_FJ_BE : 0u) |
		  (iwt_work ? AIRS_FJ_IWT : 0u) | (pl.enc[0].L << 8);
	f.first_result = job.first_result;
	f.g = pl.enc[0].g;
	f.outlier = pl.enc[0].outlier;
	f.magic = airs_fast_magic(pl.enc[0].g);
	f.job = job_index;
	f.tile_base = tile_base;
	f.n_tiles = n_tiles;
}

__device__ inline void airs_fill_tile_ext(TileExt &t, const airs_job &job, const JobPlan &pl, uint8_t *work_base, uint32_t frame_tiles)
{
	const bool model = (pl.flags & AIRS_PF_MODEL) != 0u;
	t.src_frame_stride = job.src_frame_stride;
	t.dst_frame_stride = job.dst_frame_stride;
	t.model = model ? (uint64_t)(uintptr_t)(work_base + job.work_offset) : 0u;
	t.n_frames = job.n_frames;
	t.tiles_per_frame = frame_tiles;
	t.g2 = pl.enc[1].g;
	t.outlier2 = pl.enc[1].outlier;
	t.magic2 = airs_fast_magic(pl.enc[1].g);
	t.flags2 = (pl.pre[1] == CMP_PREPROCESS_DIFF ? AIRS_TX_PRE2_DIFF : 0u) |
		   (pl.enc[1].type == CMP_ENCODER_GOLOMB_MULTI ? AIRS_TX_MULTI2 : 0u) |
		   (pl.pre[1] == CMP_PREPROCESS_MODEL ? AIRS_TX_PRE2_MODEL : 0u) | (model ? AIRS_TX_MODEL : 0u) |
		   ((pl.flags & AIRS_PF_SIGNED) ? AIRS_TX_SIGNED : 0u) | (pl.enc[1].L << 8) | (pl.rate << 16);
	t.sec_iter = pl.sec_iter;
	t.pad[0] = t.pad[1] = t.pad[2] = 0;
}
#endif

#endif /* AIRS_FAST_CUH */
