/*
 * airs_cuda_api.cu - the extern "C" boundary of the sm_100a backend
 * (include/airs_cuda.h) and the staging helper of the host shim.
 *
 * No CPU fallback lives here or anywhere else in this library: when no CUDA
 * device is usable every entry point fails with AIRS_E_NO_DEVICE.
 */
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "airs_device.cuh"
#include "airs_launch.h"
#include "airs_private.h"

namespace {

thread_local char g_err[512] = "";
thread_local int g_launches = 0;

int fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

#define CU(call)                                                                              \
	do {                                                                                  \
		cudaError_t e_ = (call);                                                      \
		if (e_ != cudaSuccess)                                                        \
			return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? \
					    AIRS_E_NO_DEVICE : AIRS_E_CUDA,                   \
				    "%s: %s", #call, cudaGetErrorString(e_));                 \
	} while (0)

constexpr size_t kScratchHeader = 512; /* ticket counters and friends: words 0..63 of the batch (word 5: the gate
					* of the two-phase CONCAT path), words 64..127 of its single-phase rerun */
constexpr uint32_t kGateWord = 5;

size_t lookback_bytes(uint32_t n_results)
{
	return (8 * ((size_t)n_results + 1) + 127) & ~(size_t)127;
}

/* CTAs that fit the device at once for the persistent job loop */
int resident_ctas(int *out)
{
	static thread_local int cached_dev = -1, cached = 0;
	int dev;

	CU(cudaGetDevice(&dev));
	if (dev != cached_dev) {
		cudaDeviceProp prop;
		CU(cudaGetDeviceProperties(&prop, dev));
		if (prop.major < 10)
			return fail(AIRS_E_NO_DEVICE, "device %d is sm_%d%d; this library holds sm_100a code only",
				    dev, prop.major, prop.minor);
		/* resident CTAs per SM as the occupancy calculator sees them (the launch bounds of
		 * airs_encode_kernel promise AIRS_CTAS_PER_SM); the grid is a multiple of the SM count */
		int per_sm = 0;
		CU(airs_encode_ctas_per_sm(&per_sm));
		if (per_sm < 1)
			return fail(AIRS_E_CUDA, "the encode kernel does not fit an SM of device %d", dev);
		cached = prop.multiProcessorCount * per_sm;
		cached_dev = dev;
	}
	*out = cached;
	return AIRS_OK;
}

/* grow-only device buffer owned by the calling thread */
struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	int reserve(size_t n)
	{
		if (n <= cap)
			return AIRS_OK;
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
		size_t want = n + n / 4 + 4096;
		cudaError_t e = cudaMalloc(&p, want);
		if (e != cudaSuccess)
			return fail(AIRS_E_NOMEM, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
		cap = want;
		return AIRS_OK;
	}
	void release()
	{
		if (p)
			cudaFree(p);
		p = nullptr;
		cap = 0;
	}
};

constexpr int kMaxGroups = 64; /* pipeline stages of one host batch */
#ifndef AIRS_HOST_GROUP_BYTES
#define AIRS_HOST_GROUP_BYTES (32u << 20) /* source bytes of a stage of the pipelined CONCAT path */
#endif

/* Everything a thread keeps between calls lives on ONE device (stream_dev): when the thread's current device
 * has changed, cache_stream() gives all of it back under the old device before anything is created on the new one. */
struct Cache {
	DevBuf src, dst, work, jobs, results, init, offs, scratch, state, tmp;
	cudaStream_t stream = nullptr;            /* compute (and everything of the unpipelined paths) */
	cudaStream_t s_in = nullptr, s_out = nullptr; /* host-to-device / device-to-host copies of the pipelined path */
	cudaEvent_t ev_in[kMaxGroups] = {}, ev_done[kMaxGroups] = {}, ev_start = nullptr;
	uint64_t *pin = nullptr;                  /* pinned host memory: 2 words per group (pipelined CONCAT) */
	int stream_dev = -1;
};
thread_local Cache g_cache;

int cache_stream(cudaStream_t *s)
{
	int dev;
	CU(cudaGetDevice(&dev));
	if (g_cache.stream_dev >= 0 && g_cache.stream_dev != dev) {
		CU(cudaSetDevice(g_cache.stream_dev));
		airs_cuda_release_cache();
		CU(cudaSetDevice(dev));
	}
	if (!g_cache.stream) {
		CU(cudaHostAlloc((void **)&g_cache.pin, 2 * kMaxGroups * sizeof(uint64_t), cudaHostAllocDefault));
		CU(cudaStreamCreateWithFlags(&g_cache.stream, cudaStreamNonBlocking));
		CU(cudaStreamCreateWithFlags(&g_cache.s_in, cudaStreamNonBlocking));
		CU(cudaStreamCreateWithFlags(&g_cache.s_out, cudaStreamNonBlocking));
		CU(cudaEventCreateWithFlags(&g_cache.ev_start, cudaEventDisableTiming));
		for (int i = 0; i < kMaxGroups; i++) {
			CU(cudaEventCreateWithFlags(&g_cache.ev_in[i], cudaEventDisableTiming));
			CU(cudaEventCreateWithFlags(&g_cache.ev_done[i], cudaEventDisableTiming));
		}
		g_cache.stream_dev = dev;
	}
	*s = g_cache.stream;
	return AIRS_OK;
}

bool on_device(const void *p)
{
	cudaPointerAttributes at;
	if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

} /* namespace */

/* for the other translation units of the library (airs_decode.cu); not exported */
extern "C" __attribute__((visibility("hidden"))) int airs_internal_fail(int code, const char *text)
{
	return fail(code, "%s", text);
}

extern "C" __attribute__((visibility("hidden"))) void airs_internal_set_launches(int n)
{
	g_launches = n;
}

extern "C" __attribute__((visibility("hidden"))) int airs_internal_check_device(void)
{
	int n = 0;
	return resident_ctas(&n);
}

extern "C" const char *airs_cuda_last_error(void)
{
	return g_err;
}

extern "C" int airs_cuda_last_launch_count(void)
{
	return g_launches;
}

extern "C" int airs_cuda_device_count(void)
{
	int n = 0, usable = 0;
	cudaError_t e = cudaGetDeviceCount(&n);

	if (e != cudaSuccess) {
		fail(AIRS_E_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
		cudaGetLastError();
		return 0;
	}
	for (int d = 0; d < n; d++) {
		int major = 0;
		if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major >= 10)
			usable++;
	}
	if (!usable)
		fail(AIRS_E_NO_DEVICE, "no sm_100 device among %d CUDA devices", n);
	return usable;
}

extern "C" int airs_cuda_concurrent_jobs(void)
{
	int n = 0;
	return resident_ctas(&n) == AIRS_OK ? n : 0;
}

extern "C" size_t airs_cuda_batch_scratch_size(uint32_t n_jobs, uint32_t n_results)
{
	/* header | one look-back word per frame (+1) | one 128-byte plan per job | two job lists | job of every frame |
	 * the job table on temporary slots (two-phase CONCAT) */
	return kScratchHeader + lookback_bytes(n_results) + 128 * (size_t)n_jobs + 8 * (size_t)n_jobs +
	       4 * (size_t)n_results + 64 + sizeof(struct airs_job) * (size_t)n_jobs + 64 +
	       airs_concat_scratch_bytes(n_jobs, n_results) + 64 + 128 * (size_t)n_jobs + 128 + (size_t)AIRS_TILE_RING_BYTES +
	       32 * (size_t)n_jobs + 64; /* (the frames of airs_iwt_kernel) */
}

extern "C" size_t airs_cuda_concat_tmp_size(uint64_t sum_of_capacities, uint32_t n_results)
{
	/* every slot is rounded up to 16 bytes; 32 bytes of slack behind the last one */
	return (size_t)(sum_of_capacities + 16ull * n_results + 64);
}

/* plan; short jobs (one warp each) and the tiles of long frames (SLOTS only); everything else, and what the
 * two fast kernels handed back, one CTA per job; checksums */
static int launch_kernels(const AirsLaunch &l_in, int resident, cudaStream_t stream)
{
	AirsLaunch l = l_in;
	unsigned int grid = l.n_jobs < (uint32_t)resident ? l.n_jobs : (unsigned int)resident;
	/* Crossover of the two kernels on long frames, measured on 4 MiB chunks: the CTA-per-job kernel takes 0.56 ms for
	 * up to one chunk per SM (0.098 / 0.195 / 0.256 / 0.338 of the roofline at 64 / 128 / 192 / 256 jobs), the tile kernel
	 * 0.154 - 0.171 whatever the count: about 110 jobs */
	l.tile_below_jobs = (uint32_t)resident * 3u / 20u;
#ifdef AIRS_DEV_ENV /* (development builds: measure the crossover with AIRS_TILE_BELOW=<jobs>) */
	if (const char *e = getenv("AIRS_TILE_BELOW"))
		l.tile_below_jobs = (uint32_t)atoi(e);
#endif
	CU(airs_launch_plan(&l, stream));
	g_launches++;
	if (l.layout == AIRS_LAYOUT_SLOTS && !l.ctx_io) {
		static thread_local int fast_dev = -1, fast_ctas = 0;
		int dev;
		CU(cudaGetDevice(&dev));
		if (dev != fast_dev) {
			CU(airs_fast_resident_ctas(&fast_ctas));
			fast_dev = dev;
		}
		if (l.work) { /* (the coefficients the two kernels below code go to work buffers: without any, no such frames) */
			CU(airs_launch_iwt(&l, stream));
			g_launches += 2;
		}
		const unsigned int want = (l.n_jobs + AIRS_FAST_THREADS / 32 - 1) / (AIRS_FAST_THREADS / 32);
		CU(airs_launch_fast(&l, want < (unsigned int)fast_ctas ? want : (unsigned int)fast_ctas, stream));
		CU(airs_launch_tile(&l, stream));
#ifndef AIRS_SKIP_RAW /* (development: the cost of the launch) */
		CU(airs_launch_raw(&l, (unsigned int)resident / AIRS_CTAS_PER_SM * 8u, stream)); /* (eight CTAs of 256 threads per SM: enough loads in flight) */
#endif
		g_launches += 4;
	}
	CU(airs_launch_encode(&l, grid, stream));
	CU(airs_launch_checksum(&l, stream));
	g_launches += 2;
	return AIRS_OK;
}

/* the batch as the kernels see it: where everything lies in the scratch memory */
static void fill_launch(AirsLaunch &l, const struct airs_batch *b, struct airs_ctx_state *ctx_io)
{
	memset(&l, 0, sizeof(l));
	l.src = (const uint8_t *)b->src;
	l.dst = (uint8_t *)b->dst;
	l.work = (uint8_t *)b->work;
	l.jobs = b->jobs;
	l.results = b->results;
	l.init_results = b->init_results;
	l.out_offsets = b->out_offsets;
	l.ticket = (uint32_t *)b->scratch;
	l.lookback = (uint64_t *)((uint8_t *)b->scratch + kScratchHeader);
	l.plans = (struct JobPlan *)((uint8_t *)b->scratch + kScratchHeader + lookback_bytes(b->n_results));
	l.big_list = (uint32_t *)((uint8_t *)l.plans + 128 * (size_t)b->n_jobs);
	l.small_list = l.big_list + b->n_jobs;
	l.result_job = l.small_list + b->n_jobs;
	{ /* from the end of the scratch memory: tile rings, in front of them the tile extensions, then the fast-job records */
		const size_t end = airs_cuda_batch_scratch_size(b->n_jobs, b->n_results);
		const size_t iwt = (end - 32 * (size_t)b->n_jobs) & ~(size_t)63;
		l.iwt_recs = (uint8_t *)b->scratch + iwt;
		const size_t ring = (iwt - (size_t)AIRS_TILE_RING_BYTES) & ~(size_t)63;
		l.tile_ring = (uint64_t *)((uint8_t *)b->scratch + ring);
		const size_t ext = (ring - 64 * (size_t)b->n_jobs) & ~(size_t)63;
		l.tile_ext = (uint8_t *)b->scratch + ext;
		l.fast_jobs = (uint8_t *)b->scratch + ((ext - 64 * (size_t)b->n_jobs) & ~(size_t)63);
	}
	l.ctx_io = ctx_io;
	l.dst_size = b->dst_size;
	l.n_jobs = b->n_jobs;
	l.n_results = b->n_results;
	l.layout = b->layout;
	l.ordered = b->layout == AIRS_LAYOUT_CONCAT;
	l.be_batch = (b->flags & AIRS_BATCH_BIG_ENDIAN) != 0u;
}

/* slice: the batch is a run of consecutive jobs of a larger CONCAT batch whose first frame is number
 * slice->result_base of the whole (results / out_offsets point at that frame) and whose streams start at
 * *slice->base.  Two-phase path only; if the gate closes nothing is written and the caller starts over. */
struct ConcatSlice {
	uint32_t result_base;
	const uint64_t *base;
};

static int launch_batch(const struct airs_batch *b, struct airs_ctx_state *ctx_io, cudaStream_t stream,
			const ConcatSlice *slice = nullptr)
{
	g_launches = 0;
	if (!b || !b->jobs || !b->results || !b->scratch)
		return fail(AIRS_E_ARGUMENT, "batch, jobs, results and scratch must be non-NULL");
	if (b->layout != AIRS_LAYOUT_SLOTS && b->layout != AIRS_LAYOUT_CONCAT)
		return fail(AIRS_E_ARGUMENT, "unknown layout %u", b->layout);
	if (b->layout == AIRS_LAYOUT_CONCAT && !b->out_offsets)
		return fail(AIRS_E_ARGUMENT, "the CONCAT layout needs out_offsets");
	if (b->n_jobs == 0)
		return AIRS_OK;

	int resident = 0;
	int rc = resident_ctas(&resident);
	if (rc != AIRS_OK)
		return rc;

	if ((uintptr_t)b->scratch & 127u)
		return fail(AIRS_E_ARGUMENT, "scratch must be 128-byte aligned");
	CU(cudaMemsetAsync(b->scratch, 0, kScratchHeader + lookback_bytes(b->n_results), stream));

	AirsLaunch l;
	fill_launch(l, b, ctx_io);
	CU(cudaMemsetAsync(l.result_job, 0xFF, 4 * (size_t)b->n_results, stream));
	if (b->layout == AIRS_LAYOUT_SLOTS || b->tmp)
		CU(cudaMemsetAsync(l.tile_ring, 0, (size_t)AIRS_TILE_RING_BYTES, stream));

	if (b->layout == AIRS_LAYOUT_CONCAT && b->tmp && b->tmp_size && b->dst && !ctx_io && !((uintptr_t)b->tmp & 15u)) {
		/* two phases (airs_concat.cu): SLOTS-style into temporary slots, scan, copy.  Everything is
		 * enqueued at once; word kGateWord of the header decides on the device which kernels run:
		 * 0 = the two-phase path, set = the single-phase path below (temporary slots or
		 * destination too small). */
		AirsConcat c;
		memset(&c, 0, sizeof(c));
		c.jobs = b->jobs;
		c.slot_jobs = (struct airs_job *)(((uintptr_t)(l.result_job + b->n_results) + 63u) & ~(uintptr_t)63u);
		c.sums = (uint64_t *)(((uintptr_t)(c.slot_jobs + b->n_jobs) + 63u) & ~(uintptr_t)63u);
		c.n_big = (uint32_t *)(c.sums + (b->n_jobs + 1023u) / 1024u + (b->n_results + 1023u) / 1024u + 4);
		c.big_list = c.n_big + 4;
		c.results = b->results;
		c.result_job = l.result_job;
		c.out_offsets = b->out_offsets;
		c.tmp = (uint8_t *)b->tmp;
		c.dst = (uint8_t *)b->dst;
		c.flag = l.ticket + kGateWord;
		c.tmp_size = b->tmp_size;
		c.dst_size = b->dst_size;
		c.n_jobs = b->n_jobs;
		c.n_results = b->n_results;
		if (slice) {
			c.result_base = slice->result_base;
			c.base = slice->base;
		}
		CU(airs_launch_concat_slots(&c, stream));
		g_launches += 3;

		AirsLaunch l1 = l;
		l1.jobs = c.slot_jobs;
		l1.dst = c.tmp;
		l1.dst_size = b->tmp_size;
		l1.out_offsets = nullptr;
		l1.layout = AIRS_LAYOUT_SLOTS;
		l1.gate = c.flag;
		l1.gate_want = 0;
		if ((rc = launch_kernels(l1, resident, stream)))
			return rc;
		/* one CTA per long stream, one warp per short one: as many CTAs as there are frames, at most
		 * a few per SM (latency bound copies: many warps in flight) */
		unsigned int cap = (unsigned int)resident * 2u;
		CU(airs_launch_concat_gather(&c, b->n_results < cap ? (b->n_results ? b->n_results : 1u) : cap, stream));
		g_launches += 4;
		if (slice)
			return AIRS_OK; /* (the caller looks at the gate word) */

		l.ticket = (uint32_t *)b->scratch + 64; /* fresh counters for the rerun */
		l.gate = c.flag;
		l.gate_want = 1;
	} else if (slice) {
		return fail(AIRS_E_ARGUMENT, "a slice needs the two-phase CONCAT path");
	}
	return launch_kernels(l, resident, stream);
}

extern "C" int airs_cuda_compress_batch(const struct airs_batch *b, void *stream)
{
	return launch_batch(b, nullptr, (cudaStream_t)stream);
}

extern "C" int airs_cuda_hash_streams(const struct airs_batch *b, uint64_t *hashes, void *stream)
{
	if (!b || !b->jobs || !b->results || !b->scratch || !hashes)
		return fail(AIRS_E_ARGUMENT, "hash: batch, jobs, results, scratch and hashes must be non-NULL");
	if (b->layout == AIRS_LAYOUT_CONCAT && !b->out_offsets)
		return fail(AIRS_E_ARGUMENT, "the CONCAT layout needs out_offsets");
	AirsLaunch l;
	fill_launch(l, b, nullptr);
	CU(airs_launch_hash(&l, hashes, (cudaStream_t)stream));
	return AIRS_OK;
}

extern "C" int airs_cuda_hash_ranges(const void *base, const uint64_t *offsets, const uint32_t *sizes, uint32_t n,
				     uint64_t *hashes, void *stream)
{
	if (n && (!base || !offsets || !sizes || !hashes))
		return fail(AIRS_E_ARGUMENT, "hash: base, offsets, sizes and hashes must be non-NULL");
	CU(airs_launch_hash_ranges((const uint8_t *)base, offsets, sizes, n, hashes, (cudaStream_t)stream));
	return AIRS_OK;
}

/*
 * SLOTS layout, large batch: the jobs are cut into groups of consecutive jobs
 * (about equal source bytes); group g's samples travel to the device while
 * group g-1 is encoded and the slots of group g-2 travel back, on three streams.
 * Each group is one airs_cuda_compress_batch launch over its slice of the job
 * table.  The device buffers are reserved by the caller.
 */
static int host_batch_pipelined(const struct airs_host_batch *hb, cudaStream_t s_comp)
{
	Cache &c = g_cache;
	const struct airs_job *jobs = hb->jobs;
	const uint64_t total_src = hb->src_size;
	int n_groups = (int)(total_src / (256u << 20)) + 1;
	if (n_groups < 4)
		n_groups = 4;
	if (n_groups > kMaxGroups)
		n_groups = kMaxGroups;
	if ((uint32_t)n_groups > hb->n_jobs)
		n_groups = (int)hb->n_jobs;

	/* everything the groups share goes first: job table, host models */
	CU(cudaMemcpyAsync(c.jobs.p, hb->jobs, (size_t)hb->n_jobs * sizeof(struct airs_job), cudaMemcpyHostToDevice, c.s_in));
	if (hb->work && hb->work_size)
		CU(cudaMemcpyAsync(c.work.p, hb->work, hb->work_size, cudaMemcpyHostToDevice, c.s_in));
	CU(cudaEventRecord(c.ev_start, c.s_in));
	CU(cudaStreamWaitEvent(s_comp, c.ev_start, 0));

	uint64_t bytes_total = 0;
	for (uint32_t j = 0; j < hb->n_jobs; j++)
		bytes_total += (uint64_t)jobs[j].src_size * jobs[j].n_frames;
	const uint64_t per_group = bytes_total / (uint64_t)n_groups + 1;

	uint32_t j0 = 0;
	int launches = 0;
	for (int g = 0; g < n_groups && j0 < hb->n_jobs; g++) {
		/* jobs of this group and the byte ranges they touch */
		uint64_t acc = 0, s_lo = ~0ull, s_hi = 0, d_lo = ~0ull, d_hi = 0;
		uint32_t j1 = j0;
		while (j1 < hb->n_jobs && (acc < per_group || g == n_groups - 1)) {
			const struct airs_job &jb = jobs[j1];
			const uint64_t nf = jb.n_frames;
			acc += (uint64_t)jb.src_size * nf;
			if (nf) {
				const uint64_t se = jb.src_offset + (nf - 1) * jb.src_frame_stride + jb.src_size;
				const uint64_t de = jb.dst_offset + (nf - 1) * jb.dst_frame_stride + jb.dst_capacity;
				if (jb.src_offset < s_lo)
					s_lo = jb.src_offset;
				if (se > s_hi)
					s_hi = se;
				if (jb.dst_offset < d_lo)
					d_lo = jb.dst_offset;
				if (de > d_hi)
					d_hi = de;
			}
			j1++;
		}
		if (s_hi > hb->src_size)
			s_hi = hb->src_size;
		if (d_hi > hb->dst_size)
			d_hi = hb->dst_size;
		if (s_lo < s_hi)
			CU(cudaMemcpyAsync((uint8_t *)c.src.p + s_lo, (const uint8_t *)hb->src + s_lo, s_hi - s_lo,
					   cudaMemcpyHostToDevice, c.s_in));
		CU(cudaEventRecord(c.ev_in[g], c.s_in));
		CU(cudaStreamWaitEvent(s_comp, c.ev_in[g], 0));

		struct airs_batch b;
		memset(&b, 0, sizeof(b));
		b.src = c.src.p;
		b.dst = c.dst.p;
		b.work = hb->work_size ? c.work.p : nullptr;
		b.jobs = (const struct airs_job *)c.jobs.p + j0;
		b.results = (uint32_t *)c.results.p;
		b.init_results = (uint32_t *)c.init.p + j0;
		b.scratch = c.scratch.p;
		b.dst_size = hb->dst_size;
		b.n_jobs = j1 - j0;
		b.n_results = hb->n_results;
		b.layout = AIRS_LAYOUT_SLOTS;
		b.flags = hb->flags;
		int rc = launch_batch(&b, nullptr, s_comp);
		if (rc)
			return rc;
		launches += g_launches;
		CU(cudaEventRecord(c.ev_done[g], s_comp));
		CU(cudaStreamWaitEvent(c.s_out, c.ev_done[g], 0));
		if (d_lo < d_hi)
			CU(cudaMemcpyAsync((uint8_t *)hb->dst + d_lo, (const uint8_t *)c.dst.p + d_lo, d_hi - d_lo,
					   cudaMemcpyDeviceToHost, c.s_out));
		j0 = j1;
	}
	g_launches = launches;
	/* results, init results and models once everything has been encoded (s_out is behind the last group) */
	CU(cudaMemcpyAsync(hb->results, c.results.p, (size_t)hb->n_results * 4, cudaMemcpyDeviceToHost, c.s_out));
	if (hb->init_results)
		CU(cudaMemcpyAsync(hb->init_results, c.init.p, (size_t)hb->n_jobs * 4, cudaMemcpyDeviceToHost, c.s_out));
	if (hb->work && hb->work_size)
		CU(cudaMemcpyAsync(hb->work, c.work.p, hb->work_size, cudaMemcpyDeviceToHost, c.s_out));
	CU(cudaStreamSynchronize(c.s_out));
	CU(cudaStreamSynchronize(s_comp));
	return AIRS_OK;
}

/*
 * CONCAT layout, large batch: groups of consecutive jobs as above, but every group is a slice of the whole
 * (two-phase path: temporary slots sized for the largest group, the scan of a group starts at the total the
 * group before left on the device) and only the bytes a group produced travel back - the host learns the
 * group's end from two words in pinned memory and issues the copy as soon as the group is done.
 * *redo is set when the batch has to take the unpipelined path after all (frames not numbered in job
 * order, no room for the temporary slots, a destination that overflows: only the single-phase path follows
 * the reference loop through that).
 */
static int host_batch_concat_pipelined(const struct airs_host_batch *hb, cudaStream_t s_comp, bool *redo)
{
	Cache &c = g_cache;
	const struct airs_job *jobs = hb->jobs;
	struct Group {
		uint32_t j0, j1, r0, r1;
		uint64_t s_lo, s_hi, caps;
	} grp[kMaxGroups];

	*redo = true;
	uint64_t bytes_total = 0, frames = 0;
	for (uint32_t j = 0; j < hb->n_jobs; j++) {
		if (jobs[j].n_frames && jobs[j].first_result != frames)
			return AIRS_OK;
		frames += jobs[j].n_frames;
		bytes_total += (uint64_t)jobs[j].src_size * jobs[j].n_frames;
	}
	if (frames != hb->n_results)
		return AIRS_OK;
	int n_groups = (int)(bytes_total / (size_t)AIRS_HOST_GROUP_BYTES) + 1; /* the first copy in and the last copy out are not overlapped */
	if (n_groups < 4)
		n_groups = 4;
	if (n_groups > kMaxGroups)
		n_groups = kMaxGroups;
	/* a group has to keep the device busy: contexts of many frames take a CTA each for as long as their frames last,
	 * however few of them a group holds */
	if ((uint32_t)n_groups > hb->n_jobs / 32u)
		n_groups = hb->n_jobs / 32u > 4u ? (int)(hb->n_jobs / 32u) : 4;
	if ((uint32_t)n_groups > hb->n_jobs)
		n_groups = (int)hb->n_jobs;
	const uint64_t per_group = bytes_total / (uint64_t)n_groups + 1;

	int ng = 0;
	size_t tmp_need = 0;
	uint32_t j0 = 0, r0 = 0;
	for (int g = 0; g < n_groups && j0 < hb->n_jobs; g++) {
		Group &G = grp[ng];
		uint64_t acc = 0;
		uint32_t j1 = j0, r1 = r0;
		G.s_lo = ~0ull;
		G.s_hi = 0;
		G.caps = 0;
		while (j1 < hb->n_jobs && (acc < per_group || g == n_groups - 1)) {
			const struct airs_job &jb = jobs[j1];
			const uint64_t nf = jb.n_frames;
			acc += (uint64_t)jb.src_size * nf;
			G.caps += nf * jb.dst_capacity;
			if (nf) {
				const uint64_t se = jb.src_offset + (nf - 1) * jb.src_frame_stride + jb.src_size;
				if (jb.src_offset < G.s_lo)
					G.s_lo = jb.src_offset;
				if (se > G.s_hi)
					G.s_hi = se;
			}
			r1 += jb.n_frames;
			j1++;
		}
		if (G.s_hi > hb->src_size)
			G.s_hi = hb->src_size;
		G.j0 = j0;
		G.j1 = j1;
		G.r0 = r0;
		G.r1 = r1;
		const size_t need = airs_cuda_concat_tmp_size(G.caps, r1 - r0);
		if (need > tmp_need)
			tmp_need = need;
		j0 = j1;
		r0 = r1;
		ng++;
	}
	if (tmp_need > c.tmp.cap) { /* (cudaMemGetInfo costs milliseconds: only when the slots have to grow) */
		size_t free_b = 0, total_b = 0;
		if (!(cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && tmp_need < free_b / 2 && c.tmp.reserve(tmp_need) == AIRS_OK)) {
			cudaGetLastError();
			return AIRS_OK;
		}
	}

	CU(cudaMemcpyAsync(c.jobs.p, hb->jobs, (size_t)hb->n_jobs * sizeof(struct airs_job), cudaMemcpyHostToDevice, c.s_in));
	if (hb->work && hb->work_size)
		CU(cudaMemcpyAsync(c.work.p, hb->work, hb->work_size, cudaMemcpyHostToDevice, c.s_in));
	CU(cudaEventRecord(c.ev_start, c.s_in));
	CU(cudaStreamWaitEvent(s_comp, c.ev_start, 0));

	int launches = 0;
	for (int g = 0; g < ng; g++) {
		const Group &G = grp[g];
		if (G.s_lo < G.s_hi)
			CU(cudaMemcpyAsync((uint8_t *)c.src.p + G.s_lo, (const uint8_t *)hb->src + G.s_lo, G.s_hi - G.s_lo,
					   cudaMemcpyHostToDevice, c.s_in));
		CU(cudaEventRecord(c.ev_in[g], c.s_in));
		CU(cudaStreamWaitEvent(s_comp, c.ev_in[g], 0));

		struct airs_batch b;
		memset(&b, 0, sizeof(b));
		b.src = c.src.p;
		b.dst = c.dst.p;
		b.work = hb->work_size ? c.work.p : nullptr;
		b.jobs = (const struct airs_job *)c.jobs.p + G.j0;
		b.results = (uint32_t *)c.results.p + G.r0;
		b.init_results = (uint32_t *)c.init.p + G.j0;
		b.out_offsets = (uint64_t *)c.offs.p + G.r0;
		b.scratch = c.scratch.p;
		b.dst_size = hb->dst_size;
		b.n_jobs = G.j1 - G.j0;
		b.n_results = G.r1 - G.r0;
		b.layout = AIRS_LAYOUT_CONCAT;
		b.flags = hb->flags;
		b.tmp = c.tmp.p;
		b.tmp_size = tmp_need;
		ConcatSlice sl = {G.r0, g ? (const uint64_t *)c.offs.p + G.r0 : nullptr};
		int rc = launch_batch(&b, nullptr, s_comp, &sl);
		if (rc)
			return rc;
		launches += g_launches;
		c.pin[2 * g + 1] = 1; /* (overwritten by the gate word) */
		CU(cudaMemcpyAsync(&c.pin[2 * g], (uint64_t *)c.offs.p + G.r1, 8, cudaMemcpyDeviceToHost, s_comp));
		CU(cudaMemcpyAsync(&c.pin[2 * g + 1], (uint32_t *)c.scratch.p + kGateWord, 4, cudaMemcpyDeviceToHost, s_comp));
		CU(cudaEventRecord(c.ev_done[g], s_comp));
	}
	g_launches = launches;

	bool ok = true;
	uint64_t lo = 0;
	for (int g = 0; g < ng; g++) {
		CU(cudaEventSynchronize(c.ev_done[g]));
		const uint64_t hi = c.pin[2 * g];
		if ((uint32_t)c.pin[2 * g + 1] != 0u || hi < lo || hi > hb->dst_size) {
			ok = false; /* the gate closed: the groups behind this one saw a base that means nothing */
			break;
		}
		if (hi > lo)
			CU(cudaMemcpyAsync((uint8_t *)hb->dst + lo, (const uint8_t *)c.dst.p + lo, hi - lo, cudaMemcpyDeviceToHost, c.s_out));
		lo = hi;
	}
	CU(cudaStreamSynchronize(s_comp));
	if (!ok) {
		CU(cudaStreamSynchronize(c.s_out));
		return AIRS_OK; /* *redo stays set */
	}
	CU(cudaMemcpyAsync(hb->results, c.results.p, (size_t)hb->n_results * 4, cudaMemcpyDeviceToHost, c.s_out));
	if (hb->init_results)
		CU(cudaMemcpyAsync(hb->init_results, c.init.p, (size_t)hb->n_jobs * 4, cudaMemcpyDeviceToHost, c.s_out));
	CU(cudaMemcpyAsync(hb->out_offsets, c.offs.p, ((size_t)hb->n_results + 1) * 8, cudaMemcpyDeviceToHost, c.s_out));
	if (hb->work && hb->work_size)
		CU(cudaMemcpyAsync(hb->work, c.work.p, hb->work_size, cudaMemcpyDeviceToHost, c.s_out));
	CU(cudaStreamSynchronize(c.s_out));
	*redo = false;
	return AIRS_OK;
}

extern "C" int airs_cuda_compress_batch_host(const struct airs_host_batch *hb)
{
	if (!hb || !hb->jobs || !hb->results || !hb->src || !hb->dst)
		return fail(AIRS_E_ARGUMENT, "host batch: src, dst, jobs and results must be non-NULL");
	Cache &c = g_cache;
	cudaStream_t s;
	int rc = cache_stream(&s);
	if (rc)
		return rc;
	size_t jobs_bytes = (size_t)hb->n_jobs * sizeof(struct airs_job);
	size_t scratch = airs_cuda_batch_scratch_size(hb->n_jobs, hb->n_results);
	if ((rc = c.src.reserve(hb->src_size + 64)) || (rc = c.dst.reserve(hb->dst_size + 64)) ||
	    (rc = c.work.reserve(hb->work_size + 64)) || (rc = c.jobs.reserve(jobs_bytes + 64)) ||
	    (rc = c.results.reserve((size_t)hb->n_results * 4 + 64)) ||
	    (rc = c.init.reserve((size_t)hb->n_jobs * 4 + 64)) ||
	    (rc = c.offs.reserve(((size_t)hb->n_results + 1) * 8 + 64)) || (rc = c.scratch.reserve(scratch)))
		return rc;

	if (hb->layout == AIRS_LAYOUT_CONCAT && !hb->out_offsets)
		return fail(AIRS_E_ARGUMENT, "host batch: the CONCAT layout needs out_offsets");
	if (hb->layout == AIRS_LAYOUT_SLOTS && hb->n_jobs >= 8 && hb->src_size >= (64u << 20))
		return host_batch_pipelined(hb, s);
	if (hb->layout == AIRS_LAYOUT_CONCAT && hb->n_jobs >= 8 && hb->src_size >= (64u << 20)) {
		bool redo = false;
		if ((rc = host_batch_concat_pipelined(hb, s, &redo)) || !redo)
			return rc;
	}

	CU(cudaMemcpyAsync(c.src.p, hb->src, hb->src_size, cudaMemcpyHostToDevice, s));
	CU(cudaMemcpyAsync(c.jobs.p, hb->jobs, jobs_bytes, cudaMemcpyHostToDevice, s));
	if (hb->work && hb->work_size)
		CU(cudaMemcpyAsync(c.work.p, hb->work, hb->work_size, cudaMemcpyHostToDevice, s));

	struct airs_batch b;
	memset(&b, 0, sizeof(b));
	b.src = c.src.p;
	b.dst = c.dst.p;
	b.work = hb->work_size ? c.work.p : nullptr;
	b.jobs = (const struct airs_job *)c.jobs.p;
	b.results = (uint32_t *)c.results.p;
	b.init_results = (uint32_t *)c.init.p;
	b.out_offsets = hb->layout == AIRS_LAYOUT_CONCAT ? (uint64_t *)c.offs.p : nullptr;
	b.scratch = c.scratch.p;
	b.dst_size = hb->dst_size;
	b.n_jobs = hb->n_jobs;
	b.n_results = hb->n_results;
	b.layout = hb->layout;
	b.flags = hb->flags;
	if (hb->layout == AIRS_LAYOUT_CONCAT) {
		/* temporary slots for the two-phase path, if the device has room for them */
		uint64_t caps = 0;
		for (uint32_t j = 0; j < hb->n_jobs; j++)
			caps += (uint64_t)hb->jobs[j].n_frames * hb->jobs[j].dst_capacity;
		const size_t need = airs_cuda_concat_tmp_size(caps, hb->n_results);
		size_t free_b = 0, total_b = 0;
		if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (need <= c.tmp.cap || need < free_b / 2) &&
		    c.tmp.reserve(need) == AIRS_OK) {
			b.tmp = c.tmp.p;
			b.tmp_size = need;
		} else {
			cudaGetLastError();
		}
	}
	rc = launch_batch(&b, nullptr, s);
	if (rc)
		return rc;

	CU(cudaMemcpyAsync(hb->results, c.results.p, (size_t)hb->n_results * 4, cudaMemcpyDeviceToHost, s));
	if (hb->init_results)
		CU(cudaMemcpyAsync(hb->init_results, c.init.p, (size_t)hb->n_jobs * 4, cudaMemcpyDeviceToHost, s));
	if (hb->layout == AIRS_LAYOUT_CONCAT) {
		/* only the bytes the scan laid out travel back */
		CU(cudaMemcpyAsync(hb->out_offsets, c.offs.p, ((size_t)hb->n_results + 1) * 8,
				   cudaMemcpyDeviceToHost, s));
		CU(cudaStreamSynchronize(s));
		uint64_t total = hb->out_offsets[hb->n_results];
		if (total > hb->dst_size)
			return fail(AIRS_E_ARGUMENT, "dst_size %llu < concatenated size %llu",
				    (unsigned long long)hb->dst_size, (unsigned long long)total);
		CU(cudaMemcpyAsync(hb->dst, c.dst.p, total, cudaMemcpyDeviceToHost, s));
	} else {
		CU(cudaMemcpyAsync(hb->dst, c.dst.p, hb->dst_size, cudaMemcpyDeviceToHost, s));
	}
	if (hb->work && hb->work_size)
		CU(cudaMemcpyAsync(hb->work, c.work.p, hb->work_size, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	return AIRS_OK;
}

extern "C" void airs_cuda_release_cache(void)
{
	Cache &c = g_cache;
	c.src.release();
	c.dst.release();
	c.work.release();
	c.jobs.release();
	c.results.release();
	c.init.release();
	c.offs.release();
	c.scratch.release();
	c.state.release();
	c.tmp.release();
	if (c.stream) {
		cudaStreamDestroy(c.stream);
		cudaStreamDestroy(c.s_in);
		cudaStreamDestroy(c.s_out);
		cudaEventDestroy(c.ev_start);
		for (int i = 0; i < kMaxGroups; i++) {
			cudaEventDestroy(c.ev_in[i]);
			cudaEventDestroy(c.ev_done[i]);
		}
	}
	if (c.pin)
		cudaFreeHost(c.pin);
	c.pin = nullptr;
	c.stream = c.s_in = c.s_out = nullptr;
	c.stream_dev = -1;
}

/* ------------------------------------------------------------------------ */
/* single-call path used by the cmp.h shim                                   */

extern "C" int airs_cuda_compress_resume(const struct airs_job *job, struct airs_ctx_state *state,
					 const void *src, void *dst, void *work, int work_is_state,
					 uint32_t *result)
{
	if (!job || !state || !result || job->n_frames != 1)
		return fail(AIRS_E_ARGUMENT, "resume: bad arguments");
	Cache &c = g_cache;
	cudaStream_t s;
	int rc = cache_stream(&s);
	if (rc)
		return rc;

	const bool src_dev = on_device(src), dst_dev = on_device(dst), work_dev = on_device(work);
	const uint32_t mis = (uint32_t)((uintptr_t)dst & 7u);
	struct airs_job j = *job;

	if ((rc = c.jobs.reserve(sizeof(j))) || (rc = c.results.reserve(64)) || (rc = c.state.reserve(sizeof(*state))) ||
	    (rc = c.scratch.reserve(airs_cuda_batch_scratch_size(1, 1))))
		return rc;

	struct airs_batch b;
	memset(&b, 0, sizeof(b));
	j.src_offset = 0;
	j.dst_offset = 0;
	j.work_offset = 0;
	j.first_result = 0;
	if (src && !src_dev) {
		if ((rc = c.src.reserve((size_t)j.src_size + 64)))
			return rc;
		CU(cudaMemcpyAsync(c.src.p, src, j.src_size, cudaMemcpyHostToDevice, s));
		b.src = c.src.p;
	} else {
		b.src = src;
	}
	if (dst && !dst_dev) {
		/* (a stream never exceeds cmp_compress_bound(): callers pass the size of whatever buffer they have) */
		const uint32_t stride = j.dtype == AIRS_DTYPE_I16_IN_I32 ? 4u : 2u;
		const uint64_t bound = 64u + 6ull * (j.src_size / stride), need = j.dst_capacity < bound ? j.dst_capacity : bound;
		if ((rc = c.dst.reserve((size_t)need + 64)))
			return rc;
		b.dst = c.dst.p; /* cudaMalloc memory is 256-byte aligned: reproduce the caller's alignment */
		j.dst_offset = mis;
	} else {
		b.dst = dst;
	}
	if (work && j.work_size && !work_dev) {
		if ((rc = c.work.reserve((size_t)j.work_size + 64)))
			return rc;
		if (work_is_state)
			CU(cudaMemcpyAsync(c.work.p, work, j.work_size, cudaMemcpyHostToDevice, s));
		b.work = c.work.p;
	} else {
		b.work = j.work_size ? work : nullptr;
	}
	CU(cudaMemcpyAsync(c.jobs.p, &j, sizeof(j), cudaMemcpyHostToDevice, s));
	CU(cudaMemcpyAsync(c.state.p, state, sizeof(*state), cudaMemcpyHostToDevice, s));
	b.jobs = (const struct airs_job *)c.jobs.p;
	b.results = (uint32_t *)c.results.p;
	b.scratch = c.scratch.p;
	b.dst_size = j.dst_capacity;
	b.n_jobs = 1;
	b.n_results = 1;
	b.layout = AIRS_LAYOUT_SLOTS;
	/* A context without secondary passes and without fallback: every call is a primary pass of a fresh context
	 * (ref cmp.c:228-262) - the frame goes through the batched path, where a long frame is spread over the whole
	 * device (airs_tile_kernel) instead of being encoded by the one CTA that continues a caller's context.  One
	 * identifier is drawn (the caller puts the real one into the stream); sequence number 0 -> 1.  A frame that
	 * fails here is done again below, where the context's state is followed step by step. */
	if (job->params.secondary_iterations == 0 && !job->params.uncompressed_fallback_enabled && state->valid) {
		rc = launch_batch(&b, nullptr, s);
		if (rc)
			return rc;
		CU(cudaMemcpyAsync(result, c.results.p, 4, cudaMemcpyDeviceToHost, s));
		CU(cudaStreamSynchronize(s));
		if (!airs_failed(*result)) {
			const uint32_t stride = j.dtype == AIRS_DTYPE_I16_IN_I32 ? 4u : 2u;
			state->counter = 1;
			state->seq = 1;
			state->model_size = j.src_size / stride * 2u;
			if (dst && !dst_dev) {
				CU(cudaMemcpyAsync(dst, (uint8_t *)c.dst.p + mis, *result, cudaMemcpyDeviceToHost, s));
				CU(cudaStreamSynchronize(s));
			}
			return AIRS_OK;
		}
	}
	rc = launch_batch(&b, (struct airs_ctx_state *)c.state.p, s);
	if (rc)
		return rc;
	CU(cudaMemcpyAsync(result, c.results.p, 4, cudaMemcpyDeviceToHost, s));
	CU(cudaMemcpyAsync(state, c.state.p, sizeof(*state), cudaMemcpyDeviceToHost, s));
	if (work && j.work_size && !work_dev && work_is_state)
		CU(cudaMemcpyAsync(work, c.work.p, j.work_size, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	if (dst && !dst_dev && !airs_failed(*result)) {
		CU(cudaMemcpyAsync(dst, (uint8_t *)c.dst.p + mis, *result, cudaMemcpyDeviceToHost, s));
		CU(cudaStreamSynchronize(s));
	}
	return AIRS_OK;
}

extern "C" int airs_cuda_patch_bytes(void *dst, const uint8_t *bytes, uint32_t offset, uint32_t count)
{
	if (!on_device(dst)) {
		memcpy((uint8_t *)dst + offset, bytes, count);
		return AIRS_OK;
	}
	CU(cudaMemcpy((uint8_t *)dst + offset, bytes, count, cudaMemcpyHostToDevice));
	return AIRS_OK;
}
