/*
 * hash_jobs.h - AIRS_DEFINE_HASH_JOBS(name, run): the batch loop `run` (oracle_run_jobs or
 * ref_run_jobs) with the streams of every job written into scratch memory of the calling
 * thread and hashed there (include/airs_stream_hash.h), so that a workload of any size can be
 * compared with the device's output through 12 bytes per stream (size and hash).
 * TEST INFRASTRUCTURE ONLY.
 */
#ifndef ORACLE_HASH_JOBS_H
#define ORACLE_HASH_JOBS_H

#include <stdlib.h>
#include <string.h>

#include "../include/airs_stream_hash.h"

#define AIRS_DEFINE_HASH_JOBS(name, run)                                                                      \
	int name(const void *src, const struct airs_job *jobs, uint32_t job_begin, uint32_t job_end,          \
		 uint32_t *results, uint64_t *hashes)                                                         \
	{                                                                                                     \
		uint8_t *dst = NULL, *work = NULL;                                                            \
		size_t dst_cap = 0, work_cap = 0;                                                             \
		uint32_t j, f, dummy_init;                                                                    \
                                                                                                              \
		for (j = job_begin; j < job_end; j++) {                                                       \
			struct airs_job job = jobs[j];                                                        \
			const uint32_t cap = job.dst_capacity > 0xFFFFFF80u ? 0u : job.dst_capacity;          \
			const size_t slot = ((size_t)cap + 15u) & ~(size_t)15u;                               \
			const size_t need = slot * job.n_frames + 64, wneed = (size_t)job.work_size + 64;     \
                                                                                                              \
			if (need > dst_cap) {                                                                 \
				free(dst);                                                                    \
				if (posix_memalign((void **)&dst, 64, need))                                  \
					return -1;                                                            \
				dst_cap = need;                                                               \
			}                                                                                     \
			if (wneed > work_cap) {                                                               \
				free(work);                                                                   \
				if (posix_memalign((void **)&work, 64, wneed))                                \
					return -1;                                                            \
				work_cap = wneed;                                                             \
			}                                                                                     \
			job.dst_offset = 0;                                                                   \
			job.dst_frame_stride = slot;                                                          \
			job.work_offset = 0;                                                                  \
			if (run(src, dst, work, &job, 0, 1, AIRS_LAYOUT_SLOTS, results, &dummy_init, NULL))   \
				return -1;                                                                    \
			for (f = 0; f < job.n_frames; f++) {                                                  \
				const uint32_t r = results[job.first_result + f];                             \
				hashes[job.first_result + f] =                                                \
					r > 0xFFFFFF80u ? 0 : airs_stream_hash(dst + (size_t)f * slot, r);    \
			}                                                                                     \
		}                                                                                             \
		free(dst);                                                                                    \
		free(work);                                                                                   \
		return 0;                                                                                     \
	}

#endif
