/*
 * airs_stream_hash.h - the 64-bit hash used to compare compressed streams at full workload
 * sizes without moving them: bench.py and the tests hash every stream where it was produced
 * (airs_cuda_hash_streams on the device, ref_hash_jobs / oracle_hash_jobs on the host cores)
 * and compare 8 bytes per stream.  Not part of the stream format; no reference counterpart
 * (SURVEY.md section 8d "Parity check": per-chunk 64-bit hash of the oracle output).
 *
 * H(b[0..n)) = mix(n) + sum over i of mix(w_i + (i + 1) * 0x9E3779B97F4A7C15)   (mod 2^64)
 * w_i = bytes 8 i .. 8 i + 7 as a little-endian word, missing bytes 0; mix = the splitmix64
 * finaliser.  A sum of position-keyed terms: any order, any number of threads.
 */
#ifndef AIRS_STREAM_HASH_H
#define AIRS_STREAM_HASH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define AIRS_HASH_FN __host__ __device__ static inline
#else
#define AIRS_HASH_FN static inline
#endif

#define AIRS_HASH_GOLDEN 0x9E3779B97F4A7C15ull

AIRS_HASH_FN uint64_t airs_hash_mix(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

AIRS_HASH_FN uint64_t airs_hash_term(uint64_t word, uint64_t index)
{
	return airs_hash_mix(word + (index + 1) * AIRS_HASH_GOLDEN);
}

#ifndef __CUDA_ARCH__
/* the whole hash, one thread (host) */
static inline uint64_t airs_stream_hash(const uint8_t *b, size_t n)
{
	uint64_t h = airs_hash_mix((uint64_t)n);
	size_t i, k;

	for (i = 0; i < n / 8; i++) {
		uint64_t w = 0;
		for (k = 0; k < 8; k++)
			w |= (uint64_t)b[8 * i + k] << (8 * k);
		h += airs_hash_term(w, i);
	}
	if (n % 8) {
		uint64_t w = 0;
		for (k = 0; k < n % 8; k++)
			w |= (uint64_t)b[8 * i + k] << (8 * k);
		h += airs_hash_term(w, i);
	}
	return h;
}
#endif

#endif /* AIRS_STREAM_HASH_H */
