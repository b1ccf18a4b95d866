"""the streams do not depend on the byte order of the machine that produces them.

The flight target of the reference is a big-endian LEON3 (cross/bcc2-gcc-gr712rc.txt); this backend and the CPU
legs here are little-endian.  The reference's hot path touches host byte order in exactly two places (DESIGN.md,
section on byte order): the 64-bit store of the bit writer (explicit cpu_to_be64, bitstream_writer.h:86-89) and
the checksum, whose big-endian FAST PATH hashes the sample buffer as it lies in memory (header.c:146-147) while
little-endian hosts swap every sample first (header.c:153-162).  These tests pin the second one: the trailer a
little-endian producer writes is the XXH32 of the byte image a big-endian machine holds.
"""
import ctypes as C

import numpy as np
import pytest

import jobgen

SEED = 419764627  # CHECKSUM_SEED, ref lib/common/header_private.h:46


def _xxh32(oracle, data):
    buf = np.ascontiguousarray(data, dtype=np.uint8)
    return int(oracle.lib.oracle_xxh32(buf.ctypes.data_as(C.c_void_p), buf.nbytes, SEED))


@pytest.mark.parametrize("dtype", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 1000, 4099])
def test_checksum_is_the_hash_of_the_big_endian_image(oracle, dtype, n):
    """cmp_checksum on a little-endian host (slow path, restated in oracle.c) == XXH32 over the bytes of the same
    samples stored big-endian, which is what the fast path of a big-endian host hashes."""
    rng = np.random.default_rng(n * 3 + dtype)
    vals = rng.integers(-32768, 32768, size=n).astype(np.int16)
    if dtype == 1:      # i16 in i32: the low 16 bits of every word are the sample, whatever the upper bits say
        src = (vals.astype(np.int32) & 0xFFFF) | (rng.integers(0, 1 << 15, size=n).astype(np.int32) << 16)
        src = src.astype("<i4")
    else:
        src = vals.view(np.uint16) if dtype == 2 else vals
    got = int(oracle.lib.oracle_checksum(np.ascontiguousarray(src).ctypes.data_as(C.c_void_p), n, dtype))
    be_image = vals.astype(">i2").view(np.uint8)     # the memory of a big-endian machine holding these samples
    assert got == _xxh32(oracle, be_image)


def test_reference_trailer_is_the_hash_of_the_big_endian_image(ref, oracle, pkg):
    """The compiled (little-endian) reference: the four trailer bytes of a stream, read big-endian, are the XXH32 of
    the big-endian sample image - so the on-board build, which hashes its memory directly, writes the same bytes."""
    abi = pkg.abi
    rng = np.random.default_rng(44)
    js = jobgen.build_jobs(rng, 40, sizes=[5, 64, 1000, 2049], max_frames=1)
    js["jobs"]["params"]["checksum_enabled"] = 1
    dst, res, init, _, _ = jobgen.run_cpu(ref, js)
    checked = 0
    for j, job in enumerate(js["jobs"]):
        r = int(res[int(job["first_result"])])
        if abi.is_error(r) or abi.is_error(int(init[j])):
            continue
        o, so, ss = int(job["dst_offset"]), int(job["src_offset"]), int(job["src_size"])
        raw = js["src"][so:so + ss]
        vals = (raw.view("<i4") & 0xFFFF).astype(np.uint16) if int(job["dtype"]) == 1 else raw.view("<u2")
        trailer = int.from_bytes(dst[o + r - 4:o + r].tobytes(), "big")
        assert trailer == _xxh32(oracle, vals.astype(">u2").view(np.uint8)), j
        checked += 1
    assert checked > 20


@pytest.mark.gpu
@pytest.mark.parametrize("n_jobs,n", [(24, 5000), (32768, 40)])
def test_gpu_trailer_is_the_hash_of_the_big_endian_image(gpu, oracle, pkg, n_jobs, n):
    """... and the same for the streams of this backend, on both checksum kernels (four lanes per frame for few
    frames, one thread per frame from 32768 frames on)."""
    abi = pkg.abi
    rng = np.random.default_rng(n)
    jobs, dsz, wsz = pkg.workloads.uniform_jobs(n_jobs, n, 1, abi.compress_bound(2 * n))
    jobs["params"] = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=9,
                                     checksum_enabled=1)
    x = (rng.integers(-30, 31, size=n_jobs * n).cumsum() & 0xFFFF).astype(np.uint16)
    js = {"src": x.view(np.uint8), "jobs": jobs, "layout": 0, "dst_size": dsz, "work_size": 0, "n_results": n_jobs}
    dst, res, _, _, _ = gpu.run_jobs_device(js)
    for j in list(range(0, n_jobs, max(1, n_jobs // 50))) + [n_jobs - 1]:
        r, o = int(res[j]), int(jobs[j]["dst_offset"])
        assert not abi.is_error(r)
        trailer = int.from_bytes(dst[o + r - 4:o + r].tobytes(), "big")
        assert trailer == _xxh32(oracle, x[j * n:(j + 1) * n].astype(">u2").view(np.uint8)), j
