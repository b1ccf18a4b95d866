"""ctypes loader for the test oracle (oracle/liboracle.so) and, where built,
the compiled reference (oracle/_ref/libcmp_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the CPU
baseline legs of bench.py.  Product code never imports this module.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcmp_ref.so")


def build(quiet=True):
    """Compile liboracle.so and, if /root/reference is mounted, _ref/libcmp_ref.so."""
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


_RUN_ARGS = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
             C.c_void_p, C.c_void_p, C.c_void_p]


class _Lib:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.run = getattr(self.lib, prefix + "_run_jobs")
        self.run.argtypes = _RUN_ARGS
        self.run.restype = C.c_int
        self.hash = getattr(self.lib, prefix + "_hash_jobs")
        self.hash.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        self.hash.restype = C.c_int
        self.kind = prefix

    def hash_jobs(self, src, jobs, threads=1, job_begin=0, job_end=None):
        """The batch loop with every stream hashed where it was produced (oracle/hash_jobs.h) instead of kept:
        returns (results, hashes) for the frames of jobs[job_begin:job_end]; other entries stay 0.
        src: contiguous numpy array, or an integer host address (pinned torch memory)."""
        n_jobs = len(jobs)
        job_end = n_jobs if job_end is None else job_end
        n_results = int(jobs["first_result"][-1] + jobs["n_frames"][-1]) if n_jobs else 0
        results = np.zeros(n_results, dtype=np.uint32)
        hashes = np.zeros(n_results, dtype=np.uint64)
        addr = src if isinstance(src, int) else src.ctypes.data
        count = max(job_end - job_begin, 0)
        threads = max(1, min(threads, count))
        edges = (job_begin + np.linspace(0, count, threads + 1)).astype(np.int64)
        with ThreadPoolExecutor(threads) as ex:
            futs = [ex.submit(self.hash, addr, jobs.ctypes.data, int(edges[t]), int(edges[t + 1]),
                              results.ctypes.data, hashes.ctypes.data) for t in range(threads)]
            assert all(f.result() == 0 for f in futs)
        return results, hashes

    def run_jobs(self, src, dst, work, jobs, layout=0, threads=1):
        """Run the batch loop.  src/dst/work: contiguous numpy uint8-viewable arrays (work may
        be None).  Returns (results, init_results, out_offsets-or-None)."""
        n_jobs = len(jobs)
        n_results = int(jobs["first_result"][-1] + jobs["n_frames"][-1]) if n_jobs else 0
        results = np.zeros(n_results, dtype=np.uint32)
        init_results = np.zeros(n_jobs, dtype=np.uint32)
        offs = np.zeros(n_results + 1, dtype=np.uint64) if layout == 1 else None
        args = (src.ctypes.data, dst.ctypes.data, work.ctypes.data if work is not None else None,
                jobs.ctypes.data)
        tail = (results.ctypes.data, init_results.ctypes.data,
                offs.ctypes.data if offs is not None else None)
        threads = min(threads, n_jobs)
        if threads <= 1 or layout == 1:
            rc = self.run(*args, 0, n_jobs, layout, *tail)
            assert rc == 0
        else:
            edges = np.linspace(0, n_jobs, threads + 1).astype(np.int64)
            with ThreadPoolExecutor(threads) as ex:
                futs = [ex.submit(self.run, *args, int(edges[t]), int(edges[t + 1]), layout, *tail)
                        for t in range(threads)]
                assert all(f.result() == 0 for f in futs)
        return results, init_results, offs


def load_oracle():
    if not os.path.exists(ORACLE_SO):
        build()
    lib = _Lib(ORACLE_SO, "oracle")
    L = lib.lib
    L.oracle_xxh32.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.oracle_xxh32.restype = C.c_uint32
    L.oracle_checksum.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.oracle_checksum.restype = C.c_uint32
    L.oracle_outlier.argtypes = [C.c_uint32] * 3
    L.oracle_outlier.restype = C.c_uint32
    L.oracle_encode_residual.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int16,
                                         C.POINTER(C.c_uint64)]
    L.oracle_encode_residual.restype = C.c_uint32
    L.oracle_iwt.argtypes = [C.c_void_p, C.c_uint32]
    L.oracle_iwt.restype = None
    L.oracle_model_update.argtypes = [C.c_uint16, C.c_uint16, C.c_uint32, C.c_uint32]
    L.oracle_model_update.restype = C.c_uint16
    L.oracle_compress_bound.argtypes = [C.c_uint32]
    L.oracle_compress_bound.restype = C.c_uint32
    L.oracle_work_buf_size.argtypes = [C.c_void_p, C.c_uint32]
    L.oracle_work_buf_size.restype = C.c_uint32
    return lib


def have_ref():
    return os.path.exists(REF_SO)


def load_ref():
    """The compiled, unmodified reference (cmp_* symbols) plus ref_run_jobs."""
    lib = _Lib(REF_SO, "ref")
    L = lib.lib
    L.ref_driver_install.restype = None
    L.ref_driver_set_counter.argtypes = [C.c_uint64]
    L.ref_driver_set_counter.restype = None
    L.ref_driver_get_counter.restype = C.c_uint64
    for name in ("cmp_compress_u16", "cmp_compress_i16", "cmp_compress_i16_in_i32"):
        f = getattr(L, name)
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        f.restype = C.c_uint32
    L.cmp_initialise.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.cmp_initialise.restype = C.c_uint32
    L.cmp_reset.argtypes = [C.c_void_p]
    L.cmp_reset.restype = C.c_uint32
    L.cmp_compress_bound.argtypes = [C.c_uint32]
    L.cmp_compress_bound.restype = C.c_uint32
    L.cmp_cal_work_buf_size.argtypes = [C.c_void_p, C.c_uint32]
    L.cmp_cal_work_buf_size.restype = C.c_uint32
    L.cmp_checksum.argtypes = [C.c_void_p]
    L.cmp_checksum.restype = C.c_uint32
    return lib
