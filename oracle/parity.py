"""Whole-workload parity of the CUDA path against the CPU reference (SURVEY.md section 8d "Parity check").

TEST INFRASTRUCTURE ONLY (used by tests/ and by bench.py's CPU legs): every stream of a device batch is
hashed where it was produced (airs_cuda_hash_streams), the reference compresses the same input on the host
cores and hashes its streams the same way (oracle/hash_jobs.h), and sizes + hashes of ALL frames are
compared; a sample of the jobs (every `sample_every`-th) is then compared byte for byte.
"""
import os
import time

import numpy as np

from . import oracle_py


def cpu_library():
    """The compiled, unmodified reference where it was built (oracle/_ref), else the oracle port."""
    if oracle_py.have_ref():
        return oracle_py.load_ref(), "reference"
    return oracle_py.load_oracle(), "port"


def check_device_batch(db, jobs, src_dev, threads=None, sample_every=100, max_jobs=None, lib=None):
    """db: batch.DeviceBatch whose run() has completed (SLOTS layout); jobs: the host job table;
    src_dev: the device uint8 tensor the batch read.  Returns a dict (all counts are per call)."""
    import torch
    lib, kind = (lib, lib.kind) if lib is not None else cpu_library()
    threads = threads or os.cpu_count() or 1
    n_jobs = len(jobs) if max_jobs is None else min(len(jobs), int(max_jobs))
    t0 = time.perf_counter()
    hashes_dev = db.hash_streams()
    src_host = torch.empty(src_dev.numel(), dtype=torch.uint8, pin_memory=True)
    src_host.copy_(src_dev.view(torch.uint8).reshape(-1), non_blocking=True)
    torch.cuda.synchronize(db.device)
    res_dev = db.results.cpu().numpy().view(np.uint32)
    h_dev = hashes_dev.cpu().numpy().view(np.uint64)
    t1 = time.perf_counter()
    res_cpu, h_cpu = lib.hash_jobs(src_host.data_ptr(), jobs, threads=threads, job_end=n_jobs)
    t2 = time.perf_counter()
    last = jobs[n_jobs - 1]
    n_frames = int(last["first_result"]) + int(last["n_frames"])            # jobs are in result order
    sizes_ok = bool(np.array_equal(res_dev[:n_frames], res_cpu[:n_frames]))
    hashes_ok = bool(np.array_equal(h_dev[:n_frames], h_cpu[:n_frames]))
    ok_frames = res_cpu[:n_frames] <= 0xFFFFFF80
    in_bytes = int((jobs["src_size"][:n_jobs].astype(np.int64) * jobs["n_frames"][:n_jobs]).sum())

    # byte-for-byte on a sample of the jobs
    sel = np.arange(0, n_jobs, max(1, int(sample_every)))
    sub = jobs[sel].copy()
    slot = ((sub["dst_capacity"].astype(np.uint64) + np.uint64(15)) // np.uint64(16) * np.uint64(16)) * sub["n_frames"]
    offs = np.concatenate([[0], np.cumsum(slot)]).astype(np.uint64)
    woffs = np.concatenate([[0], np.cumsum((sub["work_size"].astype(np.uint64) + np.uint64(15)) // np.uint64(16) * np.uint64(16))])
    nres = np.concatenate([[0], np.cumsum(sub["n_frames"])]).astype(np.uint32)
    dev_first, dev_off, dev_stride = sub["first_result"].copy(), sub["dst_offset"].copy(), sub["dst_frame_stride"].copy()
    sub["dst_offset"] = offs[:-1]
    sub["dst_frame_stride"] = (sub["dst_capacity"].astype(np.uint64) + np.uint64(15)) // np.uint64(16) * np.uint64(16)
    sub["work_offset"] = woffs[:-1]
    sub["first_result"] = nres[:-1]
    dst_h = np.zeros(int(offs[-1]) + 64, dtype=np.uint8)
    work_h = np.zeros(int(woffs[-1]) + 64, dtype=np.uint8)
    src_np = np.frombuffer((__import__("ctypes").c_uint8 * src_host.numel()).from_address(src_host.data_ptr()), dtype=np.uint8)
    res_s, _, _ = lib.run_jobs(src_np, dst_h, work_h, sub, 0, threads=threads)
    bytes_ok, compared, compared_bytes = True, 0, 0
    for i in range(len(sub)):
        for f in range(int(sub[i]["n_frames"])):
            r = int(res_s[int(sub[i]["first_result"]) + f])
            if r > 0xFFFFFF80:
                continue
            if int(res_dev[int(dev_first[i]) + f]) != r:
                bytes_ok = False
                continue
            o = int(dev_off[i]) + f * int(dev_stride[i])
            g = db.dst[o:o + r].cpu().numpy()
            c = int(sub[i]["dst_offset"]) + f * int(sub[i]["dst_frame_stride"])
            bytes_ok &= bool(np.array_equal(g, dst_h[c:c + r]))
            compared += 1
            compared_bytes += r
    t3 = time.perf_counter()
    return {
        "cpu": kind, "cpu_threads": threads, "jobs": n_jobs, "frames": n_frames, "frames_ok_on_cpu": int(ok_frames.sum()),
        "input_bytes": in_bytes, "sizes_identical": sizes_ok, "hashes_identical": hashes_ok,
        "byte_compared_frames": compared, "byte_compared_bytes": compared_bytes, "bytes_identical": bool(bytes_ok),
        "identical": bool(sizes_ok and hashes_ok and bytes_ok),
        "cpu_hash_seconds": t2 - t1, "cpu_hash_gbs": in_bytes / max(t2 - t1, 1e-9) / 1e9,
        "seconds": t3 - t0,
    }
