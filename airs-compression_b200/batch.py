"""torch-side plumbing for the batched C-ABI: device buffers, streams, the call.

torch is used for device memory, streams and (in bench.py) torch.distributed only;
all compression work happens inside libcmp_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import abi
from .loader import load_library


def _dev_u8(arr, device):
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).to(device)


class DeviceBatch:
    """A batch resident in HBM: source, destination, work buffers, job table, results."""

    def __init__(self, src, jobs, dst_size, work_size, n_results, layout=abi.LAYOUT_SLOTS,
                 device="cuda:0", fill=None, work_fill=None, concat_tmp=0):
        """concat_tmp: bytes of temporary memory for the two-phase CONCAT path (0: single phase)."""
        self.lib = load_library()
        self.device = torch.device(device)
        self.layout = layout
        self.n_jobs = len(jobs)
        self.n_results = int(n_results)
        self.src = src if isinstance(src, torch.Tensor) else _dev_u8(src, self.device)
        self.jobs = jobs if isinstance(jobs, torch.Tensor) else _dev_u8(jobs, self.device)
        self.dst = torch.empty(int(dst_size), dtype=torch.uint8, device=self.device)
        if fill is not None:
            self.dst.fill_(fill)
        self.work = torch.empty(max(int(work_size), 16), dtype=torch.uint8, device=self.device)
        if work_fill is not None:
            self.work.fill_(work_fill)
        self.results = torch.zeros(self.n_results, dtype=torch.int32, device=self.device)
        self.init_results = torch.zeros(max(self.n_jobs, 1), dtype=torch.int32, device=self.device)
        self.out_offsets = torch.zeros(self.n_results + 1, dtype=torch.int64, device=self.device)
        nscratch = self.lib.airs_cuda_batch_scratch_size(self.n_jobs, self.n_results)
        self.scratch = torch.empty(nscratch, dtype=torch.uint8, device=self.device)
        b = abi.AirsBatch()
        b.src = self.src.data_ptr()
        b.dst = self.dst.data_ptr()
        b.work = self.work.data_ptr() if int(work_size) > 0 else None   # (no work buffers: no transform kernels launched)
        b.jobs = self.jobs.data_ptr()
        b.results = self.results.data_ptr()
        b.init_results = self.init_results.data_ptr()
        b.out_offsets = self.out_offsets.data_ptr() if layout == abi.LAYOUT_CONCAT else None
        b.scratch = self.scratch.data_ptr()
        b.dst_size = int(dst_size)
        b.n_jobs = self.n_jobs
        b.n_results = self.n_results
        b.layout = layout
        b.flags = abi.BATCH_BIG_ENDIAN if bool((np.asarray(jobs)["dtype"] & abi.DT_BE).any()) else 0
        self.tmp = None
        if layout == abi.LAYOUT_CONCAT and concat_tmp:
            self.tmp = torch.empty(int(concat_tmp), dtype=torch.uint8, device=self.device)
            b.tmp = self.tmp.data_ptr()
            b.tmp_size = int(concat_tmp)
        self.desc = b

    def run(self, stream=None):
        """Enqueue the batch on `stream` (default: torch's current stream); asynchronous."""
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.airs_cuda_compress_batch(C.byref(self.desc), C.c_void_p(s.cuda_stream))
        if rc != 0:
            raise RuntimeError("airs_cuda_compress_batch failed (%d): %s"
                               % (rc, self.lib.airs_cuda_last_error().decode()))
        return self

    def launches(self):
        return int(self.lib.airs_cuda_last_launch_count())

    def hash_streams(self, stream=None):
        """64-bit hash of every stream of the last run (include/airs_stream_hash.h), a device int64 tensor."""
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        out = torch.zeros(max(self.n_results, 1), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.airs_cuda_hash_streams(C.byref(self.desc), C.c_void_p(out.data_ptr()), C.c_void_p(s.cuda_stream))
        if rc != 0:
            raise RuntimeError("airs_cuda_hash_streams failed (%d): %s" % (rc, self.lib.airs_cuda_last_error().decode()))
        return out[:self.n_results]

    def fetch(self):
        """(dst, results, init_results, out_offsets-or-None, work) as numpy, after a sync."""
        torch.cuda.synchronize(self.device)
        offs = self.out_offsets.cpu().numpy().astype(np.uint64) if self.layout == abi.LAYOUT_CONCAT else None
        return (self.dst.cpu().numpy(), self.results.cpu().numpy().view(np.uint32),
                self.init_results.cpu().numpy().view(np.uint32)[:self.n_jobs], offs,
                self.work.cpu().numpy())


def hash_ranges(base, offsets, sizes):
    """64-bit hashes (include/airs_stream_hash.h) of the byte ranges base[offsets[k] : offsets[k] + sizes[k]];
    base uint8, offsets int64, sizes int32 device tensors."""
    lib = load_library()
    n = int(sizes.numel())
    out = torch.zeros(max(n, 1), dtype=torch.int64, device=base.device)
    s = torch.cuda.current_stream(base.device)
    with torch.cuda.device(base.device):
        rc = lib.airs_cuda_hash_ranges(C.c_void_p(base.data_ptr()), C.c_void_p(offsets.data_ptr()), C.c_void_p(sizes.data_ptr()),
                                       n, C.c_void_p(out.data_ptr()), C.c_void_p(s.cuda_stream))
    if rc != 0:
        raise RuntimeError("airs_cuda_hash_ranges failed (%d): %s" % (rc, lib.airs_cuda_last_error().decode()))
    return out[:n]


def concat_tmp_size(jobs, n_results):
    """Bytes of temporary memory the two-phase CONCAT path needs for a (host-side) job table."""
    caps = int((jobs["n_frames"].astype(np.uint64) * jobs["dst_capacity"].astype(np.uint64)).sum())
    return int(load_library().airs_cuda_concat_tmp_size(caps, int(n_results)))


def run_jobs_device(js, device="cuda:0", fill=0xA5, work_fill=0x5A, concat_tmp=0):
    """Run a job set (tests/jobgen.py layout) on the GPU; same return shape as jobgen.run_cpu."""
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: this backend has no CPU path")
    db = DeviceBatch(js["src"], js["jobs"], js["dst_size"], js["work_size"], js["n_results"],
                     layout=js["layout"], device=device, fill=fill, work_fill=work_fill, concat_tmp=concat_tmp)
    db.run()
    dst, res, init, offs, work = db.fetch()
    return dst, res, init, offs, work[:js["work_size"]]


def run_jobs_host(js, fill=0xA5, work_fill=0x5A):
    """Same job set through airs_cuda_compress_batch_host (host buffers, copies inside the call)."""
    lib = load_library()
    src = np.ascontiguousarray(js["src"])
    jobs = np.ascontiguousarray(js["jobs"])
    dst = np.full(js["dst_size"], fill, dtype=np.uint8)
    work = np.full(js["work_size"], work_fill, dtype=np.uint8)
    n_jobs, n_results = len(jobs), js["n_results"]
    results = np.zeros(n_results, dtype=np.uint32)
    init = np.zeros(max(n_jobs, 1), dtype=np.uint32)
    offs = np.zeros(n_results + 1, dtype=np.uint64)
    hb = abi.AirsHostBatch()
    hb.src, hb.src_size = src.ctypes.data, src.nbytes
    hb.dst, hb.dst_size = dst.ctypes.data, dst.nbytes
    hb.work, hb.work_size = work.ctypes.data, work.nbytes
    hb.jobs, hb.results, hb.init_results = jobs.ctypes.data, results.ctypes.data, init.ctypes.data
    hb.out_offsets = offs.ctypes.data if js["layout"] == abi.LAYOUT_CONCAT else None
    hb.n_jobs, hb.n_results, hb.layout = n_jobs, n_results, js["layout"]
    hb.flags = abi.BATCH_BIG_ENDIAN if bool((jobs["dtype"] & abi.DT_BE).any()) else 0
    rc = lib.airs_cuda_compress_batch_host(C.byref(hb))
    if rc != 0:
        raise RuntimeError("airs_cuda_compress_batch_host failed (%d): %s"
                           % (rc, lib.airs_cuda_last_error().decode()))
    return dst, results, init[:n_jobs], offs if js["layout"] == abi.LAYOUT_CONCAT else None, work


class DeviceDecodeBatch:
    """Streams resident in HBM -> samples in HBM (include/airs_cuda_decode.h)."""

    def __init__(self, streams, dec_jobs, dst_size, n_results, device="cuda:0", fill=None):
        self.lib = load_library()
        self.device = torch.device(device)
        self.n_jobs = len(dec_jobs)
        self.n_results = int(n_results)
        self.src = streams if isinstance(streams, torch.Tensor) else _dev_u8(streams, self.device)
        self.jobs = dec_jobs if isinstance(dec_jobs, torch.Tensor) else _dev_u8(dec_jobs, self.device)
        self.dst = torch.empty(max(int(dst_size), 16), dtype=torch.uint8, device=self.device)
        if fill is not None:
            self.dst.fill_(fill)
        self.results = torch.zeros(max(self.n_results, 1), dtype=torch.int32, device=self.device)
        self.info = torch.zeros(max(self.n_results, 1) * abi.FRAME_INFO_DTYPE.itemsize, dtype=torch.uint8,
                                device=self.device)
        nscratch = self.lib.airs_cuda_decode_scratch_size(self.n_jobs, self.n_results)
        self.scratch = torch.empty(nscratch, dtype=torch.uint8, device=self.device)
        b = abi.AirsDecBatch()
        b.src = self.src.data_ptr()
        b.dst = self.dst.data_ptr()
        b.jobs = self.jobs.data_ptr()
        b.results = self.results.data_ptr()
        b.info = self.info.data_ptr()
        b.scratch = self.scratch.data_ptr()
        b.n_jobs = self.n_jobs
        b.n_results = self.n_results
        self.desc = b

    def run(self, stream=None):
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.airs_cuda_decompress_batch(C.byref(self.desc), C.c_void_p(s.cuda_stream))
        if rc != 0:
            raise RuntimeError("airs_cuda_decompress_batch failed (%d): %s"
                               % (rc, self.lib.airs_cuda_last_error().decode()))
        return self

    def fetch(self):
        """(samples as bytes, results, parsed headers) as numpy, after a sync."""
        torch.cuda.synchronize(self.device)
        return (self.dst.cpu().numpy(), self.results.cpu().numpy().view(np.uint32)[:self.n_results],
                self.info.cpu().numpy().view(abi.FRAME_INFO_DTYPE)[:self.n_results])


def decode_jobs_for(jobs, results, out_offsets=None):
    """Decode descriptors for the streams a compression job set produced: one per job, cut off in
    front of the first frame that failed (the encoder leaves a half-updated model behind a
    failure, ref cmp.h:300-303).  SLOTS layout unless out_offsets (CONCAT) is given.  The decoded
    frames of a job lie back to back, 16-byte aligned.  Returns (dec_jobs, dst_size, n_results)."""
    dj = np.zeros(len(jobs), dtype=abi.DEC_JOB_DTYPE)
    dst_off = 0
    for j, job in enumerate(jobs):
        first, nf = int(job["first_result"]), int(job["n_frames"])
        good = 0
        while good < nf and not abi.is_error(int(results[first + good])):
            good += 1
        stride = 4 if int(job["dtype"]) == abi.DT_I16_IN_I32 else 2
        n = int(job["src_size"]) // stride
        frame = (n * stride + 15) // 16 * 16
        if out_offsets is None:
            dj[j]["src_offset"] = job["dst_offset"]
            dj[j]["src_frame_stride"] = job["dst_frame_stride"]
            dj[j]["src_size"] = job["dst_capacity"]
        else:
            dj[j]["src_offset"] = out_offsets[first]
            dj[j]["src_frame_stride"] = 0
            dj[j]["src_size"] = int(out_offsets[first + good]) - int(out_offsets[first])
        dj[j]["dst_offset"] = dst_off
        dj[j]["dst_frame_stride"] = frame
        dj[j]["dst_capacity"] = n * stride
        dj[j]["n_frames"] = good
        dj[j]["dtype"] = job["dtype"]
        dj[j]["first_result"] = first
        dst_off += frame * nf
    return dj, dst_off + 64, int(sum(int(j["n_frames"]) for j in jobs))


def residual_stats(src, jobs, device="cuda:0"):
    """Per job: sum / count / maximum / log2 histogram of the zig-zag mapped residuals of its first
    frame (airs_cuda_residual_stats).  src, jobs: device tensors or host arrays.  Returns a numpy
    record array of abi.STATS_DTYPE."""
    lib = load_library()
    dev = torch.device(device)
    s = src if isinstance(src, torch.Tensor) else _dev_u8(src, dev)
    n_jobs = len(jobs)
    j = jobs if isinstance(jobs, torch.Tensor) else _dev_u8(jobs, dev)
    out = torch.zeros(max(n_jobs, 1) * abi.STATS_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.airs_cuda_residual_stats(C.c_void_p(s.data_ptr()), C.c_void_p(j.data_ptr()), n_jobs,
                                          C.c_void_p(out.data_ptr()),
                                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError("airs_cuda_residual_stats failed (%d): %s" % (rc, lib.airs_cuda_last_error().decode()))
    torch.cuda.synchronize(dev)
    return out.cpu().numpy().view(abi.STATS_DTYPE)[:n_jobs]


def param_candidates(stats_row, encoder_type):
    """airs_cuda_param_candidates: candidate encoders around what the statistics of one job suggest."""
    lib = load_library()
    out = np.zeros(abi.MAX_CANDIDATES, dtype=abi.CANDIDATE_DTYPE)
    st = np.ascontiguousarray(stats_row).reshape(1)
    n = lib.airs_cuda_param_candidates(C.c_void_p(st.ctypes.data), int(encoder_type), C.c_void_p(out.ctypes.data), len(out))
    return out[:n].copy()


def candidate_bits(src, jobs, cand, device="cuda:0"):
    """airs_cuda_candidate_bits: exact code bits of every job's first frame under every candidate encoder.
    Returns a uint64 array [n_jobs, n_cand] (2^64 - 1: a candidate cmp_initialise would refuse)."""
    lib = load_library()
    dev = torch.device(device)
    s = src if isinstance(src, torch.Tensor) else _dev_u8(src, dev)
    j = jobs if isinstance(jobs, torch.Tensor) else _dev_u8(jobs, dev)
    c = _dev_u8(np.ascontiguousarray(cand), dev)
    n_jobs, n_cand = len(jobs), len(cand)
    out = torch.zeros(max(n_jobs * n_cand, 1) * 8, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.airs_cuda_candidate_bits(C.c_void_p(s.data_ptr()), C.c_void_p(j.data_ptr()), n_jobs, C.c_void_p(c.data_ptr()),
                                          n_cand, C.c_void_p(out.data_ptr()),
                                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError("airs_cuda_candidate_bits failed (%d): %s" % (rc, lib.airs_cuda_last_error().decode()))
    torch.cuda.synchronize(dev)
    return out.cpu().numpy().view(np.uint64)[:n_jobs * n_cand].reshape(n_jobs, n_cand)
