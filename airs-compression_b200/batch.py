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
                 device="cuda:0", fill=None, work_fill=None):
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
        b.work = self.work.data_ptr()
        b.jobs = self.jobs.data_ptr()
        b.results = self.results.data_ptr()
        b.init_results = self.init_results.data_ptr()
        b.out_offsets = self.out_offsets.data_ptr() if layout == abi.LAYOUT_CONCAT else None
        b.scratch = self.scratch.data_ptr()
        b.dst_size = int(dst_size)
        b.n_jobs = self.n_jobs
        b.n_results = self.n_results
        b.layout = layout
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

    def fetch(self):
        """(dst, results, init_results, out_offsets-or-None, work) as numpy, after a sync."""
        torch.cuda.synchronize(self.device)
        offs = self.out_offsets.cpu().numpy().astype(np.uint64) if self.layout == abi.LAYOUT_CONCAT else None
        return (self.dst.cpu().numpy(), self.results.cpu().numpy().view(np.uint32),
                self.init_results.cpu().numpy().view(np.uint32)[:self.n_jobs], offs,
                self.work.cpu().numpy())


def run_jobs_device(js, device="cuda:0", fill=0xA5, work_fill=0x5A):
    """Run a job set (tests/jobgen.py layout) on the GPU; same return shape as jobgen.run_cpu."""
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: this backend has no CPU path")
    db = DeviceBatch(js["src"], js["jobs"], js["dst_size"], js["work_size"], js["n_results"],
                     layout=js["layout"], device=device, fill=fill, work_fill=work_fill)
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
    rc = lib.airs_cuda_compress_batch_host(C.byref(hb))
    if rc != 0:
        raise RuntimeError("airs_cuda_compress_batch_host failed (%d): %s"
                           % (rc, lib.airs_cuda_last_error().decode()))
    return dst, results, init[:n_jobs], offs if js["layout"] == abi.LAYOUT_CONCAT else None, work
