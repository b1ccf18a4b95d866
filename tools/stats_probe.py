"""Device-timed throughput of airs_cuda_residual_stats (development tool)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import perf_probe_lib as L
pkg, abi, synth = L.pkg, L.abi, L.synth
lib = pkg.load_library()
for name, n_chunks, n in (("4096 x 256 KiB", 4096, 1 << 17), ("262144 x 4 KiB", 1 << 18, 2048), ("512 x 2 MiB", 512, 1 << 20)):
    data = synth.chunks_torch(1, 0, n_chunks, n, device="cuda")
    def pf(p, idx):
        p["primary_preprocessing"] = abi.PRE_DIFF; p["primary_encoder_type"] = 1; p["primary_encoder_param"] = 16
    jobs, _, _ = L.make_uniform_jobs(n_chunks, n, 1, pf)
    jd = torch.from_numpy(jobs.view(np.uint8).reshape(-1)).cuda()
    out = torch.zeros(n_chunks * abi.STATS_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    src = data.view(torch.uint8).reshape(-1)
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: lib.airs_cuda_residual_stats(C.c_void_p(src.data_ptr()), C.c_void_p(jd.data_ptr()), n_chunks, C.c_void_p(out.data_ptr()), C.c_void_p(st))
    for _ in range(2): assert call() == 0
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record(); call(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[2]
    nb = src.numel()
    print(f"stats {name}: {ms:.3f} ms, {nb/ms/1e6:.0f} GB/s = {nb/ms/1e6/6531.6*100:.1f}% of 6531.6", flush=True)
