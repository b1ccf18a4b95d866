"""End-to-end probe (development tool): bench.py's e2e leg alone, on the library named by AIRS_PROBE_LIB."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
pkg = load_package()
if os.environ.get("AIRS_PROBE_LIB"):
    pkg.loader.library_path = lambda: os.path.join(ROOT, "airs-compression_b200", os.environ["AIRS_PROBE_LIB"])
import bench
lib = pkg.load_library()
dev = torch.device("cuda", 0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
units = {"c3": 1 << 20, "c2": lib.airs_cuda_concurrent_jobs(), "c5": 512}[wl]
r = bench.e2e_run(pkg, lib, wl, units, 0, 1, dev, lambda: torch.cuda.synchronize(dev), 8)
print(os.environ.get("AIRS_PROBE_LIB", "libcmp_b200.so"), wl, "%.1f GB/s" % r["value"], r["ms_per_step_min_max"], r.get("ms_steps"), flush=True)
