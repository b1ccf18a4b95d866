"""Development tool: per-phase clock totals of frame_fast (needs the -DAIRS_PHASE_CLOCKS build in build_dbg/)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from __graft_entry__ import load_package
pkg = load_package()
import airs_compression_b200.loader as loader
loader.library_path = lambda: os.path.join(ROOT, "airs-compression_b200", "build_dbg", "libcmp_b200.so")
lib = pkg.load_library()
import perf_probe_lib as L
names = ["wait+lds", "residual+lut", "scan", "pack", "model", "barrier2", "copyout", "tile total"]
for case in sys.argv[1:] or ["c1", "c4"]:
    buf = (C.c_ulonglong * 16)()
    lib.airs_phase_clocks(buf, 1)
    L.run_case(case, steps=1, warmup=0)
    lib.airs_phase_clocks(buf, 1)
    tot = buf[7] or 1
    print(case, {n: "%.1f%%" % (100.0 * buf[i] / tot) for i, n in enumerate(names)}, "cycles/tile(thread0 sum)=", buf[7])
    jn = ["ticket+info", "lut build", "plan_frame", "encode_pass(all)"]
    jt = sum(buf[8:12]) or 1
    print("   job level:", {n: "%.1f%%" % (100.0 * buf[8 + i] / jt) for i, n in enumerate(jn)}, "tiles share of encode_pass: %.1f%%" % (100.0 * buf[7] / (buf[11] or 1)), "job cycles sum", jt)
