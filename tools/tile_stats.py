"""Look-back statistics of airs_tile_kernel (libstats.so, built with -DAIRS_TILE_STATS): development tool."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ["AIRS_PROBE_LIB"] = "libstats.so"
import numpy as np, torch
import perf_probe_lib as L
pkg = L.pkg
orig = pkg.batch.DeviceBatch.run
last = {}
def run(self, stream=None):
    r = orig(self, stream); last["db"] = self; return r
pkg.batch.DeviceBatch.run = run
for case in sys.argv[1:] or ["c4b"]:
    L.run_case(case, steps=1, warmup=1)
    torch.cuda.synchronize()
    st = last["db"].scratch[:128].cpu().numpy().view(np.uint32)
    print(case, "look-backs %d  polls after the early one %d  tail polls %d  extra windows %d" % tuple(st[16:20]))
