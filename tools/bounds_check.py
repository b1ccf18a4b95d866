"""Run the random parity job sets through the AIRS_BOUNDS_CHECK build (libcheck.so) and report the strings that
would have been staged outside their warp's shared-memory words (development tool; stands in for compute-sanitizer,
which is closed on the GPU pool)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import jobgen
from oracle import oracle_py
pkg = jobgen.pkg
pkg.loader.library_path = lambda: os.path.join(ROOT, "airs-compression_b200", "libcheck.so")
lib = pkg.load_library()
_chk = ctypes.CDLL(os.path.join(ROOT, "airs-compression_b200", "libcheck.so"))
assert _chk.airs_fast_bounds_violations() >= 0 and _chk.airs_tile_bounds_violations() >= 0, "not a bounds-check build"


def viol():
    return _chk.airs_fast_bounds_violations() + _chk.airs_tile_bounds_violations()


SMALL = [1, 2, 3, 5, 7, 8, 63, 64, 65, 255, 256, 257, 1000, 2048, 2049, 4099, 9000, 40000, 70001]
total = 0
for seed in range(12):
    rng = np.random.default_rng(900 + seed)
    js = jobgen.build_jobs(rng, 300, sizes=SMALL, max_frames=3, allow_invalid=True)
    want = jobgen.run_cpu(oracle_py.load_oracle(), js)
    got = pkg.batch.run_jobs_device(js)
    jobgen.compare(want, got, js, "bounds-check")
    v = viol()
    total += v
    print("seed", seed, "violations", v, flush=True)
import importlib
t = importlib.import_module("test_gpu_parity")
for enc, g, outl, pre, n in [(2, 3, 107, 0, 2048), (2, 1, 200, 0, 1000), (1, 3000, 0, 0, 4104), (2, 3, 107, 1, 1531)]:
    x = np.full((64, n), 40000, dtype=np.uint16)
    p = pkg.abi.make_params(primary_preprocessing=pre, primary_encoder_type=enc, primary_encoder_param=g, primary_encoder_outlier=outl)
    js = t._uniform_jobs(pkg, 64, n, 1, p)
    js["src"] = x.view(np.uint8).reshape(-1)
    got = pkg.batch.run_jobs_device(js)
    v = viol(); total += v
    print("longest", enc, g, n, "violations", v, flush=True)
print("total violations", total)
