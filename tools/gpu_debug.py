"""Debug helper: run random job sets on the GPU, print the first mismatching frame in detail."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import jobgen
from oracle import oracle_py
pkg = jobgen.pkg
abi = pkg.abi
O = oracle_py.load_oracle()

def first_diff(js, want, got):
    jobs = js["jobs"]
    for j in range(len(jobs)):
        job = jobs[j]
        if want[2][j] != got[2][j]:
            return j, -1, "init %#x vs %#x" % (want[2][j], got[2][j])
        for f in range(int(job["n_frames"])):
            k = int(job["first_result"]) + f
            rw, rg = int(want[1][k]), int(got[1][k])
            if rw != rg:
                return j, f, "result %#x vs %#x" % (rw, rg)
            if abi.is_error(rw):
                continue
            o = int(job["dst_offset"]) + f * int(job["dst_frame_stride"])
            a, b = want[0][o:o + rw], got[0][o:o + rw]
            if not np.array_equal(a, b):
                at = int(np.nonzero(a != b)[0][0])
                return j, f, "byte %d of %d: want %s got %s" % (at, rw, a[max(0,at-4):at+8].tobytes().hex(), b[max(0,at-4):at+8].tobytes().hex())
    return None

seeds = [int(x) for x in sys.argv[1:]] or [100]
for seed in seeds:
    rng = np.random.default_rng(seed)
    if seed >= 200:
        js = jobgen.build_jobs(rng, 120, sizes=[1000, 2047, 2048, 2049, 4099, 6000, 10000], max_frames=4, allow_invalid=False)
    else:
        js = jobgen.build_jobs(rng, 300, sizes=[1, 2, 3, 5, 7, 8, 63, 64, 65, 255, 256, 257, 2049, 4099], max_frames=5, allow_invalid=True)
    want = jobgen.run_cpu(O, js)
    got = pkg.batch.run_jobs_device(js)
    nbad = 0
    jobs = js["jobs"]
    d = first_diff(js, want, got)
    print("seed", seed, "first diff:", d)
    if d:
        j = d[0]
        job = jobs[j]
        print({k: job[k] for k in job.dtype.names if k != "params"})
        print({k: job["params"][k] for k in job["params"].dtype.names})
        k0 = int(job["first_result"])
        print("want results", [hex(int(x)) for x in want[1][k0:k0 + int(job["n_frames"])]])
        print("got  results", [hex(int(x)) for x in got[1][k0:k0 + int(job["n_frames"])]])
