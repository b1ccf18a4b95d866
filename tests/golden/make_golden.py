"""Generates tests/golden/ref_streams.json from the COMPILED, UNMODIFIED reference
(oracle/_ref/libcmp_ref.so, built by oracle/Makefile from /root/reference).

Run in the build container (where /root/reference is mounted):  python tests/golden/make_golden.py
The fixture travels to the GPU box, the reference does not.  Each case stores the job
description, the source bytes, and for every frame the reference's return value and stream.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jobgen  # noqa: E402
from oracle import oracle_py  # noqa: E402

abi = jobgen.abi


def main():
    ref = oracle_py.load_ref()
    rng = np.random.default_rng(20261018)
    cases = []
    # a grid over preprocessing x encoder x dtype x checksum x fallback on small, odd sizes
    grid = []
    for pre in (0, 1, 2):
        for enc in (0, 1, 2):
            for dt in (0, 1, 2):
                grid.append((pre, enc, dt))
    sizes = [1, 2, 3, 5, 7, 8, 63, 64, 65, 257]
    for gi, (pre, enc, dt) in enumerate(grid):
        def pf(r, pre=pre, enc=enc):
            g = int(r.choice([1, 3, 8, 16, 255, 4096, 65535]))
            o = int(r.choice([1, 16, 107, 0xFFFFFFFF]))
            sec_iter = int(r.choice([0, 2, 3]))
            return abi.make_params(
                primary_preprocessing=pre, primary_encoder_type=enc, primary_encoder_param=g,
                primary_encoder_outlier=o, secondary_iterations=sec_iter,
                secondary_preprocessing=int(r.choice([0, 1, 2, 3, 3])),
                secondary_encoder_type=int(r.choice([0, 1, 2])), secondary_encoder_param=int(r.choice([1, 8, 60, 1055])),
                secondary_encoder_outlier=int(r.choice([5, 42, 200])), model_rate=int(r.choice([0, 1, 8, 11, 16])),
                checksum_enabled=int(r.integers(0, 2)), uncompressed_fallback_enabled=int(r.integers(0, 2)))
        js = jobgen.build_jobs(rng, 3, sizes=sizes, max_frames=5, params_fn=pf, dtypes=(dt,))
        dst, results, init, _, work = jobgen.run_cpu(ref, js)
        jobs = js["jobs"]
        frames = []
        for j in range(len(jobs)):
            job = jobs[j]
            for f in range(int(job["n_frames"])):
                r = int(results[int(job["first_result"]) + f])
                o = int(job["dst_offset"]) + f * int(job["dst_frame_stride"])
                frames.append({"result": r, "stream": "" if abi.is_error(r) else dst[o:o + r].tobytes().hex()})
        cases.append({"src": js["src"].tobytes().hex(), "jobs": jobs.tobytes().hex(), "dst_size": js["dst_size"],
                      "work_size": js["work_size"], "n_results": js["n_results"],
                      "init": [int(x) for x in init], "frames": frames})
    out = os.path.join(HERE, "ref_streams.json")
    with open(out, "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "airs-compression v0.6.0 (compiled)",
                   "cases": cases}, fh)
    print("wrote", out, os.path.getsize(out), "bytes,", sum(len(c["frames"]) for c in cases), "frames")


if __name__ == "__main__":
    main()
