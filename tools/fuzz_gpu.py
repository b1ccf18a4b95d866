"""Longer randomized parity campaign than the test suite runs: CUDA encoder vs oracle (SLOTS, both CONCAT
paths), decoder vs the oracle's streams.  usage: fuzz_gpu.py FIRST_SEED N_SEEDS   (development tool)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import jobgen
from oracle import oracle_py
from test_gpu_decode import decode_and_check

pkg = jobgen.pkg
gpu = pkg.batch
oracle = oracle_py.load_oracle()
first, count = int(sys.argv[1]), int(sys.argv[2])
SMALL = [1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 63, 64, 65, 255, 256, 257, 511, 1023]
MED = [1000, 2047, 2048, 2049, 4095, 4096, 4097, 4099, 6000, 8191, 8192, 10000, 16384, 20001, 33000, 40000, 70001, 131072]
n_frames = 0
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    kind = seed % 4
    sizes = SMALL if kind < 2 else MED
    layout = kind & 1
    if seed % 5 == 4:  # many contexts of whole tiles, mostly in slots of the bound: the runs of model passes (model_run_fast)
        js = jobgen.build_jobs(rng, 320, sizes=[4096, 8192, 12288, 2048, 4097, 1000], max_frames=6, layout=layout,
                               capacity_modes=["bound", "bound", "bound", "big", "raw", "tight"])
    else:
        js = jobgen.build_jobs(rng, 200 if kind < 2 else 60, sizes=sizes, max_frames=6, allow_invalid=(kind == 0), layout=layout)
    want = jobgen.run_cpu(oracle, js)
    got = gpu.run_jobs_device(js)
    jobgen.compare(want, got, js, f"seed {seed} layout {layout}")
    if seed % 3 == 0:  # the same job set with big-endian sample containers (AIRS_DTYPE_BE)
        from test_gpu_parity import _to_big_endian
        jobgen.compare(want, gpu.run_jobs_device(_to_big_endian(js)), js, f"seed {seed} big-endian")
    if layout == 1:
        got2 = gpu.run_jobs_device(js, concat_tmp=gpu.concat_tmp_size(js["jobs"], js["n_results"]))
        jobgen.compare(want, got2, js, f"seed {seed} two-phase")
    n_frames += decode_and_check(pkg, gpu, js, want, f"seed {seed} decode")
    print("seed", seed, "ok", flush=True)
print("all ok:", count, "seeds,", n_frames, "frames decoded back")
