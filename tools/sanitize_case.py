"""One small batch per kernel arm, checked against the oracle (development tool): the program compute-sanitizer
runs (tools/sanitize.sh).  Arms: airs_encode_kernel (every preprocessing / encoder, model contexts), airs_fast_kernel
(short single-frame jobs), airs_tile_kernel (a few long frames), both checksum kernels, CONCAT two-phase and
single-phase, the decoder, the cmp.h shim path (host buffers)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package(); abi = pkg.abi
from oracle import oracle_py
import jobgen

ora = oracle_py.load_oracle()
rng = np.random.default_rng(11)
which = sys.argv[1:] or ["mixed", "fast", "tile", "checksum", "concat", "decode"]

def check(js, what, **kw):
    want = jobgen.run_cpu(ora, js)
    got = pkg.batch.run_jobs_device(js, **kw)
    jobgen.compare(want, got, js, what)
    print(what, "ok:", len(js["jobs"]), "jobs", js["n_results"], "frames", flush=True)
    return got

def uniform(n_jobs, n, nf, p, layout=0):
    cap = abi.compress_bound(2 * n)
    jobs, dsz, wsz = pkg.workloads.uniform_jobs(n_jobs, n, nf, cap, model=True)
    jobs["params"] = p
    x = rng.integers(-40, 41, size=n_jobs * nf * n).cumsum().astype(np.int64) & 0xFFFF
    return {"src": x.astype(np.uint16).view(np.uint8), "jobs": jobs, "layout": layout, "dst_size": dsz,
            "work_size": wsz, "n_results": n_jobs * nf}

if "mixed" in which:      # encode kernel: all arms, generic tiles, errors
    check(jobgen.build_jobs(rng, 48, sizes=[1, 7, 64, 1000, 2048, 4099, 9000], max_frames=3, allow_invalid=True), "mixed")
if "fast" in which:       # warp per job
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=2, primary_encoder_param=7, primary_encoder_outlier=42)
    check(uniform(96, 2048, 1, p), "fast")
if "tile" in which:       # long frames cut into tiles over all CTAs (few jobs)
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16)
    check(uniform(3, 70001, 1, p), "tile")
if "checksum" in which:   # warp kernel (few frames); the thread-per-frame kernel needs >= 32768 frames
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16, checksum_enabled=1,
                        secondary_iterations=2, secondary_preprocessing=abi.PRE_MODEL, secondary_encoder_type=1,
                        secondary_encoder_param=8, model_rate=8)
    check(uniform(5, 5000, 3, p), "checksum-warp")
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=4, checksum_enabled=1)
    check(uniform(32768, 24, 1, p), "checksum-thread")
if "concat" in which:
    js = jobgen.build_jobs(rng, 60, sizes=[64, 257, 2048, 4099], max_frames=3, layout=1)
    check(js, "concat-single-phase")
    check(js, "concat-two-phase", concat_tmp=pkg.batch.concat_tmp_size(js["jobs"], js["n_results"]))
if "decode" in which:
    js = jobgen.build_jobs(rng, 40, sizes=[7, 64, 1000, 4099], max_frames=3)
    got = check(js, "decode-encode")
    dj, dst_size, n_results = pkg.batch.decode_jobs_for(js["jobs"], got[1])
    out, dres, _ = pkg.batch.DeviceDecodeBatch(got[0], dj, dst_size, n_results).run().fetch()
    bad = sum(1 for j, job in enumerate(js["jobs"]) for f in range(int(dj[j]["n_frames"]))
              if int(dres[int(job["first_result"]) + f]) > 0xFFFFFF80)
    print("decode ok: frames with decoder errors", bad, flush=True)
print("sanitize_case done", flush=True)
