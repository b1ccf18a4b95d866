"""Residual statistics and the Golomb parameter they suggest: the kernel against a
numpy restatement (integers: exact), and the suggestion against a sweep of the parameter."""
import numpy as np
import pytest

import jobgen

pytestmark = pytest.mark.gpu


def numpy_stats(js, job):
    raw = js["src"][int(job["src_offset"]):int(job["src_offset"]) + int(job["src_size"])]
    x = (raw.view("<u4") & 0xFFFF).astype(np.int64) if int(job["dtype"]) == 1 else raw.view("<u2").astype(np.int64)
    r = x.copy()
    if int(job["params"]["primary_preprocessing"]) != 0:
        r[1:] = x[1:] - x[:-1]
    r &= 0xFFFF                                                  # int16 wrap (ref preprocess.c:284-290)
    m = ((r << 1) ^ np.where(r >> 15, 0xFFFFFFFF, 0)) & 0xFFFF   # ref map_to_unsigned, encoder.c:274-286
    bins = np.where(m == 0, 0, np.floor(np.log2(np.maximum(m, 1))).astype(np.int64) + 1)
    return int(m.sum()), len(m), int(m.max()), np.bincount(bins, minlength=17)


@pytest.mark.parametrize("seed", range(3))
def test_stats_match_numpy(gpu, pkg, seed):
    rng = np.random.default_rng(700 + seed)
    js = jobgen.build_jobs(rng, 150, sizes=[1, 2, 7, 8, 9, 64, 1000, 2049, 4099, 20000], max_frames=3)
    st = gpu.residual_stats(js["src"], js["jobs"])
    for j, job in enumerate(js["jobs"]):
        s, n, mx, hist = numpy_stats(js, job)
        assert (int(st[j]["sum_mapped"]), int(st[j]["n_samples"]), int(st[j]["max_mapped"])) == (s, n, mx), j
        assert np.array_equal(st[j]["log2_hist"].astype(np.int64), hist), j


def test_suggested_parameter_is_near_the_best(gpu, pkg):
    """g = mean * ln 2 from the statistics compresses within 2 % of the best g of a sweep, on walks of
    different roughness."""
    from test_gpu_parity import _uniform_jobs
    abi, synth = pkg.abi, pkg.synth
    lib = pkg.load_library()
    n = 1 << 16
    rng = np.random.default_rng(9)
    for amp in (2, 20, 300):
        x = (30000 + np.cumsum(rng.integers(-amp, amp + 1, size=(8, n)), axis=1)).astype(np.uint16)
        p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=1)
        js = _uniform_jobs(pkg, 8, n, 1, p)
        js["src"] = x.view(np.uint8).reshape(-1)
        st = gpu.residual_stats(js["src"], js["jobs"])
        g_sug = int(lib.airs_cuda_golomb_param_for_mean(int(st["sum_mapped"].sum()), int(st["n_samples"].sum())))
        sizes = {}
        for g in sorted({1, 2, 4, 8, 16, 32, 64, 128, 256, 512, g_sug}):
            js["jobs"]["params"]["primary_encoder_param"] = g
            res = gpu.run_jobs_device(js)[1]
            assert not np.any(res > 0xFFFFFF80)
            sizes[g] = int(res.astype(np.int64).sum())
        best = min(sizes.values())
        assert sizes[g_sug] <= 1.02 * best, (amp, g_sug, sizes)
