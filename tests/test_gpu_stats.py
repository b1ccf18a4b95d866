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


@pytest.mark.parametrize("dtype", [2, 0, 1])
def test_candidate_bits_are_exact_and_the_search_pays(gpu, oracle, pkg, dtype):
    """airs_cuda_candidate_bits (the parameter search the reference leaves to the user): for candidates around the
    statistics' suggestion - GOLOMB_ZERO and GOLOMB_MULTI, NONE and DIFF preprocessing, smooth / rough / escape-heavy
    data - the bit count per candidate is EXACT: compressing the frame with that candidate (oracle = reference) gives
    CMP_HDR_SIZE + 6 + ceil(bits / 8) bytes.  Invalid candidates report 2^64 - 1.  The best candidate beats the
    plain suggestion g = mean * ln 2 where escapes matter."""
    from test_gpu_parity import _uniform_jobs
    abi = pkg.abi
    rng = np.random.default_rng(31 + dtype)
    n = 12000
    gained = []
    for pre in (abi.PRE_DIFF, abi.PRE_NONE):
        for step, esc in ((3, 0), (60, 0), (900, 0), (12, 40)):
            x = (20000 + rng.integers(-step, step + 1, size=n).cumsum()) & 0xFFFF
            if pre == abi.PRE_NONE:
                x = (rng.integers(-step, step + 1, size=n) * 3) & 0xFFFF
            if esc:
                x[rng.integers(0, n, size=n // esc)] ^= rng.integers(0, 65536, size=n // esc)
            for enc in (1, 2):
                p = abi.make_params(primary_preprocessing=pre, primary_encoder_type=enc, primary_encoder_param=1, primary_encoder_outlier=16)
                js = _uniform_jobs(pkg, 1, n, 1, p, dtype=dtype)
                js["src"] = (x.astype("<i4") if dtype == 1 else x.astype(np.uint16)).view(np.uint8)
                st = gpu.residual_stats(js["src"], js["jobs"])
                cand = gpu.param_candidates(st[0], enc)
                assert 3 <= len(cand) <= abi.MAX_CANDIDATES and cand[0]["g"] == pkg.load_library().airs_cuda_golomb_param_for_mean(
                    int(st[0]["sum_mapped"]), int(st[0]["n_samples"]))
                bad = np.zeros(1, dtype=abi.CANDIDATE_DTYPE)
                bad[0] = (enc, 0, 5, 0)                                   # g = 0: cmp_initialise refuses it
                bits = gpu.candidate_bits(js["src"], js["jobs"], np.concatenate([cand, bad]))[0]
                assert int(bits[-1]) == 2**64 - 1
                for c, b in zip(cand, bits[:-1]):
                    js["jobs"]["params"]["primary_encoder_param"] = c["g"]
                    js["jobs"]["params"]["primary_encoder_outlier"] = c["outlier"]
                    _, res, init, _, _ = jobgen.run_cpu(oracle, js)
                    assert not abi.is_error(int(init[0])) and int(res[0]) == 22 + (int(b) + 7) // 8, (pre, step, esc, enc, c)
                gained.append(int(bits[0]) / int(bits[:-1].min()))
    assert min(gained) >= 1.0 and max(gained) > 1.02, gained
