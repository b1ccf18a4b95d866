"""The BASELINE configurations at their full sizes: every stream the GPU produces against the CPU reference
(oracle/_ref/libcmp_ref.so where it was built, else the oracle port) - sizes and 64-bit hashes of ALL frames
(include/airs_stream_hash.h; device: airs_cuda_hash_streams, host: oracle/hash_jobs.h) and the bytes of every
100th job (SURVEY.md section 8d "Parity check")."""
import numpy as np
import pytest

import jobgen

pytestmark = pytest.mark.gpu


def _check(pkg, w, sample_every=100, max_jobs=None, layout=0, lib=None):
    import torch
    from oracle import parity
    data = w["data"].view(torch.uint8).reshape(-1)
    db = pkg.batch.DeviceBatch(data, w["jobs"], w["dst_size"], w["work_size"], w["n_results"], layout=layout)
    db.run()
    torch.cuda.synchronize()
    rep = parity.check_device_batch(db, w["jobs"], data, sample_every=sample_every, max_jobs=max_jobs, lib=lib)
    assert rep["sizes_identical"], rep
    assert rep["hashes_identical"], rep
    assert rep["bytes_identical"] and rep["byte_compared_frames"] > 0, rep
    return rep


def test_hash_streams_against_host_hash(gpu, oracle, pkg):
    """airs_cuda_hash_streams == the host hash of the same bytes, SLOTS and CONCAT, odd sizes and alignments."""
    import torch
    from oracle import parity
    rng = np.random.default_rng(11)
    for layout in (0, 1):
        js = jobgen.build_jobs(rng, 200, sizes=[1, 2, 7, 64, 65, 257, 1000, 2049, 4099], max_frames=3, layout=layout,
                               capacity_modes=["bound", "raw", "tight", "big"])
        db = pkg.batch.DeviceBatch(js["src"], js["jobs"], js["dst_size"], js["work_size"], js["n_results"], layout=layout,
                                   concat_tmp=pkg.batch.concat_tmp_size(js["jobs"], js["n_results"]) if layout else 0)
        db.run()
        h = db.hash_streams().cpu().numpy().view(np.uint64)
        res, hh = oracle.hash_jobs(js["src"], js["jobs"], threads=4)
        assert np.array_equal(db.results.cpu().numpy().view(np.uint32), res)
        assert np.array_equal(h, hh)
    # ranges: the streams of the CONCAT layout again, by offset and size
    offs = db.out_offsets[:-1].contiguous()
    got = pkg.batch.hash_ranges(db.dst, offs, db.results).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, hh)


def test_random_jobs_against_compiled_reference(gpu, ref):
    """The differential test of test_gpu_parity.py with the compiled, unmodified reference as the arbiter."""
    for seed in range(3):
        rng = np.random.default_rng(700 + seed)
        js = jobgen.build_jobs(rng, 250, sizes=[1, 3, 8, 63, 65, 256, 1000, 2048, 2049, 4099, 10000], max_frames=4,
                               allow_invalid=True)
        want = jobgen.run_cpu(ref, js)
        got = gpu.run_jobs_device(js)
        jobgen.compare(want, got, js, "gpu-vs-compiled-reference")


def test_config3_full_size(gpu, pkg):
    """BASELINE config 3 as stated: 1 Mi independent 4 KiB chunks, mixed parameters, escape heavy (4 GiB)."""
    rep = _check(pkg, pkg.workloads.config3(1 << 20, 0, "cuda"))
    assert rep["frames"] == 1 << 20 and rep["frames_ok_on_cpu"] == 1 << 20


def test_config2_batched_1gib(gpu, pkg):
    """Config 2: 64 contexts x 256 frames x 64 KiB (1 GiB), model update; every frame hashed, model state included
    implicitly (a wrong model shows in every later frame)."""
    rep = _check(pkg, pkg.workloads.config2(64, 0, "cuda"), sample_every=16)
    assert rep["frames"] == 64 * 256


def test_config2_batched_runs_of_model_passes(gpu, pkg):
    """Config 2 with enough contexts (320: more than a third of the resident CTAs) that they take the CTA-per-context
    kernel and, their slots holding cmp_compress_bound() bytes, the runs of model passes with the frames in the inner
    loop (model_run_fast): 5 GiB, every one of the 81920 frames hashed against the compiled reference - and the same
    contexts with slots of 2 n + 64 bytes (frame by frame)."""
    for bound in (True, False):
        rep = _check(pkg, pkg.workloads.config2(320, 5, "cuda", bound_slots=bound), sample_every=40)
        assert rep["frames"] == 320 * 256 and rep["frames_ok_on_cpu"] == 320 * 256


@pytest.mark.parametrize("row", range(15))
def test_config4_rows_1gib(gpu, pkg, row):
    """The named subset of config 4, 1 GiB per row in the default cut of 512 x 2 MiB chunks."""
    wl = pkg.workloads
    rep = _check(pkg, wl.config4_row(wl.CONFIG4_ROWS[row], device="cuda"), sample_every=64)
    assert rep["frames"] == 512 and rep["frames_ok_on_cpu"] == 512


@pytest.mark.parametrize("cut", [(16384, 1 << 15), (262144, 1 << 11)])
def test_config4_other_cuts(gpu, pkg, cut):
    """... and the other two cuts of SURVEY.md section 8 (16384 x 64 KiB, 262144 x 4 KiB) for DIFF + GOLOMB_ZERO g16,
    DIFF + GOLOMB_MULTI g3 and IWT + GOLOMB_ZERO g16."""
    wl = pkg.workloads
    for row in (5, 8, 12):
        rep = _check(pkg, wl.config4_row(wl.CONFIG4_ROWS[row], chunk_samples=cut[1], device="cuda"), sample_every=997)
        assert rep["frames"] == cut[0]


def test_config5_shard_8gib(gpu, pkg):
    """Config 5: the shard one of 8 GPUs takes of the 64 GiB stream (4096 chunks x 2 MiB), CONCAT layout laid out by
    the device-wide scan: offsets follow from the sizes, every stream hashed against the CPU reference."""
    import torch
    from oracle import parity
    w = pkg.workloads.config5(4096, 12288, "cuda")
    data = w["data"].view(torch.uint8).reshape(-1)
    tmp = pkg.batch.concat_tmp_size(w["jobs"], w["n_results"])
    db = pkg.batch.DeviceBatch(data, w["jobs"], data.numel(), w["work_size"], w["n_results"], layout=1, concat_tmp=tmp)
    db.run()
    h = db.hash_streams().cpu().numpy().view(np.uint64)
    res = db.results.cpu().numpy().view(np.uint32)
    offs = db.out_offsets.cpu().numpy()
    assert np.array_equal(np.diff(offs), res.astype(np.int64))
    lib, _ = parity.cpu_library()
    src_h = data.cpu().numpy()
    res_c, h_c = lib.hash_jobs(src_h, w["jobs"], threads=__import__("os").cpu_count())
    assert np.array_equal(res, res_c) and np.array_equal(h, h_c)


@pytest.mark.parametrize("n", [65536, 1 << 20, (1 << 20) + 4097])
def test_iwt_large_frames_bytes(gpu, oracle, pkg, n):
    """IWT at 65536 and 1 Mi samples (and an odd size above): byte for byte against the oracle, with and without
    a secondary pass over the same work buffer."""
    abi = pkg.abi
    rng = np.random.default_rng(n)
    for enc, g, sec in ((1, 16, 0), (2, 7, 0), (0, 0, 0), (1, 16, 2)):
        p = abi.make_params(primary_preprocessing=abi.PRE_IWT, primary_encoder_type=enc, primary_encoder_param=g,
                            primary_encoder_outlier=60, secondary_iterations=sec, secondary_preprocessing=abi.PRE_IWT,
                            secondary_encoder_type=enc, secondary_encoder_param=g, secondary_encoder_outlier=60)
        nf = 3 if sec else 1
        jobs, dsz, wsz = pkg.workloads.uniform_jobs(3, n, nf, abi.compress_bound(2 * n), model=True)
        jobs["params"] = p
        x = pkg.synth.chunks(7, 0, 3 * nf, n)
        js = dict(src=x.view(np.uint8).reshape(-1), jobs=jobs, dst_size=dsz, work_size=wsz, n_results=3 * nf, layout=0)
        jobgen.compare(jobgen.run_cpu(oracle, js, threads=3), gpu.run_jobs_device(js), js, "iwt-large")
