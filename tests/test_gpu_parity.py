"""CUDA path vs oracle, through the C-ABI (airs_cuda_compress_batch), byte for byte."""
import numpy as np
import pytest

import jobgen

pytestmark = pytest.mark.gpu

SMALL = [1, 2, 3, 5, 7, 8, 63, 64, 65, 255, 256, 257]
MEDIUM = [1000, 2047, 2048, 2049, 4099, 6000, 10000]


@pytest.mark.parametrize("seed", range(6))
def test_random_jobs_small(gpu, oracle, seed):
    rng = np.random.default_rng(100 + seed)
    js = jobgen.build_jobs(rng, 300, sizes=SMALL, max_frames=5, allow_invalid=True)
    want = jobgen.run_cpu(oracle, js)
    got = gpu.run_jobs_device(js)
    jobgen.compare(want, got, js, "gpu-vs-oracle")


@pytest.mark.parametrize("seed", range(4))
def test_random_jobs_multi_tile(gpu, oracle, seed):
    rng = np.random.default_rng(200 + seed)
    js = jobgen.build_jobs(rng, 120, sizes=MEDIUM, max_frames=4, allow_invalid=False)
    want = jobgen.run_cpu(oracle, js)
    got = gpu.run_jobs_device(js)
    jobgen.compare(want, got, js, "gpu-vs-oracle")


@pytest.mark.parametrize("seed", range(4))
def test_concat_layout(gpu, oracle, seed):
    """Streams laid out back to back by the device-wide scan == the oracle loop's concatenation."""
    rng = np.random.default_rng(300 + seed)
    js = jobgen.build_jobs(rng, 250, sizes=SMALL + [2048, 2049, 4099], max_frames=4,
                           allow_invalid=True, layout=1)
    want = jobgen.run_cpu(oracle, js)
    got = gpu.run_jobs_device(js)
    jobgen.compare(want, got, js, "gpu-concat")


@pytest.mark.parametrize("seed", range(4))
def test_concat_two_phase(gpu, oracle, seed):
    """CONCAT with temporary memory: SLOTS-style encoding into temporary slots, scan, copy - the
    same bytes, offsets and results as the oracle loop (and as the single-phase path)."""
    rng = np.random.default_rng(320 + seed)
    js = jobgen.build_jobs(rng, 300, sizes=SMALL + [2048, 2049, 4099, 9000], max_frames=4,
                           allow_invalid=True, layout=1)
    want = jobgen.run_cpu(oracle, js)
    tmp = gpu.concat_tmp_size(js["jobs"], js["n_results"])
    got = gpu.run_jobs_device(js, concat_tmp=tmp)
    jobgen.compare(want, got, js, "gpu-concat-two-phase")


def test_concat_two_phase_gives_way(gpu, oracle):
    """Too little temporary memory, or a destination that cannot hold all streams: the gate on the
    device hands the batch to the single-phase path, which follows the reference loop through the
    overflow (a frame that does not fit fails and its context goes on from there)."""
    rng = np.random.default_rng(340)
    js = jobgen.build_jobs(rng, 120, sizes=[64, 257, 2048, 4099], max_frames=4, layout=1,
                           capacity_modes=["bound", "raw", "big"])
    want = jobgen.run_cpu(oracle, js)
    tmp = gpu.concat_tmp_size(js["jobs"], js["n_results"])
    got = gpu.run_jobs_device(js, concat_tmp=tmp // 3)          # slots do not fit
    jobgen.compare(want, got, js, "gpu-concat-small-tmp")
    # the reference loop knows no bound of the concatenation; the single-phase path defines it
    total = int(want[3][-1])
    js2 = dict(js, dst_size=(total * 2 // 3 + 63) // 64 * 64)    # streams do not fit
    one = gpu.run_jobs_device(js2)
    assert any(int(r) == jobgen.abi.err("DST_TOO_SMALL") for r in one[1])
    two = gpu.run_jobs_device(js2, concat_tmp=tmp)
    assert np.array_equal(one[1], two[1]) and np.array_equal(one[3], two[3])
    fit = int(js2["dst_size"])
    assert np.array_equal(one[0][:fit], two[0][:fit])


def test_bad_job_table_is_refused(gpu, pkg):
    """include/airs_cuda.h, contract of the job table: result indices beyond n_results (any layout) or frames that
    are not listed in order (CONCAT) - the batch is refused on the device, every result is CMP_ERR_GENERIC and
    nothing hangs (the look-back scan of the single-phase CONCAT path would wait for an index nobody covers)."""
    abi = pkg.abi
    generic = abi.err("GENERIC")
    rng = np.random.default_rng(9)
    for layout, breakage in ((0, "range"), (1, "range"), (1, "gap"), (1, "order")):
        js = jobgen.build_jobs(rng, 40, sizes=[64, 257, 2048], max_frames=3, layout=layout)
        jobs = js["jobs"].copy()
        if breakage == "range":
            jobs["first_result"][17] = js["n_results"] - 1
            jobs["n_frames"][17] = 2
        elif breakage == "gap":
            jobs["first_result"][20:] += 1
            js["n_results"] += 1
        else:
            a, b = jobs[5].copy(), jobs[6].copy()
            jobs[5], jobs[6] = b, a
        js["jobs"] = jobs
        for tmp in ((0,) if layout == 0 else (0, gpu.concat_tmp_size(jobs, js["n_results"]))):
            got = gpu.run_jobs_device(js, concat_tmp=tmp)
            assert all(int(r) == generic for r in got[1]), (layout, breakage, tmp)


def _to_big_endian(js):
    """The same job set with the samples of every 16-bit job stored big-endian (AIRS_DTYPE_BE): what a sample file of
    the reference's front end holds before programs/file.c:337-358 swaps it to host order."""
    src = js["src"].copy()
    jobs = js["jobs"].copy()
    for job in jobs:
        if int(job["dtype"]) == 1:
            continue
        for f in range(int(job["n_frames"])):
            o, sz = int(job["src_offset"]) + f * int(job["src_frame_stride"]), int(job["src_size"]) & ~1
            if o + sz <= len(src):
                src[o:o + sz] = src[o:o + sz].reshape(-1, 2)[:, ::-1].reshape(-1)
        job["dtype"] = int(job["dtype"]) | 4
    return dict(js, src=src, jobs=jobs)


@pytest.mark.parametrize("seed", range(3))
def test_big_endian_samples_random(gpu, oracle, seed):
    """the byte swap of a big-endian sample file fused into the loads.  Random jobs (every
    preprocessing, encoder, capacity mode, checksum, fallback, model contexts): the streams, results and models of the
    big-endian job set on the GPU are those of the host-order job set on the CPU."""
    rng = np.random.default_rng(600 + seed)
    js = jobgen.build_jobs(rng, 200, sizes=SMALL + MEDIUM + [40000], max_frames=4, allow_invalid=True)
    want = jobgen.run_cpu(oracle, js)
    be = _to_big_endian(js)
    got = gpu.run_jobs_device(be)
    jobgen.compare(want, got, js, "big-endian")
    if seed == 0:      # a batch that does not declare them (AIRS_BATCH_BIG_ENDIAN): an unknown container, as dtype 3 is
        db = gpu.DeviceBatch(be["src"], be["jobs"], be["dst_size"], be["work_size"], be["n_results"])
        db.desc.flags = 0
        res = db.run().fetch()[1]
        for job in be["jobs"]:
            if int(job["dtype"]) & 4 and int(job["n_frames"]):
                assert int(res[int(job["first_result"])]) == jobgen.abi.err("SRC_SIZE_WRONG")


def test_big_endian_samples_every_kernel(gpu, oracle, pkg):
    """... and through every kernel: short chunks (warp per job), a few long frames and a few contexts with model (tiles),
    many long frames (CTA per job), each with checksums (the hash is taken of the big-endian image: of the memory as
    it is)."""
    abi = pkg.abi
    rng = np.random.default_rng(77)
    plain = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=2, primary_encoder_param=11,
                            primary_encoder_outlier=90, checksum_enabled=1)
    model = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16,
                            secondary_iterations=3, secondary_preprocessing=abi.PRE_MODEL, secondary_encoder_type=1,
                            secondary_encoder_param=8, model_rate=7, checksum_enabled=1)
    for n_jobs, n, nf, p, dtype in ((600, 2048 + 5, 1, plain, 2), (3, 70001, 1, plain, 0), (3, 32768, 5, model, 0),
                                    (460, 36000, 1, plain, 2), (40000, 24, 1, plain, 2)):
        js = _uniform_jobs(pkg, n_jobs, n, nf, p, dtype=dtype)
        x = (20000 + rng.integers(-30, 31, size=n_jobs * nf * n).cumsum()) & 0xFFFF
        js["src"] = x.astype(np.uint16).view(np.uint8)
        want = jobgen.run_cpu(oracle, js, threads=8)
        got = gpu.run_jobs_device(_to_big_endian(js))
        jobgen.compare(want, got, js, "big-endian n_jobs=%d n=%d" % (n_jobs, n))


@pytest.mark.parametrize("layout", [0, 1])
def test_host_batch(gpu, oracle, layout):
    """airs_cuda_compress_batch_host: host buffers in, host buffers out."""
    rng = np.random.default_rng(400 + layout)
    js = jobgen.build_jobs(rng, 120, sizes=SMALL + [2048, 4099], max_frames=3, layout=layout)
    want = jobgen.run_cpu(oracle, js)
    got = gpu.run_jobs_host(js)
    # slot bytes behind a stream are unspecified on the host path (the whole dst range is copied back)
    jobgen.compare(want, got, js, "gpu-host-batch", check_tail=False)


def test_host_batch_pipelined(gpu, oracle, pkg):
    """A host batch large enough (>= 64 MiB, >= 8 jobs, SLOTS) to take the pipelined path of
    airs_cuda_compress_batch_host: groups of jobs copied in, encoded and copied back on three streams."""
    abi, synth = pkg.abi, pkg.synth
    n_jobs, n, nf = 48, 1 << 18, 3                      # 48 contexts x 3 frames x 512 KiB = 72 MiB
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16,
                        secondary_iterations=2, secondary_preprocessing=abi.PRE_MODEL,
                        secondary_encoder_type=2, secondary_encoder_param=8, secondary_encoder_outlier=60,
                        model_rate=11, checksum_enabled=1)
    js = _uniform_jobs(pkg, n_jobs, n, nf, p, cap=2 * n + 64)
    x = np.stack([synth.frames(3, c, nf, n) for c in range(n_jobs)])
    js["src"] = x.view(np.uint8).reshape(-1)
    assert js["src"].nbytes >= 64 << 20
    want = jobgen.run_cpu(oracle, js, threads=8)
    got = gpu.run_jobs_host(js)
    jobgen.compare(want, got, js, "gpu-host-pipelined", check_tail=False)


def test_host_batch_concat_pipelined(gpu, oracle, pkg):
    """CONCAT host batches of >= 64 MiB take the pipelined path: every group of jobs is a slice of the whole
    batch (its scan starts at the total the group before left on the device) and only compressed bytes travel
    back.  Multi-frame contexts with models and checksums, then 4 KiB chunks with mixed parameters; a destination
    that cannot hold the streams makes the call start over on the unpipelined path, whose answer is AIRS_E_ARGUMENT."""
    abi, synth = pkg.abi, pkg.synth
    n_jobs, n, nf = 48, 1 << 18, 3
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16,
                        secondary_iterations=2, secondary_preprocessing=abi.PRE_MODEL,
                        secondary_encoder_type=2, secondary_encoder_param=8, secondary_encoder_outlier=60,
                        model_rate=11, checksum_enabled=1)
    js = _uniform_jobs(pkg, n_jobs, n, nf, p, cap=2 * n + 64)
    js["layout"] = 1
    x = np.stack([synth.frames(3, c, nf, n) for c in range(n_jobs)])
    js["src"] = x.view(np.uint8).reshape(-1)
    assert js["src"].nbytes >= 64 << 20
    want = jobgen.run_cpu(oracle, js, threads=8)
    got = gpu.run_jobs_host(js)
    jobgen.compare(want, got, js, "gpu-host-concat-pipelined", check_tail=False)
    assert pkg.load_library().airs_cuda_last_launch_count() > 30      # several groups, ~11 launches each

    total = int(want[3][-1])
    js2 = dict(js, dst_size=(total * 2 // 3 + 63) // 64 * 64)          # streams do not fit: the call starts over on the
    with pytest.raises(RuntimeError, match="concatenated size"):       # unpipelined path, which reports it
        gpu.run_jobs_host(js2)

    w = pkg.workloads.config3(1 << 15, 5)                               # 32768 chunks of 4 KiB = 128 MiB
    js3 = {"src": np.ascontiguousarray(w["data"]).view(np.uint8).reshape(-1), "jobs": w["jobs"], "layout": 1,
           "dst_size": w["dst_size"], "work_size": 0, "n_results": w["n_results"]}
    want = jobgen.run_cpu(oracle, js3, threads=8)
    got = gpu.run_jobs_host(js3)
    jobgen.compare(want, got, js3, "gpu-host-concat-pipelined-c3", check_tail=False)


@pytest.mark.parametrize("enc,g,outl", [(1, 2, 0), (2, 3, 40), (1, 16, 0)])
def test_wide_table_arm(gpu, oracle, pkg, enc, g, outl):
    """Residuals that all sit in the codeword table but cost more than 8 bits per sample: the fast path's
    wide arm (one string per pair instead of per segment), next to ordinary and escape-heavy stretches."""
    abi = pkg.abi
    rng = np.random.default_rng(77 + g)
    n = 5 * 4096 + 24
    amp = 7 if g < 16 else 31
    steps = np.where(np.arange(n) % 2 == 0, amp, -amp) + rng.integers(-1, 1, size=n)   # |diff| close to the table edge
    steps[8192:12288] = rng.integers(-2, 3, size=4096)                                  # an easy stretch
    steps[13000:13040] = rng.integers(-3000, 3000, size=40)                             # and a few escapes
    x = (20000 + np.cumsum(steps)).astype(np.int64) & 0xFFFF
    frames = np.stack([x, (x + rng.integers(-amp, amp, size=n)) & 0xFFFF]).astype(np.uint16)
    for pre2, iters in ((abi.PRE_MODEL, 1), (abi.PRE_DIFF, 0)):
        p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=enc, primary_encoder_param=g,
                            primary_encoder_outlier=outl, secondary_iterations=iters, secondary_preprocessing=pre2,
                            secondary_encoder_type=enc, secondary_encoder_param=g, secondary_encoder_outlier=outl,
                            model_rate=3)
        js = _uniform_jobs(pkg, 3, n, 2, p)
        js["src"] = np.concatenate([frames.reshape(-1)] * 3).view(np.uint8)
        jobgen.compare(jobgen.run_cpu(oracle, js), gpu.run_jobs_device(js), js, "wide-arm")


def _uniform_jobs(pkg, n_jobs, n, n_frames, params, dtype=2, cap=None):
    abi = pkg.abi
    jobs = np.zeros(n_jobs, dtype=abi.JOB_DTYPE)
    stride = 4 if dtype == 1 else 2
    fb = n * stride
    cap = cap or abi.compress_bound(2 * n)
    slot = (cap + 15) // 16 * 16
    idx = np.arange(n_jobs, dtype=np.uint64)
    jobs["src_offset"] = idx * np.uint64(fb * n_frames)
    jobs["src_frame_stride"] = fb
    jobs["dst_offset"] = idx * np.uint64(slot * n_frames)
    jobs["dst_frame_stride"] = slot
    jobs["work_offset"] = idx * np.uint64(2 * n)
    jobs["identifier_base"] = idx * np.uint64(100000) + np.uint64(7)
    jobs["src_size"] = fb
    jobs["dst_capacity"] = cap
    jobs["work_size"] = 2 * n
    jobs["n_frames"] = n_frames
    jobs["dtype"] = dtype
    jobs["first_result"] = (idx * np.uint64(n_frames)).astype(np.uint32)
    jobs["params"] = params
    return dict(jobs=jobs, dst_size=int(slot * n_frames * n_jobs + 64), work_size=int(2 * n * n_jobs + 64),
                n_results=n_jobs * n_frames, layout=0)


def test_config1_one_mebisample_buffer(gpu, oracle, pkg):
    """BASELINE config 1: one 1 Mi-sample u16 buffer, diff + Golomb (zero and multi escape)."""
    abi, synth = pkg.abi, pkg.synth
    x = synth.chunks(1, 0, 1, 1 << 20)
    for enc, g, outl, cs in [(1, 16, 0, 0), (1, 8, 0, 1), (2, 16, 200, 0)]:
        p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=enc,
                            primary_encoder_param=g, primary_encoder_outlier=outl, checksum_enabled=cs)
        js = _uniform_jobs(pkg, 1, 1 << 20, 1, p)
        js["src"] = x.view(np.uint8).reshape(-1)
        jobgen.compare(jobgen.run_cpu(oracle, js), gpu.run_jobs_device(js), js, "config1")


@pytest.mark.parametrize("dtype", [2, 0])
def test_contexts_through_the_tile_kernel(gpu, oracle, pkg, dtype):
    """A batch of few contexts of several frames: the frames go through airs_tile_kernel tile by tile (a tile of a
    frame waits for the same tile of the frame before to leave its slice of the model).  Primary passes come back
    every sec_iter + 1 frames with a new identifier; secondary passes with model (rates 0, 5, 16), with differences,
    with nothing; GOLOMB_MULTI; ragged frame ends; checksums; a frame that does not fit sends its context to
    airs_encode_kernel, with and without the uncompressed fallback."""
    abi = pkg.abi
    rng = np.random.default_rng(50 + dtype)
    cases = [  # (n, frames, sec_iter, secondary preprocessing, secondary encoder, g2, rate, checksum, capacity, fallback)
        (32768, 9, 3, abi.PRE_MODEL, 1, 8, 5, 0, None, 0),
        (5000 * 8, 7, 2, abi.PRE_MODEL, 2, 6, 0, 1, None, 0),
        (2048 + 512, 6, 255, abi.PRE_MODEL, 1, 9, 16, 0, None, 0),
        (40000, 5, 1, abi.PRE_DIFF, 2, 12, 0, 1, None, 0),
        (4096, 5, 4, abi.PRE_NONE, 1, 4000, 0, 0, None, 0),
        (16384, 4, 0, abi.PRE_MODEL, 1, 8, 8, 0, None, 0),
        (32768, 6, 5, abi.PRE_MODEL, 1, 8, 8, 0, 30000, 0),     # some frames do not fit
        (32768, 6, 5, abi.PRE_MODEL, 1, 8, 8, 1, 65536 + 20, 1),  # ... and are stored raw
    ]
    for n, nf, it, pre2, enc2, g2, rate, cs, cap, fb in cases:
        p = abi.make_params(primary_preprocessing=abi.PRE_DIFF if n != 4096 else abi.PRE_NONE, primary_encoder_type=1,
                            primary_encoder_param=16 if n != 4096 else 3000, secondary_iterations=it, secondary_preprocessing=pre2,
                            secondary_encoder_type=enc2, secondary_encoder_param=g2, secondary_encoder_outlier=60,
                            model_rate=rate, checksum_enabled=cs, uncompressed_fallback_enabled=fb)
        js = _uniform_jobs(pkg, 3, n, nf, p, dtype=dtype, cap=cap)
        walk = rng.integers(-25, 26, size=(3, 1, n)).cumsum(axis=2) + 20000
        x = walk + rng.integers(-12, 13, size=(3, nf, n)) + np.arange(nf)[None, :, None] * 3
        if cap == 30000:
            x[:, 2] = rng.integers(0, 65536, size=(3, n))        # a frame of noise: larger than the slot
        if fb:
            x[:, 3] = rng.integers(0, 65536, size=(3, n))
        js["src"] = (x & 0xFFFF).astype(np.uint16).reshape(-1).view(np.uint8)
        want = jobgen.run_cpu(oracle, js, threads=3)
        got = gpu.run_jobs_device(js)
        jobgen.compare(want, got, js, "tile-contexts n=%d frames=%d iter=%d" % (n, nf, it))


@pytest.mark.parametrize("dtype", [2, 0])
def test_config2_model_frames(gpu, oracle, pkg, dtype):
    """BASELINE config 2: model-based preprocessing with model update over 256 consecutive 64 KiB frames."""
    abi, synth = pkg.abi, pkg.synth
    for rate in (8, 11):
        p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16,
                            secondary_iterations=255, secondary_preprocessing=abi.PRE_MODEL,
                            secondary_encoder_type=1, secondary_encoder_param=8, model_rate=rate)
        x = np.stack([synth.frames(1, c, 256, 32768) for c in range(2)])
        js = _uniform_jobs(pkg, 2, 32768, 256, p, dtype=dtype)
        js["src"] = x.view(np.uint8).reshape(-1)
        jobgen.compare(jobgen.run_cpu(oracle, js), gpu.run_jobs_device(js), js, "config2")


def test_config3_many_small_chunks(gpu, oracle, pkg):
    """BASELINE config 3 (reduced count): independent 4 KiB chunks, mixed Golomb parameters, escape heavy."""
    abi, synth = pkg.abi, pkg.synth
    n_chunks, n = 8192, 2048
    x = synth.chunks(1, 0, n_chunks, n, esc=32)
    js = _uniform_jobs(pkg, n_chunks, n, 1, abi.make_params())
    idx = np.arange(n_chunks, dtype=np.uint64)
    h = synth.mix(np.uint64(1) ^ idx)
    gtab = np.array([1, 2, 4, 7, 16, 60, 255, 1055], dtype=np.uint32)
    otab = np.array([5, 42, 107, 200], dtype=np.uint32)
    P = js["jobs"]["params"]
    P["primary_preprocessing"] = abi.PRE_DIFF
    P["primary_encoder_type"] = 1 + (idx & np.uint64(1)).astype(np.uint32)
    P["primary_encoder_param"] = gtab[(h & np.uint64(7)).astype(np.int64)]
    P["primary_encoder_outlier"] = otab[((h >> np.uint64(3)) & np.uint64(3)).astype(np.int64)]
    js["src"] = x.view(np.uint8).reshape(-1)
    for layout in (0, 1):
        js["layout"] = layout
        jobgen.compare(jobgen.run_cpu(oracle, js, threads=8), gpu.run_jobs_device(js), js, "config3")


def test_synth_torch_matches_numpy(gpu, pkg):
    import torch
    a = pkg.synth.chunks(3, 5, 7, 1000, esc=32)
    b = pkg.synth.chunks_torch(3, 5, 7, 1000, esc=32, device="cuda").cpu().numpy().view(np.uint16)
    assert np.array_equal(a, b)
    a = np.stack([pkg.synth.frames(2, c, 5, 513) for c in range(3, 5)])
    b = pkg.synth.frames_torch(2, 3, 2, 5, 513, device="cuda").cpu().numpy().view(np.uint16)
    assert np.array_equal(a, b)


def test_concat_two_phase_config2_size(gpu, pkg):
    """Config 2 batched (32 contexts x 256 frames x 64 KiB, model update, checksum): the concatenation
    of the two-phase CONCAT path holds exactly the streams the SLOTS layout produces, in result order."""
    import torch
    abi, synth = pkg.abi, pkg.synth
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16,
                        secondary_iterations=255, secondary_preprocessing=abi.PRE_MODEL,
                        secondary_encoder_type=1, secondary_encoder_param=8, model_rate=8, checksum_enabled=1)
    n_ctx, nf, n = 32, 256, 32768
    x = synth.frames_torch(1, 0, n_ctx, nf, n).view(-1).view(dtype=torch.uint8)
    js = _uniform_jobs(pkg, n_ctx, n, nf, p, cap=2 * n + 64)
    slots = gpu.DeviceBatch(x, js["jobs"], js["dst_size"], js["work_size"], js["n_results"]).run()
    tmp = gpu.concat_tmp_size(js["jobs"], js["n_results"])
    cat = gpu.DeviceBatch(x, js["jobs"], js["dst_size"], js["work_size"], js["n_results"], layout=1,
                          concat_tmp=tmp).run()
    torch.cuda.synchronize()
    res_s = slots.results.cpu().numpy().view(np.uint32)
    res_c = cat.results.cpu().numpy().view(np.uint32)
    assert np.array_equal(res_s, res_c) and not np.any(res_s > 0xFFFFFF80)
    offs = cat.out_offsets.cpu().numpy()
    assert np.array_equal(np.diff(offs), res_s.astype(np.int64))
    # gather the slot streams on the device and compare with the concatenation
    stride = int(js["jobs"][0]["dst_frame_stride"])
    sizes = torch.from_numpy(res_s.astype(np.int64)).cuda()
    starts = torch.arange(js["n_results"], device="cuda", dtype=torch.int64) * stride
    total = int(offs[-1])
    idx = torch.repeat_interleave(starts - torch.from_numpy(offs[:-1]).cuda(), sizes) + torch.arange(total, device="cuda")
    assert torch.equal(slots.dst[idx], cat.dst[:total])


@pytest.mark.parametrize("enc,g,outl,n", [(1, 16, 0, 2048), (2, 8, 40, 777), (1, 255, 0, 4099), (1, 3, 0, 20000)])
def test_short_jobs_with_one_encoder(gpu, oracle, pkg, enc, g, outl, n):
    """A batch of short chunks with one parameter set takes the warp-per-job kernel with the pair table per
    CTA (airs_small_kernel<true>); one job in eight uses another encoder and falls back to its warp's own
    table inside the same kernel."""
    abi, synth = pkg.abi, pkg.synth
    n_chunks = 1500
    x = synth.chunks(5, 0, n_chunks, n, esc=1)
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=enc, primary_encoder_param=g,
                        primary_encoder_outlier=outl, checksum_enabled=1)
    js = _uniform_jobs(pkg, n_chunks, n, 1, p)
    js["jobs"]["params"]["primary_encoder_param"][7::8] = g + 1
    js["src"] = x.view(np.uint8).reshape(-1)
    jobgen.compare(jobgen.run_cpu(oracle, js, threads=8), gpu.run_jobs_device(js), js, "short-one-encoder")


@pytest.mark.parametrize("enc,g,outl,pre,n", [(2, 3, 107, 0, 2048), (2, 1, 200, 0, 1000), (1, 3000, 0, 0, 4104), (2, 3, 107, 1, 1531)])
def test_short_jobs_longest_code_words(gpu, oracle, pkg, enc, g, outl, pre, n):
    """Warp-per-job kernel at its staging limit: every sample escapes with the longest code word the encoder has
    (48 bits for GOLOMB_MULTI with a small g, 28 for GOLOMB_ZERO with g = 3000), in every warp of every CTA."""
    abi = pkg.abi
    n_chunks = 64
    rng = np.random.default_rng(n)
    if pre == 0:
        x = np.full((n_chunks, n), 40000, dtype=np.uint16)           # zig-zag 51071: level 7 escapes
        x[1::2] += rng.integers(0, 3000, size=(n_chunks // 2, n)).astype(np.uint16)
    else:
        x = np.where(np.arange(n) % 2 == 0, 100, 30000)[None, :].repeat(n_chunks, 0).astype(np.uint16)   # |diff| 29900
    p = abi.make_params(primary_preprocessing=pre, primary_encoder_type=enc, primary_encoder_param=g,
                        primary_encoder_outlier=outl)
    js = _uniform_jobs(pkg, n_chunks, n, 1, p)
    js["jobs"]["dst_offset"] += 8 * (np.arange(n_chunks) % 2).astype(np.uint64)   # half of the slots 8 mod 16
    js["dst_size"] += 16
    js["src"] = x.view(np.uint8).reshape(-1)
    want = jobgen.run_cpu(oracle, js)
    res = want[1]
    assert int(res.max()) >= n * (48 if enc == 2 else 28) // 8          # the streams really are that long
    jobgen.compare(want, gpu.run_jobs_device(js), js, "longest-code-words")


@pytest.mark.parametrize("seed", range(4))
def test_random_jobs_long_frames(gpu, oracle, seed):
    """Frames of more than 32768 samples: single-frame jobs among them are cut into tiles that all CTAs share
    (airs_tile_kernel, bit-granular seams between tiles), the others take the CTA-per-job kernel; every capacity
    mode (fallback to raw storage and DST_TOO_SMALL included) and both slot alignments."""
    rng = np.random.default_rng(500 + seed)
    js = jobgen.build_jobs(rng, 48, sizes=[32769, 33000, 40960, 65536, 70001, 131077, 300000], max_frames=2,
                           allow_invalid=False)
    want = jobgen.run_cpu(oracle, js, threads=4)
    got = gpu.run_jobs_device(js)
    jobgen.compare(want, got, js, "gpu-vs-oracle-long")


def test_capacity_sweep_around_the_exact_size(gpu, oracle, pkg):
    """dst_capacity from four bytes below to two bytes above what the stream needs, with and without checksum and
    uncompressed fallback, through the warp-per-job kernel (2048 samples), the tile kernel (70001 samples, three jobs)
    and the CTA-per-job kernel (model context): success, DST_TOO_SMALL and the raw fallback switch at exactly the
    capacity where the reference switches."""
    abi = pkg.abi
    rng = np.random.default_rng(808)
    for n, nf, n_jobs in ((2048, 1, 20), (70001, 1, 3), (5000, 3, 4)):
        for cs, fb in ((0, 0), (1, 0), (0, 1), (1, 1)):
            p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16,
                                secondary_iterations=2 if nf > 1 else 0, secondary_preprocessing=abi.PRE_MODEL,
                                secondary_encoder_type=1, secondary_encoder_param=8, model_rate=8, checksum_enabled=cs,
                                uncompressed_fallback_enabled=fb)
            x = (20000 + rng.integers(-40, 41, size=n_jobs * nf * n).cumsum()) & 0xFFFF
            if fb:                                       # noise: the compressed stream is longer than the raw one
                x = rng.integers(0, 65536, size=n_jobs * nf * n)
            src = x.astype(np.uint16).view(np.uint8)
            js = _uniform_jobs(pkg, n_jobs, n, nf, p)
            js["src"] = src
            exact = int(jobgen.run_cpu(oracle, js)[1][0])        # size of the first stream with room to spare
            assert not abi.is_error(exact)
            target = 16 + 2 * n + 4 * cs if fb else exact       # where the decision falls: raw size / stream size
            for delta in (-4, -2, -1, 0, 1, 2):
                js = _uniform_jobs(pkg, n_jobs, n, nf, p, cap=target + delta)
                js["src"] = src
                want = jobgen.run_cpu(oracle, js, threads=4)
                got = gpu.run_jobs_device(js)
                jobgen.compare(want, got, js, "capacity n=%d cs=%d fb=%d delta=%d" % (n, cs, fb, delta))


def test_frames_beyond_the_header_limits(gpu, oracle, pkg):
    """A frame of more than 2^24 - 1 bytes cannot be described by the header (HDR_ORIGINAL_TOO_LARGE, header.c:34),
    the largest one that can is compressed; both as the reference answers, through the tile kernel and (CONCAT) the
    CTA-per-job kernel."""
    abi = pkg.abi
    rng = np.random.default_rng(99)
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16)
    for n in ((1 << 23) - 1, 1 << 23):                  # 16 MiB - 2 bytes, 16 MiB
        cap = min(abi.CMP_MAX_SIZE // 8 * 8, 22 + 6 * n)
        for layout in (0, 1):
            js = _uniform_jobs(pkg, 1, n, 1, p, cap=cap)
            js["layout"] = layout
            js["src"] = ((20000 + rng.integers(-20, 21, size=n).cumsum()) & 0xFFFF).astype(np.uint16).view(np.uint8)
            want = jobgen.run_cpu(oracle, js)
            got = gpu.run_jobs_device(js)
            jobgen.compare(want, got, js, "header-limit n=%d layout=%d" % (n, layout))
            assert abi.is_error(int(want[1][0])) == (n == 1 << 23)


def test_shards_give_the_bytes_of_the_whole(gpu, pkg):
    """the output does not depend on how the jobs are cut over GPUs.  Config 3 and config 2 workloads as
    ONE batch and as the two / three contiguous shards `parallel.shard_range` gives the ranks (each shard a batch of its
    own, as on its own GPU): the same stream hashes and sizes, frame by frame."""
    import torch
    abi = pkg.abi
    for w in (pkg.workloads.config3(3 * 4096 + 17, 7, "cuda"), pkg.workloads.config2(10, 3, "cuda")):
        data = w["data"].view(torch.uint8).reshape(-1)
        whole = gpu.DeviceBatch(data, w["jobs"], w["dst_size"], w["work_size"], w["n_results"])
        whole.run()
        h_all, r_all = whole.hash_streams().cpu().numpy(), whole.results.cpu().numpy()
        nf = int(w["jobs"][0]["n_frames"])
        for world in (2, 3):
            hs, rs = [], []
            for rank in range(world):
                lo, hi = pkg.parallel.shard_range(len(w["jobs"]), rank, world)
                jobs = w["jobs"][lo:hi].copy()
                jobs["first_result"] -= np.uint32(lo * nf)
                jobs["dst_offset"] -= jobs["dst_offset"][0]
                jobs["work_offset"] -= jobs["work_offset"][0]
                dsz = int(jobs["dst_offset"][-1] + jobs["dst_frame_stride"][-1] * nf) + 64
                db = gpu.DeviceBatch(data, jobs, dsz, w["work_size"], (hi - lo) * nf)
                db.run()
                hs.append(db.hash_streams().cpu().numpy())
                rs.append(db.results.cpu().numpy())
            assert np.array_equal(np.concatenate(hs), h_all) and np.array_equal(np.concatenate(rs), r_all), world


def test_thread_changes_device(gpu, oracle, pkg):
    """One host thread, two GPUs: the per-thread cache of the host-buffer entry points (staging buffers, streams,
    events) belongs to ONE device; after cudaSetDevice(other) it is given back under the old device and rebuilt, instead
    of handing the new device pointers into the old one's memory.  Needs two devices."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    rng = np.random.default_rng(12)
    js = jobgen.build_jobs(rng, 80, sizes=[64, 2048, 4099], max_frames=3)
    want = jobgen.run_cpu(oracle, js)
    cur = torch.cuda.current_device()
    try:
        for dev in (0, 1, 0, 1):
            torch.cuda.set_device(dev)
            got = gpu.run_jobs_host(js)
            jobgen.compare(want, got, js, "host batch on device %d" % dev, check_tail=False)
    finally:
        torch.cuda.set_device(cur)


IWT_SIZES = [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 255, 256, 257, 511, 513, 1000, 4095, 7679, 7680, 7681, 8191, 8192, 8193,
             8703, 8704, 8705, 9216, 16384, 16389, 24576 + 3, 32768, 32769, 65536 + 255, 100003]


@pytest.mark.parametrize("seed", range(3))
def test_iwt_transform_kernel(gpu, oracle, pkg, seed):
    """Single IWT frames take airs_iwt_kernel (tiles of 8192 samples with halos over all CTAs, levels 1-3 in registers,
    4-8 in shared memory, the rest by one CTA per frame) in front of the warp encoders: sizes around the tile, halo and
    level boundaries, both 16-bit containers, every data style (full-range samples wrap in every level), slots that
    are too small.  Streams, results AND the coefficients left in the work buffers (ref preprocess.c:337-353) equal
    the oracle's."""
    abi = pkg.abi
    rng = np.random.default_rng(900 + seed)

    def params(r):
        enc = int(r.integers(1, 3))
        return abi.make_params(primary_preprocessing=abi.PRE_IWT, primary_encoder_type=enc,
                               primary_encoder_param=int(r.choice([1, 3, 16, 60, 1000])),
                               primary_encoder_outlier=int(r.choice([5, 60, 300])) if enc == 2 else 0,
                               checksum_enabled=int(r.integers(0, 2)))
    js = jobgen.build_jobs(rng, 90, sizes=IWT_SIZES, max_frames=1, dtypes=(0, 2), params_fn=params,
                           capacity_modes=["bound", "bound", "bound", "big", "tight"])
    jobs = js["jobs"]
    jobs["work_size"] = 2 * (jobs["src_size"] // 2)   # (build_jobs also hands out work buffers that are too small)
    jobs["work_offset"] = np.concatenate(([0], np.cumsum((jobs["work_size"].astype(np.int64) + 15) // 16 * 16)[:-1]))
    js["work_size"] = int(jobs["work_offset"][-1]) + int(jobs["work_size"][-1]) + 64
    want = jobgen.run_cpu(oracle, js, threads=4)
    got = gpu.run_jobs_device(js)
    jobgen.compare(want, got, js, "iwt-kernel")
    for j in range(len(jobs)):
        o, n2 = int(jobs[j]["work_offset"]), int(jobs[j]["src_size"])
        assert np.array_equal(want[4][o:o + n2], got[4][o:o + n2]), f"job {j}: coefficients differ (n={n2 // 2})"


def test_iwt_longest_frames_of_the_transform_kernel(gpu, oracle, pkg):
    """2 Mi samples is the longest frame airs_iwt_kernel takes (8192 coefficients left for its last levels); a longer one
    stays with the CTA-per-job kernel."""
    abi = pkg.abi
    for n in (1 << 21, (1 << 21) + 8):
        p = abi.make_params(primary_preprocessing=abi.PRE_IWT, primary_encoder_type=1, primary_encoder_param=16)
        jobs, dsz, wsz = pkg.workloads.uniform_jobs(2, n, 1, abi.compress_bound(2 * n), model=True)
        jobs["params"] = p
        x = pkg.synth.chunks(11, 0, 2, n)
        js = dict(src=x.view(np.uint8).reshape(-1), jobs=jobs, dst_size=dsz, work_size=wsz, n_results=2, layout=0)
        want, got = jobgen.run_cpu(oracle, js, threads=2), gpu.run_jobs_device(js)
        jobgen.compare(want, got, js, "iwt-longest")
        assert np.array_equal(want[4][:4 * n], got[4][:4 * n])


@pytest.mark.parametrize("dtype", [2, 0])
def test_contexts_with_frames_in_the_inner_loop(gpu, oracle, pkg, dtype):
    """Runs of secondary MODEL passes take model_run_fast() in airs_encode_kernel: all frames of the run tile position by
    tile position, the model in registers, the streams of all frames growing side by side - contexts whose slots hold
    cmp_compress_bound() bytes (no frame can fail), in batches of 300 contexts (fewer would take the tile kernel).
    Frame counts and secondary iterations that put several primary passes and several runs into a context (runs are
    at most 64 frames long), both encoders, checksums, primary passes other than DIFF, rough frames in the middle of
    a run (code words computed arithmetically), a context whose samples walk across 0 / 65535 (small residuals between
    a sample and a model on either side of the boundary) - and what stays on the frame-by-frame path: rates 0 and 16,
    smaller slots, one of them with a frame that does not fit.  Streams, results, model bytes and the bytes behind the
    streams against the oracle."""
    abi, synth = pkg.abi, pkg.synth
    cases = [  # contexts, n, frames, sec_iter, rate, pre1, enc1, g1, enc2, g2, outlier, checksum, cap, rough frame
        (300, 4096, 3, 255, 8, abi.PRE_DIFF, 1, 16, 1, 8, 0, 0, None, None),
        (300, 8192, 8, 3, 11, abi.PRE_DIFF, 1, 16, 2, 8, 40, 1, None, None),
        (300, 4096, 70, 255, 8, abi.PRE_NONE, 1, 16, 1, 8, 0, 0, None, None),     # two runs: 64 and 5 frames
        (300, 8192, 21, 5, 5, abi.PRE_DIFF, 1, 16, 1, 8, 0, 0, None, 13),         # one frame of noise: residuals outside the table
        (300, 4096, 9, 255, 8, abi.PRE_IWT, 1, 16, 1, 8, 0, 1, None, None),
        (300, 6144, 5, 255, 8, abi.PRE_DIFF, 1, 16, 1, 8, 0, 0, None, None),      # not whole tiles of 4096 samples: frame by frame
        (3, 32768, 256, 255, 8, abi.PRE_DIFF, 1, 16, 1, 8, 0, 0, None, None),     # few contexts: the tile kernel
        (300, 4096, 6, 255, 0, abi.PRE_DIFF, 2, 16, 1, 4, 60, 0, None, None),     # rates 0 and 16: frame by frame
        (300, 4096, 6, 255, 16, abi.PRE_DIFF, 1, 16, 1, 8, 0, 1, None, None),
        (300, 4096, 12, 255, 8, abi.PRE_DIFF, 1, 16, 1, 8, 0, 0, 2 * 4096 + 64, None),  # smaller slots: frame by frame
        (3, 4096, 12, 255, 8, abi.PRE_DIFF, 1, 16, 1, 8, 0, 0, 3000, None),       # the DIFF frame (about 3.3 KB) does not fit its slot
    ]
    for nctx, n, nf, sec, rate, pre1, e1, g1, e2, g2, outl, cs, cap, rough in cases:
        p = abi.make_params(primary_preprocessing=pre1, primary_encoder_type=e1, primary_encoder_param=g1,
                            primary_encoder_outlier=outl if e1 == 2 else 0, secondary_iterations=sec,
                            secondary_preprocessing=abi.PRE_MODEL, secondary_encoder_type=e2, secondary_encoder_param=g2,
                            secondary_encoder_outlier=outl if e2 == 2 else 0, model_rate=rate, checksum_enabled=cs)
        x = np.stack([synth.frames(3, c, nf, n) for c in range(3)])
        if rough is not None:
            x[1, rough] = np.random.default_rng(5).integers(0, 65536, n).astype(np.uint16)
        # the third context: flat positions with a little noise - and every 64th one walks across the 0 / 65535 boundary
        # half way through the context (even positions upwards, odd ones downwards): small residuals between a sample
        # and a model on either side of it, next to residuals that all sit in the code word table
        rng2 = np.random.default_rng(77)
        y = rng2.integers(2000, 60000, n)[None, :] + rng2.integers(-3, 4, (nf, n))
        step = np.where(np.arange(nf) >= nf // 2, 8, 0)[:, None]
        y[:, 0::64] = 65530 + rng2.integers(-3, 4, (nf, len(range(0, n, 64)))) + step
        y[:, 33::64] = 5 + rng2.integers(-3, 4, (nf, len(range(33, n, 64)))) - step
        x[2] = (y & 0xFFFF).astype(np.uint16)
        x = np.tile(x, (nctx // 3, 1, 1))
        js = _uniform_jobs(pkg, nctx, n, nf, p, dtype=dtype, cap=cap)
        js["src"] = x.view(np.uint8).reshape(-1)
        jobgen.compare(jobgen.run_cpu(oracle, js, threads=8), gpu.run_jobs_device(js), js, f"frames-inner n={n} nf={nf} rate={rate}")
        if nf == 8:  # ... and through the CONCAT layout with temporary slots (the same kernels on the temporary job table)
            js["layout"] = 1
            got = gpu.run_jobs_device(js, concat_tmp=gpu.concat_tmp_size(js["jobs"], js["n_results"]))
            jobgen.compare(jobgen.run_cpu(oracle, js, threads=8), got, js, f"frames-inner concat n={n} nf={nf}")
