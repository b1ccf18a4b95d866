"""The batched decoder (include/airs_cuda_decode.h) against the reference's own streams.

The reference ships no decoder, so the check is the round trip its format defines: streams
produced by the oracle (the compiled reference, CPU) for random parameter sets must decode, on
the GPU, to the samples that went in.  A second set of tests closes the loop on the device alone
at the sizes of the BASELINE configurations: CUDA encoder -> CUDA decoder -> the input.
"""
import numpy as np
import pytest

import jobgen

pytestmark = pytest.mark.gpu

SMALL = [1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 63, 64, 65, 255, 256, 257]
MEDIUM = [1000, 2047, 2048, 2049, 4099, 6000, 10000]


def expected_samples(js, job, f):
    """Frame f of a job as the decoder returns it: 16-bit samples, or sign-extended words."""
    o = int(job["src_offset"]) + f * int(job["src_frame_stride"])
    raw = js["src"][o:o + int(job["src_size"])]
    if int(job["dtype"]) == 1:
        lo = raw.view("<u4").astype(np.uint32) & 0xFFFF
        return lo.astype(np.uint16).view(np.int16).astype("<i4").view(np.uint8)
    return raw


def decode_and_check(pkg, gpu, js, run, what):
    abi = pkg.abi
    dst, res, _init, offs, _work = run
    concat = js["layout"] == abi.LAYOUT_CONCAT
    dj, dst_size, n_results = gpu.decode_jobs_for(js["jobs"], res, offs if concat else None)
    db = gpu.DeviceDecodeBatch(dst, dj, dst_size, n_results, fill=0xEE).run()
    out, dres, info = db.fetch()
    checked = 0
    for j, job in enumerate(js["jobs"]):
        for f in range(int(dj[j]["n_frames"])):
            k = int(job["first_result"]) + f
            want = expected_samples(js, job, f)
            assert int(dres[k]) == len(want), (
                f"{what}: job {j} frame {f}: result {int(dres[k]):#x}, want {len(want)} bytes; "
                f"params={job['params']} dtype={job['dtype']} hdr={info[k]}")
            o = int(dj[j]["dst_offset"]) + f * int(dj[j]["dst_frame_stride"])
            got = out[o:o + len(want)]
            if not np.array_equal(got, want):
                at = int(np.nonzero(got != want)[0][0])
                raise AssertionError(f"{what}: job {j} frame {f} differs at byte {at} of {len(want)}; "
                                     f"params={job['params']} dtype={job['dtype']} hdr={info[k]}")
            assert int(info[k]["compressed_size"]) == int(res[k])
            assert int(info[k]["original_size"]) == 2 * (len(want) // (4 if int(job["dtype"]) == 1 else 2))
            checked += 1
    return checked


@pytest.mark.parametrize("seed", range(6))
def test_decode_reference_streams_small(gpu, oracle, pkg, seed):
    rng = np.random.default_rng(900 + seed)
    js = jobgen.build_jobs(rng, 300, sizes=SMALL, max_frames=6, allow_invalid=True)
    n = decode_and_check(pkg, gpu, js, jobgen.run_cpu(oracle, js), "decode-small")
    assert n > 300


@pytest.mark.parametrize("seed", range(4))
def test_decode_reference_streams_multi_tile(gpu, oracle, pkg, seed):
    rng = np.random.default_rng(950 + seed)
    js = jobgen.build_jobs(rng, 120, sizes=MEDIUM, max_frames=5,
                           capacity_modes=["bound", "bound", "raw", "big"])
    n = decode_and_check(pkg, gpu, js, jobgen.run_cpu(oracle, js), "decode-medium")
    assert n > 120


@pytest.mark.parametrize("seed", range(2))
def test_decode_back_to_back(gpu, oracle, pkg, seed):
    """CONCAT layout: the decoder finds the streams of a sequence by walking their headers."""
    rng = np.random.default_rng(980 + seed)
    js = jobgen.build_jobs(rng, 200, sizes=SMALL + [2048, 4099], max_frames=4, layout=1,
                           capacity_modes=["bound", "raw", "big"])
    n = decode_and_check(pkg, gpu, js, jobgen.run_cpu(oracle, js), "decode-concat")
    assert n > 200


def test_header_fields(gpu, oracle, pkg):
    """airs_frame_info carries what cmp_hdr_deserialize reads (ref header.c:89-134; the
    reference's vectors for it are test/test_header.c:89-163)."""
    abi = pkg.abi
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=2, primary_encoder_param=10,
                        primary_encoder_outlier=42, secondary_iterations=3, secondary_preprocessing=abi.PRE_MODEL,
                        secondary_encoder_type=1, secondary_encoder_param=7, model_rate=11, checksum_enabled=1)
    rng = np.random.default_rng(5)
    js = jobgen.build_jobs(rng, 1, sizes=[777], max_frames=1, params_fn=lambda r: p, dtypes=(2,),
                           capacity_modes=["bound"], styles=("walk",))
    js["jobs"][0]["n_frames"] = 1
    js["jobs"][0]["identifier_base"] = 0x123456789A
    run = jobgen.run_cpu(oracle, js)
    dj, dst_size, n_results = gpu.decode_jobs_for(js["jobs"], run[1])
    out, dres, info = gpu.DeviceDecodeBatch(run[0], dj, dst_size, n_results).run().fetch()
    h = info[0]
    assert int(dres[0]) == 2 * 777
    assert int(h["version"]) == 0x8000 | 600
    assert int(h["compressed_size"]) == int(run[1][0]) and int(h["original_size"]) == 2 * 777
    assert int(h["identifier"]) == 0x123456789A + 1  # first identifier drawn after the base (ref cmp.c:438-450)
    assert (int(h["sequence_number"]), int(h["preprocessing"]), int(h["checksum_enabled"])) == (0, abi.PRE_DIFF, 1)
    assert (int(h["encoder_type"]), int(h["encoder_param"]), int(h["encoder_outlier"])) == (2, 10, 42)
    assert int(h["header_size"]) == 22 and int(h["model_rate"]) == 0  # rate only in MODEL passes (ref cmp.c:265-279)


def test_damaged_streams(gpu, oracle, pkg):
    """Damage is reported, never decoded silently when a checksum is there, and never crashes."""
    abi = pkg.abi
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16,
                        secondary_iterations=2, secondary_preprocessing=abi.PRE_MODEL, secondary_encoder_type=1,
                        secondary_encoder_param=8, model_rate=8, checksum_enabled=1)
    js = jobgen.build_jobs(np.random.default_rng(6), 6, sizes=[3000], max_frames=3, params_fn=lambda r: p,
                           dtypes=(2,), capacity_modes=["bound"], styles=("walk",))
    dst, res, *_ = jobgen.run_cpu(oracle, js)
    assert not any(abi.is_error(int(r)) for r in res)
    dst = dst.copy()
    jobs = js["jobs"]

    def stream(j, f=0):
        return int(jobs[j]["dst_offset"]) + f * int(jobs[j]["dst_frame_stride"])

    dst[stream(0) + 200] ^= 0x10          # a flipped bit in the code words of job 0, frame 0
    dst[stream(1) + 0] = 0x12             # bad version
    dst[stream(2) + 15] = 0x7F            # unknown preprocessing / encoder
    dst[stream(3) + 2:stream(3) + 5] = [0xFF, 0xFF, 0xFF]  # compressed size beyond the slot
    dst[stream(4) + int(res[int(jobs[4]["first_result"])]) - 1] ^= 1  # trailer
    dj, dst_size, n_results = gpu.decode_jobs_for(jobs, res)
    out, dres, info = gpu.DeviceDecodeBatch(dst, dj, dst_size, n_results).run().fetch()
    k = [int(j["first_result"]) for j in jobs]
    assert int(dres[k[0]]) in (abi.dec_err("CHECKSUM"), abi.dec_err("CORRUPT"))
    assert int(dres[k[1]]) == abi.err("INT_HDR")
    assert int(dres[k[2]]) == abi.err("INT_HDR")
    assert int(dres[k[3]]) == abi.err("SRC_SIZE_WRONG")
    assert int(dres[k[4]]) == abi.dec_err("CHECKSUM")
    # model frames behind a first frame that could not be decoded have no model (a wrong model, if
    # only the checksum noticed: then their own checksums fail)
    for j in range(4):
        for f in range(1, int(jobs[j]["n_frames"])):
            ok = (abi.dec_err("NO_MODEL"), abi.dec_err("CHECKSUM")) if j == 0 else (abi.dec_err("NO_MODEL"),)
            assert int(dres[k[j] + f]) in ok, (j, f, hex(int(dres[k[j] + f])))
    # the untouched job decodes, and so do the frames behind the damaged trailer
    for j, f0 in ((5, 0), (4, 1)):
        for f in range(f0, int(jobs[j]["n_frames"])):
            want = expected_samples(js, jobs[j], f)
            o = int(dj[j]["dst_offset"]) + f * int(dj[j]["dst_frame_stride"])
            assert int(dres[k[j] + f]) == len(want) and np.array_equal(out[o:o + len(want)], want)


def _device_round_trip(pkg, gpu, x_dev, n_jobs, n, n_frames, params, dtype=2):
    """CUDA encoder -> CUDA decoder on the device; returns True if every sample came back."""
    import torch
    from test_gpu_parity import _uniform_jobs
    abi = pkg.abi
    js = _uniform_jobs(pkg, n_jobs, n, n_frames, params, dtype=dtype, cap=16 + 2 * n + 4 + 64)
    js["jobs"]["params"]["uncompressed_fallback_enabled"] = 1
    enc = gpu.DeviceBatch(x_dev, js["jobs"], js["dst_size"], js["work_size"], js["n_results"])
    enc.run()
    torch.cuda.synchronize()
    res = enc.results.cpu().numpy().view(np.uint32)
    assert not np.any(res > 0xFFFFFF80), "encoder reported errors"
    dj = np.zeros(n_jobs, dtype=abi.DEC_JOB_DTYPE)
    jobs = js["jobs"]
    dj["src_offset"] = jobs["dst_offset"]
    dj["src_frame_stride"] = jobs["dst_frame_stride"]
    dj["src_size"] = jobs["dst_capacity"]
    dj["dst_offset"] = jobs["src_offset"]
    dj["dst_frame_stride"] = jobs["src_frame_stride"]
    dj["dst_capacity"] = jobs["src_size"]
    dj["n_frames"] = n_frames
    dj["dtype"] = dtype
    dj["first_result"] = jobs["first_result"]
    dec = gpu.DeviceDecodeBatch(enc.dst, dj, x_dev.numel(), js["n_results"]).run()
    torch.cuda.synchronize()
    dres = dec.results.cpu().numpy().view(np.uint32)[:js["n_results"]]
    assert np.all(dres == jobs["src_size"][0]), f"decoder results: {np.unique(dres)}"
    return bool(torch.equal(dec.dst[:x_dev.numel()], x_dev)), float(res.sum()) / x_dev.numel()


def test_round_trip_config2_size(gpu, pkg):
    """BASELINE config 2, batched: 64 contexts x 256 frames x 32768 samples (1 GiB), model
    preprocessing with update, checksum on; what comes out of the decoder is what went in."""
    abi, synth = pkg.abi, pkg.synth
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16,
                        secondary_iterations=255, secondary_preprocessing=abi.PRE_MODEL,
                        secondary_encoder_type=1, secondary_encoder_param=8, model_rate=11, checksum_enabled=1)
    x = synth.frames_torch(1, 0, 64, 256, 32768).view(-1).view(dtype=__import__("torch").uint8)
    same, ratio = _device_round_trip(pkg, gpu, x, 64, 32768, 256, p)
    assert same and ratio < 0.5


def test_round_trip_config4_size(gpu, pkg):
    """BASELINE config 4 default cut: 1 GiB as 512 chunks of 2 MiB, DIFF + Golomb (both escapes) and IWT."""
    import torch
    abi, synth = pkg.abi, pkg.synth
    x = synth.chunks_torch(1, 0, 512, 1 << 20).view(-1).view(dtype=torch.uint8)
    for pre, enc, g, outl in [(abi.PRE_DIFF, 1, 16, 0), (abi.PRE_DIFF, 2, 16, 200), (abi.PRE_IWT, 1, 16, 0),
                              (abi.PRE_NONE, 0, 0, 0)]:
        p = abi.make_params(primary_preprocessing=pre, primary_encoder_type=enc, primary_encoder_param=g,
                            primary_encoder_outlier=outl, checksum_enabled=1)
        same, ratio = _device_round_trip(pkg, gpu, x, 512, 1 << 20, 1, p)
        assert same, (pre, enc, g)


def test_round_trip_many_short_frames(gpu, pkg):
    """From 65536 frames on the decoder takes one thread per frame instead of one warp: 70000 chunks of
    1 KiB, DIFF + both Golomb encoders, checksum on, through both containers."""
    import torch
    abi, synth = pkg.abi, pkg.synth
    x = synth.chunks_torch(3, 0, 70000, 512, esc=4).view(-1).view(dtype=torch.uint8)
    for enc, g, outl, dtype in [(1, 7, 0, 2), (2, 16, 60, 0)]:
        p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=enc, primary_encoder_param=g,
                            primary_encoder_outlier=outl, checksum_enabled=1)
        same, ratio = _device_round_trip(pkg, gpu, x, 70000, 512, 1, p, dtype=dtype)
        assert same, (enc, g, dtype)
