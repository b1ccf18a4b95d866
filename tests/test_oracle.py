"""Pins the oracle (oracle/oracle.c): reference KATs, golden streams made by the compiled
reference, and - where oracle/_ref exists - the compiled reference itself on random jobs."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import jobgen

abi = jobgen.abi
HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat_vectors.json")))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_streams.json")))


def one_frame_js(samples, params, dtype=2, identifier_base=0, cap=None):
    x = np.asarray(samples)
    raw = (x.astype(np.int64) & 0xFFFF).astype("<u4" if dtype == 1 else "<u2")
    n = len(x)
    jobs = np.zeros(1, dtype=abi.JOB_DTYPE)
    jobs[0]["src_frame_stride"] = raw.nbytes
    jobs[0]["dst_frame_stride"] = 26 + 6 * n + 6
    jobs[0]["identifier_base"] = identifier_base
    jobs[0]["src_size"] = raw.nbytes
    jobs[0]["dst_capacity"] = cap if cap is not None else 26 + 6 * n
    jobs[0]["work_size"] = 2 * n
    jobs[0]["n_frames"] = 1
    jobs[0]["dtype"] = dtype
    jobs[0]["params"] = params
    return dict(src=np.concatenate([raw.view(np.uint8), np.zeros(16, np.uint8)]), jobs=jobs,
                dst_size=64 + 8 * n, work_size=2 * n + 64, n_results=1, layout=0)


def stream_of(lib, js):
    dst, res, init, _, _ = jobgen.run_cpu(lib, js)
    assert init[0] == 0 and not abi.is_error(res[0]), hex(int(res[0]))
    return bytes(dst[:int(res[0])])


@pytest.mark.parametrize("v", KAT["encoder"], ids=lambda v: v["src"])
def test_kat_encoder(oracle, v):
    p = abi.make_params(primary_encoder_type=v["enc"], primary_encoder_param=v["g"],
                        primary_encoder_outlier=v["outlier"])
    s = stream_of(oracle, one_frame_js(v["input"], p))
    assert len(s) == 22 + len(v["payload"]) // 2
    assert s[22:].hex() == v["payload"]
    assert int.from_bytes(s[19:22], "big") == v["hdr_outlier"]          # derived outlier in the header
    assert int.from_bytes(s[17:19], "big") == v["g"]
    assert int.from_bytes(s[2:5], "big") == len(s) and int.from_bytes(s[5:8], "big") == 2 * len(v["input"])
    assert s[0:2] == bytes([0x82, 0x58]) and s[15] == v["enc"]


def test_kat_diff(oracle):
    v = KAT["diff"][0]
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF)
    s = stream_of(oracle, one_frame_js(v["input_u16"], p))
    got = np.frombuffer(s[22:], dtype=">i2")
    assert list(got) == v["residuals"]
    assert s[15] == (1 << 4)


@pytest.mark.parametrize("v", KAT["iwt"], ids=lambda v: v["src"])
@pytest.mark.parametrize("dtype", [0, 1, 2])
def test_kat_iwt(oracle, v, dtype):
    p = abi.make_params(primary_preprocessing=abi.PRE_IWT)
    s = stream_of(oracle, one_frame_js(v["input"], p, dtype=dtype))
    assert list(np.frombuffer(s[22:], dtype=">i2")) == v["output"]
    buf = np.array(v["input"], dtype=np.int16)
    oracle.lib.oracle_iwt(buf.ctypes.data, len(buf))
    assert list(buf) == v["output"]


@pytest.mark.parametrize("v", KAT["model_update"], ids=lambda v: v["src"])
def test_kat_model_update(oracle, v):
    got = [oracle.lib.oracle_model_update(d & 0xFFFF, m & 0xFFFF, v["rate"], v["dtype"])
           for d, m in zip(v["data"], v["model"])]
    assert got == [u & 0xFFFF for u in v["updated"]]


def test_kat_example_stream_and_checksum(oracle):
    v = KAT["example_streams"][0]
    p = abi.make_params(primary_preprocessing=v["pre"], primary_encoder_type=v["enc"],
                        primary_encoder_param=v["g"], checksum_enabled=v["checksum"])
    assert stream_of(oracle, one_frame_js(v["samples"], p, identifier_base=v["identifier_base"])).hex() == v["stream"]
    x = KAT["example_streams"][1]
    be = np.array(x["xxh32_samples"], dtype=">u2")
    assert "%08x" % oracle.lib.oracle_xxh32(be.ctypes.data, be.nbytes, 419764627) == x["xxh32"]
    try:
        import xxhash
    except ImportError:
        xxhash = None
    if xxhash:
        rng = np.random.default_rng(0)
        for n in [0, 1, 3, 4, 15, 16, 17, 31, 32, 33, 1000, 4097]:
            b = rng.integers(0, 256, n, dtype=np.uint8)
            buf = np.concatenate([b, np.zeros(4, np.uint8)])
            assert oracle.lib.oracle_xxh32(buf.ctypes.data, n, 419764627) == xxhash.xxh32(b.tobytes(), seed=419764627).intdigest()
    z = KAT["example_streams"][2]
    p = abi.make_params(primary_preprocessing=z["pre"], primary_encoder_type=z["enc"], primary_encoder_param=z["g"])
    s = stream_of(oracle, one_frame_js(z["samples"], p))
    assert s[22:].hex() == z["payload"] and int.from_bytes(s[19:22], "big") == z["hdr_outlier"]


def test_kat_fallback(oracle):
    """ref test/test_cmp.c:634-724: incompressible data falls back to header(NONE,UNCOMPRESSED) + raw samples."""
    x = [int(v) for v in np.random.default_rng(9).integers(0, 65536, 32)]   # 17 bits per escape > 16 raw
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=1,
                        uncompressed_fallback_enabled=1, checksum_enabled=1)
    js = one_frame_js(x, p, identifier_base=5)
    s = stream_of(oracle, js)
    assert len(s) == 16 + 2 * len(x) + 4 and s[15] == 0x08 and s[14] == 0
    assert list(np.frombuffer(s[16:-4], dtype=">u2")) == x
    assert int.from_bytes(s[8:14], "big") == 5 + 3        # init, first attempt, reset, raw pass
    # the same data without fallback is larger than the raw form
    p["uncompressed_fallback_enabled"] = 0
    assert len(stream_of(oracle, one_frame_js(x, p))) > len(s)


def test_header_layout(oracle):
    """ref test/test_header.c:23-52 pins the byte order of every header field; check ours field by field."""
    p = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=2, primary_encoder_param=0x1112,
                        primary_encoder_outlier=0x11415, checksum_enabled=1)
    s = stream_of(oracle, one_frame_js([7] * 9, p, identifier_base=0x08090A0B0C0D - 1))
    assert s[0:2].hex() == "8258" and s[5:8].hex() == "000012" and s[8:14].hex() == "08090a0b0c0d"
    assert s[14] == 0 and s[15] == (1 << 4) | (1 << 3) | 2 and s[16] == 0
    assert s[17:19].hex() == "1112" and s[19:22].hex() == "011415"


def test_bounds(oracle):
    """ref test/test_cmp.c:435-476 and SURVEY.md 8-a18"""
    L = oracle.lib
    assert L.oracle_compress_bound(4096) == 12314
    assert L.oracle_compress_bound(65536) == 196634
    assert L.oracle_compress_bound(2 * 2796198) == 26 + 6 * 2796198
    assert L.oracle_compress_bound(2 * 2796199) == abi.err("HDR_CMP_SIZE_TOO_LARGE")
    assert L.oracle_compress_bound((1 << 24)) == abi.err("HDR_ORIGINAL_TOO_LARGE")
    for n in (1, 5, 2048):
        assert abi.compress_bound(2 * n) == L.oracle_compress_bound(2 * n)


@pytest.mark.parametrize("ci", range(len(GOLD["cases"])))
def test_golden_streams(oracle, ci):
    """Streams produced by the compiled reference (tests/golden/make_golden.py)."""
    c = GOLD["cases"][ci]
    jobs = np.frombuffer(bytes.fromhex(c["jobs"]), dtype=abi.JOB_DTYPE).copy()
    js = dict(src=np.frombuffer(bytes.fromhex(c["src"]), dtype=np.uint8).copy(), jobs=jobs,
              dst_size=c["dst_size"], work_size=c["work_size"], n_results=c["n_results"], layout=0)
    dst, res, init, _, _ = jobgen.run_cpu(oracle, js)
    assert [int(x) for x in init] == c["init"]
    k = 0
    for j in range(len(jobs)):
        for f in range(int(jobs[j]["n_frames"])):
            fr = c["frames"][k]
            assert int(res[k]) == fr["result"], f"case {ci} frame {k}"
            if fr["stream"]:
                o = int(jobs[j]["dst_offset"]) + f * int(jobs[j]["dst_frame_stride"])
                assert dst[o:o + fr["result"]].tobytes().hex() == fr["stream"], f"case {ci} frame {k}"
            k += 1


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("layout", [0, 1])
def test_oracle_vs_compiled_reference(oracle, ref, seed, layout):
    rng = np.random.default_rng(1000 + seed)
    js = jobgen.build_jobs(rng, 150, sizes=[1, 2, 3, 5, 7, 8, 63, 64, 65, 2047, 2048, 2049, 4099],
                           max_frames=6, allow_invalid=True, layout=layout)
    jobgen.compare(jobgen.run_cpu(ref, js), jobgen.run_cpu(oracle, js), js, "ref-vs-oracle")


def test_reference_threads_agree(ref):
    """The CPU baseline leg of bench.py spreads jobs over threads: same bytes as one thread."""
    rng = np.random.default_rng(5)
    js = jobgen.build_jobs(rng, 64, sizes=[2048], max_frames=2)
    jobgen.compare(jobgen.run_cpu(ref, js, threads=1), jobgen.run_cpu(ref, js, threads=4), js, "threads")
