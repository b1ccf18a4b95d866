"""The C-ABI library loads, exports every symbol include/*.h declares, and its host-only
entry points (no device work) behave like the reference's.  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import jobgen

abi = jobgen.abi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(pkg):
    return pkg.load_library()


def test_every_declared_symbol_is_exported(lib, pkg):
    declared = set()
    for h in ("cmp.h", "cmp_errors.h", "airs_cuda.h", "airs_cuda_decode.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b((?:cmp|airs_cuda)_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations found"
    assert declared == set(pkg.loader.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_nothing_else_is_exported(pkg):
    """... and the other direction: the dynamic symbol table of libcmp_b200.so holds the declared functions and
    nothing else (no launch helpers, no kernel stubs) - lib/exports.map is the list the linker applies."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.loader.library_path()], capture_output=True, text=True, check=True)
    exported = {line.split()[-1] for line in out.stdout.splitlines() if line.strip()}
    assert exported == set(pkg.loader.EXPORTS), sorted(exported ^ set(pkg.loader.EXPORTS))
    listed = set(re.findall(r"^\s+((?:cmp|airs_cuda)_[a-z0-9_]+);", open(os.path.join(
        ROOT, "airs-compression_b200", "lib", "exports.map")).read(), flags=re.M))
    assert listed == set(pkg.loader.EXPORTS)


def test_struct_layouts():
    """params 44 bytes, context 80 bytes with the probed offsets."""
    assert C.sizeof(abi.CmpParams) == 44 and C.sizeof(abi.CmpContext) == 80
    assert abi.CmpParams.checksum_enabled.offset == 40 and abi.CmpParams.uncompressed_fallback_enabled.offset == 41
    ctx = abi.CmpContext
    assert (ctx.magic.offset, ctx.params.offset, ctx.work_buf.offset, ctx.work_buf_size.offset,
            ctx.model_size.offset, ctx.identifier.offset, ctx.sequence_number.offset) == (0, 4, 48, 56, 60, 64, 72)
    assert abi.JOB_DTYPE.itemsize == 120 and C.sizeof(abi.AirsBatch) == 104 and C.sizeof(abi.AirsHostBatch) == 96
    assert abi.STATS_DTYPE.itemsize == 88
    assert abi.DEC_JOB_DTYPE.itemsize == 56 and abi.FRAME_INFO_DTYPE.itemsize == 32 and C.sizeof(abi.AirsDecBatch) == 56


def test_golomb_param_for_mean(lib):
    """g = mean * ln 2 in integers: 1 <= g <= 65535, monotone in the mean."""
    assert lib.airs_cuda_golomb_param_for_mean(0, 0) == 1 and lib.airs_cuda_golomb_param_for_mean(0, 100) == 1
    assert lib.airs_cuda_golomb_param_for_mean(100 * 23, 100) == int(23 * 0.693147)
    assert lib.airs_cuda_golomb_param_for_mean(65535 * (1 << 20), 1 << 20) == int(65535 * 45426 / 65536)
    last = 0
    for mean in (1, 2, 3, 10, 100, 1000, 30000, 65535):
        g = lib.airs_cuda_golomb_param_for_mean(mean * 977, 977)
        assert 1 <= g <= 65535 and g >= last
        last = g


def test_error_helpers(lib):
    assert lib.cmp_is_error(abi.err("GENERIC")) and not lib.cmp_is_error(0) and not lib.cmp_is_error(0xFFFFFF80)
    assert lib.cmp_is_error(0xFFFFFF81)                      # ref test_cmp_errors.c:17-33: > (uint32_t)-128
    for name, code in abi.ERRORS.items():
        if name == "MAX_CODE":
            continue
        assert lib.cmp_get_error_code(abi.err(name)) == code
        assert lib.cmp_get_error_message(abi.err(name)) == lib.cmp_get_error_string(code)
    assert lib.cmp_get_error_code(1234) == 0
    assert lib.cmp_get_error_string(77) == b"Unspecified error code"


def test_bound_and_work_buf_size(lib, oracle):
    for size in [0, 1, 2, 4096, 65536, 2 * 2796198, 2 * 2796199, (1 << 24) - 1, 1 << 24, 0xFFFFFFFF]:
        assert lib.cmp_compress_bound(size) == oracle.lib.oracle_compress_bound(size), size
    rng = np.random.default_rng(1)
    for _ in range(300):
        p = abi.params_to_ctypes(jobgen.random_params(rng, allow_invalid=True))
        size = int(rng.choice([0, 1, 2, 3, 4097, 1 << 20, 0xFFFFFFFF]))
        assert lib.cmp_cal_work_buf_size(C.byref(p), size) == oracle.lib.oracle_work_buf_size(C.byref(p), size)
    assert lib.cmp_cal_work_buf_size(None, 8) == abi.err("GENERIC")


def test_initialise_validation_matches_oracle(lib, oracle):
    """cmp_initialise runs on the host: validation order and codes as ref cmp.c:152-209."""
    L = oracle.lib
    L.oracle_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]
    L.oracle_init.restype = C.c_uint32
    rng = np.random.default_rng(2)
    work = np.zeros(64, dtype=np.uint16)
    octx = (C.c_uint8 * 128)()
    for _ in range(600):
        p = abi.params_to_ctypes(jobgen.random_params(rng, allow_invalid=True))
        wptr = [work.ctypes.data, work.ctypes.data + 1, None][int(rng.integers(0, 3))]
        wsz = int(rng.choice([0, 2, 64, abi.err("GENERIC")]))
        ctx = abi.CmpContext()
        got = lib.cmp_initialise(C.byref(ctx), C.byref(p), wptr, wsz)
        want = L.oracle_init(octx, C.byref(p), wptr, wsz, 0)
        assert got == want
        if abi.is_error(got):
            assert bytes(ctx) == bytes(80)               # ref test_initialisation.c:58-71
        else:
            assert ctx.magic == 34021395 and ctx.sequence_number == 0 and ctx.model_size == 0
    assert lib.cmp_initialise(None, None, None, 0) == abi.err("GENERIC")
    ctx = abi.CmpContext()
    assert lib.cmp_initialise(C.byref(ctx), None, None, 0) == abi.err("GENERIC")
    assert lib.cmp_reset(C.byref(ctx)) == abi.err("CONTEXT_INVALID") and lib.cmp_reset(None) == abi.err("GENERIC")


def test_identifier_from_timestamp_callback(lib):
    """ref test_cmp.c:488-509: identifier = coarse << 16 | fine."""
    TS = C.CFUNCTYPE(None, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16))

    def cb(c, f):
        c[0], f[0] = 0x12345678, 0xABCD
    fn = TS(cb)
    lib.cmp_set_timestamp_func(C.cast(fn, C.c_void_p))
    try:
        ctx = abi.CmpContext()
        p = abi.params_to_ctypes(abi.make_params())
        assert lib.cmp_initialise(C.byref(ctx), C.byref(p), None, 0) == 0
        assert ctx.identifier == 0x12345678ABCD
    finally:
        lib.cmp_set_timestamp_func(None)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the compress entry points fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.airs_cuda_device_count() == 0
    assert lib.airs_cuda_concurrent_jobs() == 0
    ctx = abi.CmpContext()
    p = abi.params_to_ctypes(abi.make_params())
    assert lib.cmp_initialise(C.byref(ctx), C.byref(p), None, 0) == 0
    dst = np.zeros(16, dtype=np.uint64)
    src = np.arange(8, dtype=np.uint16)
    assert lib.cmp_compress_u16(C.byref(ctx), dst.ctypes.data, 128, src.ctypes.data, 16) == abi.err("GENERIC")
    assert b"CUDA" in lib.airs_cuda_last_error() or b"cuda" in lib.airs_cuda_last_error()
    # ... and so does the decoder
    db = abi.AirsDecBatch()
    jobs = np.zeros(1, dtype=abi.DEC_JOB_DTYPE)
    res = np.zeros(1, dtype=np.uint32)
    scratch = np.zeros(lib.airs_cuda_decode_scratch_size(1, 1) // 8 + 2, dtype=np.uint64)
    db.src, db.dst, db.jobs, db.results = dst.ctypes.data, src.ctypes.data, jobs.ctypes.data, res.ctypes.data
    db.scratch = (scratch.ctypes.data + 15) & ~15
    db.n_jobs = db.n_results = 1
    assert lib.airs_cuda_decompress_batch(C.byref(db), None) == -1  # AIRS_E_NO_DEVICE


def test_golomb_division_trick():
    """The kernel divides by g with a multiply-high (airs_device.cuh): exact for every reachable dividend.
    Sampled here; tools/micro/div.c-style exhaustive run documented in DESIGN.md."""
    rng = np.random.default_rng(3)
    gs = np.unique(np.concatenate([np.arange(1, 300), 2 ** np.arange(16), 2 ** np.arange(1, 16) - 1,
                                   2 ** np.arange(1, 16) + 1, rng.integers(1, 65536, 400), [65535, 65534, 32769]]))
    gs = gs[(gs >= 1) & (gs <= 65535)].astype(np.uint64)
    for g in gs:
        L = int(g).bit_length() - 1
        magic = ((1 << (32 + L)) - 1) // int(g)
        assert magic < (1 << 32)
        top = int((33 - L) * int(g) + 70000)
        u = np.unique(np.concatenate([np.arange(0, min(top, 5000)), rng.integers(0, top, 3000),
                                      (np.arange(1, 40)[:, None] * int(g) + np.array([-1, 0, 1])).ravel()]))
        u = u[(u >= 0) & (u < top)].astype(np.uint64)
        q = (((u + np.uint64(1)) * np.uint64(magic)) >> np.uint64(32)) >> np.uint64(L)
        assert np.array_equal(q, u // g), int(g)


def test_fast_kernel_division():
    """airs_fast_kernel's quotient (airs_fast.cuh): q = umulhi(x, ceil(2^32 / g)) for 2 <= g <= 32767, and
    umulhi(x + 1, 2^32 - 1) for g = 1, is floor(x / g) for every dividend the encoders can form
    (0 <= x <= 65536 + g).  Every g; every dividend at and next to a multiple of g, plus random ones."""
    rng = np.random.default_rng(5)
    for g in range(1, 32768):
        m = 0xFFFFFFFF if g == 1 else 0xFFFFFFFF // g + 1
        top = 65536 + g
        k = np.arange(0, top // g + 2, dtype=np.uint64) * np.uint64(g)
        x = np.concatenate([k, k + np.uint64(1), k[1:] - np.uint64(1), rng.integers(0, top + 1, 64).astype(np.uint64)])
        x = x[x <= top]
        q = ((x + np.uint64(1 if g == 1 else 0)) * np.uint64(m)) >> np.uint64(32)
        assert np.array_equal(q, x // np.uint64(g)), g
