"""The cmp.h drop-in API of libcmp_b200.so (host shim over the CUDA backend) against the oracle.

Mirrors how the reference's own tests drive the API (test/test_cmp.c, test/test_preprocessing.c):
initialise, compress a few frames, reset, check sizes / bytes / error codes / identifiers.
"""
import ctypes as C

import numpy as np
import pytest

import jobgen

pytestmark = pytest.mark.gpu
abi = jobgen.abi

TS = C.CFUNCTYPE(None, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16))


class Clock:
    def __init__(self, start):
        self.t = start
        self.calls = 0

        def cb(coarse, fine):
            coarse[0] = (self.t >> 16) & 0xFFFFFFFF
            fine[0] = self.t & 0xFFFF
            self.t += 1
            self.calls += 1
        self.cb = TS(cb)


def _oracle_ctx(oracle, params, work, work_size, base):
    class Ctx(C.Structure):
        _fields_ = [("valid", C.c_int), ("params", abi.CmpParams), ("work", C.c_void_p), ("work_size", C.c_uint32),
                    ("model_size", C.c_uint32), ("identifier", C.c_uint64), ("seq", C.c_uint8), ("counter", C.c_uint64)]
    L = oracle.lib
    L.oracle_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]
    L.oracle_init.restype = C.c_uint32
    L.oracle_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
    L.oracle_compress.restype = C.c_uint32
    L.oracle_reset.argtypes = [C.c_void_p]
    L.oracle_reset.restype = C.c_uint32
    ctx = Ctx()
    r = L.oracle_init(C.byref(ctx), C.byref(params), work, work_size, base)
    return ctx, r


@pytest.mark.parametrize("seed", range(8))
def test_shim_sequences(gpu, oracle, pkg, seed):
    lib = pkg.load_library()
    rng = np.random.default_rng(500 + seed)
    clock = Clock(int(rng.integers(0, 1 << 40)))
    lib.cmp_set_timestamp_func(C.cast(clock.cb, C.c_void_p))
    try:
        for case in range(12):
            p = jobgen.random_params(rng, allow_invalid=(case % 4 == 3))
            cp = abi.params_to_ctypes(p)
            n = int(rng.choice([1, 5, 64, 257, 2048, 4099]))
            dtype = int(rng.choice([0, 1, 2]))
            nf = int(rng.integers(1, 6))
            x = jobgen.random_samples(rng, nf, n, str(rng.choice(["walk", "noise", "spiky", "const"])))
            frames = x.astype("<u4") if dtype == 1 else x
            work_g = np.zeros(n + 8, dtype=np.uint16)
            work_o = np.zeros(n + 8, dtype=np.uint16)
            wsz = 2 * n
            ctx = abi.CmpContext()
            base = clock.t
            r_g = lib.cmp_initialise(C.byref(ctx), C.byref(cp), work_g.ctypes.data, wsz)
            octx, r_o = _oracle_ctx(oracle, cp, work_o.ctypes.data, wsz, base)
            assert r_g == r_o, f"initialise {r_g:#x} != {r_o:#x}"
            fn = {0: lib.cmp_compress_i16, 1: lib.cmp_compress_i16_in_i32, 2: lib.cmp_compress_u16}[dtype]
            cap = int(rng.choice([26 + 6 * n, 16 + 2 * n + 4, 40]))
            for f in range(nf):
                if rng.random() < 0.15:
                    assert lib.cmp_reset(C.byref(ctx)) == oracle.lib.oracle_reset(C.byref(octx))
                dst_g = np.full(cap + 16, 0xA5, dtype=np.uint8)
                dst_o = np.full(cap + 16, 0xA5, dtype=np.uint8)
                # 8-byte aligned views
                og = (-dst_g.ctypes.data) % 8
                oo = (-dst_o.ctypes.data) % 8
                src = np.ascontiguousarray(frames[f])
                rg = fn(C.byref(ctx), dst_g.ctypes.data + og, cap, src.ctypes.data, src.nbytes)
                ro = oracle.lib.oracle_compress(C.byref(octx), dst_o.ctypes.data + oo, cap, src.ctypes.data,
                                                src.nbytes, dtype)
                assert rg == ro, f"case {case} frame {f}: {rg:#x} != {ro:#x} params={p}"
                if not abi.is_error(rg):
                    assert np.array_equal(dst_g[og:og + rg], dst_o[oo:oo + ro]), f"case {case} frame {f} bytes"
                    assert np.all(dst_g[og + rg:] == 0xA5)
                if octx.valid:
                    assert ctx.sequence_number == octx.seq and ctx.identifier == octx.identifier
                    assert ctx.model_size == octx.model_size
                    assert clock.t == octx.counter, "number of timestamp calls differs"
                    if p["secondary_preprocessing"] == 3 and p["secondary_iterations"]:
                        assert np.array_equal(work_g[:n], work_o[:n]), "model state in the caller's work buffer"
    finally:
        lib.cmp_set_timestamp_func(None)


def test_shim_error_precedence(gpu, pkg):
    """Argument errors the shim answers before any device work (ref test_cmp.c:160-342)."""
    lib = pkg.load_library()
    p = abi.params_to_ctypes(abi.make_params(primary_encoder_type=1, primary_encoder_param=4))
    ctx = abi.CmpContext()
    assert lib.cmp_initialise(C.byref(ctx), C.byref(p), None, 0) == 0
    dst = np.zeros(64, dtype=np.uint64)
    src = np.arange(8, dtype=np.uint16)
    E = abi.err
    assert lib.cmp_compress_u16(C.byref(ctx), dst.ctypes.data, 64, None, 16) == E("SRC_NULL")
    assert lib.cmp_compress_u16(C.byref(ctx), dst.ctypes.data, 64, src.ctypes.data, 0) == E("SRC_SIZE_WRONG")
    assert lib.cmp_compress_u16(C.byref(ctx), dst.ctypes.data, 64, src.ctypes.data, 3) == E("SRC_SIZE_WRONG")
    assert lib.cmp_compress_i16_in_i32(C.byref(ctx), dst.ctypes.data, 64, src.ctypes.data, 6) == E("SRC_SIZE_WRONG")
    assert lib.cmp_compress_u16(None, dst.ctypes.data, 64, src.ctypes.data, 16) == E("GENERIC")
    assert lib.cmp_compress_u16(C.byref(ctx), None, 64, src.ctypes.data, 16) == E("DST_NULL")
    assert lib.cmp_compress_u16(C.byref(ctx), dst.ctypes.data + 4, 64, src.ctypes.data, 16) == E("DST_UNALIGNED")
    assert lib.cmp_compress_u16(C.byref(ctx), dst.ctypes.data, E("GENERIC"), src.ctypes.data, 16) == E("GENERIC")
    assert lib.cmp_compress_u16(C.byref(ctx), dst.ctypes.data, 10, src.ctypes.data, 16) == E("DST_TOO_SMALL")
    lib.cmp_deinitialise(C.byref(ctx))
    assert bytes(ctx) == bytes(80)
    assert lib.cmp_compress_u16(C.byref(ctx), dst.ctypes.data, 64, src.ctypes.data, 16) == E("CONTEXT_INVALID")
