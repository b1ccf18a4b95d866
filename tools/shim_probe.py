"""Latency of one cmp_compress_u16() call through the host shim (BASELINE config 1: one 1 Mi-sample buffer),
host buffers and device buffers.  Development tool."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package(); abi = pkg.abi; synth = pkg.synth
lib = pkg.load_library()
n = 1 << 20
x = synth.chunks(1, 0, 1, n)[0]
p = abi.params_to_ctypes(abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=1, primary_encoder_param=16))
ctx = abi.CmpContext()
assert lib.cmp_initialise(C.byref(ctx), C.byref(p), None, 0) == 0
cap = lib.cmp_compress_bound(2 * n)
dst = np.zeros((cap + 7) // 8, dtype=np.uint64)
for name, src_ptr, dst_ptr in (("host buffers", x.ctypes.data, dst.ctypes.data),):
    for _ in range(3):
        r = lib.cmp_compress_u16(C.byref(ctx), dst_ptr, cap, src_ptr, 2 * n)
    t0 = time.perf_counter()
    for _ in range(20):
        r = lib.cmp_compress_u16(C.byref(ctx), dst_ptr, cap, src_ptr, 2 * n)
    dt = (time.perf_counter() - t0) / 20
    print(f"cmp_compress_u16, 1 Mi samples, {name}: {dt*1e3:.3f} ms per call = {2*n/dt/1e9:.2f} GB/s, size {r}")
xd = torch.from_numpy(x.view(np.int16)).cuda()
dd = torch.zeros(cap + 16, dtype=torch.uint8, device="cuda")
for _ in range(3):
    r = lib.cmp_compress_u16(C.byref(ctx), dd.data_ptr(), cap, xd.data_ptr(), 2 * n)
t0 = time.perf_counter()
for _ in range(20):
    r = lib.cmp_compress_u16(C.byref(ctx), dd.data_ptr(), cap, xd.data_ptr(), 2 * n)
dt = (time.perf_counter() - t0) / 20
print(f"cmp_compress_u16, 1 Mi samples, device buffers: {dt*1e3:.3f} ms per call = {2*n/dt/1e9:.2f} GB/s, size {r}")
