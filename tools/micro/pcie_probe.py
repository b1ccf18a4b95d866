"""PCIe copy rates of this box from pinned host memory (development tool): H2D alone, D2H alone, both at once."""
import time, torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunks=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = n // chunks
    for i in range(chunks):
        if h2d:
            with torch.cuda.stream(s1): d_in[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
    torch.cuda.synchronize(); return n / (time.perf_counter() - t0) / 1e9
for _ in range(2):
    print("H2D %.1f GB/s  D2H %.1f GB/s  both %.1f GB/s each  both in 16 chunks %.1f" % (run(1, 0), run(0, 1), run(1, 1), run(1, 1, 16)), flush=True)
