"""Quick device-timed throughput probe (development tool, not the bench contract)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package(); abi = pkg.abi; synth = pkg.synth
if os.environ.get("AIRS_PROBE_LIB"):  # development builds with other launch bounds
    pkg.loader.library_path = lambda: os.path.join(ROOT, "airs-compression_b200", os.environ["AIRS_PROBE_LIB"])

def make_uniform_jobs(n_jobs, n_samples, n_frames, params_fn, cap=None, model=False):
    jobs = np.zeros(n_jobs, dtype=abi.JOB_DTYPE)
    fb = n_samples * 2
    cap = cap or abi.compress_bound(fb)
    slot = (cap + 15) // 16 * 16
    idx = np.arange(n_jobs, dtype=np.uint64)
    jobs["src_offset"] = idx * np.uint64(fb * n_frames)
    jobs["src_frame_stride"] = fb
    jobs["dst_offset"] = idx * np.uint64(slot * n_frames)
    jobs["dst_frame_stride"] = slot
    jobs["work_offset"] = idx * np.uint64(fb) if model else 0
    jobs["identifier_base"] = idx * np.uint64(1000)
    jobs["src_size"] = fb
    jobs["dst_capacity"] = cap
    jobs["work_size"] = fb if model else 0
    jobs["n_frames"] = n_frames
    jobs["dtype"] = abi.DT_U16
    jobs["first_result"] = (idx * np.uint64(n_frames)).astype(np.uint32)
    params_fn(jobs["params"], idx)
    return jobs, slot * n_frames * n_jobs, (fb * n_jobs if model else 0)

def time_batch(name, data, jobs, dst_size, work_size, n_results, steps=5, warmup=2):
    layout = int(os.environ.get("AIRS_PROBE_LAYOUT", "0"))  # 1: CONCAT
    tmp = pkg.batch.concat_tmp_size(jobs, n_results) if layout == 1 and os.environ.get("AIRS_PROBE_TMP") else 0
    db = pkg.batch.DeviceBatch(data.view(torch.uint8).reshape(-1), jobs, dst_size, work_size, n_results, layout=layout,
                               concat_tmp=tmp)
    for _ in range(warmup):
        db.run()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record(); db.run(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    if os.environ.get("AIRS_PROBE_DEBUG"):  # development builds leave counters in the header of the scratch memory
        print("scratch words 40-48:", db.scratch[:512].cpu().numpy().view(np.uint32)[40:49].tolist())
    res = db.results.cpu().numpy().view(np.uint32)
    bad = int((res > 0xFFFFFF80).sum())
    out_bytes = int(res[res <= 0xFFFFFF80].astype(np.int64).sum())
    in_bytes = data.numel() * 2
    med = ms[len(ms) // 2]
    B = in_bytes + out_bytes
    print(f"{name}: {med:.3f} ms (min {ms[0]:.3f}) in {in_bytes/2**30:.2f} GiB ratio {in_bytes/max(out_bytes,1):.2f} "
          f"input {in_bytes/med/1e6:.1f} GB/s  algorithmic {B/med/1e6:.1f} GB/s = {B/med/1e6/6531.6*100:.1f}% of 6531.6  errors {bad}", flush=True)
    if os.environ.get("AIRS_PROBE_DECODE") and layout == 0:  # the same streams back through the decoder
        dj = np.zeros(len(jobs), dtype=abi.DEC_JOB_DTYPE)
        for a_, b_ in (("src_offset", "dst_offset"), ("src_frame_stride", "dst_frame_stride"), ("src_size", "dst_capacity"),
                       ("dst_offset", "src_offset"), ("dst_frame_stride", "src_frame_stride"), ("dst_capacity", "src_size"),
                       ("n_frames", "n_frames"), ("dtype", "dtype"), ("first_result", "first_result")):
            dj[a_] = jobs[b_]
        dec = pkg.batch.DeviceDecodeBatch(db.dst, dj, in_bytes, n_results)
        dec.run(); torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for a, b in evs:
            a.record(); dec.run(); b.record()
        torch.cuda.synchronize()
        dms = sorted(a.elapsed_time(b) for a, b in evs)[1]
        same = bool(torch.equal(dec.dst[:in_bytes], data.view(torch.uint8).reshape(-1)))
        print(f"{name}: decode {dms:.3f} ms, output {in_bytes/dms/1e6:.1f} GB/s, round trip identical: {same}", flush=True)
    return med


def run_case(case, steps=5, warmup=2, dev="cuda"):
    gtab = np.array([1, 2, 4, 7, 16, 60, 255, 1055], dtype=np.uint32)
    otab = np.array([5, 42, 107, 200], dtype=np.uint32)
    def p_plain(p, idx):
        p["primary_preprocessing"] = abi.PRE_DIFF; p["primary_encoder_type"] = 1; p["primary_encoder_param"] = 16
    def p_multi(p, idx):
        p["primary_preprocessing"] = abi.PRE_DIFF; p["primary_encoder_type"] = 2; p["primary_encoder_param"] = 16; p["primary_encoder_outlier"] = 200
    def p_cs(p, idx):
        p_plain(p, idx); p["checksum_enabled"] = 1
    def p_mixed(p, idx):
        h = synth.mix(np.uint64(1) ^ idx)
        p["primary_preprocessing"] = abi.PRE_DIFF
        p["primary_encoder_type"] = 1 + (idx & np.uint64(1)).astype(np.uint32)
        p["primary_encoder_param"] = gtab[(h & np.uint64(7)).astype(np.int64)]
        p["primary_encoder_outlier"] = otab[((h >> np.uint64(3)) & np.uint64(3)).astype(np.int64)]
    def p_model(p, idx):
        p_plain(p, idx)
        p["secondary_iterations"] = 255; p["secondary_preprocessing"] = abi.PRE_MODEL
        p["secondary_encoder_type"] = 1; p["secondary_encoder_param"] = 8; p["model_rate"] = 8
    if case in ("c3", "c3plain", "c3cs", "c3multi"):
        n_chunks, n = 1 << 18, 2048
        data = synth.chunks_torch(1, 0, n_chunks, n, esc=32 if case == "c3" else 0, device=dev)
        pf = {"c3": p_mixed, "c3plain": p_plain, "c3cs": p_cs, "c3multi": p_multi}[case]
        jobs, dsz, wsz = make_uniform_jobs(n_chunks, n, 1, pf)
        return time_batch(case, data, jobs, dsz, wsz, n_chunks, steps, warmup)
    def p_iwt(p, idx):
        p_plain(p, idx); p["primary_preprocessing"] = abi.PRE_IWT
    def p_none(p, idx):
        p_plain(p, idx); p["primary_preprocessing"] = abi.PRE_NONE
    def p_unc(p, idx):
        p_plain(p, idx); p["primary_encoder_type"] = 0
    if case == "sweep":
        # BASELINE config 4 (named subset): preprocessing x encoder x options, 4096 chunks x 128 KiB of u16 each
        n_chunks, n = 4096, 1 << 16
        data = synth.chunks_torch(1, 0, n_chunks, n, device=dev)
        frames2 = synth.frames_torch(1, 0, 512, 8, n, device=dev)
        rows = []
        for name, pre, enc, g, outl, cs, fb, dt in [
                ("NONE+UNCOMPRESSED", 0, 0, 0, 0, 0, 0, 2), ("DIFF+UNCOMPRESSED", 1, 0, 0, 0, 0, 0, 2),
                ("DIFF+ZERO g1", 1, 1, 1, 0, 0, 0, 2), ("DIFF+ZERO g8", 1, 1, 8, 0, 0, 0, 2), ("DIFF+ZERO g16", 1, 1, 16, 0, 0, 0, 2),
                ("DIFF+ZERO g255", 1, 1, 255, 0, 0, 0, 2), ("DIFF+MULTI g16 o200", 1, 2, 16, 200, 0, 0, 2),
                ("DIFF+MULTI g3 o16", 1, 2, 3, 16, 0, 0, 2), ("DIFF+ZERO g16 +checksum", 1, 1, 16, 0, 1, 0, 2),
                ("DIFF+ZERO g16 +fallback", 1, 1, 16, 0, 0, 1, 2), ("DIFF+ZERO g16 i16", 1, 1, 16, 0, 0, 0, 0),
                ("IWT+ZERO g16", 2, 1, 16, 0, 0, 0, 2), ("NONE+ZERO g16", 0, 1, 16, 0, 0, 0, 2)]:
            def pf(p, idx, pre=pre, enc=enc, g=g, outl=outl, cs=cs, fb=fb):
                p["primary_preprocessing"] = pre; p["primary_encoder_type"] = enc; p["primary_encoder_param"] = g
                p["primary_encoder_outlier"] = outl; p["checksum_enabled"] = cs; p["uncompressed_fallback_enabled"] = fb
            jobs, dsz, wsz = make_uniform_jobs(n_chunks, n, 1, pf, model=(pre == 2))
            jobs["dtype"] = dt
            rows.append((name, time_batch(name, data, jobs, dsz, wsz, n_chunks, steps, warmup)))
        for name, rate, g2 in [("DIFF g16 -> MODEL g8 rate 8 (8 frames)", 8, 8), ("DIFF g16 -> MODEL g16 rate 11 (8 frames)", 11, 16)]:
            def pm(p, idx, rate=rate, g2=g2):
                p_plain(p, idx); p["secondary_iterations"] = 255; p["secondary_preprocessing"] = abi.PRE_MODEL
                p["secondary_encoder_type"] = 1; p["secondary_encoder_param"] = g2; p["model_rate"] = rate
            jobs, dsz, wsz = make_uniform_jobs(512, n, 8, pm, model=True)
            rows.append((name, time_batch(name, frames2, jobs, dsz, wsz, 512 * 8, steps, warmup)))
        return rows
    def p_iwtunc(p, idx):
        p_plain(p, idx); p["primary_preprocessing"] = abi.PRE_IWT; p["primary_encoder_type"] = 0
    if case in ("iwt", "none", "unc", "iwtunc", "iwt4", "iwt32k", "none4"):
        n_chunks, n = {"iwt4": (512, 1 << 20), "none4": (512, 1 << 20), "iwt32k": (1 << 14, 1 << 15)}.get(case, (4096, 1 << 16))
        data = synth.chunks_torch(1, 0, n_chunks, n, device=dev)
        jobs, dsz, wsz = make_uniform_jobs(n_chunks, n, 1, {"iwt": p_iwt, "none": p_none, "unc": p_unc, "iwtunc": p_iwtunc, "iwt4": p_iwt, "iwt32k": p_iwt, "none4": p_none}[case],
                                           model=(case in ("iwt", "iwtunc", "iwt4", "iwt32k")))
        return time_batch(case, data, jobs, dsz, wsz, n_chunks, steps, warmup)
    if case in ("c4", "c1", "c4b", "c16k", "c32k", "c8k", "c4b_esc1", "c4b_esc8", "c256", "c4cs", "c4bcs", "c32kcs", "c128", "c64", "c192", "c32"):
        n_chunks, n = {"c128": (128, 1 << 21), "c64": (64, 1 << 21), "c192": (192, 1 << 21), "c32": (32, 1 << 21), "c4": (512, 1 << 20), "c1": (1, 1 << 20), "c4b": (4096, 1 << 17), "c16k": (1 << 15, 1 << 14),
                       "c32k": (1 << 14, 1 << 15), "c8k": (1 << 16, 1 << 13), "c4b_esc1": (4096, 1 << 17),
                       "c4b_esc8": (4096, 1 << 17), "c256": (256, 1 << 21), "c4cs": (512, 1 << 20), "c4bcs": (4096, 1 << 17),
                       "c32kcs": (1 << 14, 1 << 15)}[case]
        data = synth.chunks_torch(1, 0, n_chunks, n, esc={"c4b_esc1": 1, "c4b_esc8": 8}.get(case, 0), device=dev)
        jobs, dsz, wsz = make_uniform_jobs(n_chunks, n, 1, p_cs if case.endswith("cs") else p_plain)
        return time_batch(case, data, jobs, dsz, wsz, n_chunks, steps, warmup)
    def p_model_cs(p, idx):
        p_model(p, idx); p["checksum_enabled"] = 1
    if case in ("c2lit", "c2lit8"):   # config 2 as stated: ONE context (c2lit8: eight), slots of cmp_compress_bound() bytes
        R, F, n = (8 if case == "c2lit8" else 1), 256, 32768
        data = synth.frames_torch(1, 0, R, F, n, device=dev)
        jobs, dsz, wsz = make_uniform_jobs(R, n, F, p_model, model=True)
        return time_batch(case, data, jobs, dsz, wsz, R * F, steps, warmup)
    if case in ("c2", "c2one", "c2cs", "c2small"):   # c2small: slots of 2 n + 64 bytes (frame by frame), else cmp_compress_bound()
        R, F, n = (int(os.environ.get("AIRS_C2_CONTEXTS", "0")) or pkg.load_library().airs_cuda_concurrent_jobs()) if case != "c2one" else 1, 256, 32768
        print("c2 contexts:", R)
        data = synth.frames_torch(1, 0, R, F, n, device=dev)
        jobs, dsz, wsz = make_uniform_jobs(R, n, F, p_model_cs if case == "c2cs" else p_model, cap=(2 * n + 64) if case == "c2small" else None, model=True)
        return time_batch(case, data, jobs, dsz, wsz, R * F, steps, warmup)
    raise SystemExit("unknown case " + case)
