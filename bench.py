#!/usr/bin/env python
"""bench.py - throughput of the compression hot path on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
    python bench.py --impl reference ...                     (the reference's CPU library, rank 0 only)
    python bench.py --workload c2|c3|c4|c5 ...               (default c3)

Metric (BASELINE.json): input GB/s compressed, device-timed; one JSON line on rank 0.
A "step" is one pass of the batched driver (airs_cuda_compress_batch) over the whole workload of this
rank, inputs resident in HBM.  Default workload = BASELINE config 3 (the largest single-GPU configuration
BASELINE.json names): 1 Mi independent 4 KiB chunks per GPU, DIFF, mixed Golomb zero / multi parameters,
escape-heavy data.  Every stream of the last timed step is then compared with the CPU reference
(oracle/parity.py: sizes and 64-bit hashes of ALL frames, bytes of every 100th job) - `config.parity`.
At N = 1 the default run also times config 2 batched and the named subset of config 4 (`config.all`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

METRIC = "input GB/s compressed (device-timed)"


def build_workload(pkg, name, units, first_unit, device=None, row=None):
    wl = pkg.workloads
    if name == "c2":
        return wl.config2(units, first_unit, device)
    if name == "c3":
        return wl.config3(units, first_unit, device)
    if name == "c5":
        return wl.config5(units, first_unit, device)
    if name == "c4":
        return wl.config4_row(row, device=device, first_unit=first_unit)
    raise SystemExit("unknown workload " + name)


def default_units(pkg, lib, name, world):
    if name == "c2":
        return lib.airs_cuda_concurrent_jobs()      # one context per resident CTA: all contexts side by side
    if name == "c3":
        return 1 << 20                              # config 3 as stated, per GPU (weak scaling)
    if name == "c5":
        return (1 << 15) // max(world, 2)           # the 64 GiB stream cut over the ranks (N = 1: half of it)
    return 0


# --------------------------------------------------------------------------- clocks

class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < 5.0:   # nvidia-smi needs ~1 s to come up
            time.sleep(0.05)

    def mark(self):
        self.first = len(self.rows)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        time.sleep(0.06)
        rows = self.rows[getattr(self, "first", 0):] or self.rows[-1:]
        self.rows = rows
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


# --------------------------------------------------------------------------- CPU arm

def cpu_reference_run(pkg, workload, steps, warmup, target_s):
    """Times the reference's own CPU library (oracle/_ref, else the oracle port) on a bounded sample of the
    workload with every host core.  Returns (GB/s, info dict, ms per step)."""
    from oracle import parity  # the CPU legs are the one place bench.py executes oracle/
    lib, kind = parity.cpu_library()
    cores = os.cpu_count() or 1
    wl = pkg.workloads
    if workload == "c2":
        w = wl.config2(cores, 0)
    elif workload == "c5":
        w = wl.config5(cores, 0)
    elif workload == "c4":
        w = wl.config4_row(wl.CONFIG4_ROWS[5], total_samples=cores << 20)
    else:
        w = wl.config3(cores * 256, 0)
    src = np.ascontiguousarray(w["data"]).view(np.uint8).reshape(-1)
    dst = np.zeros(w["dst_size"], dtype=np.uint8)
    work = np.zeros(w["work_size"], dtype=np.uint8)

    def once():
        t0 = time.perf_counter()
        lib.run_jobs(src, dst, work, w["jobs"], 0, threads=cores)
        return time.perf_counter() - t0
    t_pilot = once()
    reps = max(1, int(target_s / max(t_pilot, 1e-3)))   # repeat the sample so a step lasts ~target_s
    times = []
    for s in range(warmup + steps):
        t = sum(once() for _ in range(reps))
        if s >= warmup:
            times.append(t)
    in_bytes = src.nbytes * reps
    gbs = in_bytes * len(times) / sum(times) / 1e9
    info = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind,
            "sample": "%d x (%s) per step, %.1f s per step, %d threads" % (reps, w["desc"], sum(times) / len(times), cores),
            "note": "same workload generator, parameters and per-chunk shape as the GPU arm, fewer chunks: a per-byte rate"}
    return gbs, info, sum(times) / len(times) * 1e3


# --------------------------------------------------------------------------- GPU arm

def timed_steps(db, steps, warmup, device, barrier, sampler=None):
    """W untimed steps, then K steps between CUDA events on the launching stream.  Returns (total ms, ms per
    step, kernels launched)."""
    import torch
    for _ in range(warmup):
        db.run()
    barrier()
    if sampler:
        sampler.mark()
    stream = torch.cuda.current_stream(device)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier()
    evs[0].record(stream)
    launches = 0
    for s in range(steps):
        db.run(stream)
        launches += db.launches()
        evs[s + 1].record(stream)
    barrier()
    return evs[0].elapsed_time(evs[-1]), [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)], launches


def run_workload(pkg, w, device, steps, warmup, barrier, peak, sampler=None, parity_opts=None):
    """One workload on this rank: timing, sizes, roofline numbers and (parity_opts) the comparison with the CPU
    reference.  Returns (summary dict, DeviceBatch, device data tensor)."""
    import torch
    abi = pkg.abi
    data = w["data"].view(torch.uint8).reshape(-1)
    db = pkg.batch.DeviceBatch(data, w["jobs"], w["dst_size"], w["work_size"], w["n_results"],
                               layout=abi.LAYOUT_SLOTS, device=device)
    total_ms, step_ms, launches = timed_steps(db, steps, warmup, device, barrier, sampler)
    res = db.results.cpu().numpy().view(np.uint32)
    n_err = int((res > 0xFFFFFF80).sum())
    out_bytes = int(res[res <= 0xFFFFFF80].astype(np.int64).sum())
    in_bytes = w["n_samples_total"] * 2
    alg = in_bytes + out_bytes + w["model_bytes"]
    launch_ms = float(np.mean(step_ms))
    out = {"desc": w["desc"], "in_bytes": in_bytes, "out_bytes": out_bytes, "errors": n_err, "total_ms": total_ms,
           "launch_ms": launch_ms, "launches": launches, "alg_bytes": alg,
           "input_gbs": in_bytes / (launch_ms * 1e-3) / 1e9, "frac": alg / (launch_ms * 1e-3) / 1e9 / peak}
    if parity_opts is not None:
        from oracle import parity  # CPU leg: the checker
        out["parity"] = parity.check_device_batch(db, w["jobs"], data, **parity_opts)
    return out, db, data


def e2e_run(pkg, lib, workload, units, rank, world, device, barrier, e_steps):
    """The metric end to end: airs_cuda_compress_batch_host on pinned HOST buffers (samples and job table in, streams,
    sizes and offsets out), host wall clock around e_steps calls, max over ranks."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    abi = pkg.abi
    e_units = {"c2": max(8, units // 4), "c3": 1 << 18, "c5": 512}[workload]
    e_units = max(1, min(units, e_units))
    we = build_workload(pkg, workload, e_units, rank * units, device=device)
    src_h = torch.empty(we["data"].numel() * 2, dtype=torch.uint8).pin_memory()
    src_h.copy_(we["data"].view(torch.uint8).reshape(-1).cpu())
    # every workload goes through the CONCAT layout: only the compressed bytes travel back
    dst_h = torch.empty(int(we["n_samples_total"] * 2 * 1.1) + (1 << 20), dtype=torch.uint8).pin_memory()

    def pinned(nbytes, dtype):                   # job table and result arrays in pinned memory as well
        return torch.zeros(nbytes, dtype=torch.uint8).pin_memory().numpy().view(dtype)
    jobs_h = pinned(we["jobs"].nbytes, we["jobs"].dtype)
    jobs_h[:] = we["jobs"]
    results_h = pinned(4 * we["n_results"], np.uint32)
    init_h = pinned(4 * e_units, np.uint32)
    offs_h = pinned(8 * (we["n_results"] + 1), np.uint64)
    hb = abi.AirsHostBatch()
    hb.src, hb.src_size = src_h.data_ptr(), src_h.numel()
    hb.dst, hb.dst_size = dst_h.data_ptr(), dst_h.numel()
    hb.work, hb.work_size = None, we["work_size"]
    hb.jobs, hb.results, hb.init_results = jobs_h.ctypes.data, results_h.ctypes.data, init_h.ctypes.data
    hb.out_offsets = offs_h.ctypes.data
    hb.n_jobs, hb.n_results, hb.layout = e_units, we["n_results"], abi.LAYOUT_CONCAT
    for _ in range(2):
        assert lib.airs_cuda_compress_batch_host(C.byref(hb)) == 0, lib.airs_cuda_last_error()
    barrier()
    times = []
    for _ in range(e_steps):
        t0 = time.perf_counter()
        assert lib.airs_cuda_compress_batch_host(C.byref(hb)) == 0, lib.airs_cuda_last_error()
        times.append(time.perf_counter() - t0)
    torch.cuda.synchronize(device)
    td = torch.tensor([sum(times)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
    e_in = we["n_samples_total"] * 2
    out = {"value": e_in * world * e_steps / float(td[0]) / 1e9, "unit": "GB/s",
           "h2d_bytes_per_step": int(src_h.numel() + jobs_h.nbytes),
           "d2h_bytes_per_step": int(offs_h[-1]) + offs_h.nbytes + results_h.nbytes + init_h.nbytes,
           "workload": we["desc"], "api": "airs_cuda_compress_batch_host (pinned host buffers, CONCAT layout)",
           "timing": "host wall clock, max over ranks, %d steps" % e_steps,
           "ms_per_step_min_max": [round(min(times) * 1e3, 2), round(max(times) * 1e3, 2)],
           "ms_steps": [round(t * 1e3, 1) for t in times]}
    lib.airs_cuda_release_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--units", type=int, default=0, help="contexts (c2) or chunks (c3, c5) per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--parity-jobs", type=int, default=0, help="compare only the first N jobs with the CPU reference (0: all)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    pkg = load_package()
    abi = pkg.abi

    if args.impl == "reference":
        if rank != 0:
            return 0
        gbs, info, ms = cpu_reference_run(pkg, args.workload, args.steps, args.warmup, target_s=3.0)
        print(json.dumps({
            "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic", "impl": "reference",
            "config": {"workload": info["sample"], "timing": "host wall clock around the compress calls only"},
            "cpu_baseline": info,
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this backend has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = pkg.load_library()
    wl = pkg.workloads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    cores = max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))))
    parity_opts = None if args.no_parity else {"threads": cores, "sample_every": 100,
                                               "max_jobs": args.parity_jobs or None}

    sampler = ClockSampler(local_rank)
    sampler.start()
    rows = None
    if args.workload == "c4":
        # the named subset of config 4, 1 GiB per row, default cut of 512 x 2 MiB; value = all rows together
        rows, tot_in, tot_out, tot_alg, tot_ms, tot_err, launches = [], 0, 0, 0, 0.0, 0, 0
        par_ok, par_frames = True, 0
        for row in wl.CONFIG4_ROWS:
            w = wl.config4_row(row, device=device, first_unit=rank * 512)
            r, db, data = run_workload(pkg, w, device, args.steps, args.warmup, barrier, peak, sampler, parity_opts)
            rows.append({"row": row[0], "input_gbs": r["input_gbs"], "frac": r["frac"], "ms": r["launch_ms"],
                         "ratio": r["in_bytes"] / max(r["out_bytes"], 1), "errors": r["errors"],
                         "parity": ({k: r["parity"][k] for k in ("frames", "identical", "byte_compared_frames")}
                                    if "parity" in r else None)})
            tot_in += r["in_bytes"]; tot_out += r["out_bytes"]; tot_alg += r["alg_bytes"]
            tot_ms += r["total_ms"]; tot_err += r["errors"]; launches += r["launches"]
            if "parity" in r:
                par_ok &= r["parity"]["identical"]; par_frames += r["parity"]["frames"]
            del db, data, w
            torch.cuda.empty_cache()
        main_r = {"desc": "config4: named subset of %d (preprocessing, encoder, option) rows, 1 GiB each as 512 x 2 MiB chunks "
                          "(model rows: 64 contexts x 8 frames)" % len(rows),
                  "in_bytes": tot_in, "out_bytes": tot_out, "errors": tot_err, "total_ms": tot_ms,
                  "launch_ms": tot_ms / args.steps, "launches": launches, "alg_bytes": tot_alg}
        if parity_opts is not None:
            main_r["parity"] = {"identical": par_ok, "frames": par_frames, "jobs": None}
        db = data = w = None
        units = 0
    else:
        units = args.units or default_units(pkg, lib, args.workload, world)
        first_unit = rank * units                                   # contiguous shards, identifiers by global index
        w = build_workload(pkg, args.workload, units, first_unit, device=device)
        main_r, db, data = run_workload(pkg, w, device, args.steps, args.warmup, barrier, peak, sampler, parity_opts)
    clocks = sampler.stop()

    in_bytes, out_bytes = main_r["in_bytes"], main_r["out_bytes"]
    par = main_r.get("parity")
    t = torch.tensor([main_r["total_ms"], float(in_bytes), float(out_bytes), float(main_r["errors"]),
                      0.0 if (par is None or par["identical"]) else 1.0, float(par["frames"]) if par else 0.0],
                     dtype=torch.float64, device=device)
    total_ms = main_r["total_ms"]
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_ms = float(tmax[0])
    all_in, all_out, all_err, par_bad, par_frames = float(t[1]), float(t[2]), int(t[3]), float(t[4]), int(t[5])
    value = all_in * args.steps / (total_ms * 1e-3) / 1e9

    # ---- N > 1: the epilogue of SURVEY.md 8e, reported apart from the metric - every rank lays (a part of) its
    # shard out as ONE concatenated stream (CONCAT layout, device-wide scan), then sizes and streams are gathered
    # over NCCL so that every rank holds the whole output in rank order; the gathered bytes are hashed stream by
    # stream and compared with the hashes every rank took of its own part
    gather = None
    if world > 1 and not args.no_gather and args.workload != "c4":
        try:
            g_units = {"c2": min(units, 128), "c3": min(units, 1 << 18), "c5": min(units, 4096)}[args.workload]
            wg = build_workload(pkg, args.workload, g_units, rank * units, device=device)
            gdata = wg["data"].view(torch.uint8).reshape(-1)
            tmp = pkg.batch.concat_tmp_size(wg["jobs"], wg["n_results"])
            cb = pkg.batch.DeviceBatch(gdata, wg["jobs"], gdata.numel(), wg["work_size"], wg["n_results"],
                                       layout=abi.LAYOUT_CONCAT, device=device, concat_tmp=tmp)
            cb.run()
            own_hashes = cb.hash_streams()
            torch.cuda.synchronize(device)
            total = int(cb.out_offsets[-1].item())
            # (an untimed first gather: NCCL sets up its point-to-point channels on first use; the buffer is reused)
            _, _ = pkg.parallel.allgather_sizes(cb.results)
            streams_all, _ = pkg.parallel.allgather_streams(cb.dst[:total])
            torch.cuda.synchronize(device)
            barrier()
            stream = torch.cuda.current_stream(device)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            sizes_all, counts = pkg.parallel.allgather_sizes(cb.results)
            streams_all, offs = pkg.parallel.allgather_streams(cb.dst[:total], out=streams_all)
            g1.record(stream)
            barrier()
            tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=device)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            hashes_all, _ = pkg.parallel.allgather_sizes(own_hashes)
            # hash the gathered bytes: one CONCAT "batch" over all ranks' frames
            sz = sizes_all.to(torch.int64)
            sz = torch.where(sz < 0, torch.zeros_like(sz), sz)
            gathered_ok = int(streams_all.numel()) == int(sz.sum().item())
            if gathered_ok:
                starts = torch.cumsum(sz, 0) - sz
                got = pkg.batch.hash_ranges(streams_all, starts.contiguous(), sizes_all.view(torch.int32).contiguous())
                gathered_ok = bool(torch.equal(got, hashes_all))
            gather = {"contexts_or_chunks_per_rank": g_units, "frames": int(sizes_all.numel()),
                      "gathered_bytes": int(streams_all.numel()), "ms": float(tg[0]),
                      "gbs": streams_all.numel() / (float(tg[0]) * 1e-3) / 1e9, "gathered_bytes_match_rank_hashes": gathered_ok,
                      "what": "CONCAT layout per rank, then sizes and streams sent by every rank to every peer into one buffer in rank order "
                              "(grouped ncclSend / ncclRecv), second of two gathers timed, "
                              "max over ranks; every gathered stream re-hashed and compared with its rank's hash"}
            del cb, streams_all
        except Exception as exc:  # the metric line must survive a failing epilogue
            gather = {"error": repr(exc)[:300]}

    # ---- end to end through the C-ABI with HOST buffers (copies inside the timed region)
    e2e = None
    if not args.no_e2e and args.workload != "c4":
        try:
            e2e = e2e_run(pkg, lib, args.workload, units, rank, world, device, barrier, max(3, min(args.steps, 10)))
        except Exception as exc:  # the metric line must survive (host memory, pinning)
            e2e = {"error": repr(exc)[:300]}

    # ---- N = 1, default workload: the other single-GPU configurations beside the headline
    extras = None
    if world == 1 and args.workload == "c3" and not args.no_extras:
        try:
            del db, data, w
            torch.cuda.empty_cache()
            extras = {}
            po = None if args.no_parity else {"threads": cores, "sample_every": 100, "max_jobs": 64}
            w2 = wl.config2(lib.airs_cuda_concurrent_jobs(), 0, device)
            r2, db2, d2 = run_workload(pkg, w2, device, 3, 3, barrier, peak, None, po)
            extras["c2b"] = {"workload": r2["desc"], "input_gbs": r2["input_gbs"], "frac": r2["frac"], "ms": r2["launch_ms"],
                             "parity_first_64_contexts": r2.get("parity", {}).get("identical"),
                             "parity_frames": r2.get("parity", {}).get("frames")}
            del w2, db2, d2
            torch.cuda.empty_cache()
            # configs 1 and 2 AS STATED (one chunk, one context): latency of a single call's worth of work, every stream
            # compared with the CPU reference
            po1 = None if args.no_parity else {"threads": cores, "sample_every": 1, "max_jobs": None}
            for key, w1 in (("c1", wl.config1(device)), ("c2", wl.config2(1, 0, device, bound_slots=True))):
                r1, db1, d1 = run_workload(pkg, w1, device, 10, 3, barrier, peak, None, po1)
                extras[key] = {"workload": w1["desc"], "ms": round(r1["launch_ms"], 4), "input_gbs": round(r1["input_gbs"], 1),
                               "frac": round(r1["frac"], 4), "parity": r1.get("parity", {}).get("identical"),
                               "frames": r1.get("parity", {}).get("frames")}
                del w1, db1, d1
            torch.cuda.empty_cache()
            extras["c4"] = []
            po = None if args.no_parity else {"threads": cores, "sample_every": 100, "max_jobs": None}
            for row in wl.CONFIG4_ROWS:
                w4 = wl.config4_row(row, device=device)
                r4, db4, d4 = run_workload(pkg, w4, device, 3, 3, barrier, peak, None, po)
                extras["c4"].append({"row": row[0], "input_gbs": round(r4["input_gbs"], 1), "frac": round(r4["frac"], 4),
                                     "parity": r4.get("parity", {}).get("identical"), "frames": r4.get("parity", {}).get("frames")})
                del w4, db4, d4
                torch.cuda.empty_cache()
            extras["what"] = "device-timed like the headline (3 steps after 3 warm-up steps each); c4 = named subset of config 4, " \
                             "1 GiB per row as 512 x 2 MiB chunks, every frame compared with the CPU reference"
        except Exception as exc:
            extras = {"error": repr(exc)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    alg_bytes = main_r["alg_bytes"]                                # per launch, this rank
    launch_ms = main_r["launch_ms"]
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        tr = json.load(open(tpath)).get(args.workload)
        if tr and tr.get("units") == units:
            traffic, traffic_src = tr["dram_bytes_per_launch"], tr.get("source")
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.workload == "c5" and world > 1 else "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": main_r["desc"], "per_gpu_input_bytes": in_bytes, "layout": "SLOTS",
                   "l2": "inputs (%.1f GiB per GPU) larger than the 126 MB L2; no flush needed" % (in_bytes / 2**30),
                   "compression_ratio": all_in / max(all_out, 1), "frames_with_errors": all_err},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "traffic_note": "dram__bytes_read + dram__bytes_write of one launch from a stored ncu --set full capture of "
                                     "this workload (profiles/roofline_traffic.json names the commit); not measured in this run",
                     "kernel": "airs_fast_kernel" if args.workload == "c3" else "airs_encode_kernel",
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": launch_ms,
                     "bytes_per_sample": alg_bytes / (in_bytes / 2)},
        "clocks": clocks, "gpu_launches": main_r["launches"],
    }
    if par is not None:
        line["config"]["parity"] = dict(par, frames_all_ranks=par_frames,
                                        what="every stream of the last timed step against the CPU %s: sizes and 64-bit hashes of all "
                                             "frames, bytes of every 100th job" % par.get("cpu", "reference"))
        line["parity_full"] = bool(par_bad == 0.0 and not args.parity_jobs and all_err == 0)
    if rows is not None:
        line["config"]["rows"] = rows
    if e2e:
        line["e2e"] = e2e
    if gather:
        line["config"]["gather"] = gather
    if extras:
        line["config"]["all"] = extras
    if world == 1 and not args.no_cpu:
        gbs, info, _ = cpu_reference_run(pkg, args.workload, 1, 0, target_s=args.cpu_seconds)
        line["cpu_baseline"] = info
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
