"""Random job generation for the differential tests (oracle vs reference vs CUDA)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
abi = pkg.abi

G_CHOICES = [1, 2, 3, 4, 7, 8, 10, 16, 60, 255, 256, 1055, 4096, 32767, 32768, 65535]
OUTLIER_CHOICES = [1, 5, 16, 24, 42, 107, 200, 1000, 65535, 0xFFFFFFFF]


def align(v, a):
    return (v + a - 1) // a * a


def random_params(rng, allow_invalid=False):
    def enc():
        t = int(rng.integers(0, 3))
        g = int(rng.choice(G_CHOICES)) if rng.random() < 0.8 else int(rng.integers(1, 65536))
        o = int(rng.choice(OUTLIER_CHOICES))
        return t, g, o
    pt, pg, po = enc()
    st, sg, so = enc()
    p = abi.make_params(
        primary_preprocessing=int(rng.choice([0, 1, 2])),
        primary_encoder_type=pt, primary_encoder_param=pg, primary_encoder_outlier=po,
        secondary_iterations=int(rng.choice([0, 0, 1, 2, 3, 255])),
        secondary_preprocessing=int(rng.choice([0, 1, 2, 3, 3, 3])),
        secondary_encoder_type=st, secondary_encoder_param=sg, secondary_encoder_outlier=so,
        model_rate=int(rng.choice([0, 1, 8, 11, 15, 16])),
        checksum_enabled=int(rng.integers(0, 2)) * int(rng.choice([1, 1, 7])),
        uncompressed_fallback_enabled=int(rng.integers(0, 2)))
    if allow_invalid and rng.random() < 0.15:
        which = int(rng.integers(0, 7))
        if which == 0:
            p["primary_encoder_param"] = int(rng.choice([0, 65536, 1 << 20]))
            p["primary_encoder_type"] = int(rng.choice([1, 2]))
        elif which == 1:
            p["primary_preprocessing"] = int(rng.choice([3, 4, 9]))
        elif which == 2:
            p["secondary_iterations"] = int(rng.choice([256, 1000]))
        elif which == 3:
            p["model_rate"] = 17
            p["secondary_preprocessing"] = 3
            p["secondary_iterations"] = 2
        elif which == 4:
            p["primary_encoder_type"] = int(rng.choice([3, 7]))
        elif which == 5:
            p["secondary_encoder_type"] = 2
            p["secondary_encoder_outlier"] = 0
            p["secondary_iterations"] = 1
        else:
            p["secondary_preprocessing"] = 5
            p["secondary_iterations"] = 1
    return p


def random_samples(rng, n_frames, n, style):
    """uint16 [n_frames, n] with a given statistical style."""
    if style == "walk":
        amp = int(rng.choice([1, 3, 10, 40, 300]))
        steps = rng.integers(-amp, amp + 1, size=(n_frames, n))
        base = rng.integers(0, 65536)
        x = base + np.cumsum(steps, axis=1)
        if n_frames > 1:  # frames of one scene: small changes frame to frame
            x = x[0][None, :] + rng.integers(-amp, amp + 1, size=(n_frames, n))
    elif style == "noise":
        x = rng.integers(0, 65536, size=(n_frames, n))
    elif style == "const":
        x = np.full((n_frames, n), int(rng.integers(0, 65536)))
    elif style == "spiky":
        x = 1000 + rng.integers(-5, 6, size=(n_frames, n))
        hit = rng.random(size=(n_frames, n)) < 0.1
        x = np.where(hit, rng.integers(0, 65536, size=(n_frames, n)), x)
    else:  # extremes
        x = rng.choice([0, 1, 0x7FFF, 0x8000, 0xFFFF, 0xFFFE], size=(n_frames, n))
    return (x & 0xFFFF).astype(np.uint16)


def build_jobs(rng, n_jobs, sizes, max_frames=4, allow_invalid=False, capacity_modes=None,
               styles=("walk", "noise", "const", "spiky", "extreme"), dtypes=(0, 1, 2),
               params_fn=None, layout=0):
    """Returns dict(src=u8 array, jobs=JOB_DTYPE array, dst_size, work_size, n_results)."""
    jobs = np.zeros(n_jobs, dtype=abi.JOB_DTYPE)
    src_parts, src_off, dst_off, work_off, res = [], 0, 0, 0, 0
    capacity_modes = capacity_modes or ["bound", "bound", "bound", "raw", "raw-1", "tight", "tiny", "big"]
    for j in range(n_jobs):
        n = int(rng.choice(sizes))
        nf = int(rng.integers(1, max_frames + 1))
        dt = int(rng.choice(dtypes))
        p = params_fn(rng) if params_fn else random_params(rng, allow_invalid)
        x = random_samples(rng, nf, n, str(rng.choice(styles)))
        if dt == abi.DT_I16_IN_I32:
            hi = rng.integers(0, 65536, size=x.shape).astype(np.uint32) << 16
            raw = (x.astype(np.uint32) | hi).astype("<u4")
            stride = 4
        else:
            raw = x.astype("<u2")
            stride = 2
        frame_bytes = n * stride
        fstride = align(frame_bytes, 16) if rng.random() < 0.7 else frame_bytes
        buf = np.zeros(nf * fstride + 16, dtype=np.uint8)
        for f in range(nf):
            buf[f * fstride:f * fstride + frame_bytes] = raw[f].view(np.uint8)
        src_off = align(src_off, 16 if rng.random() < 0.8 else 4)
        pad = src_off - sum(len(b) for b in src_parts)
        if pad:
            src_parts.append(np.zeros(pad, dtype=np.uint8))
        src_parts.append(buf)
        mode = str(rng.choice(capacity_modes))
        bound = 26 + 6 * n
        rawsz = 16 + 2 * n + (4 if p["checksum_enabled"] else 0)
        cap = {"bound": bound, "raw": rawsz, "raw-1": max(rawsz - 1, 0),
               "tight": int(rng.integers(16, bound + 1)), "tiny": int(rng.integers(0, 40)),
               "big": bound + int(rng.integers(0, 100))}[mode]
        slot = align(max(cap, 8), 8)
        need_work = 2 * n
        wmode = rng.random()
        work_size = need_work if wmode < 0.8 else (need_work + 2 if wmode < 0.9 else
                                                    max(need_work - 2, 0) if wmode < 0.95 else 0)
        jobs[j]["src_offset"] = src_off
        jobs[j]["src_frame_stride"] = fstride
        jobs[j]["dst_offset"] = dst_off
        jobs[j]["dst_frame_stride"] = slot
        jobs[j]["work_offset"] = work_off
        jobs[j]["identifier_base"] = int(rng.integers(0, 1 << 40))
        jobs[j]["src_size"] = frame_bytes
        jobs[j]["dst_capacity"] = cap
        jobs[j]["work_size"] = work_size
        jobs[j]["n_frames"] = nf
        jobs[j]["dtype"] = dt
        jobs[j]["first_result"] = res
        jobs[j]["params"] = p
        src_off += len(buf)
        dst_off += slot * nf
        work_off += align(max(work_size, 2), 16)
        res += nf
    src = np.concatenate(src_parts) if src_parts else np.zeros(16, dtype=np.uint8)
    return dict(src=src, jobs=jobs, dst_size=align(dst_off + 64, 64), work_size=align(work_off + 64, 64),
                n_results=res, layout=layout)


def run_cpu(lib, js, threads=1, fill=0xA5):
    """Run a job set through an oracle-like library; returns (dst, results, init, offs, work)."""
    dst = np.full(js["dst_size"], fill, dtype=np.uint8)
    work = np.full(js["work_size"], 0x5A, dtype=np.uint8)
    results, init, offs = lib.run_jobs(js["src"], dst, work, js["jobs"], js["layout"], threads)
    return dst, results, init, offs, work


def compare(a, b, js, what="", check_work=True, check_tail=True):
    """Assert two runs agree on results and on the bytes of every successful stream."""
    dst_a, res_a, init_a, offs_a, work_a = a
    dst_b, res_b, init_b, offs_b, work_b = b
    assert np.array_equal(init_a, init_b), f"{what}: init results differ {init_a} {init_b}"
    if not np.array_equal(res_a, res_b):
        k = int(np.nonzero(res_a != res_b)[0][0])
        raise AssertionError(f"{what}: result[{k}] {res_a[k]:#x} != {res_b[k]:#x}")
    jobs = js["jobs"]
    if js["layout"] == abi.LAYOUT_CONCAT:
        assert np.array_equal(offs_a, offs_b), f"{what}: out_offsets differ"
        total = int(offs_a[-1])
        assert np.array_equal(dst_a[:total], dst_b[:total]), f"{what}: concatenated stream differs"
        return
    for j in range(len(jobs)):
        job = jobs[j]
        for f in range(int(job["n_frames"])):
            r = int(res_a[int(job["first_result"]) + f])
            if abi.is_error(r):
                continue
            o = int(job["dst_offset"]) + f * int(job["dst_frame_stride"])
            sa, sb = dst_a[o:o + r], dst_b[o:o + r]
            if not np.array_equal(sa, sb):
                at = int(np.nonzero(sa != sb)[0][0])
                raise AssertionError(
                    f"{what}: job {j} frame {f} differs at byte {at} of {r}: "
                    f"{sa[at]:#x} != {sb[at]:#x}; params={job['params']} n={job['src_size']} dtype={job['dtype']}")
            # bytes behind the stream inside the slot must be untouched
            cap = int(job["dst_frame_stride"]) if check_tail else 0
            ta, tb = dst_a[o + r:o + cap], dst_b[o + r:o + cap]
            assert np.array_equal(ta, tb), f"{what}: job {j} frame {f}: bytes behind the stream differ"
    if check_work:
        for j in range(len(jobs)):
            job = jobs[j]
            if abi.is_error(int(init_a[j])):
                continue
            p = job["params"]
            if p["secondary_preprocessing"] == 3 and p["secondary_iterations"] != 0:
                o, w = int(job["work_offset"]), int(job["work_size"])
                n2 = min(w, 2 * (int(job["src_size"]) // (4 if job["dtype"] == 1 else 2)))
                if not np.array_equal(work_a[o:o + n2], work_b[o:o + n2]):
                    raise AssertionError(f"{what}: job {j}: model state differs; params={p}")
