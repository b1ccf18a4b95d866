"""The BASELINE.json configurations as synthetic job tables, shared by bench.py, the
full-size parity tests and the development probes.  Data comes from synth.py (integer-only generators, the
same numbers on the CPU and on the GPU); nothing here compresses anything.
"""
import numpy as np

from . import abi, synth

FRAMES, FRAME_SAMPLES = 256, 32768      # config 2: one context = 256 frames of 64 KiB
CHUNK_SAMPLES = 2048                    # config 3: 4 KiB chunks
BIG_CHUNK_SAMPLES = 1 << 20             # configs 1, 4 (default cut), 5: 2 MiB chunks


def uniform_jobs(n_jobs, n, n_frames, cap, first_job=0, model=False, dtype=abi.DT_U16):
    """n_jobs contexts of n_frames frames of n samples each, slots of `cap` bytes (rounded up to 16)."""
    jobs = np.zeros(n_jobs, dtype=abi.JOB_DTYPE)
    fb = (4 if dtype == abi.DT_I16_IN_I32 else 2) * n
    slot = (cap + 15) // 16 * 16
    idx = np.arange(n_jobs, dtype=np.uint64)
    jobs["src_offset"] = idx * np.uint64(fb * n_frames)
    jobs["src_frame_stride"] = fb
    jobs["dst_offset"] = idx * np.uint64(slot * n_frames)
    jobs["dst_frame_stride"] = slot
    jobs["work_offset"] = idx * np.uint64(2 * n) if model else 0
    jobs["identifier_base"] = (idx + np.uint64(first_job)) * np.uint64(1 << 16)
    jobs["src_size"] = fb
    jobs["dst_capacity"] = cap
    jobs["work_size"] = 2 * n if model else 0
    jobs["n_frames"] = n_frames
    jobs["dtype"] = dtype
    jobs["first_result"] = (idx * np.uint64(n_frames)).astype(np.uint32)
    return jobs, int(slot) * n_frames * n_jobs, (2 * n * n_jobs if model else 0)


def _pack(data, jobs, dsz, wsz, n, nf, desc, model=False):
    units = len(jobs)
    return dict(data=data, jobs=jobs, dst_size=dsz, work_size=wsz, n_results=units * nf, desc=desc,
                n_samples_total=units * nf * n, model_bytes=(2 * 2 * n * units if model else 0))


def config2(units, first_unit=0, device=None, rate=8, g2=8, checksum=0, bound_slots=True):
    """Config 2 batched: `units` independent contexts x 256 frames x 64 KiB, DIFF+GOLOMB_ZERO g16 primary pass,
    MODEL+GOLOMB_ZERO secondary passes with model update.  Slots of cmp_compress_bound() bytes - the destination size
    the reference's API names (cmp.h: cmp_compress_bound), in which no frame can fail: what lets the runs of
    secondary passes go through model_run_fast() and a batch of few contexts through the tile kernel - or
    (bound_slots=False) of 2 n + 64 bytes (round 1's layout: frame by frame)."""
    n, nf = FRAME_SAMPLES, FRAMES
    jobs, dsz, wsz = uniform_jobs(units, n, nf, abi.compress_bound(2 * n) if bound_slots else 2 * n + 64, first_unit, model=True)
    jobs["params"] = abi.make_params(primary_preprocessing=abi.PRE_DIFF, primary_encoder_type=abi.ENC_GOLOMB_ZERO,
                                     primary_encoder_param=16, secondary_iterations=255,
                                     secondary_preprocessing=abi.PRE_MODEL, secondary_encoder_type=abi.ENC_GOLOMB_ZERO,
                                     secondary_encoder_param=g2, model_rate=rate, checksum_enabled=checksum)
    if device is None:
        data = np.stack([synth.frames(1, first_unit + c, nf, n) for c in range(units)])
    else:
        data = synth.frames_torch(1, first_unit, units, nf, n, device=device)
    desc = "config2%s: %d context%s x %d frames x 64 KiB u16, DIFF+GOLOMB_ZERO g16 -> MODEL+GOLOMB_ZERO g%d, " \
           "255 secondary iterations, model_rate %d%s, slots of %s bytes" % (
               " batched" if units > 1 else " as stated", units, "s" if units > 1 else "", nf, g2, rate,
               ", checksum" if checksum else "", "cmp_compress_bound()" if bound_slots else "2 n + 64")
    return _pack(data, jobs, dsz, wsz, n, nf, desc, model=True)


G_TABLE = np.array([1, 2, 4, 7, 16, 60, 255, 1055], dtype=np.uint32)
OUTLIER_TABLE = np.array([5, 42, 107, 200], dtype=np.uint32)


def config3(units, first_unit=0, device=None, esc=32):
    """Config 3: `units` independent 4 KiB chunks, DIFF, Golomb parameter per chunk from G_TABLE, encoders
    alternating GOLOMB_ZERO / GOLOMB_MULTI (outliers from OUTLIER_TABLE), escape-heavy data (esc of 256
    samples get 16 random bits XORed in)."""
    n = CHUNK_SAMPLES
    jobs, dsz, wsz = uniform_jobs(units, n, 1, abi.compress_bound(2 * n), first_unit)
    idx = np.arange(first_unit, first_unit + units, dtype=np.uint64)
    h = synth.mix(np.uint64(1) ^ idx)
    P = jobs["params"]
    P["primary_preprocessing"] = abi.PRE_DIFF
    P["primary_encoder_type"] = 1 + (idx & np.uint64(1)).astype(np.uint32)
    P["primary_encoder_param"] = G_TABLE[(h & np.uint64(7)).astype(np.int64)]
    P["primary_encoder_outlier"] = OUTLIER_TABLE[((h >> np.uint64(3)) & np.uint64(3)).astype(np.int64)]
    if device is None:
        data = synth.chunks(1, first_unit, units, n, esc=esc)
    else:
        data = synth.chunks_torch(1, first_unit, units, n, esc=esc, device=device)
    desc = "config3: %d independent 4 KiB chunks, DIFF, mixed Golomb zero/multi parameters, escape heavy" % units
    return _pack(data, jobs, dsz, wsz, n, 1, desc)


# config 4, the named subset: preprocessing x encoder x options, each cut over 1 GiB of samples
#   name, primary preprocessing, encoder, g, outlier, checksum, fallback, container, secondary (pre, g, rate) or None
CONFIG4_ROWS = [
    ("NONE+UNCOMPRESSED", 0, 0, 0, 0, 0, 0, abi.DT_U16, None),
    ("DIFF+UNCOMPRESSED", 1, 0, 0, 0, 0, 0, abi.DT_U16, None),
    ("NONE+ZERO g16", 0, 1, 16, 0, 0, 0, abi.DT_U16, None),
    ("DIFF+ZERO g1", 1, 1, 1, 0, 0, 0, abi.DT_U16, None),
    ("DIFF+ZERO g8", 1, 1, 8, 0, 0, 0, abi.DT_U16, None),
    ("DIFF+ZERO g16", 1, 1, 16, 0, 0, 0, abi.DT_U16, None),
    ("DIFF+ZERO g255", 1, 1, 255, 0, 0, 0, abi.DT_U16, None),
    ("DIFF+MULTI g16 o200", 1, 2, 16, 200, 0, 0, abi.DT_U16, None),
    ("DIFF+MULTI g3 o16", 1, 2, 3, 16, 0, 0, abi.DT_U16, None),
    ("DIFF+ZERO g16 +checksum", 1, 1, 16, 0, 1, 0, abi.DT_U16, None),
    ("DIFF+ZERO g16 +fallback", 1, 1, 16, 0, 0, 1, abi.DT_U16, None),
    ("DIFF+ZERO g16 i16", 1, 1, 16, 0, 0, 0, abi.DT_I16, None),
    ("IWT+ZERO g16", 2, 1, 16, 0, 0, 0, abi.DT_U16, None),
    ("DIFF g16 -> MODEL g8 rate 8 (8 frames)", 1, 1, 16, 0, 0, 0, abi.DT_U16, (3, 8, 8)),
    ("DIFF g16 -> MODEL g16 rate 11 (8 frames)", 1, 1, 16, 0, 0, 0, abi.DT_U16, (3, 16, 11)),
]


def config4_row(row, total_samples=1 << 29, chunk_samples=BIG_CHUNK_SAMPLES, device=None, first_unit=0):
    """One row of CONFIG4_ROWS over total_samples samples (1 GiB of u16 by default), cut into chunks of
    chunk_samples (default 2 MiB: the default cut of SURVEY.md section 8); rows with a secondary pass take
    8 consecutive chunks as the frames of one context."""
    name, pre, enc, g, outl, cs, fb, dt, sec = row
    n = chunk_samples
    nf = 8 if sec else 1
    units = max(1, total_samples // (n * nf))
    needs_work = pre == abi.PRE_IWT or sec is not None
    jobs, dsz, wsz = uniform_jobs(units, n, nf, abi.compress_bound(2 * n), first_unit, model=needs_work, dtype=dt)
    P = jobs["params"]
    P["primary_preprocessing"], P["primary_encoder_type"] = pre, enc
    P["primary_encoder_param"], P["primary_encoder_outlier"] = g, outl
    P["checksum_enabled"], P["uncompressed_fallback_enabled"] = cs, fb
    if sec:
        P["secondary_iterations"], P["secondary_preprocessing"] = 255, sec[0]
        P["secondary_encoder_type"], P["secondary_encoder_param"], P["model_rate"] = 1, sec[1], sec[2]
    if sec:
        data = (synth.frames_torch(1, first_unit, units, nf, n, device=device) if device is not None else
                np.stack([synth.frames(1, first_unit + c, nf, n) for c in range(units)]))
    else:
        data = (synth.chunks_torch(1, first_unit, units, n, device=device) if device is not None else
                synth.chunks(1, first_unit, units, n))
    desc = "config4 %s: %d x %d frames x %d KiB" % (name, units, nf, 2 * n // 1024)
    return _pack(data, jobs, dsz, wsz, n, nf, desc, model=sec is not None)


def config1(device=None):
    """Config 1: ONE chunk of 1 Mi u16 samples (2 MiB), DIFF+GOLOMB_ZERO g16 - what a single cmp_compress_u16() call
    on a large buffer is."""
    n = BIG_CHUNK_SAMPLES
    jobs, dsz, wsz = uniform_jobs(1, n, 1, abi.compress_bound(2 * n))
    P = jobs["params"]
    P["primary_preprocessing"], P["primary_encoder_type"], P["primary_encoder_param"] = abi.PRE_DIFF, 1, 16
    data = synth.chunks_torch(1, 0, 1, n, device=device) if device is not None else synth.chunks(1, 0, 1, n)
    return _pack(data, jobs, dsz, wsz, n, 1, "config1: one chunk of 1 Mi u16 samples, DIFF+GOLOMB_ZERO g16")


def config5(units, first_unit=0, device=None):
    """Config 5: a stream of 2 MiB chunks (config-3 style generator without escapes), DIFF+GOLOMB_ZERO g16;
    the 64 GiB stream is 32768 chunks, `units` of them from chunk `first_unit` on are this rank's shard."""
    n = BIG_CHUNK_SAMPLES
    jobs, dsz, wsz = uniform_jobs(units, n, 1, 2 * n + 64, first_unit)
    P = jobs["params"]
    P["primary_preprocessing"], P["primary_encoder_type"], P["primary_encoder_param"] = abi.PRE_DIFF, 1, 16
    data = (synth.chunks_torch(1, first_unit, units, n, device=device) if device is not None else
            synth.chunks(1, first_unit, units, n))
    desc = "config5: %d chunks x 2 MiB of a 64 GiB stream, DIFF+GOLOMB_ZERO g16" % units
    return _pack(data, jobs, dsz, wsz, n, 1, desc)
