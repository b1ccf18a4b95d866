"""ctypes / numpy view of the C-ABI in include/airs_cuda.h and include/cmp.h.

Pure plumbing: struct layouts, enum values and error helpers.  The values are
interface facts of the reference (lib/cmp.h:64-137, lib/cmp_errors.h:28-60,
lib/common/sample_reader.h:9); nothing here computes anything.
"""
import ctypes as C

import numpy as np

# enum cmp_preprocessing / cmp_encoder_type / sample containers
PRE_NONE, PRE_DIFF, PRE_IWT, PRE_MODEL = 0, 1, 2, 3
ENC_UNCOMPRESSED, ENC_GOLOMB_ZERO, ENC_GOLOMB_MULTI = 0, 1, 2
DT_I16, DT_I16_IN_I32, DT_U16 = 0, 1, 2
BATCH_BIG_ENDIAN = 1                       # airs_batch.flags: the batch may hold DT_BE containers
DT_BE = 4                                  # flag: 16-bit samples big-endian in memory (AIRS_DTYPE_BE)
LAYOUT_SLOTS, LAYOUT_CONCAT = 0, 1

ERRORS = {
    "NO_ERROR": 0, "GENERIC": 1, "PARAMS_INVALID": 10, "DST_TOO_SMALL": 30, "DST_NULL": 31,
    "DST_UNALIGNED": 32, "SRC_SIZE_WRONG": 40, "SRC_NULL": 41, "SRC_SIZE_MISMATCH": 42,
    "WORK_BUF_TOO_SMALL": 50, "WORK_BUF_NULL": 51, "WORK_BUF_UNALIGNED": 52,
    "HDR_CMP_SIZE_TOO_LARGE": 60, "HDR_ORIGINAL_TOO_LARGE": 61, "CONTEXT_INVALID": 70,
    "INT_HDR": 100, "INT_ENCODER": 101, "INT_BITSTREAM": 102, "MAX_CODE": 128,
}


def err(name):
    """(uint32_t)-CMP_ERR_<name>, the way every cmp_* function reports it."""
    return (0x100000000 - ERRORS[name]) & 0xFFFFFFFF


def is_error(code):
    return int(code) > 0x100000000 - 128


CMP_HDR_SIZE, CMP_EXT_HDR_SIZE, CMP_HDR_MAX_SIZE, CMP_CHECKSUM_SIZE = 16, 6, 22, 4
CMP_MAX_SIZE = (1 << 24) - 1

PARAMS_FIELDS = [
    ("primary_preprocessing", "<u4"), ("primary_encoder_type", "<u4"),
    ("primary_encoder_param", "<u4"), ("primary_encoder_outlier", "<u4"),
    ("secondary_iterations", "<u4"), ("secondary_preprocessing", "<u4"),
    ("secondary_encoder_type", "<u4"), ("secondary_encoder_param", "<u4"),
    ("secondary_encoder_outlier", "<u4"), ("model_rate", "<u4"),
    ("checksum_enabled", "u1"), ("uncompressed_fallback_enabled", "u1"), ("_pad", "u1", (2,)),
]
PARAMS_DTYPE = np.dtype(PARAMS_FIELDS)
assert PARAMS_DTYPE.itemsize == 44

JOB_DTYPE = np.dtype([
    ("src_offset", "<u8"), ("src_frame_stride", "<u8"), ("dst_offset", "<u8"),
    ("dst_frame_stride", "<u8"), ("work_offset", "<u8"), ("identifier_base", "<u8"),
    ("src_size", "<u4"), ("dst_capacity", "<u4"), ("work_size", "<u4"), ("n_frames", "<u4"),
    ("dtype", "<u4"), ("first_result", "<u4"), ("params", PARAMS_DTYPE), ("reserved", "<u4"),
])
assert JOB_DTYPE.itemsize == 120


class CmpParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n, *_ in PARAMS_FIELDS[:10]] + [
        ("checksum_enabled", C.c_uint8), ("uncompressed_fallback_enabled", C.c_uint8)]


class CmpContext(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("params", CmpParams), ("work_buf", C.c_void_p),
                ("work_buf_size", C.c_uint32), ("model_size", C.c_uint32),
                ("identifier", C.c_uint64), ("sequence_number", C.c_uint8)]


assert C.sizeof(CmpParams) == 44 and C.sizeof(CmpContext) == 80


class AirsBatch(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("work", C.c_void_p),
                ("jobs", C.c_void_p), ("results", C.c_void_p), ("init_results", C.c_void_p),
                ("out_offsets", C.c_void_p), ("scratch", C.c_void_p), ("dst_size", C.c_uint64),
                ("n_jobs", C.c_uint32), ("n_results", C.c_uint32), ("layout", C.c_uint32),
                ("flags", C.c_uint32), ("tmp", C.c_void_p), ("tmp_size", C.c_uint64)]


class AirsHostBatch(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_size", C.c_uint64), ("dst", C.c_void_p),
                ("dst_size", C.c_uint64), ("work", C.c_void_p), ("work_size", C.c_uint64),
                ("jobs", C.c_void_p), ("results", C.c_void_p), ("init_results", C.c_void_p),
                ("out_offsets", C.c_void_p), ("n_jobs", C.c_uint32), ("n_results", C.c_uint32),
                ("layout", C.c_uint32), ("flags", C.c_uint32)]


# include/airs_cuda_decode.h
DEC_ERRORS = {"CHECKSUM": 110, "NO_MODEL": 111, "CORRUPT": 112}

DEC_JOB_DTYPE = np.dtype([
    ("src_offset", "<u8"), ("src_frame_stride", "<u8"), ("dst_offset", "<u8"),
    ("dst_frame_stride", "<u8"), ("src_size", "<u4"), ("dst_capacity", "<u4"), ("n_frames", "<u4"),
    ("dtype", "<u4"), ("first_result", "<u4"), ("reserved", "<u4"),
])
assert DEC_JOB_DTYPE.itemsize == 56

FRAME_INFO_DTYPE = np.dtype([
    ("identifier", "<u8"), ("compressed_size", "<u4"), ("original_size", "<u4"),
    ("encoder_outlier", "<u4"), ("version", "<u2"), ("encoder_param", "<u2"),
    ("sequence_number", "u1"), ("preprocessing", "u1"), ("checksum_enabled", "u1"),
    ("encoder_type", "u1"), ("model_rate", "u1"), ("header_size", "u1"), ("reserved", "u1", (2,)),
])
assert FRAME_INFO_DTYPE.itemsize == 32


class AirsDecBatch(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("jobs", C.c_void_p),
                ("results", C.c_void_p), ("info", C.c_void_p), ("scratch", C.c_void_p),
                ("n_jobs", C.c_uint32), ("n_results", C.c_uint32)]


STATS_DTYPE = np.dtype([("sum_mapped", "<u8"), ("n_samples", "<u4"), ("max_mapped", "<u4"),
                        ("log2_hist", "<u4", (17,)), ("reserved", "<u4")])
assert STATS_DTYPE.itemsize == 88
CANDIDATE_DTYPE = np.dtype([("encoder_type", "<u4"), ("g", "<u4"), ("outlier", "<u4"), ("reserved", "<u4")])
MAX_CANDIDATES = 32


def dec_err(name):
    return (0x100000000 - DEC_ERRORS[name]) & 0xFFFFFFFF


def make_params(primary_preprocessing=PRE_NONE, primary_encoder_type=ENC_UNCOMPRESSED,
                primary_encoder_param=0, primary_encoder_outlier=0, secondary_iterations=0,
                secondary_preprocessing=PRE_NONE, secondary_encoder_type=ENC_UNCOMPRESSED,
                secondary_encoder_param=0, secondary_encoder_outlier=0, model_rate=0,
                checksum_enabled=0, uncompressed_fallback_enabled=0):
    """A numpy scalar record laid out like struct cmp_params."""
    p = np.zeros((), dtype=PARAMS_DTYPE)
    loc = locals()
    for name in PARAMS_DTYPE.names:
        if name != "_pad":
            p[name] = loc[name]
    return p


def params_to_ctypes(p):
    c = CmpParams()
    for name in PARAMS_DTYPE.names:
        if name != "_pad":
            setattr(c, name, int(p[name]))
    return c


def compress_bound(packed_size):
    """cmp_compress_bound (lib/compress/cmp.c:59-74) for sizing dst slots on the host."""
    if packed_size > CMP_MAX_SIZE:
        return err("HDR_ORIGINAL_TOO_LARGE")
    n = (packed_size * 8 + 15) // 16
    b = CMP_HDR_MAX_SIZE + CMP_CHECKSUM_SIZE + (n * 48 + 7) // 8
    return b if b <= CMP_MAX_SIZE else err("HDR_CMP_SIZE_TOO_LARGE")
