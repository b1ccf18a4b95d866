"""Loads libcmp_b200.so and declares its C-ABI (include/airs_cuda.h, include/airs_cuda_decode.h, include/cmp.h)."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EXPORTS = [
    # include/cmp.h
    "cmp_set_timestamp_func", "cmp_is_error", "cmp_compress_bound", "cmp_cal_work_buf_size",
    "cmp_initialise", "cmp_compress_i16", "cmp_compress_i16_in_i32", "cmp_compress_u16",
    "cmp_reset", "cmp_deinitialise",
    # include/cmp_errors.h
    "cmp_get_error_code", "cmp_get_error_message", "cmp_get_error_string",
    # include/airs_cuda.h
    "airs_cuda_device_count", "airs_cuda_concurrent_jobs", "airs_cuda_last_error", "airs_cuda_batch_scratch_size",
    "airs_cuda_concat_tmp_size", "airs_cuda_residual_stats", "airs_cuda_golomb_param_for_mean",
    "airs_cuda_compress_batch", "airs_cuda_last_launch_count", "airs_cuda_compress_batch_host",
    "airs_cuda_release_cache", "airs_cuda_hash_streams", "airs_cuda_hash_ranges",
    "airs_cuda_candidate_bits", "airs_cuda_param_candidates",
    # include/airs_cuda_decode.h
    "airs_cuda_decode_scratch_size", "airs_cuda_decompress_batch",
]


def library_path():
    return os.path.join(_HERE, "libcmp_b200.so")


def load_library():
    """dlopen the backend.  There is no fallback: a missing library is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not built: run `make -C {_HERE}` (or __graft_entry__.build())")
    lib = C.CDLL(path)
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise RuntimeError(f"{path} lacks symbols declared in include/*.h: {missing}")
    u32, vp = C.c_uint32, C.c_void_p
    lib.cmp_is_error.argtypes = [u32]
    lib.cmp_is_error.restype = C.c_uint
    lib.cmp_compress_bound.argtypes = [u32]
    lib.cmp_compress_bound.restype = u32
    lib.cmp_cal_work_buf_size.argtypes = [vp, u32]
    lib.cmp_cal_work_buf_size.restype = u32
    lib.cmp_initialise.argtypes = [vp, vp, vp, u32]
    lib.cmp_initialise.restype = u32
    for name in ("cmp_compress_i16", "cmp_compress_i16_in_i32", "cmp_compress_u16"):
        f = getattr(lib, name)
        f.argtypes = [vp, vp, u32, vp, u32]
        f.restype = u32
    lib.cmp_reset.argtypes = [vp]
    lib.cmp_reset.restype = u32
    lib.cmp_deinitialise.argtypes = [vp]
    lib.cmp_deinitialise.restype = None
    lib.cmp_set_timestamp_func.argtypes = [vp]
    lib.cmp_set_timestamp_func.restype = None
    lib.cmp_get_error_code.argtypes = [u32]
    lib.cmp_get_error_code.restype = C.c_int
    lib.cmp_get_error_message.argtypes = [u32]
    lib.cmp_get_error_message.restype = C.c_char_p
    lib.cmp_get_error_string.argtypes = [C.c_int]
    lib.cmp_get_error_string.restype = C.c_char_p
    lib.airs_cuda_device_count.restype = C.c_int
    lib.airs_cuda_concurrent_jobs.restype = C.c_int
    lib.airs_cuda_last_error.restype = C.c_char_p
    lib.airs_cuda_batch_scratch_size.argtypes = [u32, u32]
    lib.airs_cuda_batch_scratch_size.restype = C.c_size_t
    lib.airs_cuda_concat_tmp_size.argtypes = [C.c_uint64, u32]
    lib.airs_cuda_concat_tmp_size.restype = C.c_size_t
    lib.airs_cuda_residual_stats.argtypes = [vp, vp, u32, vp, vp]
    lib.airs_cuda_residual_stats.restype = C.c_int
    lib.airs_cuda_candidate_bits.argtypes = [vp, vp, u32, vp, u32, vp, vp]
    lib.airs_cuda_candidate_bits.restype = C.c_int
    lib.airs_cuda_param_candidates.argtypes = [vp, u32, vp, u32]
    lib.airs_cuda_param_candidates.restype = u32
    lib.airs_cuda_golomb_param_for_mean.argtypes = [C.c_uint64, u32]
    lib.airs_cuda_golomb_param_for_mean.restype = u32
    lib.airs_cuda_compress_batch.argtypes = [C.POINTER(abi.AirsBatch), vp]
    lib.airs_cuda_compress_batch.restype = C.c_int
    lib.airs_cuda_last_launch_count.restype = C.c_int
    lib.airs_cuda_hash_streams.argtypes = [C.POINTER(abi.AirsBatch), vp, vp]
    lib.airs_cuda_hash_streams.restype = C.c_int
    lib.airs_cuda_hash_ranges.argtypes = [vp, vp, vp, u32, vp, vp]
    lib.airs_cuda_hash_ranges.restype = C.c_int
    lib.airs_cuda_compress_batch_host.argtypes = [C.POINTER(abi.AirsHostBatch)]
    lib.airs_cuda_compress_batch_host.restype = C.c_int
    lib.airs_cuda_release_cache.restype = None
    lib.airs_cuda_decode_scratch_size.argtypes = [u32, u32]
    lib.airs_cuda_decode_scratch_size.restype = C.c_size_t
    lib.airs_cuda_decompress_batch.argtypes = [C.POINTER(abi.AirsDecBatch), vp]
    lib.airs_cuda_decompress_batch.restype = C.c_int
    _LIB = lib
    return lib
