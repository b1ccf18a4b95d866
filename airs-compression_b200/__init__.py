"""B200-native backend of the AIRSPACE compression hot path (Python plumbing only).

The product is libcmp_b200.so (lib/cuda/*.cu + lib/host/*.c): the cmp.h API and the
batched C-ABI of include/airs_cuda.h.  This package only loads it and moves buffers.
"""
from . import abi, synth  # noqa: F401
from .loader import load_library, library_path  # noqa: F401


def __getattr__(name):
    if name in ("batch", "parallel", "workloads"):  # need torch; imported on demand
        import importlib
        return importlib.import_module(__name__ + "." + name)
    raise AttributeError(name)
