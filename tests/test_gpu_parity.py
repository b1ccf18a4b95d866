"""CUDA path vs oracle, through the C-ABI (airs_cuda_compress_batch), byte for byte."""
import numpy as np
import pytest

import jobgen

pytestmark = pytest.mark.gpu

SMALL = [1, 2, 3, 5, 7, 8, 63, 64, 65, 255, 256, 257]
MEDIUM = [1000, 2047, 2048, 2049, 4099, 6000, 10000]


@pytest.mark.parametrize("seed", range(6))
def test_random_jobs_small(gpu, oracle, seed):
    rng = np.random.default_rng(100 + seed)
    js = jobgen.build_jobs(rng, 300, sizes=SMALL, max_frames=5, allow_invalid=True)
    want = jobgen.run_cpu(oracle, js)
    got = gpu.run_jobs_device(js)
    jobgen.compare(want, got, js, "gpu-vs-oracle")


@pytest.mark.parametrize("seed", range(4))
def test_random_jobs_multi_tile(gpu, oracle, seed):
    rng = np.random.default_rng(200 + seed)
    js = jobgen.build_jobs(rng, 120, sizes=MEDIUM, max_frames=4, allow_invalid=False)
    want = jobgen.run_cpu(oracle, js)
    got = gpu.run_jobs_device(js)
    jobgen.compare(want, got, js, "gpu-vs-oracle")
