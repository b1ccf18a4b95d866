"""The reference's own command line program, unmodified, linked with libcmp_b200.so instead of the
reference library (oracle/Makefile builds both where the reference sources are mounted): same arguments,
same files, byte-identical output.  This is the drop-in claim of INTEGRATION.md section 1 as an executable;
every cmp_compress_* call of the program goes through the host shim to the GPU."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "airspace_ref")
B200 = os.path.join(ROOT, "oracle", "_ref", "airspace_b200")

PARAM_SETS = [
    "",
    "primary_preprocessing=DIFF,primary_encoder_type=GOLOMB_ZERO,primary_encoder_param=16",
    "primary_preprocessing=DIFF,primary_encoder_type=GOLOMB_ZERO,primary_encoder_param=16,secondary_iterations=5,"
    "secondary_preprocessing=MODEL,secondary_encoder_type=GOLOMB_MULTI,secondary_encoder_param=8,"
    "secondary_encoder_outlier=40,model_rate=11,checksum_enabled=true",
    "primary_preprocessing=IWT,primary_encoder_type=GOLOMB_MULTI,primary_encoder_param=5,primary_encoder_outlier=30,"
    "uncompressed_fallback_enabled=true",
]


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    if not (os.path.exists(REF) and os.path.exists(B200)):
        pytest.skip("oracle/_ref/airspace_{ref,b200} not built (reference sources absent at build time)")
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    d = tmp_path_factory.mktemp("cli")
    rng = np.random.default_rng(11)
    out = []
    for k, n in enumerate((5000, 5000, 5000, 5000)):   # big-endian u16 files, as the instrument writes them
        x = (30000 + np.cumsum(rng.integers(-3, 4, size=n)) + rng.integers(-15, 16, size=n)).astype(">u2")
        p = d / f"frame_{k}.bin"
        x.tofile(p)
        out.append(str(p))
    return out


@pytest.mark.parametrize("params", PARAM_SETS)
def test_reference_cli_on_the_gpu_library(files, params):
    args = ["-c", "--stdout"] + (["--params", params] if params else []) + files
    ref = subprocess.run([REF] + args, capture_output=True, timeout=120)
    b200 = subprocess.run([B200] + args, capture_output=True, timeout=300)
    assert ref.returncode == 0, ref.stderr[:500]
    assert b200.returncode == 0, b200.stderr[:500]
    assert len(ref.stdout) > 64
    assert b200.stdout == ref.stdout


def test_reference_cli_writes_air_files(files, tmp_path):
    out_ref, out_b200 = tmp_path / "ref.air", tmp_path / "b200.air"
    p = PARAM_SETS[1]
    assert subprocess.run([REF, "-c", "--params", p, "-o", str(out_ref)] + files[:1], timeout=120).returncode == 0
    assert subprocess.run([B200, "-c", "--params", p, "-o", str(out_b200)] + files[:1], timeout=300).returncode == 0
    assert out_b200.read_bytes() == out_ref.read_bytes()


def test_reference_example_program():
    """examples/simple_compression.c of the reference, compiled against include/cmp.h (our headers) and
    linked with libcmp_b200.so: prints the two streams the reference build prints."""
    ex_ref = os.path.join(ROOT, "oracle", "_ref", "example_ref")
    ex_b200 = os.path.join(ROOT, "oracle", "_ref", "example_b200")
    if not (os.path.exists(ex_ref) and os.path.exists(ex_b200)):
        pytest.skip("oracle/_ref/example_{ref,b200} not built")
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ref = subprocess.run([ex_ref], capture_output=True, timeout=60)
    b200 = subprocess.run([ex_b200], capture_output=True, timeout=300)
    assert ref.returncode == 0 and b200.returncode == 0, (ref.stderr[:300], b200.stderr[:300])
    assert b"82 58 00 00 1A 00 00 06 00 00 00 00 00 04 00 08 00 00 00 01 00 02 55 15 04 C6" in ref.stdout
    assert b200.stdout == ref.stdout
