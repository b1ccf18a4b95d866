import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    from __graft_entry__ import load_package
    return load_package()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    return oracle_py.load_oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle_py
    if not oracle_py.have_ref():
        pytest.skip("oracle/_ref/libcmp_ref.so not built (reference sources absent)")
    return oracle_py.load_ref()


@pytest.fixture(scope="session")
def gpu(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pkg.load_library()
    return pkg.batch
