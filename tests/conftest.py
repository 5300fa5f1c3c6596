import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def pkg():
    import pkgload
    return pkgload.load(build_if_missing=True)


@pytest.fixture(scope="session")
def gpu(pkg):
    ok, why = pkg.cuda_available()
    if not ok:
        pytest.fail("GPU test selected but no CUDA device: " + why)
    return pkg
