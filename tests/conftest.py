import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def ensure_built():
    """Fresh checkouts hold no binaries (they are git-ignored): build the product library, the host library and the
    oracle once before the first test (nvcc cross-compiles sm_100a without a GPU)."""
    need = [os.path.join(ROOT, "evplp_b200", "lib", "libevplp_b200.so"), os.path.join(ROOT, "evplp_b200", "lib", "libevplp_host.so"),
            os.path.join(ROOT, "evplp_b200", "bin", "evplp_render"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__

        __graft_entry__.build()


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_api

    return oracle_api.load()


@pytest.fixture(scope="session")
def lib():
    from evplp_b200 import _capi

    return _capi.load_library()
