import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-dvbs2rx_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the compiled reference; container only)")


@pytest.fixture(scope="session")
def built():
    """Make sure the product library and the oracle are built (idempotent, seconds)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def ref(built):
    import oracle_lib
    if not os.path.exists(oracle_lib.REF_PATH):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle_lib.Ref()


@pytest.fixture(scope="session")
def gpu(built):
    import dvbs2rx_b200 as d
    if d.device_count() <= 0:
        pytest.skip("no CUDA device")
    return d
