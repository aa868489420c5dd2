import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def load_package():
    import __graft_entry__
    return __graft_entry__.load_package()


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def das_ctx(pkg):
    """one context per test session: building the FK20 tables is the expensive part"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ctx = pkg.DASContext(use_precomp=False)
    yield ctx
    ctx.close()
