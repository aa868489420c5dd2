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


class _PrecompHolder:
    """the production-layout context (use_precomp=True: w = 14 FK20 tables + w = 13 SRS tables, ~144 GiB): created on first
    use, and releasable so that a test which needs the memory for another table layout can have it"""

    def __init__(self, pkg):
        self.pkg = pkg
        self.ctx = None

    def get(self):
        if self.ctx is None:
            self.ctx = self.pkg.DASContext(use_precomp=True)
        return self.ctx

    def release(self):
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None


@pytest.fixture(scope="session")
def precomp_holder(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    h = _PrecompHolder(pkg)
    yield h
    h.release()


@pytest.fixture
def das_ctx_precomp(precomp_holder):
    return precomp_holder.get()


@pytest.fixture(params=["w8", "precomp"])
def vec_ctx(request):
    """every consensus-vector suite runs on both table layouts: the small one (use_precomp=False: w = 8, no merged top
    window) and the production one (use_precomp=True, default windows, merged top window)"""
    if request.param == "w8":
        return request.getfixturevalue("das_ctx")
    return request.getfixturevalue("das_ctx_precomp")
