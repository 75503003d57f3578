import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout)")


def pytest_collection_modifyitems(config, items):
    """Every GPU test gets a time limit (pytest-timeout, when installed): a multi-slab test that ever blocked in
    an exchange must fail, not hang the run.  The full-size property tests are the slowest (numpy over
    33.5 M particles on the host)."""
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            limit = 1500 if "fullsize" in item.nodeid else 600
            item.add_marker(pytest.mark.timeout(limit))


@pytest.fixture(scope="session")
def oracle_lib():
    import pyoracle
    return pyoracle.lib()


@pytest.fixture(scope="session")
def cylgpu_lib():
    """The product library; built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from cylindrical_epoch_b200 import build, _lib
    if os.path.isdir("/usr/local/cuda/bin"):
        build.build()
    return _lib.load()
