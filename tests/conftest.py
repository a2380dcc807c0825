import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture
def cpu_ops(monkeypatch):
    """Replaces the CUDA operator wrappers with the pure-torch emulation in tests/cpu_shim.py so the
    host-side logic (autograd stitching, module plumbing) can be exercised without a GPU.
    TEST-ONLY: the package itself has no CPU path."""
    import cpu_shim
    cpu_shim.install(monkeypatch)
    return cpu_shim
