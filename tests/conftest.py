import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "treelet-prefetching-for-rt_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CPU-side libraries (oracle restatement, scene writer, reference .so when /root/reference exists)
    once per session; the CUDA library is built by __graft_entry__.build()."""
    import __graft_entry__ as g
    g.build_cpu()
    yield
