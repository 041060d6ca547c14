import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def refgl():
    """the compiled, unmodified reference (oracle/_ref/librsr_ref.so)"""
    from oracle import refgl as m
    if not m.available():
        m.build()
    if not m.available():
        pytest.skip("reference oracle not built (oracle/_ref/librsr_ref.so)")
    m.init(min(8, os.cpu_count() or 1))
    return m


@pytest.fixture(scope="session")
def cuda_gpu():
    import rsr_b200
    g = rsr_b200.GPU(0)
    yield g
    g.close()


@pytest.fixture(scope="session")
def ref_gpu(refgl):
    g = refgl.RefGPU()
    yield g
    g.close()
