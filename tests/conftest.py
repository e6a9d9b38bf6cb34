import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def emul_lib():
    import parity_common as pc
    return pc.build_emul()


@pytest.fixture(scope="session")
def have_reference():
    import parity_common as pc
    return pc.reference_available()


@pytest.fixture(scope="session")
def cuda_lib():
    """Path of the CUDA library; GPU tests FAIL (not skip) when it is missing or has no device."""
    import parity_common as pc
    from swmm_b200 import solver
    assert os.path.exists(solver.CUDA_LIB), "libswmm_b200.so missing: run __graft_entry__.build()"
    assert pc.cuda_available(), "no CUDA device visible to libswmm_b200.so"
    return None   # Solver(lib_path=None) loads the CUDA library
