import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def _have_gpu() -> bool:
    try:
        import bodywork_mlops_demo_b200 as b2
        return b2.native.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    """One device context shared by the GPU tests (fails loudly if the extension is missing)."""
    import bodywork_mlops_demo_b200 as b2
    c = b2.Context(0)
    yield c
    c.close()
