import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_win():
    from workloads import synthetic
    return synthetic.small_window()


@pytest.fixture(scope="session")
def small_ragged_win():
    from workloads import synthetic
    return synthetic.small_window(seed=11, ragged=True, n_frames=8, grid=(10, 14))


@pytest.fixture(scope="session")
def cfg3_win():
    """BASELINE cfg2/3: 8 frames x 4000 points x 5x5 at KITTI size (~2.5 s to render)."""
    from workloads import synthetic
    return synthetic.make_window()
