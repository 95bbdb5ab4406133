import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("oracle", "tools", "realsense-pointcloud_b200"):
    p = os.path.join(ROOT, sub)
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sweep3():
    """3-frame 640x480 synthetic sweep, seed 1 (BASELINE config 1 input)."""
    import gen_scene
    return gen_scene.make_sweep(1, 3)


@pytest.fixture(scope="session")
def pair2():
    """One 640x480 frame pair, seed 2 (BASELINE config 2 input)."""
    import gen_scene
    return gen_scene.make_sweep(2, 2)
