import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# Small grids default to the whole-step tile kernel (csrc/vof2d_tile.cuh).  The suite was written against the streaming
# kernels, on small grids for speed: keep it exercising them (tests/test_tile_gpu.py and the tile modes of
# tests/test_reference_pin_gpu.py select the tile kernel explicitly and check the default policy).
os.environ.setdefault("VOF_TILE", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """libvof.so, built in-tree (nvcc cross-compiles here without a GPU)."""
    import taichi_2d_vof_b200 as pkg
    pkg.build()
    return pkg.lib()
