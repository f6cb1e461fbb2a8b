import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """libvof.so, built in-tree (nvcc cross-compiles here without a GPU)."""
    import taichi_2d_vof_b200 as pkg
    pkg.build()
    return pkg.lib()
