"""B200-native implementation of the per-timestep hot path of houkensjtu/taichi-2d-vof.

The compute lives in libvof.so (hand-written sm_100a CUDA kernels behind the C ABI of
include/vof.h); this package is the thin host-side mirror of the reference's interface.
"""
from ._lib import VofError, VofParams, build, lib  # noqa: F401
from .solver2d import Field, VofSolver2D, VofStreamer2D, reference_params, scaled_params  # noqa: F401
from .solver3d import VofSolver3D, reference_params3d, scaled_params3d  # noqa: F401

__all__ = ["VofSolver2D", "VofStreamer2D", "VofParams", "VofError", "Field", "reference_params", "scaled_params", "build", "lib",
           "VofSolver3D", "reference_params3d", "scaled_params3d"]
