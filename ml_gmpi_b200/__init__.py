"""B200-native multiplane-image renderer: drop-in for the render path of apple/ml-gmpi
(gmpi/core/mpi.py MPI.forward + homography, driven by MPIRenderer.render)."""
from . import _lib  # noqa: F401
from ._build import build_library  # noqa: F401
from .mpi import MPI, MPIOutOfPlaneError, check_range, render_views  # noqa: F401

__all__ = ["MPI", "MPIOutOfPlaneError", "render_views", "check_range", "build_library"]
