"""B200-native batched navmesh queries: a drop-in for habitat-sim's `habitat_sim.nav`
query path (esp::nav::PathFinder over Detour) running as hand-written sm_100a CUDA kernels
behind the C ABI of include/hbn.h.  There is no CPU query path: importing works anywhere,
but creating a PathFinder needs libhbn.so and a CUDA device."""
from . import nav  # noqa: F401
from ._lib import build_library, library_path  # noqa: F401

__all__ = ["nav", "build_library", "library_path"]
