"""leven_b200: B200-native (sm_100a) chunk meshing behind the reference's compute interface.

The package holds only what the hot path needs: csrc/ (CUDA kernels + the C ABI,
built into lib/libleven_b200.so) and compute.py (host-side mirror of
leven/src/compute.h).  There is no CPU fallback.
"""
from .compute import *  # noqa: F401,F403
from . import compute  # noqa: F401
