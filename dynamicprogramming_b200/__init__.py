"""dynamicprogramming_b200 — B200-native policy iteration on regular grids.

Drop-in for the hot path of nicoRomeroCuruchet/DynamicProgramming
(src/cuda_policy_iteration.py): same plugin surface, hand-written sm_100a CUDA
behind a C ABI (include/dpb200.h, libdpb200.so).  No CPU fallback.
"""
from .engine import (  # noqa: F401
    CudaPIConfig,
    CudaPolicyIteration2D,
    CudaPolicyIteration4D,
    CudaPolicyIteration6D,
)

__version__ = "0.1.0"
