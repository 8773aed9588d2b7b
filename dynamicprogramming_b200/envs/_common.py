"""Shared pieces of the built-in environment plugins."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

import numpy as np

from ..engine import CudaPIConfig

# Angle wrap to [-pi, pi) used by every swing-up environment of the reference
# (e.g. runners/pendulum_cuda.py:73-79): fmodf keeps the sign of the dividend,
# so negative remainders are shifted up by one turn.
WRAP_SRC = r'''
#define ENV_PI 3.14159265358979323846f
__device__ float env_wrap_angle(float a) {
    float rem = fmodf(a + ENV_PI, 2.0f * ENV_PI);
    if (rem < 0.0f) rem += 2.0f * ENV_PI;
    return rem - ENV_PI;
}
'''


@dataclass
class EnvSpec:
    """Everything a runner's module-level constants + train() define."""
    name: str
    cls: type
    bounds: dict                 # axis name -> (lo, hi), insertion order = dim order
    default_bins: int
    actions: np.ndarray
    config: Callable[[], CudaPIConfig]
    reference: str               # file:line of the reference definition
    kwargs: dict = field(default_factory=dict)

    def bins_space(self, bins: int | None = None) -> dict:
        """Default grid = the runner's module-level BINS_SPACE
        (np.linspace(lo, hi, BINS_PER_DIM, dtype=float32) from Python floats, e.g.
        runners/pendulum_cuda.py:43-46); another `bins` follows the runner's `--bins`
        path: endpoints are re-read from the default float32 arrays and passed to
        np.linspace again (runners/pendulum_cuda.py:293-296)."""
        default = {k: np.linspace(lo, hi, self.default_bins, dtype=np.float32) for k, (lo, hi) in self.bounds.items()}
        if bins is None or int(bins) == self.default_bins:
            return default
        out = {}
        for k, arr in default.items():
            lo, hi = arr[0], arr[-1]
            out[k] = np.linspace(lo, hi, int(bins), dtype=np.float32)
        return out

    def make(self, bins: int | None = None, actions: np.ndarray | None = None,
             config: CudaPIConfig | None = None, **kw):
        a = self.actions if actions is None else np.asarray(actions, dtype=np.float32)
        return self.cls(self.bins_space(bins), a, config or self.config(), **{**self.kwargs, **kw})
