"""Built-in environment plugins: the nine environments of the reference's
runners/*_cuda.py restated as subclasses of this package's engine classes
(grid bounds, action sets, CUDA `step_dynamics`, terminal masks, PI configs).
They exist so tests and benchmarks can run where the reference tree is absent;
a reference runner's own subclass plugs into the engine the same way."""
from __future__ import annotations

from . import cartpole, double_cartpole, double_pendulum, mountain_car, overhead_crane, pendulum
from ._common import EnvSpec

REGISTRY: dict[str, EnvSpec] = {
    s.name: s
    for s in (
        pendulum.SPEC,
        mountain_car.SPEC_DISCRETE,
        mountain_car.SPEC_CONTINUOUS,
        cartpole.SPEC_BALANCE,
        cartpole.SPEC_SWINGUP,
        double_pendulum.SPEC,
        overhead_crane.SPEC,
        double_cartpole.SPEC_BALANCE,
        double_cartpole.SPEC_SWINGUP,
    )
}


def make(name: str, bins: int | None = None, actions=None, config=None, **kw):
    """Construct the engine for a built-in environment (allocates on the GPU)."""
    return REGISTRY[name].make(bins=bins, actions=actions, config=config, **kw)


__all__ = ["REGISTRY", "EnvSpec", "make"]
