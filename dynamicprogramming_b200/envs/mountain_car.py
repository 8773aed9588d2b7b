"""MountainCar (discrete pushes) and MountainCarContinuous (21 forces).
Restates runners/mountain_car_cuda.py:27-33,53-76,84-90 and
runners/continuous_mountain_car_cuda.py:27-33,57-85,93-99 of the reference."""
import numpy as np

from ..engine import CudaPIConfig, CudaPolicyIteration2D
from ._common import EnvSpec

_CAR_STEP = r'''
__device__ void step_dynamics(float x, float v, float push,
                              float* x_next, float* v_next, float* reward, bool* terminated)
{
    ENV_CLIP_PUSH
    v += push * ENV_POWER - 0.0025f * cosf(3.0f * x);
    v = fmaxf(-0.07f, fminf(0.07f, v));
    x += v;
    x = fmaxf(-1.2f, fminf(0.6f, x));
    if (x <= -1.2f) v = 0.0f;            // inelastic left wall
    bool at_goal = (x >= ENV_GOAL_X) && (v >= 0.0f);
    *x_next = x;
    *v_next = v;
    *terminated = at_goal;
    *reward = ENV_REWARD;
}
'''


class MountainCarCuda(CudaPolicyIteration2D):
    def _dynamics_cuda_src(self) -> str:
        return (
            "#define ENV_CLIP_PUSH\n#define ENV_POWER 0.001f\n#define ENV_GOAL_X 0.5f\n"
            "#define ENV_REWARD -1.0f\n" + _CAR_STEP
        )

    def _terminal_fn(self, states: np.ndarray):
        return (states[:, 0] >= 0.5) & (states[:, 1] >= 0.0), 0.0


class ContinuousMountainCarCuda(CudaPolicyIteration2D):
    def _dynamics_cuda_src(self) -> str:
        return (
            "#define ENV_CLIP_PUSH push = fmaxf(-1.0f, fminf(1.0f, push));\n#define ENV_POWER 0.0015f\n"
            "#define ENV_GOAL_X 0.45f\n"
            "#define ENV_REWARD (-0.1f * push * push + (at_goal ? 100.0f : 0.0f))\n" + _CAR_STEP
        )

    def _terminal_fn(self, states: np.ndarray):
        return (states[:, 0] >= 0.45) & (states[:, 1] >= 0.0), 0.0


def _cfg():
    return CudaPIConfig(gamma=0.99, theta=1e-4, max_eval_iter=5_000, max_pi_iter=50, log_interval=200)


_BOUNDS = {"position": (-1.2, 0.6), "velocity": (-0.07, 0.07)}

SPEC_DISCRETE = EnvSpec(
    name="mountain_car", cls=MountainCarCuda, bounds=_BOUNDS, default_bins=200,
    actions=np.array([-1.0, 0.0, 1.0], dtype=np.float32), config=_cfg,
    reference="runners/mountain_car_cuda.py:27-33,53-76,84-90",
)
SPEC_CONTINUOUS = EnvSpec(
    name="continuous_mountain_car", cls=ContinuousMountainCarCuda, bounds=_BOUNDS, default_bins=200,
    actions=np.linspace(-1.0, 1.0, 21, dtype=np.float32), config=_cfg,
    reference="runners/continuous_mountain_car_cuda.py:27-33,57-85,93-99",
)
