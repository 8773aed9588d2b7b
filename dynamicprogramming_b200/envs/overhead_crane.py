"""Overhead-crane anti-sway positioning (x, x_dot, theta, theta_dot).
Restates runners/overhead_crane_cuda.py:47-64 (grid, forces), :85-157
(dynamics with the target baked into the source), :175-206 (goal states start
at 1/(1-gamma)), :248-262 (config, target_x = -2.5) of the reference."""
import numpy as np

from ..engine import CudaPIConfig, CudaPolicyIteration4D, logger
from ._common import EnvSpec

_RAIL_HALF = 3.0
_SWING_BOUND = (np.pi / 2.0) * 1.1


class OverheadCraneCuda(CudaPolicyIteration4D):
    def __init__(self, bins_space, action_space, config=None, target_x: float = 0.0, **kw):
        self.target_x = float(target_x)
        super().__init__(bins_space, action_space, config, **kw)

    def _dynamics_cuda_src(self) -> str:
        src = r'''
#define CR_GRAV     9.81f
#define CR_TARGET   @TARGET@f
#define CR_M_TROLLEY 1.0f
#define CR_M_LOAD   5.0f
#define CR_ROPE     1.5f
#define CR_TAU      0.02f
#define CR_X_LIMIT  3.0f
#define CR_TH_SCALE 0.52360f
#define CR_THD_SCALE 4.0f
#define CR_XD_SCALE 4.0f
#define CR_TOL_X    0.20f
#define CR_TOL_TH   0.10f
#define CR_TOL_XD   0.20f
__device__ void step_dynamics(float x, float xd, float th, float thd, float push,
                              float* x1, float* xd1, float* th1, float* thd1,
                              float* reward, bool* terminated)
{
    float c = cosf(th);
    float s = sinf(th);
    // H = [[h11, h12], [h12, h22]]
    float h11 = CR_M_TROLLEY + CR_M_LOAD;
    float h12 = CR_M_LOAD * CR_ROPE * c;
    float h22 = CR_M_LOAD * CR_ROPE * CR_ROPE;
    float q1 = push + CR_M_LOAD * CR_ROPE * thd * thd * s;
    float q2 = -CR_M_LOAD * CR_GRAV * CR_ROPE * s;
    float det = h11 * h22 - h12 * h12;
    float inv = 1.0f / det;
    float x_acc  = ( h22 * q1 - h12 * q2) * inv;
    float th_acc = (-h12 * q1 + h11 * q2) * inv;

    *x1   = x   + CR_TAU * xd;
    *xd1  = xd  + CR_TAU * x_acc;
    *th1  = th  + CR_TAU * thd;
    *thd1 = thd + CR_TAU * th_acc;

    float ex   = (*x1 - CR_TARGET) / (2.0f * CR_X_LIMIT);
    float exd  = *xd1  / CR_XD_SCALE;
    float eth  = *th1  / CR_TH_SCALE;
    float ethd = *thd1 / CR_THD_SCALE;
    *reward = 1.0f - 0.15f * ex   * ex
                   - 0.15f * exd  * exd
                   - 0.45f * eth  * eth
                   - 0.25f * ethd * ethd;

    bool off_rail = (*x1 <= -CR_X_LIMIT) || (*x1 >= CR_X_LIMIT);
    bool parked = (fabsf(*x1 - CR_TARGET) <= CR_TOL_X)
               && (fabsf(*th1)            <= CR_TOL_TH)
               && (fabsf(*xd1)            <= CR_TOL_XD);
    *terminated = off_rail || parked;
}
'''
        return src.replace("@TARGET@", f"{self.target_x:.6f}")

    def _terminal_fn(self, states: np.ndarray):
        x, xd, th = states[:, 0], states[:, 1], states[:, 2]
        failed = (x <= -_RAIL_HALF) | (x >= _RAIL_HALF)
        goal = (np.abs(x - self.target_x) <= 0.20) & (np.abs(th) <= 0.10) & (np.abs(xd) <= 0.15)
        self._goal_mask = goal
        return failed | goal, 0.0

    def _allocate_tensors_and_compile(self) -> None:
        super()._allocate_tensors_and_compile()
        if hasattr(self, "_goal_mask") and np.any(self._goal_mask):
            goal_value = float(1.0 / (1.0 - self.config.gamma))
            self.set_values(self._goal_mask, goal_value)
            logger.info(f"Goal states: {int(self._goal_mask.sum()):,} (value={goal_value:.1f} = 1/(1-gamma))")

    def save(self, filepath) -> None:
        from pathlib import Path
        super().save(filepath)
        filepath = Path(filepath).with_suffix(".npz")
        data = dict(np.load(filepath))
        data["target_x"] = np.float32(self.target_x)
        np.savez(filepath, **data)

    @classmethod
    def load(cls, filepath):
        from pathlib import Path
        instance = super().load(filepath)
        data = np.load(Path(filepath).with_suffix(".npz"))
        instance.target_x = float(data["target_x"]) if "target_x" in data else 0.0
        return instance


SPEC = EnvSpec(
    name="overhead_crane", cls=OverheadCraneCuda,
    bounds={"x": (-_RAIL_HALF, _RAIL_HALF), "x_dot": (-4.0, 4.0), "theta": (-_SWING_BOUND, _SWING_BOUND),
            "theta_dot": (-4.0, 4.0)},
    default_bins=30,
    actions=np.array([-30.0, -20.0, -10.0, 0.0, 10.0, 20.0, 30.0], dtype=np.float32),
    config=lambda: CudaPIConfig(gamma=0.999, theta=1e-4, max_eval_iter=10_000, max_pi_iter=100, log_interval=500),
    reference="runners/overhead_crane_cuda.py:47-64,85-206,248-262",
    kwargs={"target_x": -2.5},
)
