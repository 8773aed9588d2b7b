"""Pendulum swing-up on (theta, theta_dot); theta = 0 is upright.
Restates runners/pendulum_cuda.py:44-49 (grid, actions), :81-108 (dynamics),
:119-125 (config) of the reference.  Reward is taken on the pre-step state."""
import numpy as np

from ..engine import CudaPIConfig, CudaPolicyIteration2D
from ._common import WRAP_SRC, EnvSpec


class PendulumCuda(CudaPolicyIteration2D):
    def _dynamics_cuda_src(self) -> str:
        return WRAP_SRC + r'''
#define PN_GRAV   10.0f
#define PN_MASS   1.0f
#define PN_LEN    1.0f
#define PN_STEP   0.05f
#define PN_WMAX   8.0f
#define PN_UMAX   2.0f
__device__ void step_dynamics(float th, float w, float u,
                              float* th_next, float* w_next, float* reward, bool* terminated)
{
    u = fmaxf(-PN_UMAX, fminf(PN_UMAX, u));
    float thn = env_wrap_angle(th);
    *reward = -(thn * thn + 0.1f * w * w + 0.001f * u * u);
    float acc = (3.0f * PN_GRAV / (2.0f * PN_LEN)) * sinf(th) + (3.0f / (PN_MASS * PN_LEN * PN_LEN)) * u;
    float w1 = w + acc * PN_STEP;
    w1 = fmaxf(-PN_WMAX, fminf(PN_WMAX, w1));
    *th_next = env_wrap_angle(th + w1 * PN_STEP);
    *w_next = w1;
    *terminated = false;
}
'''


SPEC = EnvSpec(
    name="pendulum", cls=PendulumCuda,
    bounds={"theta": (-np.pi, np.pi), "theta_dot": (-8.0, 8.0)},
    default_bins=200,
    actions=np.linspace(-2.0, 2.0, 21, dtype=np.float32),
    config=lambda: CudaPIConfig(gamma=0.99, theta=1e-4, max_eval_iter=5_000, max_pi_iter=50, log_interval=200),
    reference="runners/pendulum_cuda.py:44-49,81-108,119-125",
)
