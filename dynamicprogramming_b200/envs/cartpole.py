"""CartPole balancing and CartPole swing-up (x, x_dot, theta, theta_dot).
Restates runners/cartpole_cuda.py:46-54,80-123,131-137 and
runners/cartpole_swingup_cuda.py:43-52,88-133,177-183 of the reference
(Euler step of the Barto-Sutton-Anderson cart-pole, tau = 0.02)."""
import numpy as np

from ..engine import CudaPIConfig, CudaPolicyIteration4D
from ._common import WRAP_SRC, EnvSpec

_POLE_CONSTS = r'''
#define CPL_GRAV      9.8f
#define CPL_M_CART    1.0f
#define CPL_M_POLE    0.1f
#define CPL_M_TOTAL   1.1f
#define CPL_HALF_LEN  0.5f
#define CPL_ML        0.05f
#define CPL_TAU       0.02f
#define CPL_X_LIMIT   2.4f
'''

_THETA_LIMIT = 12.0 * 2.0 * np.pi / 360.0  # 12 degrees


class CartPoleCuda(CudaPolicyIteration4D):
    def _dynamics_cuda_src(self) -> str:
        return _POLE_CONSTS + r'''
#define CPL_TH_LIMIT  0.20943951f
__device__ void step_dynamics(float x, float xd, float th, float thd, float push,
                              float* x1, float* xd1, float* th1, float* thd1,
                              float* reward, bool* terminated)
{
    float c = cosf(th);
    float s = sinf(th);
    float tmp = (push + CPL_ML * thd * thd * s) / CPL_M_TOTAL;
    float th_acc = (CPL_GRAV * s - c * tmp)
                 / (CPL_HALF_LEN * (4.0f / 3.0f - CPL_M_POLE * c * c / CPL_M_TOTAL));
    float x_acc = tmp - CPL_ML * th_acc * c / CPL_M_TOTAL;

    float xn   = x   + CPL_TAU * xd;
    float xdn  = xd  + CPL_TAU * x_acc;
    float thn  = th  + CPL_TAU * thd;
    float thdn = thd + CPL_TAU * th_acc;
    *x1 = xn; *xd1 = xdn; *th1 = thn; *thd1 = thdn;
    *reward = 1.0f;
    *terminated = (xn < -CPL_X_LIMIT) || (xn > CPL_X_LIMIT)
               || (thn < -CPL_TH_LIMIT) || (thn > CPL_TH_LIMIT);
}
'''

    def _terminal_fn(self, states: np.ndarray):
        x, th = states[:, 0], states[:, 2]
        return (x < -2.4) | (x > 2.4) | (th < -_THETA_LIMIT) | (th > _THETA_LIMIT), 0.0


class CartPoleSwingUpCuda(CudaPolicyIteration4D):
    def _dynamics_cuda_src(self) -> str:
        return WRAP_SRC + _POLE_CONSTS + r'''
#define CPL_E_GOAL (CPL_M_POLE * CPL_GRAV * CPL_HALF_LEN)
__device__ void step_dynamics(float x, float xd, float th, float thd, float push,
                              float* x1, float* xd1, float* th1, float* thd1,
                              float* reward, bool* terminated)
{
    float c = cosf(th);
    float s = sinf(th);
    float tmp = (push + CPL_ML * thd * thd * s) / CPL_M_TOTAL;
    float th_acc = (CPL_GRAV * s - c * tmp)
                 / (CPL_HALF_LEN * (4.0f / 3.0f - CPL_M_POLE * c * c / CPL_M_TOTAL));
    float x_acc = tmp - CPL_ML * th_acc * c / CPL_M_TOTAL;

    *x1   = x   + CPL_TAU * xd;
    *xd1  = xd  + CPL_TAU * x_acc;
    *th1  = env_wrap_angle(th + CPL_TAU * thd);
    *thd1 = thd + CPL_TAU * th_acc;

    // energy-shaped reward on the successor state
    float tip_speed = CPL_HALF_LEN * (*thd1);
    float energy = 0.5f * CPL_M_POLE * tip_speed * tip_speed
                 + CPL_M_POLE * CPL_GRAV * CPL_HALF_LEN * cosf(*th1);
    float e_err = fabsf(energy - CPL_E_GOAL) / (2.0f * CPL_E_GOAL);
    e_err = fminf(e_err, 1.0f);
    float xr = *x1 / CPL_X_LIMIT;
    *reward = cosf(*th1) - 0.5f * e_err - 0.1f * xr * xr;
    *terminated = (*x1 < -CPL_X_LIMIT) || (*x1 > CPL_X_LIMIT);
}
'''

    def _terminal_fn(self, states: np.ndarray):
        x = states[:, 0]
        return (x < -2.4) | (x > 2.4), 0.0


SPEC_BALANCE = EnvSpec(
    name="cartpole", cls=CartPoleCuda,
    bounds={"x": (-2.5, 2.5), "x_dot": (-5.0, 5.0), "theta": (-0.25, 0.25), "theta_dot": (-5.0, 5.0)},
    default_bins=30,
    actions=np.array([-10.0, 10.0], dtype=np.float32),
    config=lambda: CudaPIConfig(gamma=0.99, theta=1e-4, max_eval_iter=10_000, max_pi_iter=100, log_interval=500),
    reference="runners/cartpole_cuda.py:46-54,80-123,131-137",
)
SPEC_SWINGUP = EnvSpec(
    name="cartpole_swingup", cls=CartPoleSwingUpCuda,
    bounds={"x": (-2.5, 2.5), "x_dot": (-5.0, 5.0), "theta": (-np.pi, np.pi), "th_dot": (-10.0, 10.0)},
    default_bins=50,
    actions=np.array([-20.0, -10.0, 0.0, 10.0, 20.0], dtype=np.float32),
    config=lambda: CudaPIConfig(gamma=0.999, theta=1e-4, max_eval_iter=15_000, max_pi_iter=500, log_interval=500),
    reference="runners/cartpole_swingup_cuda.py:43-52,88-133,177-183",
)
