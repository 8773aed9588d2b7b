"""Double inverted pendulum on a cart (x, x_dot, theta1, w1, theta2, w2):
the balancing task and the swing-up task share one Lagrangian step.
Restates runners/double_cartpole_cuda.py:53-72,100-182,243-249 and
runners/double_cartpole_swingup_cuda.py:62-74,114-244,349-355 of the reference."""
import numpy as np

from ..engine import CudaPIConfig, CudaPolicyIteration6D
from ._common import WRAP_SRC, EnvSpec

_CONSTS = r'''
#define DC_GRAV   9.8f
#define DC_M_CART 1.0f
#define DC_MA     0.1f
#define DC_MB     0.1f
#define DC_LA     0.5f
#define DC_LB     0.5f
#define DC_TAU    0.02f
#define DC_X_LIMIT 2.4f
'''

# Shared body: symmetric 3x3 mass matrix H = [[a,b,c],[b,d,e],[c,e,f]], generalised
# forces, cofactor inverse.  Defines xacc / a1acc / a2acc.
_ACCEL = r'''
    float msum = DC_MA + DC_MB;
    float c1 = cosf(a1), s1 = sinf(a1);
    float c2 = cosf(a2), s2 = sinf(a2);
    float diff = a1 - a2;
    float cd = cosf(diff), sd = sinf(diff);

    float ha = DC_M_CART + msum;
    float hb = msum  * DC_LA * c1;
    float hc = DC_MB * DC_LB * c2;
    float hd = msum  * DC_LA * DC_LA;
    float he = DC_MB * DC_LA * DC_LB * cd;
    float hf = DC_MB * DC_LB * DC_LB;

    float q1 = push
             + msum  * DC_LA * w1 * w1 * s1
             + DC_MB * DC_LB * w2 * w2 * s2;
    float q2 = msum  * DC_GRAV * DC_LA * s1
             - DC_MB * DC_LA * DC_LB * w2 * w2 * sd;
    float q3 = DC_MB * DC_GRAV * DC_LB * s2
             + DC_MB * DC_LA * DC_LB * w1 * w1 * sd;

    float k11 = hd * hf - he * he;
    float k12 = he * hc - hb * hf;
    float k13 = hb * he - hd * hc;
    float k22 = ha * hf - hc * hc;
    float k23 = hb * hc - ha * he;
    float k33 = ha * hd - hb * hb;
    float det = ha * k11 + hb * k12 + hc * k13;
    float inv = 1.0f / det;

    float xacc  = (k11 * q1 + k12 * q2 + k13 * q3) * inv;
    float a1acc = (k12 * q1 + k22 * q2 + k23 * q3) * inv;
    float a2acc = (k13 * q1 + k23 * q2 + k33 * q3) * inv;
'''

_TH_LIMIT = 20.0 * np.pi / 180.0
_TH_GRID = _TH_LIMIT * 1.15


class DoubleCartPoleCuda(CudaPolicyIteration6D):
    def _dynamics_cuda_src(self) -> str:
        return _CONSTS + r'''
#define DC_TH_LIMIT 0.34906585f
#define DC_X_WEIGHT 0.0f
__device__ void step_dynamics(float x, float xd, float a1, float w1, float a2, float w2, float push,
                              float* xn, float* xdn, float* a1n, float* w1n, float* a2n, float* w2n,
                              float* reward, bool* terminated)
{''' + _ACCEL + r'''
    *xn  = x  + DC_TAU * xd;
    *xdn = xd + DC_TAU * xacc;
    *a1n = a1 + DC_TAU * w1;
    *w1n = w1 + DC_TAU * a1acc;
    *a2n = a2 + DC_TAU * w2;
    *w2n = w2 + DC_TAU * a2acc;

    float xr = *xn / DC_X_LIMIT;
    *reward = 1.0f - DC_X_WEIGHT * xr * xr;
    *terminated = (*xn  < -DC_X_LIMIT)  || (*xn  > DC_X_LIMIT)
               || (*a1n < -DC_TH_LIMIT) || (*a1n > DC_TH_LIMIT)
               || (*a2n < -DC_TH_LIMIT) || (*a2n > DC_TH_LIMIT);
}
'''

    def _terminal_fn(self, states: np.ndarray):
        x, a1, a2 = states[:, 0], states[:, 2], states[:, 4]
        mask = ((x < -2.4) | (x > 2.4) | (a1 < -_TH_LIMIT) | (a1 > _TH_LIMIT)
                | (a2 < -_TH_LIMIT) | (a2 > _TH_LIMIT))
        return mask, 0.0


class DoubleCartPoleSwingUpCuda(CudaPolicyIteration6D):
    def _dynamics_cuda_src(self) -> str:
        return WRAP_SRC + _CONSTS + r'''
#define DC_E_UP ((DC_MA + DC_MB) * DC_GRAV * DC_LA \
               + DC_MB * DC_GRAV * DC_LB)
__device__ void step_dynamics(float x, float xd, float a1, float w1, float a2, float w2, float push,
                              float* xn, float* xdn, float* a1n, float* w1n, float* a2n, float* w2n,
                              float* reward, bool* terminated)
{''' + _ACCEL + r'''
    *xn  = x  + DC_TAU * xd;
    *xdn = xd + DC_TAU * xacc;
    *a1n = env_wrap_angle(a1 + DC_TAU * w1);
    *w1n = w1 + DC_TAU * a1acc;
    *a2n = env_wrap_angle(a2 + DC_TAU * w2);
    *w2n = w2 + DC_TAU * a2acc;

    // cosine + energy shaping on the successor state
    float c1n = cosf(*a1n);
    float c2n = cosf(*a2n);
    float cdn = cosf(*a1n - *a2n);
    float v1 = *w1n;
    float v2 = *w2n;
    float kin = 0.5f * msum  * DC_LA * DC_LA * v1 * v1
              + 0.5f * DC_MB * DC_LB * DC_LB * v2 * v2
              +        DC_MB * DC_LA * DC_LB * v1 * v2 * cdn;
    float pot = msum  * DC_GRAV * DC_LA * c1n
              + DC_MB * DC_GRAV * DC_LB * c2n;
    float e_gap = (kin + pot) - DC_E_UP;
    float e_err = (e_gap < 0.0f)
                ? 2.5f * (-e_gap) / (2.0f * DC_E_UP)
                : 1.5f *   e_gap  / (2.0f * DC_E_UP);
    float up1 = fmaxf(0.0f, c1n);
    float up2 = fmaxf(0.0f, c2n);
    float gate = up1 * up2;
    float spin_pen = 0.1f * gate * (v1 * v1 + v2 * v2);
    float deep = (c1n > 0.7f && c2n > 0.7f) ? 6.0f : 0.0f;
    float xdr = *xdn / 8.0f;
    float xr  = *xn  / DC_X_LIMIT;

    *reward = 0.5f
            + 0.5f * (c1n + c2n)
            + 0.5f * up1
            + 1.0f * up2
            + 6.0f * gate
            + deep
            - 1.0f * e_err
            - 0.5f * xr * xr
            - 0.2f * xdr * xdr
            - spin_pen;
    if ((*xn < -DC_X_LIMIT) || (*xn > DC_X_LIMIT)) {
        *reward -= 100.0f;
    }
    *terminated = (*xn < -DC_X_LIMIT) || (*xn > DC_X_LIMIT);
}
'''

    def _terminal_fn(self, states: np.ndarray):
        x = states[:, 0]
        return (x < -2.4) | (x > 2.4), 0.0


SPEC_BALANCE = EnvSpec(
    name="double_cartpole", cls=DoubleCartPoleCuda,
    bounds={"x": (-2.5, 2.5), "x_dot": (-5.0, 5.0), "theta1": (-_TH_GRID, _TH_GRID), "th1_dot": (-5.0, 5.0),
            "theta2": (-_TH_GRID, _TH_GRID), "th2_dot": (-5.0, 5.0)},
    default_bins=15,
    actions=np.array([-10.0, 0.0, 10.0], dtype=np.float32),
    config=lambda: CudaPIConfig(gamma=0.999, theta=1e-4, max_eval_iter=10_000, max_pi_iter=200, log_interval=500),
    reference="runners/double_cartpole_cuda.py:53-72,100-182,243-249",
)
SPEC_SWINGUP = EnvSpec(
    name="double_cartpole_swingup", cls=DoubleCartPoleSwingUpCuda,
    bounds={"x": (-2.5, 2.5), "x_dot": (-8.0, 8.0), "theta1": (-np.pi, np.pi), "th1_dot": (-15.0, 15.0),
            "theta2": (-np.pi, np.pi), "th2_dot": (-15.0, 15.0)},
    default_bins=20,
    actions=np.array([-60.0, -30.0, -10.0, -3.0, 0.0, 3.0, 10.0, 30.0, 60.0], dtype=np.float32),
    config=lambda: CudaPIConfig(gamma=0.999, theta=1e-4, max_eval_iter=20_000, max_pi_iter=300, log_interval=500),
    reference="runners/double_cartpole_swingup_cuda.py:62-74,114-244,349-355",
)
