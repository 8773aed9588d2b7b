"""Base-actuated two-link pendulum swing-up (theta1, w1, theta2, w2).
Restates runners/double_pendulum_swingup_cuda.py:53-66 (grid, torques),
:94-195 (dynamics + shaped reward), :288-294 (config) of the reference."""
import numpy as np

from ..engine import CudaPIConfig, CudaPolicyIteration4D
from ._common import WRAP_SRC, EnvSpec


class DoublePendulumSwingUpCuda(CudaPolicyIteration4D):
    def _dynamics_cuda_src(self) -> str:
        return WRAP_SRC + r'''
#define P2_GRAV 9.8f
#define P2_MA   0.1f
#define P2_MB   0.1f
#define P2_LA   0.5f
#define P2_LB   0.5f
#define P2_TAU  0.02f
#define P2_E_UP ((P2_MA + P2_MB) * P2_GRAV * P2_LA + P2_MB * P2_GRAV * P2_LB)

__device__ void step_dynamics(float a1, float w1, float a2, float w2, float torque,
                              float* a1n, float* w1n, float* a2n, float* w2n,
                              float* reward, bool* terminated)
{
    float msum = P2_MA + P2_MB;
    float diff = a1 - a2;
    float cd = cosf(diff);
    float sd = sinf(diff);

    // mass matrix [[m11, m12], [m12, m22]] and right-hand side
    float m11 = msum * P2_LA * P2_LA;
    float m12 = P2_MB * P2_LA * P2_LB * cd;
    float m22 = P2_MB * P2_LB * P2_LB;
    float r1 = torque
             + msum * P2_GRAV * P2_LA * sinf(a1)
             - P2_MB * P2_LA * P2_LB * sd * w2 * w2;
    float r2 =        P2_MB * P2_GRAV * P2_LB * sinf(a2)
             + P2_MB * P2_LA * P2_LB * sd * w1 * w1;
    float det = m11 * m22 - m12 * m12;
    float acc1 = (m22 * r1 - m12 * r2) / det;
    float acc2 = (m11 * r2 - m12 * r1) / det;

    *a1n = env_wrap_angle(a1 + P2_TAU * w1);
    *w1n = w1 + P2_TAU * acc1;
    *a2n = env_wrap_angle(a2 + P2_TAU * w2);
    *w2n = w2 + P2_TAU * acc2;

    // shaped reward on the successor
    float c1 = cosf(*a1n);
    float c2 = cosf(*a2n);
    float cdn = cosf(*a1n - *a2n);
    float v1 = *w1n;
    float v2 = *w2n;
    float kin = 0.5f * msum  * P2_LA * P2_LA * v1 * v1
              + 0.5f * P2_MB * P2_LB * P2_LB * v2 * v2
              +        P2_MB * P2_LA * P2_LB * v1 * v2 * cdn;
    float pot = msum  * P2_GRAV * P2_LA * c1
              + P2_MB * P2_GRAV * P2_LB * c2;
    float e_gap = (kin + pot) - P2_E_UP;
    float e_err = (e_gap < 0.0f)
                ? 1.5f * (-e_gap) / (2.0f * P2_E_UP)
                :         e_gap   / (2.0f * P2_E_UP);
    float up1 = fmaxf(0.0f, c1);
    float up2 = fmaxf(0.0f, c2);
    float gate = up1 * up2;
    float mis = c1 - c2;
    float mis_pen = 0.5f * mis * mis;
    float spin_pen = 0.1f * gate * (v1 * v1 + v2 * v2);
    float gate_sq = gate * gate;
    float bonus = 4.0f * gate_sq;
    float spin = v1 * v1 + v2 * v2;
    float calm = fmaxf(0.0f, 1.0f - spin / 2.5f);
    float calm_sq = calm * calm;
    float gate_4 = gate_sq * gate_sq;
    float deep = 5.0f * gate_4 * calm_sq;

    *reward = 0.5f
            + 0.5f * (c1 + c2)
            + bonus
            + deep
            - 1.0f * e_err
            - mis_pen
            - spin_pen;
    *terminated = false;
}
'''

    def _terminal_fn(self, states: np.ndarray):
        return np.zeros(len(states), dtype=bool), 0.0


SPEC = EnvSpec(
    name="double_pendulum_swingup", cls=DoublePendulumSwingUpCuda,
    bounds={"theta1": (-np.pi, np.pi), "th1_dot": (-15.0, 15.0), "theta2": (-np.pi, np.pi), "th2_dot": (-15.0, 15.0)},
    default_bins=15,
    actions=np.array([-3.0, -1.5, -0.5, -0.15, -0.05, 0.0, 0.05, 0.15, 0.5, 1.5, 3.0], dtype=np.float32),
    config=lambda: CudaPIConfig(gamma=0.999, theta=1e-4, max_eval_iter=15_000, max_pi_iter=300, log_interval=500),
    reference="runners/double_pendulum_swingup_cuda.py:53-66,94-195,288-294",
)
