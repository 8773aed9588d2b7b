"""
Test fixture: a runner script written the way the reference's runners/*_cuda.py are written
(same imports, same module-level names, same train()/__main__ shape, and the overhead crane's
trick of storing a goal value through a cupy boolean mask in an `_allocate_tensors_and_compile`
override).  The environment itself — a double integrator that has to park at the origin — is ours.
It is executed UNMODIFIED through dynamicprogramming_b200.compat by tests/test_compat.py.
"""
import sys
from pathlib import Path

import numpy as np
import matplotlib.pyplot as plt

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from src.cuda_policy_iteration import CudaPolicyIteration2D, CudaPIConfig

BINS_PER_DIM = 41

BINS_SPACE = {
    "pos": np.linspace(-2.0, 2.0, BINS_PER_DIM, dtype=np.float32),
    "vel": np.linspace(-3.0, 3.0, BINS_PER_DIM, dtype=np.float32),
}
ACTION_SPACE = np.array([-1.0, -0.25, 0.0, 0.25, 1.0], dtype=np.float32)


class DoubleIntegratorCuda(CudaPolicyIteration2D):
    def _dynamics_cuda_src(self) -> str:
        return r'''
        #define DI_DT 0.05f
        #define DI_XMAX 2.0f
        __device__ void step_dynamics(
            float pos, float vel, float action,
            float* npos, float* nvel, float* reward, bool* terminated
        ) {
            float a = fminf(fmaxf(action, -1.0f), 1.0f);
            *nvel = vel + DI_DT * a;
            *npos = pos + DI_DT * (*nvel);
            *reward = -(pos * pos + 0.1f * vel * vel + 0.01f * a * a);
            *terminated = (*npos <= -DI_XMAX) || (*npos >= DI_XMAX);
        }
        '''

    def _terminal_fn(self, states: np.ndarray):
        pos, vel = states[:, 0], states[:, 1]
        fail = (pos <= -2.0) | (pos >= 2.0)
        goal = (np.abs(pos) <= 0.11) & (np.abs(vel) <= 0.16)
        self._goal_mask = goal
        return (fail | goal), -50.0

    def _allocate_tensors_and_compile(self) -> None:
        import cupy as cp
        super()._allocate_tensors_and_compile()
        if np.any(self._goal_mask):
            d_goal = cp.asarray(self._goal_mask, dtype=cp.bool_)
            self.d_value_function[d_goal] = 7.5
            self.d_new_value_function[d_goal] = 7.5


def train(save_path: Path = Path("results/double_integrator_cuda_policy.npz")):
    config = CudaPIConfig(gamma=0.97, theta=1e-4, max_eval_iter=2000, max_pi_iter=40, log_interval=100)
    pi = DoubleIntegratorCuda(BINS_SPACE, ACTION_SPACE, config)
    pi.run()
    pi.save(save_path)
    return pi


if __name__ == "__main__":
    import argparse

    parser = argparse.ArgumentParser()
    parser.add_argument("--bins", type=int, default=BINS_PER_DIM)
    parser.add_argument("--no-plot", action="store_true")
    parser.add_argument("--retrain", action="store_true")
    parser.add_argument("--save-path", type=Path, default=Path("results/double_integrator_cuda_policy.npz"))
    args = parser.parse_args()

    if args.bins != BINS_PER_DIM:
        for key in BINS_SPACE:
            lo, hi = BINS_SPACE[key][0], BINS_SPACE[key][-1]
            BINS_SPACE[key] = np.linspace(lo, hi, args.bins, dtype=np.float32)
    if args.save_path.exists() and not args.retrain:
        pi = DoubleIntegratorCuda.load(args.save_path)
    else:
        pi = train(args.save_path)
    print(f"states={pi.policy.size} V=[{pi.value_function.min():.3f},{pi.value_function.max():.3f}]")
    if not args.no_plot:
        plt.figure()
