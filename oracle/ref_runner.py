"""
oracle/ref_runner.py — run the REFERENCE's own kernels (oracle/_ref/*.cubin,
built by oracle/build_ref.py from the sources under /root/reference) on a GPU,
following the reference's host loop line by line.

TEST INFRASTRUCTURE / BASELINE ONLY.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py (--impl reference); never by the
product package.  It is the bit-exact oracle for transition rows, value
functions and policies, and the "reference cupy path on one B200" baseline.

cupy is not installed in this image, so the three things cupy does for the
reference are done here with torch + the CUDA driver API through ctypes:
  * device arrays              -> torch tensors (legacy default stream)
  * RawModule / get_function   -> cuModuleLoadData / cuModuleGetFunction
  * ReductionKernel max|x-y|   -> oracle_max_abs_diff, a one-pass fused reduction compiled into the same
                                  cubin (exact: max and abs do not round); torch ops only when an old cubin lacks it
The kernel launches use the reference's launch shape: 256 threads per block,
ceil(N/256) blocks (src/cuda_policy_iteration.py:293-296).
"""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path

import numpy as np

REF_DIR = Path(__file__).resolve().parent / "_ref"

_cuda = None


def _drv():
    global _cuda
    if _cuda is None:
        _cuda = C.CDLL("libcuda.so.1")
        _cuda.cuGetErrorString.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    return _cuda


def _ck(rc: int, what: str) -> None:
    if rc != 0:
        s = C.c_char_p()
        _drv().cuGetErrorString(rc, C.byref(s))
        raise RuntimeError(f"{what} failed: {rc} {s.value.decode() if s.value else ''}")


def available(env: str | None = None) -> bool:
    if not (REF_DIR / "manifest.json").exists():
        return False
    return env is None or (REF_DIR / f"{env}.cubin").exists()


def manifest() -> dict:
    return json.loads((REF_DIR / "manifest.json").read_text())


class RefModule:
    """cupy.RawModule stand-in: one loaded cubin, functions by name."""

    def __init__(self, env: str) -> None:
        import torch

        torch.cuda.init()
        torch.zeros(1, device="cuda")  # make the primary context current
        data = (REF_DIR / f"{env}.cubin").read_bytes()
        self._buf = C.create_string_buffer(data, len(data))
        self.mod = C.c_void_p()
        _ck(_drv().cuModuleLoadData(C.byref(self.mod), self._buf), "cuModuleLoadData")

    def get_function(self, name: str):
        fn = C.c_void_p()
        _ck(_drv().cuModuleGetFunction(C.byref(fn), self.mod, name.encode()), f"cuModuleGetFunction({name})")
        return fn


def launch(fn, n_threads_total: int, args: list, blocks: int | None = None) -> None:
    """fn<<<ceil(n/256), 256>>>(*args) on the legacy default stream (`blocks` overrides the grid)."""
    holders = []
    for a in args:
        if hasattr(a, "data_ptr"):
            holders.append(C.c_void_p(a.data_ptr()))
        elif isinstance(a, (int, np.integer)):
            holders.append(C.c_int(int(a)))
        elif isinstance(a, (float, np.floating)):
            holders.append(C.c_float(float(a)))
        else:
            raise TypeError(type(a))
    arr = (C.c_void_p * len(holders))(*[C.addressof(h) for h in holders])
    if blocks is None:
        blocks = (n_threads_total + 255) // 256
    _ck(_drv().cuLaunchKernel(fn, blocks, 1, 1, 256, 1, 1, 0, None, arr, None), "cuLaunchKernel")


class RefPolicyIteration:
    """The reference's CudaPolicyIteration{2,4,6}D object, driving the reference's
    compiled kernels.  Attribute names follow src/cuda_policy_iteration.py."""

    SYNC_INTERVAL = 25  # :303

    def __init__(self, env: str, axes: list[np.ndarray], actions: np.ndarray, gamma: float, theta: float,
                 max_eval_iter: int, max_pi_iter: int, terminal_mask: np.ndarray | None = None,
                 terminal_value: float = 0.0, extra_values: tuple | None = None) -> None:
        import torch

        self.torch = torch
        meta = manifest()["envs"][env]
        self.D = int(meta["D"])
        assert len(axes) == self.D
        self.gamma, self.theta = float(gamma), float(theta)
        self.max_eval_iter, self.max_pi_iter = int(max_eval_iter), int(max_pi_iter)
        dev = "cuda"

        # __init__ (:81-88): meshgrid(indexing="ij") -> (N, D) float32
        grids = np.meshgrid(*[np.asarray(a, dtype=np.float32) for a in axes], indexing="ij")
        states = np.column_stack([g.ravel() for g in grids]).astype(np.float32)
        self.n_states = len(states)
        # _precompute_grid_metadata (:95-109)
        self.bounds_low = states.min(axis=0).astype(np.float32)
        self.bounds_high = states.max(axis=0).astype(np.float32)
        self.grid_shape = np.array([len(a) for a in axes], dtype=np.int32)
        strides = np.ones(self.D, dtype=np.int64)
        for d in range(self.D - 2, -1, -1):
            strides[d] = strides[d + 1] * self.grid_shape[d + 1]
        self.strides = strides.astype(np.int32)
        self.action_space = np.ascontiguousarray(actions, dtype=np.float32)
        self.n_actions = len(self.action_space)

        # _allocate_tensors_and_compile (:142-161)
        t = torch.from_numpy
        self.d_states = t(states).to(dev)
        self.d_actions = t(self.action_space).to(dev)
        self.d_bounds_low = t(self.bounds_low).to(dev)
        self.d_bounds_high = t(self.bounds_high).to(dev)
        self.d_grid_shape = t(self.grid_shape).to(dev)
        self.d_strides = t(self.strides).to(dev)
        self.d_policy = torch.zeros(self.n_states, dtype=torch.int32, device=dev)
        self.d_value_function = torch.zeros(self.n_states, dtype=torch.float32, device=dev)
        self.d_new_value_function = torch.zeros(self.n_states, dtype=torch.float32, device=dev)
        if terminal_mask is None:
            terminal_mask = np.zeros(self.n_states, dtype=bool)
        self.d_terminal_mask = t(np.ascontiguousarray(terminal_mask, dtype=np.bool_)).to(dev)
        if terminal_mask.any():
            self.d_value_function[self.d_terminal_mask] = float(terminal_value)
        if extra_values is not None:  # crane goal init (runners/overhead_crane_cuda.py:193-206)
            m, v = extra_values
            dm = t(np.ascontiguousarray(m, dtype=np.bool_)).to(dev)
            self.d_value_function[dm] = float(v)
        self.d_new_value_function[:] = self.d_value_function[:]

        self.module = RefModule(env)
        self.eval_kernel = self.module.get_function(meta["eval_kernel"])
        self.improve_kernel = self.module.get_function(meta["improve_kernel"])
        self.probe_kernel = self.module.get_function(meta["probe_kernel"])
        self.mad_kernel = self.module.get_function(meta["max_abs_diff_kernel"]) if "max_abs_diff_kernel" in meta else None
        self.d_delta = torch.zeros(1, dtype=torch.float32, device=dev)
        self._mad_blocks = min((self.n_states + 255) // 256, 8 * torch.cuda.get_device_properties(0).multi_processor_count)
        self.total_sweeps = 0
        self.pi_iterations = 0

    # -- one launch each -----------------------------------------------------
    def eval_launch(self) -> None:
        launch(self.eval_kernel, self.n_states, [
            self.d_states, self.d_actions, self.d_policy, self.d_value_function, self.d_new_value_function,
            self.d_terminal_mask, self.d_bounds_low, self.d_bounds_high, self.d_grid_shape, self.d_strides,
            np.int32(self.n_states), np.float32(self.gamma)])

    def improve_launch(self) -> None:
        launch(self.improve_kernel, self.n_states, [
            self.d_states, self.d_actions, self.d_policy, self.d_value_function, self.d_terminal_mask,
            self.d_bounds_low, self.d_bounds_high, self.d_grid_shape, self.d_strides,
            np.int32(self.n_states), np.int32(self.n_actions), np.float32(self.gamma)])

    def _max_abs_diff(self):
        """max|new_V - V| as ONE fused pass (what cupy's ReductionKernel costs, :164-172)."""
        if self.mad_kernel is None:
            return (self.d_new_value_function - self.d_value_function).abs().max()
        self.d_delta.zero_()
        launch(self.mad_kernel, self.n_states, [self.d_new_value_function, self.d_value_function, np.int32(self.n_states),
                                                self.d_delta], blocks=self._mad_blocks)
        return self.d_delta

    # -- reference host loop (:300-370) ---------------------------------------
    def policy_evaluation(self) -> float:
        delta = float("inf")
        for i in range(self.max_eval_iter):
            self.eval_launch()
            d_delta = self._max_abs_diff()
            self.d_value_function, self.d_new_value_function = self.d_new_value_function, self.d_value_function
            self.total_sweeps += 1
            if i % self.SYNC_INTERVAL == 0 or i == self.max_eval_iter - 1:
                delta = float(d_delta.item())
                if delta < self.theta:
                    self.last_eval_sweeps = i + 1
                    return delta
        self.last_eval_sweeps = self.max_eval_iter
        return delta

    def policy_improvement(self) -> bool:
        old_policy = self.d_policy.clone()
        self.improve_launch()
        return bool(self.torch.all(self.d_policy == old_policy).item())

    def run(self) -> None:
        self.converged = False
        for n in range(self.max_pi_iter):
            self.policy_evaluation()
            self.pi_iterations = n + 1
            if self.policy_improvement():
                self.converged = True
                break
        self.value_function = self.d_value_function.cpu().numpy()
        self.policy = self.d_policy.cpu().numpy()

    # -- probes ---------------------------------------------------------------
    def sweep_once(self) -> np.ndarray:
        """One evaluation sweep from the current V; returns new_V (host) without swapping."""
        self.eval_launch()
        return self.d_new_value_function.cpu().numpy()

    def probe_rows(self, a_idx: int, s_begin: int = 0, count: int | None = None):
        """(idx, w, reward, terminated, next_state) produced by the reference's
        step_dynamics + get_barycentric_Nd for action `a_idx` at states [s_begin, s_begin + count)
        (default: every state)."""
        torch = self.torch
        Cn = 1 << self.D
        n = self.n_states - s_begin if count is None else int(count)
        states = self.d_states[s_begin:s_begin + n]
        idx = torch.empty((n, Cn), dtype=torch.int32, device="cuda")
        w = torch.empty((n, Cn), dtype=torch.float32, device="cuda")
        r = torch.empty(n, dtype=torch.float32, device="cuda")
        tm = torch.empty(n, dtype=torch.uint8, device="cuda")
        nxt = torch.empty((n, self.D), dtype=torch.float32, device="cuda")
        launch(self.probe_kernel, n, [
            states, self.d_actions, np.int32(a_idx), self.d_bounds_low, self.d_bounds_high,
            self.d_grid_shape, self.d_strides, np.int32(n), idx, w, r, tm, nxt])
        torch.cuda.synchronize()
        return idx.cpu().numpy(), w.cpu().numpy(), r.cpu().numpy(), tm.cpu().numpy(), nxt.cpu().numpy()


def from_engine_env(env_name: str, bins: int | None = None, actions=None, config=None) -> RefPolicyIteration:
    """Build the reference object for one of the package's built-in environment
    specs (grid / actions / config / terminal mask come from the spec, the
    kernels from the reference cubin)."""
    from dynamicprogramming_b200 import envs

    spec = envs.REGISTRY[env_name]
    bins_space = spec.bins_space(bins)
    axes = [np.asarray(v, dtype=np.float32) for v in bins_space.values()]
    cfg = config or spec.config()
    acts = spec.actions if actions is None else np.asarray(actions, dtype=np.float32)
    inst = spec.cls.__new__(spec.cls)
    for k, v in spec.kwargs.items():
        setattr(inst, k, v)
    grids = np.meshgrid(*axes, indexing="ij")
    states = np.column_stack([g.ravel() for g in grids]).astype(np.float32)
    mask, value = inst._terminal_fn(states)
    extra = None
    if env_name == "overhead_crane":
        extra = (inst._goal_mask, float(1.0 / (1.0 - cfg.gamma)))
    return RefPolicyIteration(env_name, axes, acts, cfg.gamma, cfg.theta, cfg.max_eval_iter, cfg.max_pi_iter,
                              np.asarray(mask, dtype=bool), float(value), extra)
