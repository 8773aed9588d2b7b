#!/usr/bin/env python3
"""K2 (BASELINE configs[1]): Continuous Mountain Car --bins 400, fine action grid (201 actions) — the argmax-heavy
improvement step.  Ours vs the reference's own kernels (oracle/_ref) on the same GPU: per-pass improvement time,
per-sweep evaluation time, full run."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger
logger.remove()
import numpy as np, torch
from dynamicprogramming_b200 import envs
from oracle import ref_runner

for env, bins, A in (("continuous_mountain_car", 400, 201), ("continuous_mountain_car", 400, 21), ("pendulum", 200, 21), ("cartpole", 30, 2)):
    spec = envs.REGISTRY[env]
    actions = np.linspace(spec.actions[0], spec.actions[-1], A, dtype=np.float32) if A != len(spec.actions) else spec.actions
    eng = spec.make(bins=bins, actions=actions)
    eng.build_table()
    eng.sweeps(50)
    t = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); eng.policy_improvement(); torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
    st0 = eng.engine_stats()["improve_ms"]; eng.policy_improvement(); imp_dev = eng.engine_stats()["improve_ms"] - st0
    _, ms = eng.sweeps(200)
    eng.close()
    eng = spec.make(bins=bins, actions=actions)
    t0 = time.perf_counter(); eng.run(); run_s = time.perf_counter() - t0
    ours = dict(improve_wall_us=min(t) * 1e6, improve_dev_us=imp_dev * 1e3, sweep_us=ms / 200 * 1e3, run_s=run_s, pi=eng.pi_iterations, sweeps=eng.total_eval_sweeps)
    ref = ref_runner.from_engine_env(env, bins=bins, actions=actions)
    for _ in range(50):
        ref.eval_launch(); ref.d_value_function, ref.d_new_value_function = ref.d_new_value_function, ref.d_value_function
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ref.improve_launch(); torch.cuda.synchronize()
    e0.record(); ref.improve_launch(); e1.record(); torch.cuda.synchronize(); ref_imp = e0.elapsed_time(e1) * 1e3
    e0.record()
    for _ in range(200):
        ref.eval_launch(); ref._max_abs_diff(); ref.d_value_function, ref.d_new_value_function = ref.d_new_value_function, ref.d_value_function
    e1.record(); torch.cuda.synchronize(); ref_sw = e0.elapsed_time(e1) / 200 * 1e3
    ref2 = ref_runner.from_engine_env(env, bins=bins, actions=actions)
    torch.cuda.synchronize(); t0 = time.perf_counter(); ref2.run(); ref_run = time.perf_counter() - t0
    same = bool(np.array_equal(ref2.policy, eng.policy) and np.array_equal(ref2.value_function.view(np.uint32), eng.value_function.view(np.uint32)))
    print(f"{env}@{bins} A={A}: ours {ours} | reference improve {ref_imp:.1f} us, sweep {ref_sw:.2f} us, run {ref_run:.3f} s, pi {ref2.pi_iterations}, sweeps {ref2.total_sweeps} | identical={same}", flush=True)
