#!/usr/bin/env python3
"""Where does the wall time of a small-grid run() go?  (K1 pendulum 200^2, K2, K3)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger
logger.remove()
from dynamicprogramming_b200 import envs, _ffi
import numpy as np
for env, bins in (("pendulum", 200), ("pendulum", 200), ("continuous_mountain_car", 400), ("cartpole", 30)):
    T = {}
    t = time.perf_counter(); eng = envs.make(env, bins=bins); T["create"] = time.perf_counter() - t
    t = time.perf_counter(); eng.build_table(); T["build_table+graphs"] = time.perf_counter() - t
    te = ti = 0.0; sweeps = 0
    for n in range(eng.config.max_pi_iter):
        t = time.perf_counter(); eng.policy_evaluation(); te += time.perf_counter() - t; sweeps += eng.last_eval_sweeps
        t = time.perf_counter(); st = eng.policy_improvement(); ti += time.perf_counter() - t
        if st: break
    T["eval_wall"] = te; T["improve_wall"] = ti
    st = eng.engine_stats()
    t = time.perf_counter(); eng._pull_tensors_from_gpu(); T["pull+destroy"] = time.perf_counter() - t
    print(env, bins, "PI", n + 1, "sweeps", sweeps, {k: round(v * 1e3, 2) for k, v in T.items()}, "gpu eval_ms", round(st["eval_ms"], 2), "improve_ms", round(st["improve_ms"], 2), flush=True)
