#!/usr/bin/env python3
"""Tiny driver for ncu: build one environment's table and run a few sweeps (+ one improvement)."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger

logger.remove()
from dynamicprogramming_b200 import envs

ap = argparse.ArgumentParser()
ap.add_argument("--env", default="double_cartpole_swingup")
ap.add_argument("--bins", type=int, default=20)
ap.add_argument("--sweeps", type=int, default=8)
ap.add_argument("--improve", type=int, default=1)
ap.add_argument("--pre-sweeps", type=int, default=0, help="sweeps of the initial policy before the improvement (bench.py uses 50)")
a = ap.parse_args()
eng = envs.make(a.env, bins=a.bins)
eng.build_table()
if a.pre_sweeps:
    eng.sweeps(a.pre_sweeps)
for _ in range(a.improve):
    eng.policy_improvement()
d, ms = eng.sweeps(a.sweeps)
print(eng.layout()); print(eng.eval_kernel_info()); print(f"{a.env}@{a.bins}: {ms / a.sweeps:.4f} ms/sweep, {eng.n_states / (ms / a.sweeps) / 1e6:.2f} G backups/s, delta={d}")
eng.close()
