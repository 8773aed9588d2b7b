#!/usr/bin/env python3
"""Integrated-path check: what the autotune picks and what pi_sweeps then delivers."""
import argparse, sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger
logger.remove(); logger.add(sys.stderr, level="INFO")
from dynamicprogramming_b200 import envs
ap = argparse.ArgumentParser()
ap.add_argument("--env", default="double_cartpole_swingup")
ap.add_argument("--bins", type=int, default=20)
a = ap.parse_args()
eng = envs.make(a.env, bins=a.bins)
eng.build_table()
print(eng.eval_kernel_info(), flush=True)
eng.policy_improvement()
for i in range(3):
    d, ms = eng.sweeps(25)
    print(f"sweeps(25): {ms/25:.4f} ms/sweep  {eng.n_states/(ms/25)/1e6:.1f} G/s delta={d}", flush=True)
info = eng.eval_kernel_info()
if info["xline"]:
    import re
    m = re.search(r"K=(\d+) LV=(\d+) PF=(\d+) warps=(\d+) minb=(\d+) roll=(\d+) tile=([\d,]+)", info["kernel"])
    cfg = ",".join(m.group(i) for i in range(1, 7)) + ":" + m.group(7)
    print(cfg, eng.debug_xline(cfg, iters=10), flush=True)
eng.close()
