#!/usr/bin/env python3
"""Packed-pair generic sweep vs scalar sweep under a realistic (irregular) policy."""
import os, sys, argparse
from pathlib import Path
os.environ.setdefault("DPB200_XLINE", "off"); os.environ.setdefault("DPB200_PAIR", "off")
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger
logger.remove()
from dynamicprogramming_b200 import envs
ap = argparse.ArgumentParser()
ap.add_argument("--env", default="double_cartpole_swingup"); ap.add_argument("--bins", type=int, default=20)
ap.add_argument("--cfgs", default="256,2;128,4;256,1;512,1;128,3;64,8")
a = ap.parse_args()
eng = envs.make(a.env, bins=a.bins)
eng.build_table()
eng.sweeps(50)
eng.policy_improvement()
eng.sweeps(10)
print(eng.layout())
for c in a.cfgs.split(";"):
    t, m = map(int, c.split(","))
    print(a.env, a.bins, c, eng.debug_pair(t, m, iters=10), flush=True)
eng.close()
