#!/usr/bin/env python3
"""ncu target: K5 under the bench policy, a few sweeps of the plane-staged sweep (forced) and of the gather sweep.
    ncu --set full -k regex:ps_sweep -c 1 ... python scripts/prof_plane.py [bins] [cfg] [policy0]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
bins = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cfg = sys.argv[2] if len(sys.argv) > 2 else ""
os.environ["DPB200_PLANE"] = "force" + (":" + cfg if cfg else "")
os.environ.setdefault("DPB200_XLINE", "off")
from dynamicprogramming_b200 import envs
eng = envs.make("double_cartpole_swingup", bins=bins)
eng.build_table()
if len(sys.argv) <= 3:
    eng.sweeps(50)
    eng.policy_improvement()
print(eng.eval_kernel_info())
d, ms = eng.sweeps(25)
print("ms/sweep", ms / 25)
