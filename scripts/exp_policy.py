#!/usr/bin/env python3
"""How regular are real policies along the fast dimension?  Window-path fraction and sweep time of the
x-line sweep for the greedy policy after N evaluation sweeps of the initial policy (K5)."""
import os, sys
from pathlib import Path
os.environ.setdefault("DPB200_XLINE", "off")
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger
logger.remove()
from dynamicprogramming_b200 import envs
import numpy as np
eng = envs.make("double_cartpole_swingup", bins=20)
eng.build_table()
done = 0
for target in (200, 1000, 3000, 6000):
    eng.sweeps(target - done); done = target
    # improvement changes the policy; evaluate regularity, then restore policy 0 to keep evaluating it
    eng.policy_improvement()
    for cfg in ("2,0,4,8,2,1:1,1,1,2,10", "4,0,4,8,1,1:1,1,1,4,10"):
        print(target, cfg, eng.debug_xline(cfg, iters=3), flush=True)
    v, p = eng.download()
    p6 = p.reshape([20] * 6)
    same4 = (p6.reshape(5, 4, -1) == p6.reshape(5, 4, -1)[:, :1]).all(axis=1).mean()
    same2 = (p6.reshape(10, 2, -1) == p6.reshape(10, 2, -1)[:, :1]).all(axis=1).mean()
    print(target, "uniform 4-groups", round(float(same4), 4), "2-groups", round(float(same2), 4), "hist", (np.bincount(p, minlength=9) / p.size).round(3), flush=True)
    eng.upload_policy(np.zeros(eng.n_states, np.int32))
eng.close()
