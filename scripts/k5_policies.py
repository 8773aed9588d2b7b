#!/usr/bin/env python3
"""Run K5 policy iteration and keep the policy after selected PI iterations (uint8, compressed) for
offline locality analysis (which V lines does a tile of states gather from under a REAL policy?)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger
logger.remove()
import numpy as np
from dynamicprogramming_b200 import envs

keep = {int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,5,16").split(",")}
eng = envs.make("double_cartpole_swingup", bins=20)
eng.build_table()
out = {}
t0 = time.perf_counter()
for n in range(1, max(keep) + 1):
    eng.policy_evaluation()
    stable = eng.policy_improvement()
    if n in keep or stable:
        _, p = eng.download()
        out[f"policy_{n}"] = p.astype(np.uint8)
        print(n, "kept", eng.eval_kernel_info()["kernel"], round(time.perf_counter() - t0, 1), flush=True)
    if stable:
        break
Path("gpurun_out").mkdir(exist_ok=True)
np.savez_compressed("gpurun_out/k5_policies.npz", **out)
print({k: v.shape for k, v in out.items()}, Path("gpurun_out/k5_policies.npz").stat().st_size)
eng.close()
