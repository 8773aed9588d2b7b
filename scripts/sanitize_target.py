#!/usr/bin/env python3
"""One sweep family on a small grid, for compute-sanitizer (tests/test_gpu_sanitizer.py):
    compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitize_target.py <family>
families: aot | gp_single | gp_pair | xline | persistent | plane | plane_small | plane_scalar | plane_items_small | lookup"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
family = sys.argv[1]
env, bins = "double_cartpole_swingup", 6
cfg = {
    "aot": {"DPB200_PLANE": "off", "DPB200_XLINE": "off", "DPB200_PAIR": "off"},
    "gp_single": {"DPB200_PLANE": "off", "DPB200_XLINE": "off", "DPB200_PAIR": "force:128,8,2,8,1"},
    "gp_pair": {"DPB200_PLANE": "off", "DPB200_XLINE": "off", "DPB200_PAIR": "force:64,8,2,8,0"},
    "xline": {"DPB200_PLANE": "off", "DPB200_FAST_DIM": "0", "DPB200_XLINE": "force:2,0,4,8,2,1:1,1,1,2,6"},
    "persistent": {"DPB200_PLANE": "off", "DPB200_XLINE": "off", "DPB200_PAIR": "off", "DPB200_PERSIST": "on"},
    "plane": {"DPB200_PLANE": "force"},
    "plane_small": {"DPB200_PLANE": "force:34,5,2,1,1"},
    "plane_scalar": {"DPB200_PLANE": "force:0,0,2,2,0"},
    "plane_items_small": {"DPB200_PLANE": "force:34,5,2,2,2"},
    "lookup": {},
}[family]
os.environ.update(cfg)
os.environ["DPB200_CACHE"] = "off"
if family == "persistent":
    env, bins = "cartpole", 10
import numpy as np

from dynamicprogramming_b200 import envs

eng = envs.make(env, bins=bins)
eng.build_table()
info = eng.eval_kernel_info()["kernel"]
want = {"aot": "eval_sweep_kernel", "gp_single": "gp_sweep", "gp_pair": "gp_sweep", "xline": "xl_sweep",
        "persistent": "eval_persistent_kernel", "plane": "ps_sweep", "plane_small": "ps_sweep", "plane_scalar": "ps_sweep", "plane_items_small": "ps_sweep", "lookup": ""}[family]
assert want in info, (family, info)
eng.sweeps(3)
eng.policy_improvement()
d, _ = eng.sweeps(26)                 # a whole sync interval + a check sweep
eng.upload_policy(np.random.default_rng(1).integers(0, eng.n_actions, eng.n_states).astype(np.int32))
eng.sweeps(2)
if family == "lookup":
    pts = np.random.default_rng(2).uniform(-1, 1, (257, eng.N_DIMS)).astype(np.float32)
    eng.lookup_actions(pts)
v, p = eng.download()
assert np.isfinite(v).all()
eng.close()
print("SANITIZE_TARGET_OK", family, info[:60])
