#!/usr/bin/env python3
"""torchrun --nproc-per-node N scripts/multi_gpu_check.py — sharded == unsharded, bit for bit.

Every rank builds its shard; rank 0 additionally runs the same problem unsharded
and compares V, policy, PI iterations and sweep counts (SURVEY §8e invariant)."""
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger

logger.remove()
import torch

from dynamicprogramming_b200 import _ffi, dist as pdist, envs

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
td = pdist.init_process_group()
# (env, bins, max_pi_iter, engine environment of the SHARDED run); the 1-GPU comparison run always uses the
# scalar sweep, so the x-line cases also prove: sharded x-line sweep == unsharded scalar sweep, bit for bit
XL = {"DPB200_FAST_DIM": "0", "DPB200_XLINE": "force:2,0,4,8,2,1:1,1,1,8,8"}
XL4 = {"DPB200_FAST_DIM": "0", "DPB200_XLINE": "force:4,0,4,8,1,1:1,6,12"}
# the JIT generic sweeps (one state per thread with grouped gathers — the default of large 6-D grids — and packed pairs)
GP1 = {"DPB200_XLINE": "off", "DPB200_PAIR": "force:128,8,2,8,1"}
GP2 = {"DPB200_XLINE": "off", "DPB200_PAIR": "force:64,8,2,8,0"}
# the plane-staged sweep (TMA-staged V-planes; needs plane-aligned shard boundaries), default and with few slots / ragged chunks
PL = {"DPB200_PLANE": "force"}
PL2 = {"DPB200_PLANE": "force:40,7,2,2,1"}
# ... "force" is item mode (two states per thread); one state per thread, and item mode with few slots / ragged chunks
PL0 = {"DPB200_PLANE": "force:0,0,2,2,0"}
PL3 = {"DPB200_PLANE": "force:40,7,2,2,2"}
# the gather sweep with the opt-in DMA exchange (boundary part, copy engines, interior part) instead of in-kernel peer stores,
# and both JIT sweeps over NCCL send/recv
GP1d = {"DPB200_XLINE": "off", "DPB200_PAIR": "force:128,8,2,8,1", "DPB200_EXCHANGE": "dma"}
PLn = {"DPB200_PLANE": "force", "DPB200_EXCHANGE": "nccl"}
cases = [("cartpole", 12, 4, {}), ("mountain_car", 60, None, {}), ("double_pendulum_swingup", 10, 3, {}),
         ("double_cartpole_swingup", 7, 2, {}), ("pendulum", 33, 6, {}),
         ("double_cartpole_swingup", 8, 2, XL), ("cartpole_swingup", 12, 3, XL4),
         ("double_cartpole_swingup", 7, 2, GP1), ("cartpole_swingup", 11, 3, GP1), ("double_cartpole", 6, 2, GP2),
         ("double_cartpole_swingup", 8, 3, PL), ("cartpole", 12, 4, PL), ("double_cartpole", 6, 2, PL2),
         ("double_cartpole_swingup", 10, 2, PL2), ("double_cartpole_swingup", 8, 3, PLn), ("double_cartpole_swingup", 7, 2, GP1d),
         ("double_cartpole_swingup", 8, 3, PL0), ("double_cartpole_swingup", 10, 2, PL3), ("cartpole_swingup", 12, 3, PL3)]
ok = True
for env, bins, max_pi, engine_env in cases:
    for k in ("DPB200_FAST_DIM", "DPB200_XLINE", "DPB200_PAIR", "DPB200_PLANE", "DPB200_EXCHANGE"):
        os.environ.pop(k, None)
    os.environ.update(engine_env)
    spec = envs.REGISTRY[env]
    cfg = spec.config()
    if max_pi:
        cfg.max_pi_iter = max_pi
    shard = pdist.make_shard(local)
    eng = spec.make(bins=bins, config=cfg, device=local, shard=shard)
    eng.build_table()
    kernel = eng.eval_kernel_info()["kernel"]
    if "DPB200_PLANE" in engine_env:
        assert "ps_sweep" in kernel, kernel
    elif "DPB200_PAIR" in engine_env:
        assert "gp_sweep" in kernel, kernel
    elif engine_env:
        assert "xl_sweep" in kernel, kernel
    eng.run()
    res = dict(env=env, bins=bins, N=eng.n_states, pi=eng.pi_iterations, sweeps=eng.total_eval_sweeps, kernel=kernel[:40])
    if rank == 0:
        os.environ["DPB200_XLINE"] = "off"
        os.environ["DPB200_PAIR"] = "off"
        os.environ["DPB200_PLANE"] = "off"
        one = spec.make(bins=bins, config=cfg, device=local)
        one.run()
        res.update(pi_1gpu=one.pi_iterations, sweeps_1gpu=one.total_eval_sweeps,
                   policy_diff=int(np.sum(one.policy != eng.policy)),
                   v_bitdiff=int(np.sum(one.value_function.view(np.uint32) != eng.value_function.view(np.uint32))))
        good = (res["policy_diff"] == 0 and res["v_bitdiff"] == 0 and res["pi"] == res["pi_1gpu"]
                and res["sweeps"] == res["sweeps_1gpu"])
        ok &= good
        print(("OK   " if good else "FAIL ") + json.dumps(res), flush=True)
    td.barrier()
flag = torch.tensor([1 if ok else 0], device="cuda")
td.broadcast(flag, src=0)
td.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
