#!/usr/bin/env python3
"""K4 (BASELINE configs[3]): Double Pendulum swing-up 4-D --bins 50 (6.25 M states x 11 actions), state grid
sharded across N GPUs (torchrun --nproc-per-node N scripts/k4_scale.py) — evaluation-sweep throughput under the
greedy policy after 50 sweeps of policy 0, like bench.py does for K5."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger
logger.remove()
import torch
from dynamicprogramming_b200 import dist as pdist, envs

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
td = pdist.init_process_group() if world > 1 else None
shard = pdist.make_shard(local) if world > 1 else None
env, bins = os.environ.get("EXP_ENV", "double_pendulum_swingup"), int(os.environ.get("EXP_BINS", "50"))
eng = envs.make(env, bins=bins, device=local, shard=shard)
eng.build_table()
eng.sweeps(50); eng.policy_improvement()
for _ in range(3): eng.sweeps(100)
ms = min(eng.sweeps(200)[1] / 200 for _ in range(4))
if td is not None:
    t = torch.tensor([ms], device="cuda", dtype=torch.float64); td.all_reduce(t, op=td.ReduceOp.MAX); ms = float(t.item())
if rank == 0:
    print(json.dumps({"env": env, "bins": bins, "n_states": eng.n_states, "n_gpus": world, "us_per_sweep": ms * 1e3,
                      "backups_per_s": eng.n_states / (ms * 1e-3), "kernel": eng.eval_kernel_info()["kernel"][:60],
                      "exchange": os.environ.get("DPB200_EXCHANGE", "p2p")}), flush=True)
eng.close()
if td is not None:
    td.barrier(); td.destroy_process_group()
