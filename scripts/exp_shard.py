#!/usr/bin/env python3
"""torchrun --nproc-per-node N scripts/exp_shard.py: K5 sharded, exchange plan (log) and ms per sweep under the bench policy."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from loguru import logger
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
logger.remove()
if rank == 0: logger.add(sys.stderr, level="INFO", filter=lambda r: "Shard" in r["message"] or "timing" in r["message"] or "evaluation sweeps" in r["message"])
from dynamicprogramming_b200 import dist as pdist, envs
torch.cuda.set_device(local)
td = pdist.init_process_group()
eng = envs.make("double_cartpole_swingup", bins=20, device=local, shard=pdist.make_shard(local))
eng.build_table(); eng.sweeps(50); eng.policy_improvement()
for _ in range(3): eng.sweeps(25)
ms_all = []
for _ in range(5):
    _, ms = eng.sweeps(25); ms_all.append(ms / 25)
t = torch.tensor([min(ms_all)], device="cuda"); td.all_reduce(t, op=td.ReduceOp.MAX)
if rank == 0: print("RESULT", os.environ.get("TAG", ""), "ms/sweep %.4f" % t.item(), eng.eval_kernel_info()["kernel"][:30], flush=True)
eng.close(); td.barrier(); td.destroy_process_group()
