#!/usr/bin/env python3
"""K5 (double cartpole swing-up, 6-D, --bins 20) run to the end — or to a wall-clock cap — with one
line per PI iteration: sweeps, device time of the evaluation, which sweep kernel ran, changed states.
Answers "which policies does a real run spend its sweeps on" (regular ones -> x-line sweep, or not).

    python scripts/k5_run.py [--bins 20] [--cap-s 300] [--env double_cartpole_swingup] [--out gpurun_out/k5_run.json]
"""
import argparse
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger

logger.remove()
if "--verbose" in sys.argv:      # engine log lines (build phases, kernel selection) of rank 0 to stderr
    sys.argv.remove("--verbose")
    import os as _os
    if int(_os.environ.get("RANK", "0")) == 0:
        logger.add(sys.stderr, level="INFO")
from dynamicprogramming_b200 import envs

ap = argparse.ArgumentParser()
ap.add_argument("--env", default="double_cartpole_swingup")
ap.add_argument("--bins", type=int, default=20)
ap.add_argument("--cap-s", type=float, default=300.0)
ap.add_argument("--out", default="gpurun_out/k5_run.json")
ap.add_argument("--checksum", action="store_true", help="SHA-1 of the final policy and V (compare 1-GPU and N-GPU runs)")
a = ap.parse_args()

import os
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
shard = None
if world > 1:   # torchrun --nproc-per-node N scripts/k5_run.py: state-range sharding, every rank ends with the full result
    import torch
    from dynamicprogramming_b200 import dist as pdist
    torch.cuda.set_device(local)
    td = pdist.init_process_group()
    shard = pdist.make_shard(local)
t00 = time.perf_counter()
eng = envs.make(a.env, bins=a.bins, device=local, shard=shard)
t_create = time.perf_counter() - t00
t = time.perf_counter()
eng.build_table()
t_build = time.perf_counter() - t
rows = []
stable = False
ev_prev = 0.0
for n in range(eng.config.max_pi_iter):
    k = eng.eval_kernel_info()
    t = time.perf_counter()
    delta = eng.policy_evaluation()
    te = time.perf_counter() - t
    st = eng.engine_stats()
    ev_ms = st["eval_ms"] - ev_prev
    ev_prev = st["eval_ms"]
    t = time.perf_counter()
    stable = eng.policy_improvement()
    ti = time.perf_counter() - t
    row = {"pi": n + 1, "sweeps": eng.last_eval_sweeps, "eval_wall_s": round(te, 3), "eval_dev_ms": round(ev_ms, 1),
           "ms_per_sweep": round(ev_ms / max(eng.last_eval_sweeps, 1), 4), "delta": delta, "xline": k["xline"],
           "improve_wall_ms": round(ti * 1e3, 1), "changed": eng.last_changed}
    rows.append(row)
    if rank == 0:
        print(row, flush=True)
    if stable or time.perf_counter() - t00 > a.cap_s:
        break
total = time.perf_counter() - t00
st = eng.engine_stats()
out = {"env": a.env, "bins": a.bins, "n_states": eng.n_states, "stable": bool(stable), "pi_iterations": len(rows),
       "total_sweeps": int(sum(r["sweeps"] for r in rows)), "wall_s": round(total, 2), "create_s": round(t_create, 2),
       "build_s": round(t_build, 2), "eval_ms": st["eval_ms"], "improve_ms": st["improve_ms"],
       "xline_sweeps": int(sum(r["sweeps"] for r in rows if r["xline"])), "rows": rows}
out["n_gpus"] = world
if a.checksum:
    import hashlib
    v, p = eng.download()
    out["policy_sha1"] = hashlib.sha1(p.tobytes()).hexdigest()
    out["value_sha1"] = hashlib.sha1(v.tobytes()).hexdigest()
if rank == 0:
    print({k: v for k, v in out.items() if k != "rows"}, flush=True)
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(out))
eng.close()
if world > 1:
    td.barrier()
    td.destroy_process_group()
