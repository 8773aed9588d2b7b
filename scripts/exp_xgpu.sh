#!/bin/bash
# per-sweep cost of the sharded exchange on N GPUs: peer stores and cross-GPU barrier switched off one at a time
# (DPB200_XDEBUG: timing only, results wrong).  usage: exp_xgpu.sh N
N=${1:-2}
mkdir -p gpurun_out
for x in ${2:-normal nostore nobarrier none}; do
  DPB200_XDEBUG=${x/normal/} timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-stable --no-extras > gpurun_out/xgpu_${N}_${x}.json 2> gpurun_out/xgpu_${N}_${x}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/xgpu_${N}_${x}.json")); print("${x}", "N=$N", "%.4f ms/sweep" % (d["ms_per_step"]/25), "%.1f G" % (d["value"]/1e9), d["roofline"]["kernel"][:40])
PY
done
