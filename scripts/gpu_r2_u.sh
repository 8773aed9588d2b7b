#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2u_multi.log 2>&1; grep -E "passed|failed" gpurun_out/r2u_multi.log
for cfg in "TAG=default" "TAG=plane_off DPB200_PLANE=off" "TAG=ps_nostore DPB200_PLANE=force DPB200_XDEBUG=nostore" "TAG=nobarrier DPB200_XDEBUG=nobarrier"; do
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/exp_shard.py 2>&1 | grep -E "RESULT" | cut -c1-220
done
