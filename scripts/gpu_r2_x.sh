#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2x_multi.log 2>&1; grep -E "passed|failed|FAIL|Error" gpurun_out/r2x_multi.log | tail -5
for cfg in "TAG=default" "TAG=ps DPB200_PLANE=force" "TAG=gp_dma DPB200_PLANE=off" "TAG=ps_nostore DPB200_PLANE=force DPB200_XDEBUG=nostore"; do
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/exp_shard.py 2>&1 | grep -E "RESULT|rror" | cut -c1-220
done
