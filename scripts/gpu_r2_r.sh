#!/bin/bash
N=${1:-2}
for cfg in "TAG=default" "TAG=norotate DPB200_ROTATE=off" "TAG=plane_off DPB200_PLANE=off" "TAG=plane_off_norotate DPB200_PLANE=off DPB200_ROTATE=off" "TAG=nostore DPB200_XDEBUG=nostore" "TAG=nobarrier DPB200_XDEBUG=nobarrier"; do
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/exp_shard.py 2>&1 | grep -E "RESULT|receives" | cut -c1-220
done
