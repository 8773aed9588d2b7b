#!/bin/bash
N=${1:-8}
for cfg in "TAG=default" "TAG=ps DPB200_PLANE=force" "TAG=gp_dma DPB200_PLANE=off" "TAG=gp_p2p DPB200_PLANE=off DPB200_EXCHANGE=p2p" "TAG=ps_nostore DPB200_PLANE=force DPB200_XDEBUG=nostore"; do
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/exp_shard.py 2>&1 | grep -E "RESULT|rror|timing on" | cut -c1-250
done
