#!/bin/bash
N=${1:-2}
for cfg in "TAG=gp_default DPB200_PLANE=off" "TAG=gp_localstore DPB200_PLANE=off DPB200_XDEBUG=localstore" "TAG=ps_localstore DPB200_PLANE=force DPB200_XDEBUG=localstore" "TAG=ps_nostore DPB200_PLANE=force DPB200_XDEBUG=nostore"; do
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/exp_shard.py 2>&1 | grep -E "RESULT" | cut -c1-220
done
