#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
N=${1:-2}
if [ "$N" = 2 ]; then
DPB200_EXCHANGE=p2p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 scripts/multi_gpu_check.py 2>&1 | grep -E "^OK|^FAIL|rror|Traceback|timed out" | cut -c1-120
fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29503 scripts/k5_run.py --verbose --cap-s 4 --out gpurun_out/h_k5_${N}gpu.json 2>gpurun_out/h_k5_${N}gpu.err | tail -4 | cut -c1-400
grep -E "build phases|Shard 0|JIT sweep|x-line sweep|rror" gpurun_out/h_k5_${N}gpu.err | cut -c1-260 | tail -12
