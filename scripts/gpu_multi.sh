#!/bin/bash
# usage: gpu_multi.sh [N]   (run under gpurun --gpus N)
N=${1:-2}
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | head -$N
for mode in p2p nccl; do
  echo "== exchange=$mode: sharded == unsharded, bit for bit"
  DPB200_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 scripts/multi_gpu_check.py 2>&1 | grep -E "^OK|^FAIL|rror|Traceback|timed out" | cut -c1-250
done
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-converge 2>gpurun_out/bench_n1.err | tee gpurun_out/bench_n1.json | cut -c1-200
for mode in p2p nccl; do
  DPB200_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_n${N}_$mode.err | tee gpurun_out/bench_n${N}_$mode.json | cut -c1-200
  grep -iE "error|timed out" gpurun_out/bench_n${N}_$mode.err | tail -3
done
