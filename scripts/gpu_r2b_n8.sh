#!/bin/bash
# K5 bench sharded over 8 x B200 (item-mode plane sweep): gpurun --gpus 8 -- bash scripts/gpu_r2b_n8.sh
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-converge --no-extras 2>gpurun_out/r2b_bench_n$N.err | tee gpurun_out/r2b_bench_n$N.json | cut -c1-300
grep -iE "error|timed out|Traceback" gpurun_out/r2b_bench_n$N.err | tail -3
