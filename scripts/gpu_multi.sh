#!/bin/bash
# usage: gpu_multi.sh [N]   (run under gpurun --gpus N)
N=${1:-2}
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 scripts/multi_gpu_check.py 2>&1 | tail -12
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-converge 2>gpurun_out/bench_n1.err | tee gpurun_out/bench_n1.json | cut -c1-330
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-330
grep -i "shard\|error" gpurun_out/bench_n$N.err | tail -6
