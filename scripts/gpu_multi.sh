#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 scripts/multi_gpu_check.py 2>&1 | tail -15
for n in 1 2; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-converge 2>gpurun_out/bench_n1.err | tee gpurun_out/bench_n1.json | cut -c1-330
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/bench_n$n.err | tee gpurun_out/bench_n$n.json | cut -c1-330
    tail -5 gpurun_out/bench_n$n.err
  fi
done
