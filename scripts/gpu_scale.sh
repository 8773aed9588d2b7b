#!/bin/bash
# usage: gpu_scale.sh "<N list>" [modes]   (run under gpurun --gpus max(N))
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in $1; do
  for mode in ${2:-p2p}; do
    if [ $N = 1 ]; then
      python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-converge 2>gpurun_out/scale_n1.err | tee gpurun_out/scale_n1.json | cut -c1-180
    else
      DPB200_EXCHANGE=$mode timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/scale_n${N}_$mode.err | tee gpurun_out/scale_n${N}_$mode.json | cut -c1-180
      grep -iE "error|timed out|Traceback" gpurun_out/scale_n${N}_$mode.err | tail -3
    fi
  done
done
