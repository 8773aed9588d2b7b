#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
python scripts/k4_scale.py 2>/dev/null | tee gpurun_out/k4_scale.log
for N in 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N scripts/k4_scale.py 2>gpurun_out/k4_n$N.err | grep n_gpus | tee -a gpurun_out/k4_scale.log
done
