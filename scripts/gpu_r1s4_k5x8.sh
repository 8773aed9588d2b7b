#!/bin/bash
# K5 to a stable policy on 8 x B200 (state-range sharding, fused p2p V exchange); run under gpurun --gpus 8
cd /root/repo; mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29503 scripts/k5_run.py --cap-s 200 --checksum --out gpurun_out/g_k5_run_8gpu.json 2>gpurun_out/g_k5_8gpu.err | tail -20
grep -iE "error|timed out|Traceback" gpurun_out/g_k5_8gpu.err | tail -3
