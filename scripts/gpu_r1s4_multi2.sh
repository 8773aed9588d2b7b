#!/bin/bash
# 2-GPU bench after the residual-buffer fix + compute-sanitizer memcheck of the small-block JIT sweeps
cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/m_bench_n2.err | tee gpurun_out/m_bench_n2.json | cut -c1-700
grep -iE "error|timed out|Traceback" gpurun_out/m_bench_n2.err | tail -3
for cfg in "force:32,16,1,8,1" "force:64,8,2,8,0" "off"; do
  echo "== memcheck DPB200_PAIR=$cfg"
  DPB200_PAIR=$cfg DPB200_XLINE=off CUDA_VISIBLE_DEVICES=0 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/prof_eval.py --env double_cartpole --bins 6 --pre-sweeps 26 --sweeps 26 --improve 1 2>&1 | grep -E "ERROR SUMMARY|Invalid|ms/sweep|rror" | head -5
done
echo "== memcheck x-line"
DPB200_FAST_DIM=0 DPB200_XLINE="force:2,0,4,8,2,1:1,1,1,2,3" CUDA_VISIBLE_DEVICES=0 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/prof_eval.py --env double_cartpole_swingup --bins 6 --pre-sweeps 26 --sweeps 26 --improve 1 2>&1 | grep -E "ERROR SUMMARY|Invalid|ms/sweep|rror" | head -5
