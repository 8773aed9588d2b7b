#!/bin/bash
# session 4: 2-GPU bit-identity check (incl. the JIT generic sweeps under sharding) + 2-GPU bench, run under gpurun --gpus 2
cd /root/repo; mkdir -p gpurun_out
for mode in p2p nccl; do
  echo "== exchange=$mode: sharded == unsharded, bit for bit"
  DPB200_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 scripts/multi_gpu_check.py 2>&1 | grep -E "^OK|^FAIL|rror|Traceback|timed out" | cut -c1-250
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/m_bench_n2.err | tee gpurun_out/m_bench_n2.json | cut -c1-400
grep -iE "error|timed out|Traceback" gpurun_out/m_bench_n2.err | tail -3
