#!/bin/bash
# 8 GPUs: bench at N = 8 (with the K5 run to a stable policy) and N = 4
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
  extra="--no-extras"; [ $N = 4 ] && extra="--no-extras --no-stable"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --steps 10 --warmup 3 $extra > gpurun_out/r2l_bench_n$N.json 2> gpurun_out/r2l_bench_n$N.err
  grep -iE "error|timed out|Traceback" gpurun_out/r2l_bench_n$N.err | tail -3
done
python - <<'PY'
import json
for f in ("n8","n4"):
    try:
        d=json.load(open(f"gpurun_out/r2l_bench_{f}.json")); print(f, d["value"]/1e9, d["e2e"]["value"]/1e9, d.get("v_checksum"), d.get("roofline",{}).get("kernel","")[:80], d.get("k5_to_stable"))
    except Exception as ex: print(f, "ERR", ex)
PY
grep -h "timing on this policy\|evaluation sweeps:" gpurun_out/r2l_bench_n8.err | head -5
