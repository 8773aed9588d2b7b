#!/bin/bash
# 2 GPUs: sharded == unsharded for every sweep family incl. the plane-staged sweep; bench at N = 2 / 1 / reference with the V checksum
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2j_multi.log 2>&1; tail -12 gpurun_out/r2j_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 10 --warmup 3 --no-stable --no-extras > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; cut -c1-200 gpurun_out/r2j_bench_n2.json; grep -iE "error|timed out|Traceback" gpurun_out/r2j_bench_n2.err | tail -3
python bench.py --steps 10 --warmup 3 --no-stable --no-extras --no-cpu-baseline --no-converge > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; cut -c1-200 gpurun_out/r2j_bench_n1.json
python bench.py --impl reference --steps 10 --warmup 3 --no-extras --no-converge > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err; cut -c1-200 gpurun_out/r2j_bench_ref.json
python - <<'PY'
import json
for f in ("n2","n1","ref"):
    try:
        d=json.load(open(f"gpurun_out/r2j_bench_{f}.json")); print(f, d["value"]/1e9, d["e2e"]["value"]/1e9, d.get("v_checksum"), d.get("v_checksum_after_sweeps"), d.get("roofline",{}).get("kernel","")[:60])
    except Exception as ex: print(f, "ERR", ex)
PY
