#!/bin/bash
mkdir -p gpurun_out
# launch list of the bench command; the build-time autotune compares event timings, which ncu's per-launch
# serialisation distorts, so the kernel the un-profiled bench selects is forced here
DPB200_PAIR=force:128,8,2,8,1 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01s4_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-converge > gpurun_out/bench_under_ncu.log 2>&1
tail -c 200 gpurun_out/bench_under_ncu.log
timeout 400 python scripts/k5_run.py --cap-s 300 --out gpurun_out/e_k5_run.json > gpurun_out/e_k5_run.log 2>&1; tail -20 gpurun_out/e_k5_run.log
