#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + both bench arms (numbers under ncu are never bench values)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-converge > gpurun_out/bench_under_ncu.log 2>&1
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_ours.json | cut -c1-600
python bench.py --impl reference --steps 5 --warmup 3 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json | cut -c1-300
# first evaluation of a full run (policy 0 is regular along x): the x-line kernel is selected
DPB200_XLINE=auto python scripts/exp_auto.py 2>&1 | grep -E "x-line|sweeps\(|xline" | cut -c1-400
