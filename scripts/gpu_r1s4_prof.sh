#!/bin/bash
# session 4 profiles: ncu --set full of the default generic sweep of K5 (gp_sweep, one state per thread) under the
# bench policy, and the launch list of the bench command
mkdir -p gpurun_out
export DPB200_XLINE=off
ncu --set full --clock-control none --import-source on -k regex:gp_sweep -s 62 -c 1 -f -o gpurun_out/r01s4_gp_single_benchpolicy \
    python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --pre-sweeps 50 --sweeps 12 --improve 1 > gpurun_out/ncu_gp_single.log 2>&1
tail -2 gpurun_out/ncu_gp_single.log
unset DPB200_XLINE
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01s4_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-converge > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r01s4_bench_launches.csv
