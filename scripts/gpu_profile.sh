#!/bin/bash
# ncu evidence for profiles/: full capture of the dominant kernel + launch list of the bench command
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:eval_sweep_kernel -s 30 -c 2 -f -o gpurun_out/r01_eval6d \
    python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 40 --improve 1 > gpurun_out/ncu_eval6d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_sweep_kernel -s 30 -c 2 -f -o gpurun_out/r01_eval4d \
    python scripts/prof_eval.py --env double_pendulum_swingup --bins 50 --sweeps 40 --improve 1 > gpurun_out/ncu_eval4d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:improve_kernel -c 1 -f -o gpurun_out/r01_improve6d \
    python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 2 --improve 1 > gpurun_out/ncu_imp6d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pi_build_rows -s 6 -c 1 -f -o gpurun_out/r01_build6d \
    python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 2 --improve 0 > gpurun_out/ncu_build6d.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-converge > gpurun_out/bench_under_ncu.log 2>&1
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_ours.json | cut -c1-400
python bench.py --impl reference --steps 5 --warmup 3 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json | cut -c1-300
ls -la gpurun_out
