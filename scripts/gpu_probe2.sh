#!/bin/bash
# timing after the atomics fix + ncu capture of the 6-D evaluation sweep
set -x
mkdir -p gpurun_out
python scripts/prof_eval.py --env cartpole --bins 30 --sweeps 200
python scripts/prof_eval.py --env double_pendulum_swingup --bins 50 --sweeps 100
python scripts/prof_eval.py --env double_cartpole_swingup --bins 12 --sweeps 100
python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 50
ncu --set full --clock-control none --import-source on -k regex:eval_sweep_kernel -s 2 -c 2 -f -o gpurun_out/prof_eval6d \
    python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 6 --improve 1 > gpurun_out/ncu_eval6d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_sweep_kernel -s 2 -c 2 -f -o gpurun_out/prof_eval4d \
    python scripts/prof_eval.py --env double_pendulum_swingup --bins 50 --sweeps 6 --improve 1 > gpurun_out/ncu_eval4d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:improve_kernel -c 1 -f -o gpurun_out/prof_improve6d \
    python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 2 --improve 1 > gpurun_out/ncu_imp6d.log 2>&1
tail -3 gpurun_out/ncu_eval6d.log
ls -la gpurun_out
