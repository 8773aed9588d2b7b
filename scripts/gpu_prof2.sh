#!/bin/bash
mkdir -p gpurun_out
export DPB200_XLINE=off
ncu --set full --clock-control none --import-source on -k regex:eval_sweep_kernel -s 60 -c 1 -f -o gpurun_out/r01_eval6d_benchpolicy \
    python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --pre-sweeps 50 --sweeps 12 --improve 1 > gpurun_out/ncu_eval6d_bp.log 2>&1
tail -2 gpurun_out/ncu_eval6d_bp.log
