#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:eval_sweep -s 30 -c 1 -f -o gpurun_out/p7_persist \
    python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 40 --improve 1 > gpurun_out/p7a.log 2>&1
DPB200_EVAL_KERNEL=pair ncu --set full --clock-control none --import-source on -k regex:eval_sweep -s 30 -c 1 -f -o gpurun_out/p7_pair \
    python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 40 --improve 1 > gpurun_out/p7b.log 2>&1
tail -2 gpurun_out/p7a.log gpurun_out/p7b.log
