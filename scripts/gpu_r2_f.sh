#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_plane.py -x -q ) > gpurun_out/r2f_plane_tests.log 2>&1; grep -E "passed|failed|Error" gpurun_out/r2f_plane_tests.log | tail -4
for am in 50 0 100000; do
  echo "== DPB200_PLANE_ALIGN_MIN=$am"
  ( DPB200_PLANE_ALIGN_MIN=$am timeout 600 python scripts/exp_plane.py 20 "0,0,2,2,0" "0,40,2,2,0" ) > gpurun_out/r2f_exp_plane_$am.log 2>&1; grep -v INFO gpurun_out/r2f_exp_plane_$am.log | tail -2
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ps_sweep -s 56 -c 1 -o gpurun_out/r2f_ps python scripts/prof_plane.py 20 "0,0,2,2,0" > gpurun_out/r2f_prof.log 2>&1; tail -2 gpurun_out/r2f_prof.log
