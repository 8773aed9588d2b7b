#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_plane.py -x -q ) > gpurun_out/r2e_plane_tests.log 2>&1; tail -4 gpurun_out/r2e_plane_tests.log
( time timeout 600 python scripts/exp_plane.py 20 "0,0,2,2,0" "0,0,2,2,1" "0,40,2,2,0" "60,0,2,2,0" ) > gpurun_out/r2e_exp_plane.log 2>&1; grep -v INFO gpurun_out/r2e_exp_plane.log | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ps_sweep -s 56 -c 1 -o gpurun_out/r2e_ps python scripts/prof_plane.py 20 "0,0,2,2,0" > gpurun_out/r2e_prof.log 2>&1; tail -2 gpurun_out/r2e_prof.log
