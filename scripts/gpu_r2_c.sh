#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_plane.py -x -q ) > gpurun_out/r2c_plane_tests.log 2>&1; tail -12 gpurun_out/r2c_plane_tests.log
( time timeout 600 python scripts/exp_plane.py 20 "" "0,0,2,2,0" "0,0,2,1,1" "0,40,2,2,1" "64,0,2,2,1" ) > gpurun_out/r2c_exp_plane.log 2>&1; grep -v INFO gpurun_out/r2c_exp_plane.log | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ps_sweep -s 60 -c 1 -o gpurun_out/r2c_ps python scripts/prof_plane.py 20 > gpurun_out/r2c_prof.log 2>&1; tail -3 gpurun_out/r2c_prof.log
