#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_plane.py -x -q ) > gpurun_out/r2d_plane_tests.log 2>&1; tail -12 gpurun_out/r2d_plane_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ps_sweep -s 56 -c 1 -o gpurun_out/r2d_ps python scripts/prof_plane.py 20 "0,0,2,2,0" > gpurun_out/r2d_prof.log 2>&1; tail -3 gpurun_out/r2d_prof.log
