#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ps_sweep -s 8 -c 1 -o gpurun_out/r2b_ps python scripts/prof_plane.py 20 > gpurun_out/r2b_prof.log 2>&1; tail -5 gpurun_out/r2b_prof.log
timeout 300 python scripts/prof_plane.py 20 "" policy0 2>&1 | tail -2
