#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_plane.py 20 "0,0,2,2,0" policy0 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none -k regex:ps_sweep -s 6 -c 1 -o gpurun_out/r2o_ps_p0 python scripts/prof_plane.py 20 "0,0,2,2,0" policy0 > gpurun_out/r2o_prof.log 2>&1; tail -1 gpurun_out/r2o_prof.log
