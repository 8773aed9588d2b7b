#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python scripts/exp_plane.py 20 "0,0,2,2,0,0" "0,0,2,2,0,1" "0,0,2,2,0,2" "0,0,2,2,0,3" "0,0,2,2,0,5" "0,40,2,2,0,3" ) > gpurun_out/r2p_exp_plane.log 2>&1; grep -v INFO gpurun_out/r2p_exp_plane.log | tail -7 | cut -c1-120
timeout 300 python scripts/prof_plane.py 20 "0,0,2,2,0,3" policy0 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none -k regex:ps_sweep -s 56 -c 1 -o gpurun_out/r2p_ps python scripts/prof_plane.py 20 "0,0,2,2,0,3" > gpurun_out/r2p_prof.log 2>&1; tail -1 gpurun_out/r2p_prof.log
