#!/bin/bash
# ncu --set full of the plane-staged K5 sweep under the bench policy (the 70th ps_sweep launch of scripts/prof_plane.py:
# 4 autotune probes + 50 sweeps of policy 0 come first) -> profiles/r02_ncu_summary_ps_sweep.txt
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-id ::ps_sweep:70 -o gpurun_out/r02_ps_sweep python scripts/prof_plane.py 20 > gpurun_out/r02_prof.log 2>&1; tail -2 gpurun_out/r02_prof.log
