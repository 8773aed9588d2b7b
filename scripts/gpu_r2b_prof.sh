#!/bin/bash
# ncu --set full of the item-mode plane-staged K5 sweep under the bench policy (one launch)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-id ::ps_sweep:70 -o gpurun_out/r2b_ps_items -f python scripts/prof_plane.py 20 "${1:-0,0,2,2,2}" > gpurun_out/r2b_prof.log 2>&1; tail -3 gpurun_out/r2b_prof.log
