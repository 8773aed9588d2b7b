#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_plane.py -x -q ) > gpurun_out/r2w_plane_tests.log 2>&1; grep -E "passed|failed|Error" gpurun_out/r2w_plane_tests.log | tail -3
( timeout 600 python scripts/exp_plane.py 20 "0,20,2,2,0" "0,10,2,2,0" "0,5,2,2,0" "0,40,2,2,0" ) > gpurun_out/r2w_exp_plane.log 2>&1; grep ms_plane gpurun_out/r2w_exp_plane.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['cfg'], round(d['ms_plane'],4), round(d['ms_base'],4), round(d['loads_per_plane'],1), round(d['late_per_plane'],1), d['slots'], d['mismatches'])"
