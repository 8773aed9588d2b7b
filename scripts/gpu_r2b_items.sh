#!/bin/bash
# item-mode plane sweep: bit-parity tests on small grids, then K5 timing vs the other modes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_plane.py -x -q -m gpu -k "bit_identical or falls_back" > gpurun_out/r2b_plane_tests.log 2>&1
tail -5 gpurun_out/r2b_plane_tests.log
timeout 400 python scripts/exp_plane.py 20 "0,0,2,2,2" "0,0,2,2,0" "0,0,2,1,2" "0,0,2,3,2" "0,10,2,2,2" > gpurun_out/r2b_exp_plane.log 2>&1
grep -v "^layout" gpurun_out/r2b_exp_plane.log | cut -c1-400 | tail -12
