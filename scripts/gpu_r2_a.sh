#!/bin/bash
# round 2, call A: plane-staged sweep correctness on small grids, K5 timings, then the whole GPU suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 600 python -m pytest tests/test_gpu_plane.py -x -q ) > gpurun_out/r2a_plane_tests.log 2>&1; tail -15 gpurun_out/r2a_plane_tests.log
( time timeout 600 python scripts/exp_plane.py 20 ) > gpurun_out/r2a_exp_plane.log 2>&1; tail -30 gpurun_out/r2a_exp_plane.log
( time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_plane.py ) > gpurun_out/r2a_pytest.log 2>&1; tail -8 gpurun_out/r2a_pytest.log
