#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2q_multi.log 2>&1; tail -3 gpurun_out/r2q_multi.log
bash scripts/exp_xgpu.sh 2 "normal nobarrier"
