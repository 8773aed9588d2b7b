#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_xline.py -x -q 2>&1 | tail -15
python -m pytest tests/ -x -q -m gpu --deselect tests/test_gpu_xline.py 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_ours.json | cut -c1-1800
tail -3 gpurun_out/bench_err.log
