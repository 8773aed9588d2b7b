#!/bin/bash
# what the driver runs at round end, plus both bench arms
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -25
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_ours.json | cut -c1-1500
tail -3 gpurun_out/bench_err.log
python bench.py --impl reference --steps 5 --warmup 3 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json | cut -c1-800
