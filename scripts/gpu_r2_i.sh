#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2i_pytest.log 2>&1; tail -16 gpurun_out/r2i_pytest.log
( time python bench.py --impl reference --steps 5 ) > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench_ref.err; cut -c1-600 gpurun_out/r2i_bench_ref.json; tail -3 gpurun_out/r2i_bench_ref.err
( time python bench.py --steps 10 ) > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; cut -c1-1500 gpurun_out/r2i_bench.json; tail -5 gpurun_out/r2i_bench.err
