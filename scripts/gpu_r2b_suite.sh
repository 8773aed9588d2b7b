#!/bin/bash
# full GPU suite, smoke(), both bench arms (what the driver runs at round end on one GPU)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r2b_pytest.log 2>&1; grep -E "passed|failed|rror" gpurun_out/r2b_pytest.log | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py --impl reference ) > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; cut -c1-200 gpurun_out/r2b_bench_ref.json
( time python bench.py ) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; cut -c1-300 gpurun_out/r2b_bench.json; tail -2 gpurun_out/r2b_bench.err
