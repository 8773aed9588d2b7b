#!/bin/bash
# session 4, call A: full GPU suite, bench (both arms), K5 run profile
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1; tail -5 gpurun_out/a_pytest.log
python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; cat gpurun_out/a_bench.json
python bench.py --impl reference --steps 5 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; cat gpurun_out/a_bench_ref.json
timeout 400 python scripts/k5_run.py --cap-s 240 --out gpurun_out/a_k5_run.json > gpurun_out/a_k5_run.log 2>&1; tail -40 gpurun_out/a_k5_run.log
