#!/bin/bash
# round-end rehearsal: full GPU suite, smoke, both bench arms
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/z_pytest.log 2>&1; tail -5 gpurun_out/z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 5 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err; cut -c1-250 gpurun_out/z_bench_ref.json; grep -o '"time_to_converge".*' gpurun_out/z_bench_ref.json | cut -c1-300
python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; cat gpurun_out/z_bench.json
