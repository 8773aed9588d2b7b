#!/bin/bash
( time python -m pytest tests/test_gpu_xline.py tests/test_gpu_fullsize.py tests/test_compat.py tests/test_gpu_edge.py -m gpu -x -q ) > gpurun_out/d_pytest.log 2>&1; tail -6 gpurun_out/d_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; cat gpurun_out/d_bench.json; tail -3 gpurun_out/d_bench.err
