#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_compat.py tests/test_gpu_lookup.py -m gpu -x -q ) > gpurun_out/e_pytest.log 2>&1; tail -6 gpurun_out/e_pytest.log
python scripts/exp_k2.py 2>&1 | tee gpurun_out/e_k2.log | cut -c1-420
DPB200_PERSIST=off python scripts/exp_k2.py 2>&1 | tee gpurun_out/e_k2_off.log | cut -c1-260
