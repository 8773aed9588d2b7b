#!/bin/bash
mkdir -p gpurun_out
for pdl in off on; do
  echo "== DPB200_PDL=$pdl"
  DPB200_PDL=$pdl python scripts/exp_k2.py 2>&1 | cut -c1-330
done | tee gpurun_out/f_pdl.log
DPB200_PDL=on python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -3
