#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_fullsize_parity.py -x -q -k "slab" ) > gpurun_out/r2_par.log 2>&1; grep -E "passed|failed" gpurun_out/r2_par.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pi_build_rows|improve_kernel|compact_rows_kernel|plane_cells_kernel|plane_slots_kernel" -s 144 -c 8 --csv --log-file gpurun_out/r02_k5_kernels.csv python scripts/prof_plane.py 20 > /dev/null 2>&1
grep -c . gpurun_out/r02_k5_kernels.csv
