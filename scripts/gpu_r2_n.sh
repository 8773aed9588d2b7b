#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_plane.py tests/test_gpu_retrain.py -x -q ) > gpurun_out/r2n_plane_tests.log 2>&1; grep -E "passed|failed|Error" gpurun_out/r2n_plane_tests.log | tail -4
echo "== plane auto"; python scripts/exp_e2e.py 2>&1 | grep -v INFO | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"plane_cells_kernel|plane_slots_kernel" -c 6 --csv --log-file gpurun_out/r2n_plan.csv python scripts/prof_plane.py 20 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2n_plan.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:30], r[-3], r[-2], r[-1])
PY
