#!/bin/bash
mkdir -p gpurun_out
echo "== plane auto"; python scripts/exp_e2e.py 2>&1 | grep -v INFO | tail -5
echo "== plane off"; DPB200_PLANE=off python scripts/exp_e2e.py 2>&1 | grep -v INFO | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:plane_plan_kernel -c 4 --csv --log-file gpurun_out/r2m_plan.csv python scripts/prof_plane.py 20 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2m_plan.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:30], r[-3], r[-2], r[-1])
PY
