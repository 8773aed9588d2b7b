#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_retrain.py -x -q -s ) > gpurun_out/r2k_retrain.log 2>&1; grep -E "passed|failed|trial ready|Error|assert" gpurun_out/r2k_retrain.log | tail -8
for tool in memcheck racecheck synccheck; do
  for fam in aot gp_single gp_pair xline persistent plane plane_small lookup; do
    ( timeout 300 compute-sanitizer --tool $tool --error-exitcode 1 python scripts/sanitize_target.py $fam ) > gpurun_out/r2k_san_${tool}_$fam.log 2>&1
    echo "$tool $fam rc=$? $(grep -c 'ERROR SUMMARY: 0 errors\|RACECHECK SUMMARY: 0 hazards' gpurun_out/r2k_san_${tool}_$fam.log) $(grep -E 'SUMMARY' gpurun_out/r2k_san_${tool}_$fam.log | tail -1)"
  done
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"plane_plan|compact_rows|improve_kernel|pi_build_rows|count_reduce|eval_reduce|partial_" -c 40 --csv --log-file gpurun_out/r2k_launches.csv python scripts/prof_plane.py 20 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r2k_launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4].split("(")[0]].append(float(r[-1]))
for k,v in agg.items(): print(k, len(v), "launches, mean", sum(v)/len(v), r[-2] if rows else "")
PY
