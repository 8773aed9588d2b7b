#!/bin/bash
mkdir -p gpurun_out
for x in "" fakepeers; do
  for pl in force off; do
    echo "== XDEBUG=$x PLANE=$pl"
    DPB200_XDEBUG=$x DPB200_PLANE=$pl DPB200_XLINE=off timeout 300 python - <<'PY' 2>&1 | grep -v INFO | tail -2
import os, sys
sys.path.insert(0, ".")
from dynamicprogramming_b200 import envs
eng = envs.make("double_cartpole_swingup", bins=20)
eng.build_table(); eng.sweeps(50); eng.policy_improvement()
eng.sweeps(25)
best = min(eng.sweeps(25)[1] / 25 for _ in range(4))
print("ms/sweep %.4f" % best, eng.eval_kernel_info()["kernel"][:30])
PY
  done
done
DPB200_XDEBUG=fakepeers timeout 600 ncu --set full --clock-control none -k regex:ps_sweep -s 56 -c 1 -o gpurun_out/r2t_ps_fake python scripts/prof_plane.py 20 "0,0,2,2,0" > gpurun_out/r2t_prof.log 2>&1; tail -1 gpurun_out/r2t_prof.log
