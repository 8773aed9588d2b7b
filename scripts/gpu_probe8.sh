#!/bin/bash
cd /root/repo
for la in 0 148 592 1184 2368 4736; do
  echo "== LOOKAHEAD=$la"
  DPB200_LOOKAHEAD=$la python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 50 2>&1 | tail -1
done
for la in 0 auto; do
  if [ $la = auto ]; then unset DPB200_LOOKAHEAD; else export DPB200_LOOKAHEAD=$la; fi
  echo "== LOOKAHEAD=$la"
  python scripts/prof_eval.py --env double_pendulum_swingup --bins 50 --sweeps 100 | tail -1
  python scripts/prof_eval.py --env cartpole --bins 30 --sweeps 200 | tail -1
done
unset DPB200_LOOKAHEAD
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
