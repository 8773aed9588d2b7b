#!/bin/bash
cd /root/repo
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
for b in auto 2 3 4 6; do
  echo "== BLOCKS_PER_SM=$b"
  if [ $b = auto ]; then unset DPB200_BLOCKS_PER_SM; else export DPB200_BLOCKS_PER_SM=$b; fi
  python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 50 2>&1 | tail -1
done
unset DPB200_BLOCKS_PER_SM
python scripts/prof_eval.py --env double_pendulum_swingup --bins 50 --sweeps 100 | tail -1
python scripts/prof_eval.py --env cartpole --bins 30 --sweeps 200 | tail -1
python scripts/prof_eval.py --env pendulum --bins 200 --sweeps 400 | tail -1
python scripts/prof_eval.py --env continuous_mountain_car --bins 400 --sweeps 400 | tail -1
