#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8
for f in ref auto 0 1; do
  echo "== FAST_DIM=$f"
  DPB200_FAST_DIM=$f python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 50
done
for f in ref auto; do
  echo "== FAST_DIM=$f"
  DPB200_FAST_DIM=$f python scripts/prof_eval.py --env double_pendulum_swingup --bins 50 --sweeps 100
  DPB200_FAST_DIM=$f python scripts/prof_eval.py --env cartpole --bins 30 --sweeps 200
done
