#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8
for k in scalar pair; do
  for f in auto 1; do
    echo "== KERNEL=$k FAST_DIM=$f"
    DPB200_EVAL_KERNEL=$k DPB200_FAST_DIM=$f python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 50
  done
  DPB200_EVAL_KERNEL=$k python scripts/prof_eval.py --env double_pendulum_swingup --bins 50 --sweeps 100
  DPB200_EVAL_KERNEL=$k python scripts/prof_eval.py --env cartpole --bins 30 --sweeps 200
  DPB200_EVAL_KERNEL=$k python scripts/prof_eval.py --env pendulum --bins 200 --sweeps 400
done
