#!/bin/bash
for k in scalar pair; do
  for imp in 0 1 3; do
    echo "== KERNEL=$k improve=$imp"
    DPB200_EVAL_KERNEL=$k python scripts/prof_eval.py --env double_cartpole_swingup --bins 20 --sweeps 50 --improve $imp 2>&1 | tail -1
  done
done
