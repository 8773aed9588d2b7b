#!/bin/bash
# usage: gpu_exp_prof.sh "<cfg>" name [env] [bins]
mkdir -p gpurun_out
CFG=${1:-"2,0,4,8,2:1,1,1,4,5"}; NAME=${2:-xl}; ENVN=${3:-double_cartpole_swingup}; BINS=${4:-20}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xl_sweep -s 2 -c 1 -f -o gpurun_out/$NAME \
  python scripts/exp_xline.py --env $ENVN --bins $BINS --iters 2 --warm-sweeps 5 --configs "$CFG" > gpurun_out/$NAME.log 2>&1
tail -2 gpurun_out/$NAME.log
