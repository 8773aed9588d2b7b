#!/bin/bash
( python scripts/exp_variants.py "DPB200_PAIR=force:128,8,1,8,1" "DPB200_PAIR=force:128,8,1,16,1" "DPB200_PAIR=force:64,16,1,16,1" "DPB200_PAIR=force:256,4,1,16,1" "DPB200_PAIR=force:128,8,1,32,1" "DPB200_PAIR=force:128,8,1,4,1" "DPB200_PAIR=force:96,10,1,8,1" "DPB200_PAIR=force:192,5,1,8,1"
EXP_ENV=double_pendulum_swingup EXP_BINS=50 python scripts/exp_variants.py "DPB200_EVAL_VARIANT=0" "DPB200_PAIR=force:128,8,1,8,1" "DPB200_PAIR=force:128,8,1,4,1" "DPB200_PAIR=force:128,8,1,0,1" "DPB200_PAIR=force:256,4,1,8,1" "DPB200_PAIR=force:128,12,1,8,1" "DPB200_PAIR=force:128,16,1,8,1"
EXP_ENV=cartpole_swingup EXP_BINS=50 python scripts/exp_variants.py "DPB200_EVAL_VARIANT=0" "DPB200_PAIR=force:128,8,1,8,1" "DPB200_PAIR=force:128,8,1,0,1" "DPB200_PAIR=force:128,12,1,8,1" "DPB200_PAIR=force:128,16,1,0,1"
EXP_ENV=double_cartpole EXP_BINS=15 python scripts/exp_variants.py "DPB200_EVAL_VARIANT=0" "DPB200_EVAL_VARIANT=64" "DPB200_PAIR=force:128,8,1,8,1" "DPB200_PAIR=force:128,8,1,16,1"
) 2>&1 | tee gpurun_out/c_variants10.log
