#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python scripts/train_step_c5.py --kernels > $OUT/r2_train_step_c5.json 2> $OUT/r2_train_step_c5_kernels.txt; echo "rc=$?"; cat $OUT/r2_train_step_c5.json | cut -c1-500; grep " ms " $OUT/r2_train_step_c5_kernels.txt | cut -c1-150
