#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/train_step_time.py --kernels > $OUT/s24_train.json 2> $OUT/s24_train_kernels.txt; echo "train rc=$?"; cat $OUT/s24_train.json | cut -c1-400; grep " ms " $OUT/s24_train_kernels.txt | cut -c1-160
