#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_backward_tc.py -m gpu -q -x 2>&1 | tail -8 | cut -c1-220
timeout 300 python scripts/train_step_time.py --kernels > $OUT/s27_train.json 2> $OUT/s27_train_kernels.txt; echo "train rc=$?"; cat $OUT/s27_train.json | cut -c1-420; grep " ms " $OUT/s27_train_kernels.txt | cut -c1-150 | head -8
