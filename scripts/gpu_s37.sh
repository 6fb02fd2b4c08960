#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_train_backward.py tests/test_gpu_train_epilogue.py tests/test_gpu_parity.py tests/test_gpu_gat.py -m gpu -q 2>&1 | tail -3 | cut -c1-200
timeout 300 python scripts/train_step_time.py --kernels > $OUT/s37_train.json 2> $OUT/s37_train_kernels.txt; cut -c1-330 $OUT/s37_train.json; grep " ms " $OUT/s37_train_kernels.txt | cut -c1-130 | head -12
