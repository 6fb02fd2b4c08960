#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/train_step_time.py --kernels > $OUT/s26_train.json 2> $OUT/s26_train_kernels.txt; echo "train rc=$?"; cat $OUT/s26_train.json | cut -c1-420; grep " ms " $OUT/s26_train_kernels.txt | cut -c1-150 | head -12
timeout 900 python -m pytest tests -m gpu -q > $OUT/s26_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/s26_pytest.log | cut -c1-250
