#!/bin/bash
# dW kernel A/B: parity tests, then the training-step kernel table with 8 and 4 feature blocks per CTA
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_backward_tc.py tests/test_gpu_train_backward.py -m gpu -q -x 2>&1 | tail -8 | cut -c1-250
for mb in 8 4; do
  echo "== KAGNN_DW_MBLOCKS=$mb"
  KAGNN_DW_MBLOCKS=$mb timeout 300 python scripts/train_step_time.py --kernels > $OUT/dw_mb$mb.json 2> $OUT/dw_mb$mb.txt; cut -c1-300 $OUT/dw_mb$mb.json; grep " ms " $OUT/dw_mb$mb.txt | cut -c1-130 | head -5
done
