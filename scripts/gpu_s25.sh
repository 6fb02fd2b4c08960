#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for sw in 0 1; do
echo "== swap $sw"; KAGNN_LIB=kagnn_b200/lib/libkagnn_b200_dbg.so KAGNN_DEBUG_DW_SWAP=$sw timeout 300 python -m pytest tests/test_gpu_backward_tc.py -m gpu -q -x 2>&1 | tail -12 | cut -c1-220
done
