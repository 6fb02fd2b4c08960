#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in "" _g16; do KAGNN_LIB=kagnn_b200/lib/libkagnn_b200$v.so timeout 300 python scripts/layer_probe.py "r$v" 2>&1 | cut -c1-100 | tail -8; done
KAGNN_LIB=kagnn_b200/lib/libkagnn_b200_g16.so timeout 600 python -m pytest tests/test_gpu_tc_fused.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-200
