#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_windows.py tests/test_gpu_backward_tc.py tests/test_gpu_train_backward.py -m gpu -q -x 2>&1 | tail -25 | cut -c1-220
