#!/bin/bash
# gradient kernels: parity tests (tensor-core vs fp32 vs autograd, the reference's gradient fixtures, GAT) + training-step timings
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_backward_tc.py tests/test_gpu_train_backward.py tests/test_gpu_gat.py -m gpu -q 2>&1 | tail -8 | cut -c1-250
timeout 300 python scripts/train_step_time.py --kernels > $OUT/s39_train.json 2> $OUT/s39_train_kernels.txt; cut -c1-420 $OUT/s39_train.json; grep " ms " $OUT/s39_train_kernels.txt | cut -c1-130 | head -4
timeout 300 python scripts/train_step_c5.py --kernels > $OUT/s39_c5.json 2> $OUT/s39_c5_kernels.txt; cut -c1-420 $OUT/s39_c5.json; grep " ms " $OUT/s39_c5_kernels.txt | cut -c1-130 | head -4
