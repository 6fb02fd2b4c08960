#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_backward_tc.py tests/test_gpu_train_backward.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-250
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kan_bwd -s 13 -c 4 -f -o $OUT/r2_bwd_prof python scripts/train_step_time.py > $OUT/r2_bwd_prof.log 2>&1; echo "ncu rc=$?"
