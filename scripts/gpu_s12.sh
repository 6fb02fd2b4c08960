#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc_fused.py tests/test_gpu_wide_bf16.py -m gpu -q > $OUT/s12_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/s12_pytest.log | cut -c1-200
KAGNN_LIB=kagnn_b200/lib/libkagnn_b200_u16.so timeout 300 python bench.py --config rmat > $OUT/s12_rmat_u16.json 2> $OUT/s12_rmat_u16.err; cut -c1-420 $OUT/s12_rmat_u16.json | cut -c150-
timeout 600 ncu --set full --clock-control none -k regex:"aggregate_only|fused_tc2" -s 8 -c 2 -f -o $OUT/s12_rmat_prof python bench.py --config rmat > $OUT/s12_prof.log 2>&1; echo "ncu rc=$?"
