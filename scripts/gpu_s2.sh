#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/s2_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/s2_pytest.log
KAGNN_EXPERIMENTAL_GINE_BACKWARD=1 timeout 300 python -m pytest tests/test_gpu_train_backward.py -q > $OUT/s2_pytest_bwd.log 2>&1; echo "bwd rc=$?"; tail -4 $OUT/s2_pytest_bwd.log
: > $OUT/s2_probe.jsonl
for v in "" _nogpr _nomath; do
  KAGNN_LIB=kagnn_b200/lib/libkagnn_b200$v.so timeout 300 python scripts/layer_probe.py "gpr$v" >> $OUT/s2_probe.jsonl 2>> $OUT/s2_probe.err
done
cat $OUT/s2_probe.jsonl | cut -c1-120
tail -3 $OUT/s2_probe.err
KAGNN_LIB=kagnn_b200/lib/libkagnn_b200_trace.so timeout 200 python scripts/trace_tc2.py layer0 > $OUT/s2_trace_layer0.txt 2>&1
sed -n 1,14p $OUT/s2_trace_layer0.txt; grep -A12 "tile-level" $OUT/s2_trace_layer0.txt
