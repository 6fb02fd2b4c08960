#!/bin/bash
# session 1 of round 2: GPU suite (GINE backward enabled), layer probe over kernel variants, light trace
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/s1_smi.txt 2>&1
KAGNN_EXPERIMENTAL_GINE_BACKWARD=1 timeout 600 python -m pytest tests -m gpu -x -q > $OUT/s1_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/s1_pytest.log
: > $OUT/s1_probe.jsonl
for v in "" _ea _sl400 _sl20 _easl _u8 _nomath; do
  KAGNN_LIB=kagnn_b200/lib/libkagnn_b200$v.so timeout 300 python scripts/layer_probe.py "base$v" >> $OUT/s1_probe.jsonl 2>> $OUT/s1_probe.err
done
cat $OUT/s1_probe.jsonl
KAGNN_LIB=kagnn_b200/lib/libkagnn_b200_trace.so timeout 200 python scripts/trace_tc2.py layer0 > $OUT/s1_trace_layer0.txt 2>&1
head -50 $OUT/s1_trace_layer0.txt
