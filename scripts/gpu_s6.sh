#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/s6_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/s6_pytest.log | cut -c1-200
: > $OUT/s6_probe.jsonl
for v in "" _cpr1 _cpr2ea; do
  KAGNN_LIB=kagnn_b200/lib/libkagnn_b200$v.so timeout 300 python scripts/layer_probe.py "y$v" >> $OUT/s6_probe.jsonl 2>> $OUT/s6_probe.err
done
cat $OUT/s6_probe.jsonl | cut -c1-100; tail -3 $OUT/s6_probe.err
