#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/s8_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/s8_pytest.log | cut -c1-200
: > $OUT/s8_probe.jsonl
for v in "" _nox2 _cpr1; do
  KAGNN_LIB=kagnn_b200/lib/libkagnn_b200$v.so timeout 300 python scripts/layer_probe.py "z$v" >> $OUT/s8_probe.jsonl 2>> $OUT/s8_probe.err
done
cat $OUT/s8_probe.jsonl | cut -c1-100; tail -3 $OUT/s8_probe.err
