#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/s4_probe.jsonl
for v in _nogpr _wq _wqnm; do
  KAGNN_LIB=kagnn_b200/lib/libkagnn_b200$v.so timeout 300 python scripts/layer_probe.py "x$v" >> $OUT/s4_probe.jsonl 2>> $OUT/s4_probe.err
done
cat $OUT/s4_probe.jsonl | cut -c1-100; tail -3 $OUT/s4_probe.err
