#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/s9_probe.jsonl
for v in "" _nomma _nommanm _nomath; do
  KAGNN_LIB=kagnn_b200/lib/libkagnn_b200$v.so timeout 300 python scripts/layer_probe.py "w$v" >> $OUT/s9_probe.jsonl 2>> $OUT/s9_probe.err
done
cat $OUT/s9_probe.jsonl | cut -c1-100; tail -3 $OUT/s9_probe.err
