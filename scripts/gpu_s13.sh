#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in "" _mb4 _mb5u4 _mb6u4; do
KAGNN_LIB=kagnn_b200/lib/libkagnn_b200$v.so timeout 300 python bench.py --config rmat > $OUT/s13_rmat$v.json 2> $OUT/s13_rmat$v.err; echo "$v"; cut -c260-420 $OUT/s13_rmat$v.json
KAGNN_LIB=kagnn_b200/lib/libkagnn_b200$v.so timeout 300 python scripts/layer_probe.py "a$v" 2>/dev/null | grep agg | cut -c1-100
done
