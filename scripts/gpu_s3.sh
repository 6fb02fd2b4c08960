#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
export KAGNN_LIB=kagnn_b200/lib/libkagnn_b200_dbg.so
: > $OUT/s3_agg.jsonl
for cfg in "0 0" "148 0" "148 65536" "148 131072" "148 204800" "148 230000" "296 0" "296 100000" "592 0" "592 50000" "0 204800"; do
  set -- $cfg
  if [ "$1" != "0" ]; then export KAGNN_DEBUG_AGG_GRID=$1; else unset KAGNN_DEBUG_AGG_GRID; fi
  if [ "$2" != "0" ]; then export KAGNN_DEBUG_AGG_SMEM=$2; else unset KAGNN_DEBUG_AGG_SMEM; fi
  timeout 120 python scripts/agg_probe.py >> $OUT/s3_agg.jsonl 2>> $OUT/s3_agg.err
done
cat $OUT/s3_agg.jsonl; tail -3 $OUT/s3_agg.err
