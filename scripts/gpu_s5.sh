#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
export KAGNN_LIB=kagnn_b200/lib/libkagnn_b200_nogpr.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_tc2 -s 6 -c 3 -f -o $OUT/s5_prof python scripts/probe_once.py 3 > $OUT/s5_prof.log 2>&1; echo "ncu rc=$?"
tail -3 $OUT/s5_prof.log; ls -la $OUT/s5_prof.ncu-rep
