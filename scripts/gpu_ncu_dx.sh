#!/bin/bash
# ncu --set full (with source-level sampling) of the dX kernel: the first step's first three launches
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bwd_input -c 3 -f -o $OUT/r2_dx_prof \
    python scripts/train_step_time.py > $OUT/r2_dx_prof.log 2>&1 ; echo "ncu rc=$?"
