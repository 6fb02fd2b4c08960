#!/bin/bash
# gather4 microbenchmark + round-2 ncu evidence (launch list and full capture of one step's 4 launches)
OUT=gpurun_out; mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/tma_gather4 scripts/experiments/tma_gather4.cu -lcuda && \
  timeout 120 /tmp/tma_gather4 > $OUT/r2_tma_gather4.jsonl 2> $OUT/r2_tma_gather4.err; echo "gather4 rc=$?"; cat $OUT/r2_tma_gather4.jsonl | cut -c1-400; tail -3 $OUT/r2_tma_gather4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/r2_launches.log 2>&1 ; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_tc -s 16 -c 4 -f -o $OUT/r2_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/r2_prof.log 2>&1 ; echo "ncu rc=$?"
ls -la $OUT | tail -8
