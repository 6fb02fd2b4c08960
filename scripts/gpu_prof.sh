#!/bin/bash
# ncu --set full of the fused kernels of one bench step (4 launches).  usage: gpurun --timeout 900 -- 'bash scripts/gpu_prof.sh <tag>'
TAG=${1:-p}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_tc -s 16 -c 4 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_prof.log 2>&1 ; echo "ncu rc=$?"
ls -la gpurun_out | tail -5
