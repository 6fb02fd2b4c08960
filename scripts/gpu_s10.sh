#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/s10_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/s10_smoke.log
timeout 900 python bench.py > $OUT/s10_bench.json 2> $OUT/s10_bench.err; echo "bench rc=$?"; tail -c 6000 $OUT/s10_bench.json; tail -5 $OUT/s10_bench.err
