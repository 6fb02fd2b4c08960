#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list and one full capture of the dominant kernel.
# Usage (from the build container):  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
# Everything lands in gpurun_out/<tag>_*; nothing printed under ncu is a bench value.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -5 $OUT/${TAG}_pytest.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 $OUT/${TAG}_smoke.log
echo "== bench" ; timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ; echo "bench rc=$?" ; tail -c 3000 $OUT/${TAG}_bench.json ; tail -5 $OUT/${TAG}_bench.err
echo "== launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1 ; echo "launches rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_tc -s 16 -c 4 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_prof.log 2>&1 ; echo "ncu rc=$?"
ls -la $OUT | tail -20
