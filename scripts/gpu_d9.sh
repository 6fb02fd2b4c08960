#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for m in push pull; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 scripts/dist_timeline.py $m > $OUT/d9_timeline_$m.txt 2> $OUT/d9_timeline_$m.err; echo "rc=$?"; tail -3 $OUT/d9_timeline_$m.err | cut -c1-300
grep -v Warning $OUT/d9_timeline_$m.txt | cut -c1-170 | head -60
done
