#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 scripts/pull_probe.py > $OUT/d6_pull.jsonl 2> $OUT/d6_pull.err; cat $OUT/d6_pull.jsonl; tail -3 $OUT/d6_pull.err
