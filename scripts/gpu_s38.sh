#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/train_step_time.py > $OUT/s38_train.json 2> $OUT/s38_train.err; cut -c1-700 $OUT/s38_train.json; tail -3 $OUT/s38_train.err
