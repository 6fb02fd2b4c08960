#!/bin/bash
# 2 GPUs: NVLink push / pull microbenchmark + the bench at N=2 in every transport mode (kernels of this commit)
OUT=gpurun_out; mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/nvlink_push scripts/experiments/nvlink_push.cu && \
  timeout 240 /tmp/nvlink_push > $OUT/r2_nvlink_push.jsonl 2> $OUT/r2_nvlink_push.err; echo "push rc=$?"; tail -3 $OUT/r2_nvlink_push.err
grep -E "148|copy_engine|check" $OUT/r2_nvlink_push.jsonl | cut -c1-200
for m in pull peer halo; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-extras --dist-mode $m > $OUT/d7_n2_$m.json 2> $OUT/d7_n2_$m.err
  echo "$m rc=$?"; python -c "
import json,sys
j=json.loads(open('$OUT/d7_n2_$m.json').read().strip().splitlines()[-1]); print('$m', j['ms_per_step'], j['value'], j['e2e']['ms_per_step'], [(k['label'],k['ms']) for k in j.get('kernels',[])])"
done
