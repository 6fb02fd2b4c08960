#!/bin/bash
# 8 GPUs: the scaling line the driver will run (auto = push), the graph-sharded lines of C5 / C3, push with the input halo resident
OUT=gpurun_out; mkdir -p $OUT
run() { # name, args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 50)) \
      bench.py --gpus 8 $2 2> $OUT/f8_$1.err > $OUT/f8_$1.json
  echo "$1 rc=$?"; python - <<PY
import json
try:
    j=json.loads(open("$OUT/f8_$1.json").read().strip().splitlines()[-1])
    print("$1", "ms/step", round(j["ms_per_step"],4), "value", round(j["value"]/1e6,1), "M nodes/s; e2e", j.get("e2e") and round(j["e2e"].get("ms_per_step", 0),3), [(k["label"],k["ms"]) for k in j.get("kernels",[])])
except Exception as e:
    print("$1 failed", e); print(open("$OUT/f8_$1.err").read()[-1500:])
PY
}
run auto "--steps 20 --warmup 5 --no-cpu-baseline --no-extras"
run mutag "--config mutag --no-cpu-baseline"
run zinc "--config zinc --no-cpu-baseline"
