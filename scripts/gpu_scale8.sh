#!/bin/bash
TAG=${1:-sc8}
mkdir -p gpurun_out
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + $1 + RANDOM % 50)) \
      bench.py --gpus $1 --no-cpu-baseline --dist-mode $2 2> gpurun_out/${TAG}_n$1_$2.err > gpurun_out/${TAG}_n$1_$2.json
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/${TAG}_n$1_$2.json").read().strip().splitlines()[-1])
    print("N=$1 $2", "ms/step", round(j["ms_per_step"],4), "value", round(j["value"]/1e6,1), "M nodes/s; e2e ms", round(j["e2e"]["ms_per_step"],3), [(k["label"],k["ms"]) for k in j["kernels"]])
except Exception as e:
    print("N=$1 $2 failed", e); print(open("gpurun_out/${TAG}_n$1_$2.err").read()[-1500:])
PY
}
run 8 auto; run 8 pull
