#!/bin/bash
# Scaling runs on one multi-GPU box: bench at N = 8 (pull / peer / halo), 4, 2 + the multi-GPU parity test.
# usage: gpurun --gpus 8 --timeout 1200 -- 'bash scripts/gpu_scale.sh <tag>'
TAG=${1:-sc}
mkdir -p gpurun_out
run() {  # n mode
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
nvidia-smi -L | wc -l
timeout 500 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -3
run 8 auto; run 8 peer; run 8 halo; run 4 auto; run 2 auto
