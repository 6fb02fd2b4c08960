#!/bin/bash
# 2-GPU session: distributed parity (all transports) + bench lines per transport
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_hostside.py -m gpu -q > $OUT/d2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/d2_pytest.log | cut -c1-300
for mode in auto pull peer halo; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --dist-mode $mode > $OUT/d2_bench_$mode.json 2> $OUT/d2_bench_$mode.err
  echo "mode $mode rc=$?"; python - <<PY
import json
try:
    j = json.loads(open("$OUT/d2_bench_$mode.json").read().strip().splitlines()[-1])
    print("$mode ms/step", round(j["ms_per_step"], 4), "value", round(j["value"]/1e6,1), "M nodes/s  e2e ms", round(j["e2e"]["ms_per_step"], 3), [(k["label"], k["ms"]) for k in j["kernels"]], j["roofline"].get("nvlink", {}).get("achieved_gbs_over_step"))
except Exception as e:
    print("parse failed", e)
PY
  tail -2 $OUT/d2_bench_$mode.err
done
