#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/d5_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/d5_pytest.log | cut -c1-300
for sms in 8 16 24; do
  KAGNN_PULL_SMS=$sms timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --dist-mode pull > $OUT/d5_bench_pull$sms.json 2> $OUT/d5_bench_pull$sms.err
  python - <<PY
import json
try:
    j = json.loads(open("$OUT/d5_bench_pull$sms.json").read().strip().splitlines()[-1])
    print("pull sms=$sms ms/step", round(j["ms_per_step"], 4), "value", round(j["value"]/1e6,1), "M nodes/s  e2e ms", round(j["e2e"]["ms_per_step"], 3), [(k["label"], k["ms"]) for k in j["kernels"]])
except Exception as e:
    print("parse failed", e)
PY
done
