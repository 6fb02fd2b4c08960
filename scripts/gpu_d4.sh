#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_peer_single.py tests/test_gpu_dist.py -m gpu -q > $OUT/d4_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/d4_pytest.log | cut -c1-300
KAGNN_EXPERIMENTAL_GINE_BACKWARD=1 timeout 300 python -m pytest tests/test_gpu_train_backward.py -q > $OUT/d4_pytest_bwd.log 2>&1; echo "bwd rc=$?"; tail -3 $OUT/d4_pytest_bwd.log | cut -c1-200
for v in "" _s3; do KAGNN_LIB=kagnn_b200/lib/libkagnn_b200$v.so timeout 300 python scripts/layer_probe.py "p$v" 2>/dev/null | cut -c1-100; done
for sms in 8 16 24; do
  KAGNN_PULL_SMS=$sms timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --dist-mode pull > $OUT/d4_bench_pull$sms.json 2> $OUT/d4_bench_pull$sms.err
  python - <<PY
import json
try:
    j = json.loads(open("$OUT/d4_bench_pull$sms.json").read().strip().splitlines()[-1])
    print("pull sms=$sms ms/step", round(j["ms_per_step"], 4), "value", round(j["value"]/1e6,1), "M nodes/s  e2e ms", round(j["e2e"]["ms_per_step"], 3), [(k["label"], k["ms"]) for k in j["kernels"]])
except Exception as e:
    print("parse failed", e)
PY
done
