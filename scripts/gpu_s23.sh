#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_peer_single.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-250
timeout 300 python scripts/layer_probe.py r 2>&1 | cut -c1-110 | tail -8
timeout 900 python -m pytest tests -m gpu -q > $OUT/s23_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/s23_pytest.log | cut -c1-250
timeout 600 python bench.py --no-extras > $OUT/s23_bench.json 2> $OUT/s23_bench.err; echo "bench rc=$?"
python - <<PY
import json
j=json.loads(open("$OUT/s23_bench.json").read().strip().splitlines()[-1])
print("ms/step", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["ms_per_step"])
print([(k["label"],k["ms"]) for k in j["kernels"]])
PY
