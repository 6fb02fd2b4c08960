#!/bin/bash
# quick GPU check: parity tests + bench line (no ncu).  usage: gpurun --timeout 900 -- 'bash scripts/gpu_quick.sh <tag>'
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("ms/step", round(j["ms_per_step"], 4), "e2e ms", round(j["e2e"]["ms_per_step"], 3), [(k["label"], k["ms"]) for k in j["kernels"]])
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/${TAG}_bench.err
