#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/s21_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/s21_pytest.log | cut -c1-250
timeout 900 python bench.py > $OUT/s21_bench.json 2> $OUT/s21_bench.err; echo "bench rc=$?"; tail -3 $OUT/s21_bench.err
python - <<PY
import json
j=json.loads(open("$OUT/s21_bench.json").read().strip().splitlines()[-1])
print("ms/step", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["ms_per_step"])
print([(k["label"],k["ms"]) for k in j["kernels"]])
print("roofline frac", j["roofline"]["frac"], j["roofline"]["kernel"], j["roofline"]["launch_ms"])
for k,v in (j.get("extra") or {}).items():
    print(k, {a:(round(b,5) if isinstance(b,float) else b) for a,b in v.items() if a not in ("workload","roofline")})
PY
