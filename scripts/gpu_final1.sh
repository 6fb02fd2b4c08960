#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/f1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/f1_smoke.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q > $OUT/f1_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/f1_pytest.log | cut -c1-250
timeout 900 python bench.py > $OUT/f1_bench.json 2> $OUT/f1_bench.err; echo "bench rc=$?"; tail -2 $OUT/f1_bench.err
python - <<PY
import json
j=json.loads(open("$OUT/f1_bench.json").read().strip().splitlines()[-1])
print("ms/step", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["ms_per_step"], "frac", j["roofline"]["frac"], "cpu", j["cpu_baseline"]["value"])
print([(k["label"],k["ms"]) for k in j["kernels"]])
for k,v in (j.get("extra") or {}).items():
    print(k, {a:(round(b,5) if isinstance(b,float) else b) for a,b in v.items() if a not in ("workload","roofline")})
PY
