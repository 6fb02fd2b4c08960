#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/d12_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/d12_pytest.log | cut -c1-300
timeout 600 python bench.py > $OUT/d12_n1.json 2> $OUT/d12_n1.err; echo "n1 rc=$?"
python - <<PY
import json
j=json.loads(open("$OUT/d12_n1.json").read().strip().splitlines()[-1])
print("N=1 ms/step", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["ms_per_step"], [(k["label"],k["ms"]) for k in j["kernels"]])
for k,v in (j.get("extra") or {}).items():
    print(k, {a:(round(b,5) if isinstance(b,float) else b) for a,b in v.items() if a not in ("workload","roofline")})
PY
run() { # name, extra args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-extras $2 > $OUT/d12_n2_$1.json 2> $OUT/d12_n2_$1.err
  echo "$1 rc=$?"; grep -v "Warning\|symm_mem.enable" $OUT/d12_n2_$1.err | tail -2 | cut -c1-200; python -c "
import json,sys
j=json.loads(open('$OUT/d12_n2_$1.json').read().strip().splitlines()[-1]); print('$1', round(j['ms_per_step'],4), round(j['value']/1e6,1), 'e2e', round(j['e2e']['ms_per_step'],3), [(k['label'],k['ms']) for k in j.get('kernels',[])])"
}
run push "--dist-mode push"
run peer "--dist-mode peer"
