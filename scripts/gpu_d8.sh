#!/bin/bash
# 2 GPUs: push transport -- single-GPU simulated test, distributed parity, bench at N=1 (regression check) and N=2 (pull vs push)
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_peer_single.py tests/test_gpu_dist.py -m gpu -q > $OUT/d8_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/d8_pytest.log | cut -c1-300
timeout 600 python bench.py --no-extras --no-cpu-baseline > $OUT/d8_n1.json 2> $OUT/d8_n1.err; echo "n1 rc=$?"
python -c "
import json
j=json.loads(open('$OUT/d8_n1.json').read().strip().splitlines()[-1]); print('N=1', j['ms_per_step'], j['value'], [(k['label'],k['ms']) for k in j['kernels']])"
for m in pull push; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-extras --dist-mode $m > $OUT/d8_n2_$m.json 2> $OUT/d8_n2_$m.err
  echo "$m rc=$?"; tail -2 $OUT/d8_n2_$m.err; python -c "
import json,sys
j=json.loads(open('$OUT/d8_n2_$m.json').read().strip().splitlines()[-1]); print('$m', j['ms_per_step'], j['value'], j['e2e']['ms_per_step'], [(k['label'],k['ms']) for k in j.get('kernels',[])])"
done
