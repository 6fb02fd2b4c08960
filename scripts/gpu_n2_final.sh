#!/bin/bash
# 2-GPU sanity of HEAD: the multi-GPU tests, then the bench the way the driver launches it (default mode) and the reference arm's N>1 behaviour
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_peer_single.py tests/test_gpu_hostside.py -m gpu -q > $OUT/n2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/n2f_pytest.log | cut -c1-250
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/n2f_bench.json 2> $OUT/n2f_bench.err
echo "bench rc=$?"; grep -v "Warning\|symm_mem.enable" $OUT/n2f_bench.err | tail -2 | cut -c1-200
python - <<PY
import json
j=json.loads(open("$OUT/n2f_bench.json").read().strip().splitlines()[-1])
print("N=2", j["config"].get("dist_mode"), "ms/step", round(j["ms_per_step"],4), "M nodes/s", round(j["value"]/1e6,1), "e2e", j["e2e"]["ms_per_step"], j["e2e"].get("spread_rank0"), "nvlink", j["roofline"].get("nvlink"))
PY
