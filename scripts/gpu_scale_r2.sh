#!/bin/bash
# usage: gpurun --gpus N -- 'bash scripts/gpu_scale_r2.sh N "mode1 mode2 ..."'   (mode "push_res" = push + resident input halo)
N=${1:-8}; MODES=${2:-"auto"}
OUT=gpurun_out; mkdir -p $OUT
for m in $MODES; do
  extra="--dist-mode $m"; [ "$m" = "push_res" ] && extra="--dist-mode push --resident-x-halo"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N + RANDOM % 50)) \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras $extra 2> $OUT/r2_n${N}_$m.err > $OUT/r2_n${N}_$m.json
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/r2_n${N}_$m.json").read().strip().splitlines()[-1])
    print("N=$N $m", "ms/step", round(j["ms_per_step"],4), "value", round(j["value"]/1e6,1), "M nodes/s; e2e ms", round(j["e2e"]["ms_per_step"],3), [(k["label"],k["ms"]) for k in j["kernels"]], "nvlink GB/s", round(j["roofline"]["nvlink"]["achieved_gbs_over_step"],1))
except Exception as e:
    print("N=$N $m failed", e); print(open("$OUT/r2_n${N}_$m.err").read()[-1500:])
PY
done
