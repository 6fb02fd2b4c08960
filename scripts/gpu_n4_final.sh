OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/n4f_bench.json 2> $OUT/n4f_bench.err
echo "bench rc=$?"; grep -v "Warning\|symm_mem.enable\|OMP_NUM\|\*\*\*" $OUT/n4f_bench.err | tail -3 | cut -c1-200
python - <<PY
import json
j=json.loads(open("$OUT/n4f_bench.json").read().strip().splitlines()[-1])
print("N=4 ms/step", round(j["ms_per_step"],4), "M nodes/s", round(j["value"]/1e6,1), "e2e", round(j["e2e"]["ms_per_step"],3), j["e2e"].get("spread_rank0"))
PY
