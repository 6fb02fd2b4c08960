#!/bin/bash
# pinned host <-> device copy bandwidth of this box, then the bench's end-to-end leg
OUT=gpurun_out; mkdir -p $OUT
python - <<PY
import torch, time
x = torch.empty(87*2**20//4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for name, f in (("h2d", lambda: d.copy_(x, non_blocking=True)), ("d2h", lambda: x.copy_(d, non_blocking=True))):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): f()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(name, "%.1f GB/s" % (x.numel() * 4 / dt / 1e9), "%.3f ms" % (dt * 1e3))
PY
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/e2e_probe.json 2> $OUT/e2e_probe.err
python - <<PY
import json
j=json.loads(open("$OUT/e2e_probe.json").read().strip().splitlines()[-1])
print("ms/step", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"])
PY
