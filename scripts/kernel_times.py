"""Print the per-launch times of the bench model's forward (development aid)."""
import os, sys, json, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu-baseline"],
                     capture_output=True, text=True)
try:
    j = json.loads(out.stdout.strip().splitlines()[-1])
    print(os.environ.get("KAGNN_LIB", "default"), "ms/step", round(j["ms_per_step"], 4), [(k["label"], k["ms"]) for k in j["kernels"]])
except Exception as e:
    print("bench failed", e, out.stdout[-500:], out.stderr[-1500:])
