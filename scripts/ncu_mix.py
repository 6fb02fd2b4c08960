"""Dynamic instruction mix of fused_tc2_kernel by category (math / sync overhead / spin / gather / MMA ...), from an ncu
source-page dump joined with nvdisasm line info (development aid).
usage: python scripts/ncu_mix.py <rep> <kernel-block-index> <lib.so or .o> <kernel-substring>"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, idx, obj, pat = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
rows = rows[starts[idx]:(starts[idx + 1] if idx + 1 < len(starts) else len(rows))]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
base = int(data[0][col["Address"]], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur = fn = None
        for l in dis.split("\n"):
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*\.section\s+(\.text\.\S+)", l)
            if m:
                fn = m.group(1)
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
            if m and fn and pat in fn:
                line_of[int(m.group(1), 16)] = cur
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kagnn_b200", "csrc", "fused_tc2.cu")).read().split("\n")
def find(s):
    return next(i + 1 for i, l in enumerate(src) if s in l)
L_GATHER_FN0, L_KERNEL = find("__device__ __forceinline__ void ldp4("), find("fused_tc2_kernel(const __grid_constant__")
L_PROD, L_GATHER, L_MMA, L_LOAD = (find(s) for s in ("BASIS PRODUCERS / EPILOGUE", "= GATHER =", "MMA ISSUER", "W LOADER"))
L_EPI0, L_EPI1 = find("auto epilogue = [&]"), find("long long pend_row0")      # the epilogue lambda sits inside the producer section
cat = collections.Counter()
samp = collections.Counter()
role = "setup"
for r in data:
    off = int(r[col["Address"]], 16) - base
    f, ln = line_of.get(off) or ("?", 0)
    n = int(r[col["Instructions Executed"]] or 0)
    s_ = int(r[col["# Samples"]] or 0)
    text = r[col["Source"]]
    if f == "fused_tc2.cu":
        if ln >= L_LOAD: role = "loader"
        elif ln >= L_MMA: role = "mma"
        elif ln >= L_GATHER: role = "gather"
        elif L_EPI0 <= ln < L_EPI1: role = "epilogue"
        elif ln >= L_PROD: role = "producer"
        elif ln >= L_KERNEL: role = "setup"
        elif ln >= L_GATHER_FN0: role = "gather"
        elif ln > 100 and role not in ("producer", "epilogue"): role = "producer"
    spin = f == "tc_common.cuh" and ("TRYWAIT" in text or "YIELD" in text or ("BRA" in text) or "VIADD" in text or "ISETP.GE.U32" in text or "NANOSLEEP" in text)
    sub = "spin" if spin else ("math" if (f == "fused_tc2.cu" and ln < L_GATHER_FN0 and ln > 100) else "other")
    cat[(role, sub)] += n
    samp[(role, sub)] += s_
tot, tots = sum(cat.values()), sum(samp.values())
for k, v in sorted(cat.items(), key=lambda kv: -kv[1]):
    print(f"{k[0]:10s} {k[1]:6s} inst {v/1e6:8.2f}M {100*v/tot:5.1f}%   samples {100*samp[k]/tots:5.1f}%")
