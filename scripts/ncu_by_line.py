"""Join an `ncu --page source --csv` dump (SASS view) with nvdisasm line info: per source line, the executed warp
instructions, stall samples and the dominant stall reasons (development aid).
usage: python scripts/ncu_by_line.py <src.csv> <lib.so> <kernel-substring> <source-file> [top]"""
import collections, csv, os, re, subprocess, sys, tempfile

csvf, obj, pat, srcf = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 45
rows = list(csv.reader(open(csvf)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break          # only the first kernel instance of the dump
    if len(r) >= len(hdr):
        data.append(r)
base = int(data[0][col["Address"]], 16)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
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
if not line_of:
    sys.exit("kernel not found in " + obj)

agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_inst = tot_samp = 0
for r in data:
    off = int(r[col["Address"]], 16) - base
    key = line_of.get(off, ("?", 0))
    inst = int(r[col["Instructions Executed"]] or 0)
    samp = int(r[col["# Samples"]] or 0)
    a = agg[key]
    a[0] += inst
    a[1] += samp
    for s in stall_cols:
        v = int(r[col[s]] or 0)
        if v:
            a[2][s[6:]] += v
    tot_inst += inst
    tot_samp += samp
src = open(srcf).read().split("\n")
print(f"total warp instructions {tot_inst}, samples {tot_samp}")
print("---- by samples")
for key, (inst, samp, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    f, ln = key
    text = src[ln - 1].strip()[:70] if f == os.path.basename(srcf) and ln > 0 else ""
    print(f"{f}:{ln:4d} inst {100*inst/tot_inst:5.1f}% samp {100*samp/tot_samp:5.1f}%  {dict(st.most_common(3))}  | {text}")
print("---- by instructions")
for key, (inst, samp, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, ln = key
    text = src[ln - 1].strip()[:70] if f == os.path.basename(srcf) and ln > 0 else ""
    print(f"{f}:{ln:4d} inst {100*inst/tot_inst:5.1f}% samp {100*samp/tot_samp:5.1f}%  | {text}")
