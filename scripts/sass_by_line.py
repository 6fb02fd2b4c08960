"""Attribute SASS instruction counts of a kernel in a .o/.cubin to source lines (development aid).
usage: python scripts/sass_by_line.py <obj> <kernel-substring> [source-file]"""
import collections, os, re, subprocess, sys, tempfile
obj, pat = sys.argv[1], sys.argv[2]
srcf = sys.argv[3] if len(sys.argv) > 3 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
cur = fn = None
cnt = collections.Counter()
for l in dis.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*\.section\s+(\.text\.\S+)", l)
    if m:
        fn = m.group(1)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l) and fn and pat in fn:
        cnt[cur] += 1
print("total", sum(cnt.values()))
src = open(srcf).read().split("\n") if srcf else None
byfile = collections.Counter()
for (f, ln), c in cnt.items():
    byfile[f] += c
print(byfile.most_common())
for (f, ln), c in sorted(cnt.items(), key=lambda x: -x[1])[:40]:
    print(f, ln, c, src[ln - 1][:100] if src and f == os.path.basename(srcf) else "")
