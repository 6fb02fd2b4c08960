"""Per-role summary of an ncu source-page dump of fused_tc2_kernel: samples / instructions / stall mix per warp role,
plus the hottest instructions (development aid).
usage: python scripts/ncu_roles.py <rep> <launch-index> """
import collections, csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[idx]
hi_end = starts[idx + 1] if idx + 1 < len(starts) else len(rows)
rows = rows[lo:hi_end]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
S = lambda r: int(r[col["# Samples"]] or 0)
I = lambda r: int(r[col["Instructions Executed"]] or 0)
tot_s, tot_i = sum(map(S, data)), sum(map(I, data))
print(f"kernel: {rows[0][1][:80]}  samples {tot_s} warp-instructions {tot_i/1e6:.1f}M  sass lines {len(data)}")
# hottest instructions
print("-- hottest instructions")
for i in sorted(range(len(data)), key=lambda i: -S(data[i]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
    r = data[i]
    top = sorted(((int(r[col[s]] or 0), s[6:]) for s in stall), reverse=True)[:2]
    prev = data[i - 1][col['Source']].strip()[:60] if "BRA" in r[col['Source']] and i else ""
    print(f"{i:6d} {r[col['Source']].strip()[:70]:70s} samp {S(r):5d} ({100*S(r)/tot_s:4.1f}%) inst {I(r):8d} {[t for t in top if t[0]]} {prev}")
if len(sys.argv) > 4: sys.exit(0)
print("-- blocks of 100 SASS lines")
for b in range(0, len(data), 100):
    blk = data[b:b + 100]
    s, n = sum(map(S, blk)), sum(map(I, blk))
    if s or n:
        c = collections.Counter()
        for r in blk:
            for st in stall:
                c[st[6:]] += int(r[col[st]] or 0)
        print(f"{b:6d} samp {100*s/tot_s:5.1f}% inst {n/1e6:7.2f}M {c.most_common(3)}")
