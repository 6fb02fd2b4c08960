"""SASS evidence for profiles/: per kernel of an object file, the count of the Blackwell-specific mnemonics (tcgen05 = UTC*MMA /
LDTM / STTM / UTCBAR, bulk TMA = UBLKCP, mbarrier = SYNCS, packed fp32 = FFMA2 / FMUL2 / FADD2, cp.async = LDGSTS) and the first
lines around the first tcgen05.mma of each kernel.   usage: python scripts/sass_extract.py <obj> [<obj> ...] > profiles/<name>.txt"""
import collections, re, subprocess, sys

KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "LDGSTS", "HMMA", "LDG.E.128",
        "STS.128", "LDS.128", "NANOSLEEP", "BAR.SYNC", "MUFU.EX2"]
for obj in sys.argv[1:]:
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    print(f"==== {obj}")
    fn, body = None, collections.defaultdict(list)
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            body[fn].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).strip())
    for fn, lines in body.items():
        demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        counts = {k: sum(1 for l in lines if re.search(r"\b" + re.escape(k), l)) for k in KEYS}
        print(f"-- {demangled}\n   instructions {len(lines)}  " + "  ".join(f"{k} {v}" for k, v in counts.items() if v))
        idx = next((i for i, l in enumerate(lines) if "UTCHMMA" in l), None)
        if idx is not None:
            print("   around the first tcgen05.mma:")
            for l in lines[max(0, idx - 6): idx + 10]:
                print("      " + l)
