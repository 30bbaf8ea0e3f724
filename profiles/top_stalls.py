import csv, sys, subprocess, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# find header
hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampling" in c or "stall" in c.lower() for c in r))
h = rows[hi]
print("COLUMNS:", [c for c in h][:60], file=sys.stderr)
src = h.index("Source")
def col(name):
    for i, c in enumerate(h):
        if c == name: return i
    return None
cands = [c for c in h if "long_sb" in c or "long_scoreboard" in c or "short_sb" in c or "short_scoreboard" in c or c in ("# Samples", "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Cycles)")]
print("CANDS:", cands, file=sys.stderr)
tot = col("Warp Stall Sampling (All Samples)") or col("# Samples")
lsb = next((i for i, c in enumerate(h) if "stall_long_sb" == c or c.endswith("long_sb")), None)
ssb = next((i for i, c in enumerate(h) if "stall_short_sb" == c or c.endswith("short_sb")), None)
data = []
for r in rows[hi + 1:]:
    if len(r) <= src: continue
    def f(i):
        try: return float(r[i].replace(",", "")) if i is not None and r[i] != "" else 0.0
        except: return 0.0
    data.append((f(lsb), f(ssb), f(tot), r[src][:90]))
S = sum(d[2] for d in data) or 1
print("total samples", S)
print("== top by long scoreboard")
for d in sorted(data, key=lambda d: -d[0])[:18]: print(f"{d[0]:9.0f} {d[1]:9.0f} {d[2]:9.0f}  {d[3]}")
print("== top by all samples")
for d in sorted(data, key=lambda d: -d[2])[:18]: print(f"{d[0]:9.0f} {d[1]:9.0f} {d[2]:9.0f}  {d[3]}")

import re
agg = collections.Counter()
for d in data:
    t = re.sub(r"^@!?U?P\d+\s+", "", d[3].strip())
    op = t.split()[0] if t else "?"
    base = op.split(".")[0]
    if base == "IMAD" and ".WIDE" in op: base = "IMAD.WIDE"
    agg[base] += d[2]
print("== samples by opcode")
for k, v in agg.most_common(14): print(f"{v:10.0f} {100 * v / S:5.1f} %  {k}")
