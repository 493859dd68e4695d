"""Top stall locations of one kernel from an ncu report captured with --import-source on:
    ncu -i rep --page source --csv --kernel-name regex:<k> --launch-skip N --launch-count 1 > src.csv
    python tools/ncu_stalls.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[h]
idx = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
S = idx["# Samples"]
tot = sum(int(r[S]) for r in data) or 1
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
print(rows[0][1] if len(rows[0]) > 1 else "", "samples", tot, "instructions", len(data))
for i, r in enumerate(data):
    r.append(i)
for r in sorted(sorted(data, key=lambda r: -int(r[S]))[:top_n], key=lambda r: r[-1]):
    n = int(r[S])
    st = sorted(((int(r[idx[c]] or 0), c.replace("stall_", "")) for c in stall_cols), reverse=True)[:2]
    print(f"{r[-1]:5d} {n:6d} {100 * n / tot:5.1f}%  {r[idx['Source']].strip()[:60]:60s} {st}")
