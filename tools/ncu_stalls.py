#!/usr/bin/env python
"""Top stall sites per launch from `ncu -i rep --page source --csv` output (one section per launch).
    python tools/ncu_stalls.py src.csv [launch index] [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = heads[which]
end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
print("launches in file:", len(heads), "| kernel:", rows[hi - 1][1][:90])
h = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(h)]
iS, isrc = h.index("# Samples"), h.index("Source")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for i in stall_cols:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print("by reason:", sorted(agg.items(), key=lambda x: -x[1])[:8])
for r in sorted(data, key=lambda r: -int(r[iS]))[:N]:
    st = sorted(((h[i], int(r[i])) for i in stall_cols if int(r[i]) > 0), key=lambda x: -x[1])[:3]
    print(f"{int(r[iS]):6d} {100 * int(r[iS]) / tot:5.1f}% {r[0][-5:]} {r[isrc].strip()[:64]:64s} {st}")
