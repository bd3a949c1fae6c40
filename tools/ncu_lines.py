#!/usr/bin/env python
"""Stall samples per SOURCE LINE: pairs the i-th SASS row of an `ncu --page source --csv` section with the i-th instruction
of `nvdisasm -g -c` of the same kernel (cubin extracted with `cuobjdump -xelf all obj.o`).
    python tools/ncu_lines.py src.csv <section index> disasm.txt <mangled-name substring> [N]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2])
dis = open(sys.argv[3]).read().split("\n")
key = sys.argv[4]
N = int(sys.argv[5]) if len(sys.argv) > 5 else 40
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = heads[which]
end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
h = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(h)]
iS = h.index("# Samples")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
# instructions of the function in the disassembly, each with the most recent //## File ..., line N marker
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and key in l and l.rstrip().endswith(":"))
insts = []
line = None
for l in dis[start + 1:]:
    if l.startswith("//---------------------") or l.startswith("\t.section"):
        if insts:
            break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        insts.append(line)
print("ncu rows", len(data), "disasm instructions", len(insts))
agg = {}
for r, ln in zip(data, insts):
    a = agg.setdefault(ln, [0, {}])
    a[0] += int(r[iS])
    for i in stall_cols:
        v = int(r[i])
        if v:
            a[1][h[i]] = a[1].get(h[i], 0) + v
tot = sum(a[0] for a in agg.values())
src_cache = {}
for ln, (n, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:N]:
    text = ""
    if ln:
        try:
            if ln[0] not in src_cache:
                src_cache[ln[0]] = open(f"/root/repo/pcss-unet_b200/csrc/{ln[0]}").read().split("\n")
            text = src_cache[ln[0]][ln[1] - 1].strip()[:70]
        except Exception:
            pass
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{n:6d} {100 * n / tot:5.1f}% {str(ln):28s} {text:70s} {top}")
