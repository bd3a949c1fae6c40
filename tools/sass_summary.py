#!/usr/bin/env python
"""Counts of the Blackwell-specific SASS instructions per kernel in libnsm_b200.so -> profiles/rNN_sass_summary.txt.

    python tools/sass_summary.py [profiles/r02_sass_summary.txt]

UTCHMMA / UTCQMMA = tcgen05.mma kind::f16 / kind::f8f6f4 (".2CTA" = cta_group::2), LDTM / STTM = tcgen05.ld / tcgen05.st (TMEM <-> registers; STTM: the fused blocks write the 1x1 GEMM's A operand to TMEM),
UTMALDG / UTMASTG = TMA tensor loads / stores (cp.async.bulk.tensor), UBLKCP = 1-D bulk copies (cp.async.bulk: the TMA-staged
BatchNorm backward), UTCBAR = tcgen05.commit, HMMA = mma.sync (warp-level
tensor cores of the 16-channel head / tail stages), SYNCS = mbarrier operations, BRA.U.ANY = the loop the compiler wraps around a
uniform-datapath instruction (UTC*MMA, UTMALDG, UTCBAR ...) issued under divergent control flow such as `if (lane == 0)`: one
per tcgen05 / TMA instruction before the MMA and producer warps went to warp-uniform control flow + elect.sync, 0 there now
(the rest: per-warp TMA loads / stores of the epilogues).  Runs on the CPU box (cuobjdump only)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pcss-unet_b200", "libnsm_b200.so")
PATS = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCQMMA.2CTA", r"\bUTCQMMA\.2CTA"),
        ("UTCQMMA", r"\bUTCQMMA\b(?!\.2CTA)"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"),
        ("UTCBAR", r"\bUTCBAR"), ("HMMA", r"\bHMMA"), ("SYNCS", r"\bSYNCS"), ("STL/LDL", r"\b(STL|LDL)\b"),
        ("BRA.U.ANY", r"\bBRA\.U\.ANY")]


def main():
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATS:
            if re.search(pat, line):
                counts[cur][name] += 1
    rows = []
    for fn, c in counts.items():
        if not any(c[k] for k, _ in PATS[:10]):
            continue
        d = demangle(fn)
        d = re.sub(r">\(.*$", ">", d) if ">(" in d else re.sub(r"\(.*$", "", d)
        name = d.replace("void ", "").replace("nsm::", "").replace("(int)", "").replace("(bool)", "")
        rows.append((name, c))
    with open(dst, "w") as f:
        f.write(f"# cuobjdump -sass pcss-unet_b200/libnsm_b200.so (sm_100a), instruction counts per kernel (tools/sass_summary.py)\n")
        f.write("# " + f"{'kernel':66s}" + "".join(f"{k:>13s}" for k, _ in PATS) + "\n")
        tot = collections.Counter()
        for name, c in sorted(rows):
            f.write(f"{name[:68]:68s}" + "".join(f"{c[k]:13d}" for k, _ in PATS) + "\n")
            tot.update(c)
        f.write(f"{'TOTAL':68s}" + "".join(f"{tot[k]:13d}" for k, _ in PATS) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    main()
