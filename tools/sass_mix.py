"""Aggregate an `ncu --page source --csv` dump into an opcode mix (warp-level executed counts
and stall samples).  Usage: python tools/sass_mix.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    toks = r[iS].split()
    if not toks:
        continue
    t = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    parts = t.split(".")
    op = parts[0]
    if op in ("I2F", "F2I", "FRND", "MUFU", "LDG", "LDS", "F2F", "STG", "STS"):
        op = ".".join(parts[:3])
    e = int(r[iE])
    ops[op] += e
    samp[op] += int(r[iSm]) if r[iSm].isdigit() else 0
    tot += e
print("total warp instructions", tot)
ts = sum(samp.values())
for op, c in ops.most_common(45):
    print(f"{op:24s} {c:12d} {100 * c / tot:5.1f}%   stall samples {samp[op]:7d} {100 * samp[op] / max(ts, 1):5.1f}%")
