"""Aggregate an `ncu --page source --csv` dump by SASS opcode: python tools/sass_mix.py file.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for r in rows[2:]:
    if len(r) <= iE:
        continue
    src = r[iS].strip()
    parts = src.split()
    if not parts:
        continue
    op = parts[0]
    if op.startswith("@"):
        op = parts[1] if len(parts) > 1 else op
    base = op.split(".")[0]
    n = int(r[iE] or 0)
    agg[base][0] += n
    agg[base][1] += int(r[iSamp] or 0)
    agg[base][2] += 1
    tot += n
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print("total warp instructions executed: %d, static SASS lines: %d" % (tot, len(rows) - 2))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-12s exec=%12d  %5.1f%%  samples=%7d static=%5d" % (k, v[0], 100.0 * v[0] / tot, v[1], v[2]))
