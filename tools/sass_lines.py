"""Attribute ncu per-SASS-instruction executed counts to CUDA source lines.
usage: sass_lines.py <ncu --page source --csv dump> <nvdisasm --print-line-info dump> <mangled kernel substring> [top]"""
import collections
import csv
import re
import sys

ncu_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]
iE, iS = hdr.index("Instructions Executed"), hdr.index("Source")
counts = [(int(r[iE] or 0), r[iS].strip()) for r in rows[2:] if len(r) > iE]
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
cur = ("?", 0)
seq = []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2).strip()))
print("sass in disasm: %d, in ncu: %d" % (len(seq), len(counts)))
agg = collections.Counter()
aggop = collections.defaultdict(collections.Counter)
for (src, ins), (n, s) in zip(seq, counts):
    agg[src] += n
    aggop[src][ins.split()[0].split(".")[0] if not ins.startswith("@") else ins.split()[1].split(".")[0]] += n
tot = sum(agg.values())
byfile = collections.Counter()
for (f, ln), n in agg.items():
    byfile[f] += n
print({f: "%.1f%%" % (100.0 * n / tot) for f, n in byfile.most_common()})
for src, n in agg.most_common(top):
    ops = ", ".join("%s %.0f%%" % (o, 100.0 * c / n) for o, c in aggop[src].most_common(5))
    print("%-22s:%4d  %5.1f%%  %s" % (src[0], src[1], 100.0 * n / tot, ops))
