"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/summarize_launches.py f.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    k = row["Kernel Name"][:86]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("kernel launches: %d, summed device time %.1f us (ncu: cold cache, serialised -- compare SHARES)" % (sum(v[0] for v in agg.values()), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-88s n=%4d total=%10.1f us share=%5.1f%% avg=%8.1f us" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
