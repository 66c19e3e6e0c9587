"""Summarise an `ncu --page source --csv` dump: opcode mix (executed), stall totals, top stall sites.
usage: python tools/ncu_src_summary.py file.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
mix = collections.Counter(); tot = 0
stall_tot = collections.Counter()
sites = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].split()
    if not src:
        continue
    op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    mix[op.split(".")[0]] += n; tot += n
    st = {c: int(r[ix[c]] or 0) for c in stall_cols}
    for c, v in st.items():
        stall_tot[c] += v
    sites.append((int(r[ix["# Samples"]] or 0), r[ix["Address"]], " ".join(src)[:70], max(st, key=st.get) if st else ""))
print("warp instructions executed: %d" % tot)
fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("FP64 share: %.1f%%" % (100.0 * fp64 / tot))
for k, v in mix.most_common(top):
    print("  %-10s %12d %5.1f%%" % (k, v, 100.0 * v / tot))
s = sum(stall_tot.values())
print("stall samples by reason:")
for k, v in stall_tot.most_common(12):
    print("  %-24s %8d %5.1f%%" % (k, v, 100.0 * v / max(s, 1)))
print("top stall sites:")
for n, a, t, why in sorted(sites, reverse=True)[:top]:
    print("  %6d %s %-70s %s" % (n, a, t, why))
