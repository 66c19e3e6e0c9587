"""one-screen summary of a bench.py JSON line: python tools/bench_brief.py file.json"""
import json
import sys

for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print("n_gpus %s  ms/step %.3f  value %.4g  e2e %s  tendency launch %.3f ms (frac %.3f)" % (
        d.get("n_gpus"), d.get("ms_per_step", float("nan")), d.get("value", float("nan")), (d.get("e2e") or {}).get("value"),
        (d.get("roofline") or {}).get("avg_launch_ms", float("nan")), (d.get("roofline") or {}).get("frac", float("nan"))))
    print("  phases", {k: round(v, 3) for k, v in (d.get("phases_ms_per_step") or {}).items()})
