"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total ns, share.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
agg = defaultdict(lambda: [0, 0.0])
for name, ns, grid, blk in rows:
    short = re.sub(r"\(.*$", "", name)
    short = re.sub(r"^void ", "", short)
    agg[short][0] += 1
    agg[short][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"launches: {len(rows)}  total device time: {tot/1e6:.3f} ms (cold-cache, serialised under ncu: compare SHARES)\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"| `{k[:110]}` | {n} | {ns/1e6:.3f} | {100*ns/tot:.1f}% |")
