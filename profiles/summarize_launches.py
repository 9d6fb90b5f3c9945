"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [only_ours]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
only_ours = len(sys.argv) > 2
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
agg = defaultdict(lambda: [0, 0.0])
for name, ns in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    if only_ours and "scdef" not in name:
        continue
    agg[short][0] += 1
    agg[short][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>10s} {'avg_us':>10s} {'share':>7s}")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:8d} {ns / 1e3:10.1f} {ns / 1e3 / n:10.1f} {ns / tot:7.3f}")
print(f"{'TOTAL':60s} {sum(v[0] for v in agg.values()):8d} {tot / 1e3:10.1f}")
