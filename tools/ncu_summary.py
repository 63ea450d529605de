"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and
mean duration and share.  Usage: python tools/ncu_summary.py gpurun_out/launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, v * scale, r.get("Grid Size", ""), r.get("Block Size", "")))
rows = rows[skip:]
agg = defaultdict(lambda: [0, 0.0])
for n, us, *_ in rows:
    agg[n][0] += 1
    agg[n][1] += us
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.1f} us total")
print(f"{'kernel':60s} {'count':>7s} {'total_us':>12s} {'mean_us':>10s} {'share':>7s}")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} {c:7d} {t:12.1f} {t / c:10.2f} {100 * t / tot:6.1f}%")
