"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X <cmd>`): launches, total and
average device time and share per kernel.  usage: python profiles/summarize_launches.py launches.csv [> summary.txt]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("mcacb::", "").replace("void ", "").strip()
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    rows.append((name, us))
tot = sum(u for _, u in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, u in rows:
    agg[n][0] += 1
    agg[n][1] += u
print(f"# total kernel time {tot / 1e3:.3f} ms over {len(rows)} launches")
print(f"{'kernel':44s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:44]:44s} {c:8d} {u:12.1f} {u / c:10.2f} {100 * u / tot:6.1f}%")
