"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches,
total / mean duration and share of the listed time.  usage: ncu_launch_summary.py launches.csv [topN]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as fh:
    lines = [ln for ln in fh if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg[name][0] += 1
    agg[name][1] += v_us
tot = sum(v[1] for v in agg.values())
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total listed: %d launches, %.1f us" % (sum(v[0] for v in agg.values()), tot))
print("%-90s %6s %10s %9s %6s" % ("kernel", "n", "total_us", "mean_us", "share"))
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-90s %6d %10.1f %9.2f %5.1f%%" % (name[:90], n, t, t / n, 100 * t / tot))
