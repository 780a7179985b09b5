"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r and "Metric Value" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("<unnamed>::", "")
    val = float(d["Metric Value"].replace(",", ""))
    unit = d.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    agg[name][0] += 1
    agg[name][1] += ns
total = sum(v[1] for v in agg.values())
print("%-52s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-52s %8d %12.1f %6.1f%%" % (name[:52], cnt, ns / 1e3, 100 * ns / total))
print("%-52s %8d %12.1f" % ("TOTAL", sum(v[0] for v in agg.values()), total / 1e3))
