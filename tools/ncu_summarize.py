"""tools/ncu_summarize.py — aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, v * scale))
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, v in rows:
    agg[n][0] += 1
    agg[n][1] += v
print("%d launches, %.3f ms total (cold-cache, serialised: compare SHARES)" % (len(rows), tot))
print("%-70s %8s %12s %8s" % ("kernel", "launches", "total ms", "share"))
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %8d %12.3f %7.2f%%" % (n[:70], c, v, 100 * v / tot))
