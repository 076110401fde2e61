"""Summarise an ncu --csv launch list: per kernel name, launches and total device time of the LAST step."""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, v * scale))
# the bench runs warm-up steps first: keep the launches after the last k_init
last = max((i for i, (n, _) in enumerate(rows) if n.startswith("k_bounds_partial")), default=0)
rows = rows[last:]
agg = OrderedDict()
for n, ms in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'ms':>9s} {'share':>7s}")
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} {c:8d} {ms:9.3f} {ms / tot:7.1%}")
print(f"{'total':60s} {len(rows):8d} {tot:9.3f}")
