"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` launch list: time share per kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name] += val * scale
    cnt[name] += 1
total = sum(tot.values())
print(f"total kernel time {total:.3f} ms over {sum(cnt.values())} launches")
for name, t in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{t:10.3f} ms  {100 * t / total:5.1f}%  x{cnt[name]:<5d} {name[:110]}")
