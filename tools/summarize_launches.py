"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name (and per
grid size for the conv kernel), launches, total time and share of the step."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = r["Kernel Name"].split("(")[0]
    rows.append((name, r.get("Grid Size", ""), r.get("Block Size", ""), ns))
tot = sum(r[3] for r in rows)
agg = defaultdict(lambda: [0, 0.0])
for name, grid, blk, ns in rows:
    a = agg[name]
    a[0] += 1
    a[1] += ns
print(f"# {len(rows)} launches, total {tot / 1e6:.3f} ms (cold-cache, serialised: compare shares)")
print(f"{'kernel':60s} {'launches':>8s} {'ms':>9s} {'share':>7s}")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {n:8d} {ns / 1e6:9.3f} {100 * ns / tot:6.1f}%")
if len(sys.argv) > 2:
    print("\n# individual launches >= %s us" % sys.argv[2])
    for i, (name, grid, blk, ns) in enumerate(rows):
        if ns / 1e3 >= float(sys.argv[2]):
            print(f"{i:4d} {name[:44]:44s} grid={grid:14s} {ns / 1e3:9.1f} us")
