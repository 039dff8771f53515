#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and, with
--step MARKER, the kernels of one step in order (between two launches whose name contains MARKER)."""
import collections
import csv
import io
import re
import sys

path = sys.argv[1]
marker = sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--step" else None
txt = open(path).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
names = [re.sub(r"\(.*", "", r["Kernel Name"])[:64] for r in rows]
vals = []
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    vals.append(v if u.startswith("us") else v / 1000 if u.startswith("ns") else v * 1000)
agg = collections.OrderedDict()
for n, v in zip(names, vals):
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{path}: {len(rows)} launches, {tot:.1f} us")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"  {t:10.1f} us {c:4d}x {t / c:9.2f} us/launch {100 * t / tot:5.1f}%  {k}")
if marker:
    idx = [i for i, n in enumerate(names) if marker in n]
    a, b = idx[-2], idx[-1]
    s = 0.0
    print(f"-- one step ({marker} .. {marker}):")
    for i in range(a, b):
        print(f"  {vals[i]:8.2f} us  {names[i]}  grid={rows[i]['Grid Size']}")
        s += vals[i]
    print(f"  sum {s:.1f} us over {b - a} launches")
