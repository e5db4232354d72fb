#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/launch_list_summary.py in.csv [out.txt]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, agg = None, {}
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    name = r[hdr.index("Kernel Name")]
    val = float(r[hdr.index("Metric Value")])
    unit = r[hdr.index("Metric Unit")]
    val *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += val
tot = sum(v[1] for v in agg.values())
lines = [f"# per-kernel device time (cold-cache, serialised under ncu: compare SHARES, not absolutes); total {tot:.3f} ms"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{v[0]:5d} launches {v[1]:10.3f} ms {100 * v[1] / tot:6.2f}%  {k[:140]}")
text = "\n".join(lines)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
