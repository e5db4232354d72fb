#!/usr/bin/env python3
"""profiles/traffic.json from ncu metric passes over tools/stats_dump.py:

    python tools/traffic_from_ncu.py <stats.jsonl> <ncu csv pattern with {workload} {frame}> [out json]

For every (workload, frame) of the stats file: dram__bytes_read.sum + dram__bytes_write.sum of EVERY kernel of the last product
frame in the capture (everything launched between the previous trace kernel and the last trace kernel of the product flavour)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head = rows[0]
    iid, iname, imet, ival = head.index("ID"), head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value")
    out = {}
    for r in rows[1:]:
        out.setdefault(int(r[iid]), {"name": r[iname]})[r[imet]] = float(r[ival].replace(",", ""))
    return [out[k] for k in sorted(out)]


def short(name):
    n = name.split("::")[-1]
    return n.split("(")[0]


def main():
    pattern = sys.argv[2]
    out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "traffic.json")
    res = {"_format": "\"<workload>:<mode>\" -> DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of EVERY kernel of one resident "
                      "frame (work-counter / coverage fills, coverage raster K7, classify+fill K0, trace K1), from an ncu pass over "
                      "tools/stats_dump.py (tools/traffic_from_ncu.py); bench.py reports dram_bytes_per_frame as roofline.traffic"}
    for line in open(sys.argv[1]):
        c = json.loads(line)
        w, f = c["_workload"], c["_frame"]
        key = f"{w}:strict-accel"
        path = pattern.format(workload=w, frame=f)
        if key in res or not os.path.exists(path):
            continue
        ls = launches(path)
        traces = [i for i, l in enumerate(ls) if "trace_primary" in l["name"]]
        names = [ls[i]["name"] for i in traces]
        product = max(set(names), key=names.count)
        last = max(i for i in traces if ls[i]["name"] == product)
        prev = max([i for i in traces if i < last], default=-1)
        frame = ls[prev + 1:last + 1]
        total = sum(l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0) for l in frame)
        res[key] = {"dram_bytes_per_frame": int(total), "kernels": [short(l["name"]) for l in frame], "frame": f,
                    "k1_dram_bytes": int(ls[last].get("dram__bytes_read.sum", 0.0) + ls[last].get("dram__bytes_write.sum", 0.0)),
                    "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum of every kernel of one resident frame ({os.path.relpath(path, ROOT)})"}
        print(key, res[key]["dram_bytes_per_frame"], res[key]["kernels"])
    json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
