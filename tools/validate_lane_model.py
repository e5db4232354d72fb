#!/usr/bin/env python3
"""bench.py's performed-work estimate (work counters x profiles/lane_op_weights.json) against ncu's
smsp__thread_inst_executed.sum of the product trace kernel for the same frames.

    python tools/validate_lane_model.py <stats.jsonl from tools/stats_dump.py> <ncu csv pattern with {workload} {frame}>

The ncu files come from `ncu --metrics smsp__thread_inst_executed.sum,... -k regex:trace_primary|... python tools/stats_dump.py W F`."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = 148 * 128 * 1.965e9


def ncu_k1(path):
    """metrics of the LAST product trace_primary launch in the file (the instrumented build's kernel has `Stat` in its name)"""
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head = rows[0]
    iid, iname, imet, ival = head.index("ID"), head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value")
    launches = {}
    for r in rows[1:]:
        if "trace_primary" in r[iname]:
            launches.setdefault((int(r[iid]), r[iname]), {})[r[imet]] = float(r[ival].replace(",", ""))
    # the product kernel: the launch that executed the fewest thread instructions among the trace_primary launches is the
    # uninstrumented one only by accident; take the most common (name, duration class) instead: stats_dump runs it 5 times
    by_name = {}
    for (i, name), m in launches.items():
        by_name.setdefault(name, []).append(m)
    name = max(by_name, key=lambda n: len(by_name[n]))
    return name, by_name[name][-1]


def main():
    weights = json.load(open(os.path.join(ROOT, "profiles", "lane_op_weights.json")))["weights_thread_instructions_per_event"]
    for line in open(sys.argv[1]):
        c = json.loads(line)
        w, f = c["_workload"], c["_frame"]
        path = sys.argv[2].format(workload=w, frame=f)
        if not os.path.exists(path):
            continue
        name, m = ncu_k1(path)
        model = sum(weights[k] * c[k] for k in weights)
        ti, wi, us = m["smsp__thread_inst_executed.sum"], m["smsp__inst_executed.sum"], m["gpu__time_duration.sum"] / 1e3
        print(f"{w:20s} frame {f:2d}: ncu thread instructions {ti:.4g}  model {model:.4g}  model/ncu {model / ti:.3f}  | K1 {us:7.1f} us  "
              f"lanes/instr {ti / wi:.2f}  issue_active {m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', float('nan')):.1f} %  "
              f"-> lane-issue fraction ncu {ti / (us * 1e-6) / PEAK:.3f}  model {model / (us * 1e-6) / PEAK:.3f}")


if __name__ == "__main__":
    main()
