#!/usr/bin/env python3
"""Print the markdown tables of DESIGN.md 5 / 7 from the bench lines kept under profiles/:
    python tools/design_tables.py workloads profiles/r02c_bench_{}_n1.json
    python tools/design_tables.py scaling   profiles/r02c_bench_{}_n{}.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORDER = ["cube", "two_armadillos", "sixteen_armadillos", "trippy_teapots", "big_ben_clock"]


def load(p):
    return json.loads(open(os.path.join(ROOT, p)).read().strip().splitlines()[-1])


def workloads(pattern):
    for w in ORDER:
        p = pattern.format(w)
        if not os.path.exists(os.path.join(ROOT, p)):
            continue
        d = load(p)
        c, r, e, cb = d["config"], d["roofline"], d["e2e"], d.get("cpu_baseline") or {}
        frac = f"{r['frac']:.2f}" if r.get("frac") is not None else "—"
        cpu = f"{cb['value']:.3g} ({cb['cores']}) / {cb['single_thread']['value']:.3g} (1)" if cb else "—"
        print(f"| {c['baseline_config']} `{w}` | {c['width']}×{c['height']} | {d['value'] / 1e3:.2f} ({d['ms_per_step']:.3f}) | "
              f"{e['value'] / 1e3:.2f} ({e['ms_per_step']:.3f}) | {e['d2h_bytes_per_step'] / 1e6:.1f} MB | {frac} | {cpu} |")


def scaling(pattern):
    for w in ("sixteen_armadillos", "big_ben_clock"):
        base = None
        print(f"| `{w}` | N | `value` Grays/s (ms/frame) | speed-up | `e2e` Grays/s (ms/frame) | speed-up | frames equal 1-GPU frame |")
        print("|---|---|---|---|---|---|---|")
        for n in (1, 2, 4, 8):
            p = pattern.format(w, n)
            if not os.path.exists(os.path.join(ROOT, p)):
                continue
            d = load(p)
            if base is None:
                base = d
            ok = "—" if n == 1 else f"{d['gathered_device_frame_equals_single_gpu']} / {d['sharded_frame_equals_single_gpu']}"
            print(f"| | {n} | {d['value'] / 1e3:.2f} ({d['ms_per_step']:.3f}) | {d['value'] / base['value']:.2f}× | "
                  f"{d['e2e']['value'] / 1e3:.2f} ({d['e2e']['ms_per_step']:.3f}) | {d['e2e']['value'] / base['e2e']['value']:.2f}× | {ok} |")
        print()


if __name__ == "__main__":
    {"workloads": workloads, "scaling": scaling}[sys.argv[1]](sys.argv[2])
