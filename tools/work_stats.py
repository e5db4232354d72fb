#!/usr/bin/env python3
"""Per-ray work of the trace kernel (instrumented strict build) for each config: python tools/work_stats.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvhtracer_b200 import examples, host  # noqa: E402

CASES = [("two_armadillos", examples.two_armadillos()), ("sixteen_armadillos f4", examples.sixteen_armadillos(4)),
         ("trippy_teapots f10", examples.trippy_teapots(10)), ("big_ben_clock", examples.big_ben_clock())]
for name, spec in CASES:
    w, h = spec.bench_size
    scene, models = host.build_scene(spec)
    for mode, flags in (("brute", 0), ("accel", 2)):
        r = host.Renderer(flags=flags)
        r.sync_scene(scene)
        c = r.engine().debug_trace_stats(scene.camera(), w, h)
        n = c["rays"]
        print(f"{name:24s} {w}x{h} {mode}: " + " ".join(f"{k}={v / n:.2f}" for k, v in c.items() if k != "rays"), flush=True)
