#!/usr/bin/env python3
"""Device time of the RESIDENT frame (bvht_render_frame_device) when its pixel blocks are pulled band by band in a given order
instead of row-major: python tools/resident_order_probe.py [workload] bands:order ...   (order 0 image, 1 cheapest first,
2 cheap ascending + expensive descending, 3 descending cost)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from bvhtracer_b200 import _ffi  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "sixteen_armadillos"
specs = [tuple(int(x) for x in a.split(":")) for a in sys.argv[2:]] or [(1, 0)]
w, h = bench.frame_size(workload, 1, "strong")
wl = bench.GpuWorkload(workload, 2, 0)
d = wl.eng.device_alloc(w * h * 16)
for bands, order in specs:
    wl.goto(0)
    wl.eng.set_option(_ffi.OPT_BANDS, bands if bands > 1 else -1)
    wl.eng.set_option(_ffi.OPT_BAND_ORDER, order)
    ms = []
    for f in range(1, 26):
        wl.advance()
        wl.renderer.sync_scene(wl.scene)
        best = 1e9
        for _ in range(3):
            wl.eng.render_frame_device(wl.cam, w, h, None, bench.TILE, None, None, d)
            wl.eng.sync()
            best = min(best, wl.eng.stats()["last_trace_ms"])
        if f >= 6:
            ms.append(best)
    print(f"{workload} {w}x{h} frames 6..25  bands {bands:2d} order {order}: mean {np.mean(ms):.4f} ms  min {min(ms):.4f}  max {max(ms):.4f}", flush=True)
