#!/usr/bin/env python3
"""A/B of the scheduling options (bvht_set_option) per workload: device time of the resident frame with the coverage raster
(K7) forced off / on and with the library's own rule.  python tools/cover_ab.py [workloads...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from bvhtracer_b200 import _ffi  # noqa: E402

for workload in (sys.argv[1:] or list(bench.WORKLOADS)):
    w, h = bench.frame_size(workload, 1, "strong")
    res = {}
    for label, val in (("off", 0), ("on", 1), ("rule", -1)):
        wl = bench.GpuWorkload(workload, 2, 0)
        wl.eng.set_option(_ffi.OPT_COVER, val)
        d = wl.eng.device_alloc(w * h * 16)
        ms = []
        for f in range(1, 13):
            wl.advance()
            wl.renderer.sync_scene(wl.scene)
            best = 1e9
            for _ in range(3):
                wl.eng.render_frame_device(wl.cam, w, h, None, bench.TILE, None, None, d)
                wl.eng.sync()
                best = min(best, wl.eng.stats()["last_trace_ms"])
            if f >= 5:
                ms.append(best)
        res[label] = float(np.mean(ms))
        wl.eng.device_free(d)
        del wl
    print(f"{workload:20s} {w}x{h}: cover off {res['off']:.4f} ms  on {res['on']:.4f} ms  rule {res['rule']:.4f} ms", flush=True)
