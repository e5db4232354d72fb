#!/usr/bin/env python3
"""Work counters (instrumented strict build, launched like the product frame) of one frame of a workload, as JSON:
    python tools/stats_dump.py <workload> <frame> [mode]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

workload, frame = sys.argv[1], int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "strict-accel"
wl = bench.GpuWorkload(workload, bench.MODES[mode], 0)
w, h = bench.frame_size(workload, 1, "strong")
wl.goto(frame)
wl.renderer.sync_scene(wl.scene)
c = wl.eng.debug_trace_stats(wl.cam, w, h, bench.TILE)
# one product frame for its device time
d = wl.eng.device_alloc(w * h * 16)
ms = []
for _ in range(5):
    wl.eng.render_frame_device(wl.cam, w, h, None, bench.TILE, None, None, d)
    wl.eng.sync()
    ms.append(wl.eng.stats()["last_trace_ms"])
c["_workload"], c["_frame"], c["_size"], c["_trace_ms_best_of_5"] = workload, frame, [w, h], min(ms)
print(json.dumps(c))
