#!/usr/bin/env python3
"""Run a few resident 4K frames of one mode (for ncu captures): python tools/prof_one.py strict-accel [frames] [workload]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvhtracer_b200 import examples, host  # noqa: E402

MODES = {"strict-brute": 0x0, "strict-accel": 0x2, "fast-brute": 0x1, "fast-accel": 0x3}
mode = sys.argv[1] if len(sys.argv) > 1 else "strict-accel"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
workload = sys.argv[3] if len(sys.argv) > 3 else "sixteen_armadillos"
spec = examples.CONFIGS[workload](2) if workload in ("sixteen_armadillos", "trippy_teapots") else examples.CONFIGS[workload]()
w, h = spec.bench_size
scene, models = host.build_scene(spec)
renderer = host.Renderer(flags=MODES[mode])
eng = renderer.engine()
renderer.sync_scene(scene)
d = eng.device_alloc(w * h * 16)
for _ in range(frames):
    eng.render_frame_device(scene.camera(), w, h, None, 8, None, None, d)
    eng.sync()
print(mode, workload, w, h, renderer.stats()["last_trace_ms"], "ms")
eng.device_free(d)
