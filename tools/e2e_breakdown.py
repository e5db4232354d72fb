#!/usr/bin/env python3
"""Host-side breakdown of one e2e frame (GPU box): python tools/e2e_breakdown.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bvhtracer_b200 import examples, host

anim = examples.GridAnimation()
scene, models = host.build_scene(examples.sixteen_armadillos(0))
r = host.Renderer(flags=2)
eng = r.engine()
w, h = 3840, 2160
state = host.RendererState(host.depth_pipeline(), w, h, keep_hits=False)
d = eng.device_alloc(w * h * 16)
for _ in range(3):
    r.render(state, scene)
T = {"update": [], "sync_scene": [], "render_total": [], "device_only": []}
for f in range(12):
    t0 = time.perf_counter()
    anim.update()
    for i, o in enumerate(anim.objects()):
        scene.set_transform(i, host.object_transform(o))
    scene.rebuild()
    t1 = time.perf_counter()
    r.sync_scene(scene)
    eng.sync()
    t2 = time.perf_counter()
    r.render(state, scene)
    t3 = time.perf_counter()
    eng.render_frame_device(scene.camera(), w, h, None, 8, None, None, d)
    eng.sync()
    t4 = time.perf_counter()
    T["update"].append(t1 - t0); T["sync_scene"].append(t2 - t1); T["render_total"].append(t3 - t2); T["device_only"].append(t4 - t3)
for k, v in T.items():
    print(f"{k:14s} median {np.median(v) * 1e3:8.3f} ms   min {min(v) * 1e3:8.3f} ms")
print("last_trace_ms", r.stats()["last_trace_ms"])
