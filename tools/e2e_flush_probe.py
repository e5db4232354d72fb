#!/usr/bin/env python3
"""Exploration helper: Renderer::render (RGBA frame to pinned host memory) timed with CUDA events, with and without the
256 MiB L2-flush memset in front, static and animated scene."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bvhtracer_b200 import examples, host

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
anim = examples.GridAnimation()
scene, models = host.build_scene(examples.sixteen_armadillos(0))
r = host.Renderer(flags=2)
r.set_stream(stream.cuda_stream)
w, h = 3840, 2160
state = host.RendererState(host.depth_pipeline(), w, h, keep_hits=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    r.render(state, scene)


def run(do_flush, animate, n=24):
    ev, wall = [], []
    for _ in range(n):
        if animate:
            anim.update()
            for i, o in enumerate(anim.objects()):
                scene.set_transform(i, host.object_transform(o))
            scene.rebuild()
        if do_flush:
            flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream); r.render(state, scene); e1.record(stream)
        wall.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
        ev.append(e0.elapsed_time(e1))
    return np.median(ev[4:]), np.median(wall[4:]) * 1e3, r.stats()["last_trace_ms"]


for animate in (False, True):
    for do_flush in (False, True):
        e, wl, tr = run(do_flush, animate)
        print(f"animate={animate!s:5s} flush={do_flush!s:5s} events {e:7.3f} ms  wall {wl:7.3f} ms  (last band trace_ms {tr:.3f})", flush=True)
