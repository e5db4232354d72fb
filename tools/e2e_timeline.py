#!/usr/bin/env python3
"""Where an end-to-end frame goes (GPU box): host time of the uploads and of the render call, and the device timeline of the
bands (bvht_debug_frame_timeline).   python tools/e2e_timeline.py [workload] [bands ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from bvhtracer_b200 import _ffi, host  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "sixteen_armadillos"
band_list = [tuple(int(x) for x in b.split(":")) for b in sys.argv[2:]] or [(-1,)]      # bands[:order policy[:copy streams]]
w, h = bench.frame_size(workload, 1, "strong")
wl = bench.GpuWorkload(workload, 2, 0)
state = host.RendererState(wl.pipeline, w, h, keep_hits=False)
for spec in band_list:
    bands, policy, streams = spec[0], (spec[1] if len(spec) > 1 else -1), (spec[2] if len(spec) > 2 else -1)
    wl.goto(0)                                    # every configuration sees the same frames 1..24
    wl.eng.set_option(_ffi.OPT_BANDS, bands)
    wl.eng.set_option(_ffi.OPT_BAND_ORDER, policy)
    wl.eng.set_option(_ffi.OPT_COPY_STREAMS, streams)
    wl.eng.set_option(_ffi.OPT_TIMELINE, 1)
    T = {"sync_scene (host)": [], "render (host, incl. final sync)": [], "device frame (ev_a..ev_b)": []}
    for f in range(24):
        wl.advance()
        wl.eng.sync()
        t0 = time.perf_counter()
        wl.renderer.sync_scene(wl.scene)
        t1 = time.perf_counter()
        wl.renderer.render(state, wl.scene)
        t2 = time.perf_counter()
        if f >= 6:
            T["sync_scene (host)"].append(t1 - t0)
            T["render (host, incl. final sync)"].append(t2 - t1)
            T["device frame (ev_a..ev_b)"].append(wl.renderer.stats()["last_trace_ms"] * 1e-3)
    print(f"== {workload} {w}x{h}  bands option {bands}  order policy {policy}  copy streams {streams}")
    for k, v in T.items():
        print(f"  {k:34s} median {np.median(v) * 1e3:7.3f} ms   min {min(v) * 1e3:7.3f} ms")
    tl = wl.eng.debug_frame_timeline()
    print(f"  last frame: coverage raster done {tl['cover_done_ms']:.3f} ms, frame done {tl['frame_done_ms']:.3f} ms")
    print(f"  all kernels done {tl['bands'][0][0]:.3f} ms; copies in pull order (first 16 bands):")
    print("    copy done: " + " ".join(f"{c:.3f}" for _, c, _ in tl["bands"]))
