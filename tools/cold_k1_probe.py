#!/usr/bin/env python3
"""Trace kernel (K1) and whole-frame device time with a warm L2 and after bench.py's 256 MiB flush: python tools/cold_k1_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "sixteen_armadillos"
w, h = bench.frame_size(workload, 1, "strong")
wl = bench.GpuWorkload(workload, 2, 0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
wl.renderer.set_stream(stream.cuda_stream)
d = wl.eng.device_alloc(w * h * 16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
reader = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in ("warm", "flush: 256 MiB memset (bench.py)", "flush: 256 MiB read (clean lines)"):
    wl.goto(0)
    k1, fr = [], []
    for f in range(1, 26):
        wl.advance()
        wl.renderer.sync_scene(wl.scene)
        if mode.startswith("flush: 256 MiB memset"):
            flush.zero_()
        elif mode.startswith("flush: 256 MiB read"):
            reader.sum()
        wl.eng.render_frame_device(wl.cam, w, h, None, bench.TILE, None, None, d)
        wl.eng.sync()
        st = wl.renderer.stats()
        if f >= 6:
            k1.append(st["last_k1_ms"]); fr.append(st["last_trace_ms"])
    print(f"{workload} {mode:36s}: K1 {np.mean(k1):.4f} ms   frame (K7 + K0 + K1) {np.mean(fr):.4f} ms   rebakes {st['rebakes']} d_max {st['bake_d_max']:.3f} o_max {st['bake_o_max']:.3f}", flush=True)

# does a preceding end-to-end phase (bvht_render_frame, banded) change the resident K1 time?  (bench.py measures K1 after its e2e loops)
from bvhtracer_b200 import host  # noqa: E402
state = host.RendererState(wl.pipeline, w, h, keep_hits=False)
for _ in range(5):
    wl.advance()
    wl.renderer.render(state, wl.scene)
wl.goto(0)
k1, fr = [], []
for f in range(1, 26):
    wl.advance()
    wl.renderer.sync_scene(wl.scene)
    flush.zero_()
    wl.eng.render_frame_device(wl.cam, w, h, None, bench.TILE, None, None, d)
    wl.eng.sync()
    st = wl.renderer.stats()
    if f >= 6:
        k1.append(st["last_k1_ms"]); fr.append(st["last_trace_ms"])
print(f"{workload} after an e2e phase, flushed            : K1 {np.mean(k1):.4f} ms   frame {np.mean(fr):.4f} ms", flush=True)
print("per frame K1:", [round(x, 4) for x in k1], "rebakes", st["rebakes"], st["bake_d_max"], st["bake_o_max"])
