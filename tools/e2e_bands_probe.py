#!/usr/bin/env python3
"""End-to-end frame time of Renderer::render (C3 4K, frames 9..32) for given band counts: python tools/e2e_bands_probe.py 1 4 7
(PROBE_KEEP_HITS=1 also returns the hit records; with an experiment build of the library -- tools/build_variants.py
exp:BVHT_EXPERIMENT -- BVHT_BAND_SHAPE, BVHT_BANDS_IMAGE_ORDER and BVHT_DEBUG_NO_D2H select what else is compared)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bvhtracer_b200 import examples, host
anim = examples.GridAnimation()
scene, models = host.build_scene(examples.sixteen_armadillos(0))
r = host.Renderer(flags=2)
eng = r.engine()
w, h = 3840, 2160
keep_hits = os.environ.get("PROBE_KEEP_HITS") == "1"
state = host.RendererState(host.depth_pipeline(), w, h, keep_hits=keep_hits)
for _ in range(8):
    anim.update()
for bands in sys.argv[1:]:
    eng.set_option(3, int(bands))      # BVHT_OPT_BANDS
    a2 = examples.GridAnimation()
    for _ in range(8): a2.update()
    ts, dev = [], []
    for f in range(24):
        a2.update()
        for i, o in enumerate(a2.objects()):
            scene.set_transform(i, host.object_transform(o))
        scene.rebuild()
        eng.sync()
        t0 = time.perf_counter()
        r.render(state, scene)
        ts.append(time.perf_counter() - t0)
        dev.append(r.stats()["last_trace_ms"])
    print(f"bands={bands:3s} render() wall median {np.median(ts[4:])*1e3:.3f} ms  min {min(ts[4:])*1e3:.3f}   device ev_a..ev_b median {np.median(dev[4:]):.3f} ms")
