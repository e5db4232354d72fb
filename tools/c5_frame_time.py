#!/usr/bin/env python3
"""C5 big_ben_clock per-frame path (GPU box): animate -> upload vertices -> device refit (BLAS + sub-BVH) -> trace 8K."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvhtracer_b200 import examples, host  # noqa: E402

w, h = (7680, 4320) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
scene, models = host.build_scene(examples.big_ben_clock())
anim = examples.BigBenAnimation(models[0].primitives())
r = host.Renderer(flags=2)
st = host.RendererState(host.intersection_pipeline(), w, h)
r.render(st, scene)
for f in range(8):
    v = anim.animate()
    t0 = time.perf_counter()
    models[0].set_primitives(v)
    models[0].refit()
    t1 = time.perf_counter()
    r.sync_scene(scene)          # bvht_blas_update_vertices + bvht_blas_refit + read back node boxes
    t2 = time.perf_counter()
    r.render(st, scene)          # trace + shade + D2H of the Rgba<u8> frame
    t3 = time.perf_counter()
    s = r.stats()
    print("frame %d: host copy %.2f ms | upload+repack+sub-refit+bake+refit %.2f ms | render %dx%d %.2f ms (trace %.2f ms) | %.0f Mrays/s whole frame"
          % (f, 1e3 * (t1 - t0), 1e3 * (t2 - t1), w, h, 1e3 * (t3 - t2), s["last_trace_ms"], w * h / (t3 - t1) / 1e6), flush=True)
