#!/usr/bin/env python3
"""Exploration helper: device time of bvht_trace_rays_device (the Scene::intersect seam) for ray batches whose origins lie
well outside / inside the default leaf-accelerator limits, accel vs brute force, with a bit-for-bit comparison."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from bvhtracer_b200 import FLAG_LEAF_ACCEL, FLAG_STRICT, _ffi, examples, host

F = np.float32


def batch(rng, n, spread, scale_dirs):
    o = (rng.normal(size=(n, 3)) * spread).astype(F)
    target = rng.uniform(-4, 4, (n, 3)).astype(F)
    d = (target - o).astype(F)
    if scale_dirs:
        d = (d * rng.uniform(0.01, 40.0, (n, 1))).astype(F)
    else:
        d /= np.linalg.norm(d, axis=1, keepdims=True).astype(F)
    return np.concatenate([o, d, np.full((n, 1), 3.4028235e38, F)], axis=1).astype(F)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
    scene, _ = host.build_scene(examples.sixteen_armadillos(0))
    rng = np.random.default_rng(5)
    cases = {"near (|o|~6, unit d)": batch(rng, n, 6.0, False), "far (|o|~300, unit d)": batch(rng, n, 300.0, False),
             "far, |d| in [0.01,40]x": batch(rng, n, 300.0, True)}
    for cname, rays in cases.items():
        outs = {}
        for mname, flags in (("brute", FLAG_STRICT), ("accel", FLAG_STRICT | FLAG_LEAF_ACCEL)):
            renderer = host.Renderer(flags=flags)
            renderer.sync_scene(scene)
            eng = renderer.engine()
            if True:
                drays = eng.device_alloc(rays.nbytes)
                dout = eng.device_alloc(n * 16)
                eng.memcpy_h2d(drays, rays)
                ms = []
                for _ in range(3):
                    eng.trace_rays_device(drays, n, dout)
                    eng.sync()
                    ms.append(eng.stats()["last_trace_ms"])
                hits = np.zeros(n, dtype=_ffi.HIT)
                eng.memcpy_d2h(hits, dout)
                eng.device_free(drays); eng.device_free(dout)
            del eng, renderer
            outs[mname] = hits
            print(f"{cname:26s} {mname:6s} {min(ms):9.3f} ms {n / min(ms) / 1e3:10.1f} Mrays/s hits={(hits['id'] != 0xFFFFFFFF).mean():.3f}", flush=True)
        print(f"{cname:26s} accel == brute bit-for-bit: {outs['accel'].tobytes() == outs['brute'].tobytes()}", flush=True)


if __name__ == "__main__":
    main()
