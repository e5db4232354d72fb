"""Per-frame scene update: host (C++ mirror: set_transform x n + Tlas::rebuild, then bvht_tlas_set) against the device
(bvht_scene_set_transforms, K6) for growing instance counts.  Wall-clock per call, median of `reps`.

    python tools/scene_update_bench.py [--reps 20]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvhtracer_b200 import Engine, FLAG_LEAF_ACCEL, FLAG_STRICT, examples, host  # noqa: E402


def grid_transforms(rng, n):
    side = int(np.ceil(n ** (1.0 / 3.0)))
    out = []
    for i in range(n):
        x, y, z = i % side, (i // side) % side, i // (side * side)
        t = (np.array([x, y, z], "f4") - (side - 1) / 2) * 2.5 + rng.uniform(-0.3, 0.3, 3).astype("f4")
        out.append(host.Transform3.new((0.75, 0.75, 0.75), tuple(t), float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3))))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--counts", default="1,2,16,64,150,400,1000,2000,4096")
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    spec = examples.cube()
    mesh_scene, models = host.build_scene(spec)
    model = models[0]
    renderer = host.Renderer(flags=FLAG_STRICT | FLAG_LEAF_ACCEL)
    cam = host.Camera.from_spec(spec.camera)
    print("%6s %14s %14s %14s %8s" % ("n", "host_ms", "host+set_ms", "device_ms", "kernel+copies_ms"))
    for n in [int(c) for c in a.counts.split(",")]:
        sb = host.SceneBuilder(cam)
        tf = grid_transforms(rng, n)
        for t in tf:
            sb.with_object(model, t)
        scene = sb.build()
        renderer.sync_scene(scene)
        reps = a.reps if n <= 1000 else max(3, a.reps // 5)
        th, ths, td, tk = [], [], [], []
        for r in range(reps):
            tf = grid_transforms(rng, n)
            t0 = time.perf_counter()
            for i, t in enumerate(tf):
                scene.set_transform(i, t)
            scene.rebuild()
            t1 = time.perf_counter()
            renderer.sync_scene(scene)
            t2 = time.perf_counter()
            th.append(t1 - t0); ths.append(t2 - t0)
            t0 = time.perf_counter()
            renderer.update_transforms(scene, tf)
            td.append(time.perf_counter() - t0)
            tk.append(renderer.stats()["last_upload_ms"])
        print("%6d %14.4f %14.4f %14.4f %8.4f" % (n, 1e3 * np.median(th), 1e3 * np.median(ths), 1e3 * np.median(td), np.median(tk)))


if __name__ == "__main__":
    main()
