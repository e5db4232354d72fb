#!/usr/bin/env python3
"""Exploration helper: fixed per-ray cost of the trace kernel.  Times the C3 frame with the camera turned away from the
scene (every ray leaves at the root tight-box test: ray generation + 16 B store only) next to the normal frame."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dataclasses

import numpy as np

from bvhtracer_b200 import FLAG_LEAF_ACCEL, FLAG_STRICT, examples, host


def main():
    w, h = 3840, 2160
    spec = examples.sixteen_armadillos(0)
    away = dataclasses.replace(spec, camera=dataclasses.replace(spec.camera, forward=tuple(-np.array(spec.camera.forward, np.float32))))
    for name, sp in (("normal", spec), ("camera turned away", away)):
        scene, _ = host.build_scene(sp)
        cam = scene.camera()
        for mode, flags in (("strict-brute", FLAG_STRICT), ("strict-accel", FLAG_STRICT | FLAG_LEAF_ACCEL)):
            renderer = host.Renderer(flags=flags)
            renderer.sync_scene(scene)
            eng = renderer.engine()
            if True:
                d = eng.device_alloc(w * h * 16)
                ms = []
                for _ in range(5):
                    eng.trace_primary_device(cam, w, h, 8, None, d)
                    eng.sync()
                    ms.append(eng.stats()["last_trace_ms"])
                c = eng.debug_trace_stats(cam, w, h) if mode == "strict-accel" else None
            del eng, renderer
            extra = "" if c is None else " ".join(f"{k}={v / c['rays']:.2f}" for k, v in c.items() if k != "rays")
            print(f"{name:20s} {mode:13s} {min(ms[1:]):7.3f} ms  {extra}", flush=True)


if __name__ == "__main__":
    main()
