#!/usr/bin/env python3
"""Device time of the resident trace for the bench workloads with whatever libbvht_cuda.so is installed (no torch).
C3 / C4: the animated frames 6..25 of bench.py (host set_transform + Tlas::rebuild per frame); hit buffers of a few frames are
hashed and compared with /tmp/variant_ref/<case>.json (written by the first run)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bvhtracer_b200 import _ffi, examples, host  # noqa: E402

REF_DIR = "/tmp/variant_ref"
FLAGS = int(os.environ.get("VB_FLAGS", "2"))


def check(case, hashes):
    os.makedirs(REF_DIR, exist_ok=True)
    p = os.path.join(REF_DIR, case + ".json")
    if not os.path.exists(p):
        json.dump(hashes, open(p, "w"))
        return "ref written"
    ref = json.load(open(p))
    bad = [k for k in hashes if ref.get(k) != hashes[k]]
    return "identical" if not bad else f"MISMATCH {bad}"


def animated(case, mesh_spec, frames=(6, 25), hash_frames=(6, 15, 25)):
    spec = mesh_spec(0)
    w, h = spec.bench_size
    scene, models = host.build_scene(spec)
    r = host.Renderer(flags=FLAGS)
    eng = r.engine()
    cam = scene.camera()
    anim = examples.GridAnimation()
    d = eng.device_alloc(w * h * 16)
    ms, hashes = [], {}
    for f in range(1, frames[1] + 1):
        anim.update()
        for i, o in enumerate(anim.objects()):
            scene.set_transform(i, host.object_transform(o))
        scene.rebuild()
        r.sync_scene(scene)
        reps = 3 if f >= frames[0] else 1
        best = 1e9
        for _ in range(reps):
            eng.render_frame_device(cam, w, h, None, 8, None, None, d)
            eng.sync()
            best = min(best, eng.stats()["last_trace_ms"])
        if f >= frames[0]:
            ms.append(best)
        if f in hash_frames:
            hits = np.zeros(w * h, dtype=_ffi.HIT)
            eng.memcpy_d2h(hits, d)
            hashes[str(f)] = hashlib.sha1(hits.tobytes()).hexdigest()
    eng.device_free(d)
    print(f"{case:20s} {w}x{h} frames {frames[0]}..{frames[1]}: mean {np.mean(ms):.4f} ms  min {min(ms):.4f} max {max(ms):.4f}  "
          f"{w * h / np.mean(ms) / 1e6:.2f} Grays/s  {check(case, hashes)}", flush=True)


def static(case, spec, reps=8):
    w, h = spec.bench_size
    scene, models = host.build_scene(spec)
    r = host.Renderer(flags=FLAGS)
    eng = r.engine()
    cam = scene.camera()
    r.sync_scene(scene)
    d = eng.device_alloc(w * h * 16)
    ms = []
    for _ in range(reps):
        eng.render_frame_device(cam, w, h, None, 8, None, None, d)
        eng.sync()
        ms.append(eng.stats()["last_trace_ms"])
    hits = np.zeros(w * h, dtype=_ffi.HIT)
    eng.memcpy_d2h(hits, d)
    eng.device_free(d)
    print(f"{case:20s} {w}x{h}: best of last 3 {min(ms[-3:]):.4f} ms  {w * h / min(ms[-3:]) / 1e6:.2f} Grays/s  "
          f"{check(case, {'0': hashlib.sha1(hits.tobytes()).hexdigest()})}", flush=True)


CASES = {
    "c3": lambda: animated("c3", examples.sixteen_armadillos),
    "c4": lambda: animated("c4", examples.trippy_teapots),
    "c2": lambda: static("c2", examples.two_armadillos()),
    "c2i": lambda: static("c2i", examples.two_armadillos("initial")),
    "c5": lambda: static("c5", examples.big_ben_clock()),
    "c1": lambda: static("c1", examples.cube()),
}

if __name__ == "__main__":
    for c in (sys.argv[1:] or ["c3", "c2", "c4", "c5"]):
        CASES[c]()
