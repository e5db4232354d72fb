#!/usr/bin/env python3
"""Regenerates the golden fixtures in this directory from the CPU oracle (oracle/bvht_oracle.c).

The reference is a Rust crate and cannot run in the build image (no cargo), so the fixtures are the ORACLE's outputs; the
oracle itself is pinned by the reference's own known-answer tests (tests/test_oracle_kat.py, reference_kats.json here).
They serve two purposes: (1) the oracle cannot drift silently between rounds (tests/test_golden.py re-derives every
fixture on the CPU and compares bytes), (2) the CUDA path is compared against committed bytes, not only against whatever
the oracle computes on the day.

    python tests/golden/make_golden.py        # rewrites hits.npz, structures.json
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

import oracle_lib as O
import scene_build as SB
from bvhtracer_b200 import examples

# (fixture name, spec factory, width, height): small frames of every config (SURVEY.md section 6, C1..C5 + the quad test scene)
FRAMES = [
    ("cube_64x64", lambda: examples.cube(), 64, 64),
    ("quad_64x64", lambda: examples.quad(), 64, 64),
    ("two_armadillos_canonical_96x54", lambda: examples.two_armadillos("canonical"), 96, 54),
    ("two_armadillos_initial_96x54", lambda: examples.two_armadillos("initial"), 96, 54),
    ("sixteen_armadillos_f0_96x54", lambda: examples.sixteen_armadillos(0), 96, 54),
    ("sixteen_armadillos_f37_96x54", lambda: examples.sixteen_armadillos(37), 96, 54),
    ("trippy_teapots_f10_96x54", lambda: examples.trippy_teapots(10), 96, 54),
    ("big_ben_clock_96x54", lambda: examples.big_ben_clock(), 96, 54),
]
ASSETS = ["cube.obj", "teapot.obj", "armadillo.tri", "bigben.tri", "unity.tri"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def frames():
    out = {}
    for name, make, w, h in FRAMES:
        scene, cam = SB.oracle_scene(make())
        out[name] = scene.render(cam, w, h, threads=max(1, O.max_threads()))
    # big_ben_clock after two animate + refit steps (examples/big_ben_clock.rs:67-103)
    blas = O.Blas(O.load_asset("bigben.tri"))
    anim = examples.BigBenAnimation(blas.tris)
    for _ in range(2):
        blas.tris[:] = anim.animate()
        blas.refit()
    _, cam = SB.oracle_scene(examples.big_ben_clock())
    scene = O.Scene([blas], [(0, O.mat4_identity())], with_transform=False)
    out["big_ben_clock_refit2_96x54"] = scene.render(cam, 96, 54, threads=max(1, O.max_threads()))
    return out


def structures():
    out = {"assets": {}, "tlas": {}}
    for a in ASSETS:
        b = O.Blas(O.load_asset(a))
        out["assets"][a] = {"n_tris": int(b.n_tris), "nodes_used": int(b.nodes_used), "nodes_sha256": sha(b.nodes[:b.nodes_used]),
                            "reordered_tris_sha256": sha(b.tris)}
    for name, make in (("sixteen_armadillos_f37", lambda: examples.sixteen_armadillos(37)),
                       ("trippy_teapots_f10", lambda: examples.trippy_teapots(10)),
                       ("two_armadillos_canonical", lambda: examples.two_armadillos("canonical"))):
        scene, cam = SB.oracle_scene(make())
        out["tlas"][name] = {"nodes_used": int(scene.tlas_used), "nodes_sha256": sha(scene.tlas[:scene.tlas_used]),
                             "inverse_transforms_sha256": sha(scene.inst["inv"]), "camera_sha256": sha(SB.to_ffi_camera(cam))}
    return out


def main():
    f = frames()
    np.savez_compressed(os.path.join(HERE, "hits.npz"), **{k: v.view(np.uint8) for k, v in f.items()})
    s = structures()
    s["frames_sha256"] = {k: sha(v) for k, v in f.items()}
    with open(os.path.join(HERE, "structures.json"), "w") as fh:
        json.dump(s, fh, indent=1, sort_keys=True)
    for k, v in f.items():
        print(f"{k:36s} {v.size:6d} rays  hits {(v['id'] != O.MISS_ID).mean():.3f}  sha256 {sha(v)[:16]}")


if __name__ == "__main__":
    main()
