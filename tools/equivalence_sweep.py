#!/usr/bin/env python3
"""Empirical backing of the leaf accelerator's equivalence claim (GPU box): strict-accel vs strict-brute on many
frames / cameras / random rays, bit-for-bit.  python tools/equivalence_sweep.py [quick]

Also usable as a margin study with an EXPERIMENT build of the library (tools/build_variants.py exp:BVHT_EXPERIMENT; the
product library reads no environment): BVHT_C_MT=<c> scales the Moeller-Trumbore residual term of the box inflation
(shipped value 80); the number of mismatching hit records as c goes to 0 shows how much slack the bound has."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bvhtracer_b200 import Engine, FLAG_FAST, FLAG_LEAF_ACCEL, FLAG_STRICT, _ffi, examples, host

quick = "quick" in sys.argv[1:]
fast = "fast" in sys.argv[1:]          # compare the FAST build (accel) against strict-brute instead: how many records differ at all
rng = np.random.default_rng(2026)
total = mism = 0
t_start = time.time()


def compare(tag, brute, accel, cam, w, h):
    global total, mism
    a = brute.trace_primary(cam, w, h)
    b = accel.trace_primary(cam, w, h)
    bad = int((a.view(np.uint8).reshape(-1, 16) != b.view(np.uint8).reshape(-1, 16)).any(axis=1).sum())
    total += w * h
    mism += bad
    if bad:
        print(f"  MISMATCH {tag}: {bad} of {w * h} records", flush=True)


def upload(eng, scene, models):
    ids = {}
    inst = np.zeros(len(scene), _ffi.INSTANCE)
    for i in range(len(scene)):
        inv, _ = scene.instance(i)
        m = scene._models[i]
        if id(m) not in ids:
            nodes, used = m.nodes()
            ids[id(m)] = eng.blas_create(m.primitives(), nodes[:used], used)
        inst["transform_inv"][i] = inv
        inst["blas_id"][i] = ids[id(m)]
    tl, used = scene.tlas()
    eng.tlas_set(tl[:used], used, inst)
    return ids


def update_tlas(eng, scene, ids):
    inst = np.zeros(len(scene), _ffi.INSTANCE)
    for i in range(len(scene)):
        inst["transform_inv"][i] = scene.instance(i)[0]
        inst["blas_id"][i] = ids[id(scene._models[i])]
    tl, used = scene.tlas()
    eng.tlas_set(tl[:used], used, inst)


with Engine(flags=FLAG_STRICT) as brute, Engine(flags=(FLAG_FAST if fast else FLAG_STRICT) | FLAG_LEAF_ACCEL) as accel:
    # 1. animated frames of sixteen_armadillos and trippy_teapots at 4K / 1080p
    for name, size, frames in (("sixteen_armadillos", (3840, 2160), 12 if quick else 60), ("trippy_teapots", (3840, 2160), 6 if quick else 30)):
        anim = examples.GridAnimation()
        scene, models = host.build_scene(examples.CONFIGS[name](0))
        ib, ia = upload(brute, scene, models), upload(accel, scene, models)
        for f in range(frames):
            anim.update()
            for i, o in enumerate(anim.objects()):
                scene.set_transform(i, host.object_transform(o))
            scene.rebuild()
            update_tlas(brute, scene, ib); update_tlas(accel, scene, ia)
            compare(f"{name} frame {f + 1}", brute, accel, scene.camera(), *size)
        print(f"{name}: {frames} frames done, rays so far {total:,}, mismatches {mism}", flush=True)
    # 2. random cameras around two_armadillos (near, far, grazing), 1080p
    scene, models = host.build_scene(examples.two_armadillos("canonical"))
    ib, ia = upload(brute, scene, models), upload(accel, scene, models)
    for k in range(10 if quick else 60):
        pos = rng.normal(size=3)
        pos = pos / np.linalg.norm(pos) * rng.choice([1.6, 2.5, 4.0, 9.0, 30.0]) + np.array([0, 0.4, 0])
        fwd = (np.array([rng.uniform(-1.3, 1.3), rng.uniform(-0.5, 1.2), 0]) - pos)
        fwd /= np.linalg.norm(fwd)
        up0 = np.array([0, 1, 0]) if abs(fwd[1]) < 0.95 else np.array([1, 0, 0])
        right = np.cross(fwd, up0); right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        cam = host.Camera.symmetric_fov(float(rng.choice([40.0, 90.0, 120.0])), 1.0, 0.5, 1000.0, pos, fwd, right, up).to_ffi()
        compare(f"two_armadillos random camera {k}", brute, accel, cam, 1920, 1080)
    print(f"two_armadillos random cameras done, rays so far {total:,}, mismatches {mism}", flush=True)
    # 3. big_ben with vertex animation (device refits on both sides), 4K
    scene, models = host.build_scene(examples.big_ben_clock())
    ib, ia = upload(brute, scene, models), upload(accel, scene, models)
    bb = examples.BigBenAnimation(models[0].primitives())
    for f in range(6 if quick else 40):
        v = bb.animate()
        for eng, ids in ((brute, ib), (accel, ia)):
            bid = ids[id(models[0])]
            eng.blas_update_vertices(bid, v)
            eng.blas_refit(bid)
        compare(f"big_ben frame {f}", brute, accel, scene.camera(), 3840, 2160)
    print(f"big_ben animated done, rays so far {total:,}, mismatches {mism}", flush=True)
print(f"{'FAST-accel' if fast else 'strict-accel'} vs strict-brute: TOTAL rays {total:,}  mismatching records {mism}  c_mt={os.environ.get('BVHT_C_MT', 'shipped value (leaf_accel.hpp)')}  "
      f"wall {time.time() - t_start:.0f} s")
