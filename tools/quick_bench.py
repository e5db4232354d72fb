#!/usr/bin/env python3
"""Exploration helper (not the judged bench): device time of the trace kernel for every config and mode.
Eight frames per mode, best of the last three: the first five let the library settle its measured choices (coverage raster)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from bvhtracer_b200 import FLAG_FAST, FLAG_LEAF_ACCEL, FLAG_STRICT, _ffi, examples, host


def compare_hits(got, ref):
    """id mismatches and worst relative t error between two hit buffers"""
    both = (got["id"] != 0xFFFFFFFF) & (ref["id"] != 0xFFFFFFFF) & (got["id"] == ref["id"])
    rel = np.abs(got["t"][both].astype(np.float64) - ref["t"][both]) / np.maximum(np.abs(ref["t"][both].astype(np.float64)), 1e-30)
    return {"id_mismatch": int((got["id"] != ref["id"]).sum()), "max_rel_t": float(rel.max()) if rel.size else 0.0,
            "bit_identical": got.tobytes() == ref.tobytes()}

MODES = {"strict-brute": FLAG_STRICT, "strict-accel": FLAG_STRICT | FLAG_LEAF_ACCEL,
         "fast-brute": FLAG_FAST, "fast-accel": FLAG_FAST | FLAG_LEAF_ACCEL}


def main():
    only = sys.argv[1:] or None
    cases = [("cube", examples.cube(), (640, 640)),
             ("two_armadillos", examples.two_armadillos(), (1920, 1080)),
             ("sixteen_armadillos", examples.sixteen_armadillos(0), (3840, 2160)),
             ("sixteen_armadillos_f30", examples.sixteen_armadillos(30), (3840, 2160)),
             ("trippy_teapots", examples.trippy_teapots(10), (3840, 2160)),
             ("big_ben_clock", examples.big_ben_clock(), (7680, 4320))]
    for name, spec, (w, h) in cases:
        if only and name not in only:
            continue
        scene, _ = host.build_scene(spec)                     # host side: the C++ mirror of the reference's builders
        fcam = scene.camera()
        ref = None
        for mname, flags in MODES.items():
            renderer = host.Renderer(flags=flags)
            eng = renderer.engine()
            if True:
                t0 = time.time()
                renderer.sync_scene(scene)
                up = time.time() - t0
                dout = eng.device_alloc(w * h * 16)
                ms = []
                for it in range(8):
                    eng.trace_primary_device(fcam, w, h, 8, None, dout)
                    eng.sync()
                    ms.append(eng.stats()["last_trace_ms"])
                hits = np.zeros(w * h, dtype=_ffi.HIT)
                eng.memcpy_d2h(hits, dout)
                eng.device_free(dout)
                st = eng.stats()
            del eng, renderer
            if ref is None:
                ref = hits
                cmp_s = ""
            else:
                r = compare_hits(hits, ref)
                cmp_s = f" vs strict-brute: id_mismatch={r['id_mismatch']} max_rel_t={r['max_rel_t']:.2e} identical={r['bit_identical']}"
            best = min(ms[5:])
            print(f"{name:24s} {w}x{h} {mname:13s} {best:9.3f} ms  {w * h / best / 1e3:10.1f} Mrays/s  grid={st['trace_grid']} "
                  f"upload={up * 1e3:.0f}ms hits={(hits['id'] != 0xFFFFFFFF).mean():.3f}{cmp_s}", flush=True)


if __name__ == "__main__":
    main()
