#!/usr/bin/env python3
"""Exploration helper (not the judged bench): device time of the trace kernel for every config and mode."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

import scene_build as SB
from bvhtracer_b200 import Engine, FLAG_FAST, FLAG_LEAF_ACCEL, FLAG_STRICT, examples

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
        scene, cam = SB.oracle_scene(spec)
        fcam = SB.to_ffi_camera(cam)
        ref = None
        for mname, flags in MODES.items():
            with Engine(flags=flags) as eng:
                t0 = time.time()
                SB.upload_scene(eng, scene)
                up = time.time() - t0
                dout = eng.device_alloc(w * h * 16)
                ms = []
                for it in range(4):
                    eng.trace_primary_device(fcam, w, h, 8, None, dout)
                    eng.sync()
                    ms.append(eng.stats()["last_trace_ms"])
                host = np.zeros(w * h, dtype=SB._ffi.HIT)
                eng.memcpy_d2h(host, dout)
                eng.device_free(dout)
                st = eng.stats()
            if ref is None:
                ref = host
                cmp_s = ""
            else:
                r = SB.compare_hits(host, ref)
                cmp_s = f" vs strict-brute: id_mismatch={r['id_mismatch']} max_rel_t={r['max_rel_t']:.2e} identical={r['bit_identical']}"
            best = min(ms[1:])
            print(f"{name:24s} {w}x{h} {mname:13s} {best:9.3f} ms  {w * h / best / 1e3:10.1f} Mrays/s  grid={st['trace_grid']} "
                  f"upload={up * 1e3:.0f}ms hits={(host['id'] != 0xFFFFFFFF).mean():.3f}{cmp_s}", flush=True)


if __name__ == "__main__":
    main()
