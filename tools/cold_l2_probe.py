#!/usr/bin/env python3
"""Exploration helper: where does the cold-L2 penalty of the 4K trace come from?  Times the resident C3 frame (16 B hit
records to HBM) after (a) nothing, (b) a 256 MiB memset (dirty L2, what bench.py does), (c) a 256 MiB read (clean L2),
(d) memset + a tiny low-resolution render that re-touches the scene."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from bvhtracer_b200 import FLAG_LEAF_ACCEL, FLAG_STRICT, examples, host


def main():
    w, h = 3840, 2160
    scene, _ = host.build_scene(examples.sixteen_armadillos(0))
    fcam = scene.camera()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    renderer = host.Renderer(flags=FLAG_STRICT | FLAG_LEAF_ACCEL)
    renderer.set_stream(stream.cuda_stream)
    renderer.sync_scene(scene)
    eng = renderer.engine()
    if True:
        dout = eng.device_alloc(w * h * 16)
        dsmall = eng.device_alloc(128 * 72 * 16)

        def frame():
            eng.trace_primary_device(fcam, w, h, 8, None, dout)

        def timed(pre):
            ms = []
            for _ in range(8):
                pre()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); frame(); e1.record(stream)
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            return np.median(ms[2:]), min(ms[2:])

        def touch():
            buf.zero_()
            eng.trace_primary_device(fcam, 128, 72, 8, None, dsmall)

        for name, pre in (("warm (no flush)", lambda: None), ("memset 256 MiB (dirty L2)", lambda: buf.zero_()),
                          ("read 256 MiB (clean L2)", lambda: buf.sum()), ("memset + 128x72 pre-render", touch)):
            med, best = timed(pre)
            print(f"{name:30s} median {med:7.3f} ms  min {best:7.3f} ms", flush=True)


if __name__ == "__main__":
    main()
