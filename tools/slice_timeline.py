#!/usr/bin/env python3
"""Where does a frame's time go between the warps?  Needs the profiling build (BVHT_SLICE_LOG=1 python -m bvhtracer_b200.build
--force): the trace kernel then logs start and duration (%globaltimer) of every 32-pixel block.  Prints, per config: the span of
the launch, the sum of block times / (warps x span) = how busy the resident warps were, how long the tail is (time after the
first warp ran out of work), and the distribution of block times.

    BVHT_SLICE_LOG=1 python -m bvhtracer_b200.build --force; BVHT_K0=0 python tools/slice_timeline.py; python -m bvhtracer_b200.build --force
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from bvhtracer_b200 import FLAG_LEAF_ACCEL, FLAG_STRICT, examples, host


def main():
    cases = [("sixteen_armadillos", examples.sixteen_armadillos(0), (3840, 2160)),
             ("sixteen_armadillos_f30", examples.sixteen_armadillos(30), (3840, 2160)),
             ("trippy_teapots", examples.trippy_teapots(10), (3840, 2160)),
             ("big_ben_clock", examples.big_ben_clock(), (7680, 4320))]
    only = sys.argv[1:]
    for name, spec, (w, h) in cases:
        if only and name not in only:
            continue
        scene, _ = host.build_scene(spec)
        cam = scene.camera()
        renderer = host.Renderer(flags=FLAG_STRICT | FLAG_LEAF_ACCEL)
        renderer.sync_scene(scene)
        eng = renderer.engine()
        n_items = ((w + 7) // 8) * ((h + 7) // 8) * 2
        dout = eng.device_alloc(w * h * 16)
        dlog = eng.device_alloc(n_items * 16)
        for it in range(3):
            if it == 2:
                os.environ["BVHT_SLICE_LOG_PTR"] = str(dlog)
            eng.trace_primary_device(cam, w, h, 8, None, dout)
            eng.sync()
        os.environ.pop("BVHT_SLICE_LOG_PTR", None)
        ms = eng.stats()["last_trace_ms"]
        grid = eng.stats()["trace_grid"]
        log = np.zeros(n_items * 2, "<u8")
        eng.memcpy_d2h(log, dlog)
        hits = np.zeros(w * h, dtype=[("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("id", "<u4")])
        eng.memcpy_d2h(hits, dout)
        eng.device_free(dlog); eng.device_free(dout)
        # per 8x4 block: does any of its rays hit?  (item = (tile_y * ntx + tile_x) * 2 + half)
        hit_px = (hits["id"] != 0xFFFFFFFF).reshape(h // 8, 2, 4, w // 8, 8)          # ty, half, row, tx, col
        block_hit = hit_px.any(axis=(2, 4)).transpose(0, 2, 1).reshape(-1)             # -> (ty, tx, half) flattened = item order
        del hits, hit_px
        t0 = log[0::2].astype(np.int64)
        dur = (log[1::2] & ((1 << 44) - 1)).astype(np.int64)
        smid = (log[1::2] >> 48).astype(np.int64)
        warp = ((log[1::2] >> 44) & 15).astype(np.int64)
        ok = t0 > 0
        miss_time = float(dur[ok & ~block_hit].sum()); all_time = float(dur[ok].sum())
        n_miss_blocks = int((ok & ~block_hit).sum())
        t0, dur, smid, warp = t0[ok], dur[ok], smid[ok], warp[ok]
        start = t0.min(); end = (t0 + dur).max(); span = end - start
        warps = grid * 4
        # per resident warp slot: identify by (smid, warp-in-cta) is ambiguous across CTAs of an SM; use the timeline instead
        ev = np.concatenate([np.stack([t0, np.ones_like(t0)], 1), np.stack([t0 + dur, -np.ones_like(t0)], 1)])
        ev = ev[np.argsort(ev[:, 0], kind="stable")]
        active = np.cumsum(ev[:, 1])
        tgrid = ev[:, 0]
        # time-weighted mean of blocks in flight, and the moment the number in flight last drops below 90 % / 50 % of its plateau
        dt = np.diff(tgrid)
        mean_active = float((active[:-1] * dt).sum() / max(span, 1))
        plateau = np.percentile(active, 90)
        def last_above(frac):
            idx = np.nonzero(active >= frac * plateau)[0]
            return (tgrid[idx[-1]] - start) / span if idx.size else 0.0
        q = np.percentile(dur, [50, 90, 99, 100]) / 1e3
        print(f"{name:24s} launch {ms:.3f} ms, logged span {span / 1e6:.3f} ms, {len(dur)} blocks traced by K1, grid {grid} ({warps} warps)")
        print(f"    blocks in flight: mean {mean_active:.0f} of {warps} warp slots ({100 * mean_active / warps:.0f} %), plateau(p90) {plateau:.0f}")
        print(f"    in-flight >= 90 % of plateau until {100 * last_above(0.9):.0f} % of the span, >= 50 % until {100 * last_above(0.5):.0f} %")
        print(f"    block time us: median {q[0]:.1f}  p90 {q[1]:.1f}  p99 {q[2]:.1f}  max {q[3]:.1f};  sum {dur.sum() / 1e6:.1f} ms = {dur.sum() / 1e6 / warps:.3f} ms per warp slot")
        print(f"    blocks in which no ray hits anything: {n_miss_blocks} of {len(dur)} traced, {100 * miss_time / all_time:.1f} % of the summed block time")
        per_sm = np.bincount(smid, weights=dur.astype(np.float64), minlength=148) / 1e6
        print(f"    busy time per SM (sum of its blocks' times / 32 warp slots): min {per_sm.min() / 32:.3f}  mean {per_sm.mean() / 32:.3f}  max {per_sm.max() / 32:.3f} ms")
        del eng, renderer


if __name__ == "__main__":
    main()
