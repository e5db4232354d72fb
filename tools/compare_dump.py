#!/usr/bin/env python3
"""Diff the reference's own answers (rust/bvhtracer/examples/dump_hits.rs, run wherever cargo exists) with this repository's.

    python tools/compare_dump.py <dump_dir> [--against gpu|oracle] [--mode strict-accel]
    python tools/compare_dump.py --write-oracle <dump_dir> [case ...]      # same files, written by the CPU oracle (plumbing test)

Every `<scene>_f<frame>_<W>x<H>.hits` file holds W * H records {t, u, v: f32, id: u32} (= bvht_hit).  The same case is traced
here -- through the C ABI on the GPU (default) or with the CPU oracle -- and compared byte for byte; differing records are
broken down by field.  Exit status 1 when any record differs.  This is the check that turns "parity against the restated
oracle" into "parity against the Rust reference": it needs a machine with cargo for the dump and one with a GPU for the trace.
"""
import argparse
import glob
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from bvhtracer_b200 import _ffi, examples  # noqa: E402

HIT = _ffi.HIT
DEFAULT_CASES = ["cube:0:640x640", "quad:0:640x640", "two_armadillos:0:480x270", "two_armadillos:1:480x270",
                 "sixteen_armadillos:0:480x270", "sixteen_armadillos:1:480x270", "sixteen_armadillos:37:480x270",
                 "trippy_teapots:10:480x270", "big_ben_clock:3:480x270"]


def spec_for(scene, frame):
    if scene == "cube":
        return examples.cube()
    if scene == "quad":
        return examples.quad()
    if scene == "two_armadillos":
        return examples.two_armadillos("initial" if frame == 0 else "canonical")
    if scene == "sixteen_armadillos":
        return examples.sixteen_armadillos(frame)
    if scene == "trippy_teapots":
        return examples.trippy_teapots(frame)
    if scene == "big_ben_clock":
        return examples.big_ben_clock()
    raise SystemExit(f"unknown scene {scene}")


def oracle_hits(scene, frame, w, h):
    import oracle_lib as O
    import scene_build as SB
    spec = spec_for(scene, frame)
    if scene == "big_ben_clock":
        blas = O.Blas(O.load_asset("bigben.tri"))
        sc = O.Scene([blas], [(0, O.mat4_identity())], with_transform=False)
        _, cam = SB.oracle_scene(spec)
        anim = examples.BigBenAnimation(blas.tris)
        for _ in range(frame):
            blas.tris[:] = anim.animate()
        blas.refit()
        sc.refresh_blas()
    else:
        sc, cam = SB.oracle_scene(spec)
    return sc.render(cam, w, h, threads=max(1, O.max_threads()))


def gpu_hits(scene, frame, w, h, flags):
    from bvhtracer_b200 import host
    spec = spec_for(scene, frame)
    sc, models = host.build_scene(spec)
    r = host.Renderer(flags=flags)
    if scene == "big_ben_clock":
        anim = examples.BigBenAnimation(models[0].primitives())
        for _ in range(frame):
            models[0].set_primitives(anim.animate())
        if frame:
            models[0].refit()
    st = host.RendererState(host.depth_pipeline(), w, h, keep_hits=True)
    r.render(st, sc)
    return st.hits().copy()


def parse_name(path):
    m = re.match(r"(.+)_f(\d+)_(\d+)x(\d+)\.hits$", os.path.basename(path))
    if not m:
        return None
    return m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))


def ulps(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dump_dir")
    ap.add_argument("cases", nargs="*")
    ap.add_argument("--against", default="gpu", choices=["gpu", "oracle"])
    ap.add_argument("--mode", default="strict-accel", choices=["strict-brute", "strict-accel", "fast-brute", "fast-accel"])
    ap.add_argument("--write-oracle", action="store_true")
    a = ap.parse_args()
    if a.write_oracle:
        os.makedirs(a.dump_dir, exist_ok=True)
        for case in (a.cases or DEFAULT_CASES):
            scene, frame, size = case.split(":")
            w, h = (int(x) for x in size.split("x"))
            hits = oracle_hits(scene, int(frame), w, h)
            p = os.path.join(a.dump_dir, f"{scene}_f{frame}_{w}x{h}.hits")
            hits.tofile(p)
            print(f"{p}: {w} x {h} rays, {(hits['id'] != 0xFFFFFFFF).sum()} hits")
        return 0
    flags = {"strict-brute": 0, "strict-accel": 2, "fast-brute": 1, "fast-accel": 3}[a.mode]
    files = sorted(glob.glob(os.path.join(a.dump_dir, "*.hits")))
    if not files:
        raise SystemExit(f"no .hits files in {a.dump_dir}")
    bad_total = 0
    for f in files:
        meta = parse_name(f)
        if meta is None:
            continue
        scene, frame, w, h = meta
        ref = np.fromfile(f, dtype=HIT)
        if ref.size != w * h:
            raise SystemExit(f"{f}: {ref.size} records, expected {w * h}")
        got = oracle_hits(scene, frame, w, h) if a.against == "oracle" else gpu_hits(scene, frame, w, h, flags)
        got = np.asarray(got).reshape(-1)
        diff = (got.view(np.uint8).reshape(-1, 16) != ref.view(np.uint8).reshape(-1, 16)).any(axis=1)
        n_bad = int(diff.sum())
        bad_total += n_bad
        line = f"{os.path.basename(f):44s} rays {w * h:9d}  reference hits {(ref['id'] != 0xFFFFFFFF).sum():8d}  differing records {n_bad}"
        if n_bad:
            ids = int((got["id"] != ref["id"]).sum())
            both = (got["id"] == ref["id"]) & (ref["id"] != 0xFFFFFFFF)
            line += (f"  [ids {ids}; same id: max ulp t {int(ulps(got['t'][both], ref['t'][both]).max()) if both.any() else 0}"
                     f" u {int(ulps(got['u'][both], ref['u'][both]).max()) if both.any() else 0}"
                     f" v {int(ulps(got['v'][both], ref['v'][both]).max()) if both.any() else 0}]")
            first = np.flatnonzero(diff)[:5]
            line += "  first at pixels " + ", ".join(f"({i % w},{i // w})" for i in first)
        print(line, flush=True)
    print(f"TOTAL differing records: {bad_total} ({a.against}, {a.mode if a.against == 'gpu' else 'oracle'})")
    return 1 if bad_total else 0


if __name__ == "__main__":
    sys.exit(main())
