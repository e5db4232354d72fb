#!/usr/bin/env python3
"""Exploration helper: BvhBuilder::build_for on the device (bvht_blas_build) vs the host build (the C++ mirror)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from bvhtracer_b200 import Engine, FLAG_LEAF_ACCEL, FLAG_STRICT, host


def main():
    for asset in ("teapot.obj", "bigben.tri", "armadillo.tri", "unity.tri"):
        mesh = host.load_asset_mesh(asset)
        tris = mesh.primitives()                          # file order
        t0 = time.perf_counter(); built = host.ModelBuilder().with_mesh(mesh).build(); t_host = time.perf_counter() - t0
        nodes_used = built.nodes()[1]
        for label, flags in (("tree only", FLAG_STRICT), ("tree + leaf accel", FLAG_STRICT | FLAG_LEAF_ACCEL)):
            with Engine(flags=flags) as eng:
                eng.blas_build(tris)                          # warm-up (allocations)
                wall, dev = [], []
                for _ in range(3):
                    t0 = time.perf_counter(); eng.blas_build(tris); wall.append(time.perf_counter() - t0)
                    dev.append(eng.stats()["last_build_ms"])
                st = eng.stats()
            print(f"{asset:14s} {len(tris):6d} tris nodes_used={nodes_used:4d}  {label:18s} device build {min(dev):7.3f} ms "
                  f"(levels {st['last_build_levels']}), whole bvht_blas_build {min(wall) * 1e3:8.2f} ms | host C++ {t_host * 1e3:7.2f} ms", flush=True)


if __name__ == "__main__":
    main()
