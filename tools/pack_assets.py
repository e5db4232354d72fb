#!/usr/bin/env python3
"""Pack the reference's mesh assets into raw little-endian f32 triangle soups (N x 9 floats) with the PRODUCT's own decoders
(the C++ host mirror's TriMeshDecoder / ObjMeshDecoder, bvhtracer_b200/host/bvhtracer.hpp; mesh/decoders.rs:102-216,
tri_loader/src/{lexer,loader}.rs).  Run where the reference tree exists (the GPU box has none, so the packed files under
assets/ travel); every triangle is kept, the 999-sentinel that ends each .tri asset included; triangle order is file order.

    python tools/pack_assets.py [/root/reference] [--check]      # --check: compare with assets/ instead of writing
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bvhtracer_b200 import host  # noqa: E402

ASSETS = {
    # output name          : path relative to the reference root
    "cube.obj.f32": "examples/assets/cube.obj",
    "teapot.obj.f32": "examples/assets/teapot.obj",
    "armadillo.tri.f32": "examples/assets/armadillo.tri",
    "bigben.tri.f32": "examples/assets/bigben.tri",
    "unity.tri.f32": "bvhtracer/assets/unity.tri",
}


def packed(ref_root):
    """-> {file name under assets/: bytes}"""
    out = {}
    for name, rel in ASSETS.items():
        mesh = host.read_mesh_file(os.path.join(ref_root, rel))
        out[name] = np.ascontiguousarray(mesh.primitives(), "<f4").tobytes()
        if rel.endswith(".obj"):          # the `vn` normals of each face corner (NormalMappingAccumulator; .tri normals are derived)
            out[name.replace(".f32", ".normals.f32")] = np.ascontiguousarray(mesh.normals(), "<f4").tobytes()
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ref = args[0] if args else "/root/reference"
    check = "--check" in sys.argv
    bad = 0
    for name, blob in packed(ref).items():
        path = os.path.join(ROOT, "assets", name)
        if check:
            same = os.path.exists(path) and open(path, "rb").read() == blob
            print(f"{name}: {len(blob)} bytes {'identical' if same else 'DIFFERS'}")
            bad += not same
        else:
            open(path, "wb").write(blob)
            print(f"{name}: {len(blob)} bytes written")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
