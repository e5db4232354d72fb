#!/usr/bin/env python3
"""Pack the reference's mesh assets into raw little-endian f32 triangle soups (N x 9 floats).

Run HERE (the dev container), where /root/reference exists; the GPU box has no /root/reference, so the
packed files under assets/ are what tests, smoke() and bench.py read.  Decoding is done by the oracle's
restatement of the reference decoders (tri_loader/src/{lexer,loader}.rs; mesh/decoders.rs:108-133,157-215):
.tri values go text -> f32 (strtof), .obj values go text -> f64 -> f32, and every triangle is kept
(including the 999-sentinel that ends each .tri asset).  Triangle order is file order (pre-BVH).

    python oracle/tools/pack_assets.py [/root/reference]
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

ASSETS = {
    # output name          : path relative to the reference root
    "cube.obj.f32": "examples/assets/cube.obj",
    "teapot.obj.f32": "examples/assets/teapot.obj",
    "armadillo.tri.f32": "examples/assets/armadillo.tri",
    "bigben.tri.f32": "examples/assets/bigben.tri",
    "unity.tri.f32": "bvhtracer/assets/unity.tri",
}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    lib.orc_load_mesh_file.restype = ctypes.c_int64
    lib.orc_load_mesh_file.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_float))]
    lib.orc_free.argtypes = [ctypes.c_void_p]
    os.makedirs(os.path.join(ROOT, "assets"), exist_ok=True)
    for out_name, rel in ASSETS.items():
        p = ctypes.POINTER(ctypes.c_float)()
        n = lib.orc_load_mesh_file(os.path.join(ref, rel).encode(), ctypes.byref(p))
        if n < 0:
            raise SystemExit(f"decode failed for {rel}: {n}")
        arr = np.ctypeslib.as_array(p, shape=(n, 9)).astype("<f4").copy()
        lib.orc_free(p)
        arr.tofile(os.path.join(ROOT, "assets", out_name))
        print(f"{out_name}: {n} triangles, {arr.nbytes} bytes")
        if rel.endswith(".obj"):
            # per-vertex normals of the OBJ models (NormalMappingAccumulator needs them; .tri normals are derived from positions)
            text = open(os.path.join(ref, rel), "rb").read()
            q = ctypes.POINTER(ctypes.c_float)()
            lib.orc_parse_obj_normals.restype = ctypes.c_int64
            lib.orc_parse_obj_normals.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.POINTER(ctypes.c_float))]
            m = lib.orc_parse_obj_normals(text, len(text), ctypes.byref(q))
            assert m == n, (m, n)
            nrm = np.ctypeslib.as_array(q, shape=(m, 9)).astype("<f4").copy()
            lib.orc_free(q)
            nrm.tofile(os.path.join(ROOT, "assets", out_name.replace(".f32", ".normals.f32")))
            print(f"{out_name.replace('.f32', '.normals.f32')}: {m} x 9 normals")


if __name__ == "__main__":
    main()
