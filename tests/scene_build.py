"""Build a SceneSpec (bvhtracer_b200/examples.py) with the CPU oracle, and hand the result to an Engine.

TEST INFRASTRUCTURE: the oracle plays the role of the Rust host (BvhBuilder, Tlas::rebuild, Camera) and of
the checker.  The product's own host mirror is exercised separately (tests/test_host_mirror.py).
"""
import numpy as np

import oracle_lib as O
from bvhtracer_b200 import _ffi, examples

_BLAS_CACHE = {}


def oracle_blas(asset):
    """Built once per asset per process (the armadillo build takes ~0.1 s, copies are cheap)."""
    if asset not in _BLAS_CACHE:
        tris = examples.QUAD_TRIS if asset == "<quad>" else O.load_asset(asset)
        _BLAS_CACHE[asset] = O.Blas(tris)
    return _BLAS_CACHE[asset]


def oracle_camera(c):
    if c.box is not None:
        l, r, b, t = c.box
        return O.camera_box(l, r, b, t, c.near, c.position, c.forward, c.right, c.up)
    return O.camera_symmetric_fov(c.fovy_deg, c.aspect, c.near, c.position, c.forward, c.right, c.up)


def object_matrix(o):
    return O.transform_new_rot_xz(o.scale, o.translation, o.angle_x, o.angle_z)


def oracle_scene(spec):
    blases = [oracle_blas(a) for a in spec.meshes]
    with_transform = all(o.with_transform for o in spec.objects)
    scene = O.Scene(blases, [(o.model, object_matrix(o)) for o in spec.objects], with_transform=with_transform)
    return scene, oracle_camera(spec.camera)


def to_ffi_camera(cam):
    out = np.zeros(1, _ffi.CAMERA)
    out["top_left_eye"] = cam["tl"]
    out["top_right_eye"] = cam["tr"]
    out["bottom_left_eye"] = cam["bl"]
    out["view_matrix_inv"] = cam["view_inv"]
    return out


def upload_scene(engine, scene, blas_ids=None):
    """Upload an oracle-built scene through the C ABI.  Returns the blas ids (reused when given)."""
    if blas_ids is None:
        blas_ids = [engine.blas_create(b.tris, b.nodes.view(_ffi.BVH_NODE), b.nodes_used) for b in scene.blases]
    inst = np.zeros(len(scene.inst), _ffi.INSTANCE)
    inst["transform_inv"] = scene.inst["inv"]
    inst["blas_id"] = [blas_ids[int(b)] for b in scene.inst["blas_id"]]
    engine.tlas_set(scene.tlas.view(_ffi.TLAS_NODE), scene.tlas_used, inst)
    return blas_ids


def compare_hits(gpu, ref):
    """-> dict with id mismatches, max ulp distance of t/u/v over pixels with identical ids."""
    gpu = np.asarray(gpu).reshape(-1)
    ref = np.asarray(ref).reshape(-1)
    same_id = gpu["id"] == ref["id"]
    res = {"n": int(gpu.size), "id_mismatch": int((~same_id).sum())}
    hit = same_id & (ref["id"] != O.MISS_ID)
    for f in ("t", "u", "v"):
        a = gpu[f][hit].view(np.int32).astype(np.int64)
        b = ref[f][hit].view(np.int32).astype(np.int64)
        res["max_ulp_" + f] = int(np.abs(a - b).max()) if a.size else 0
    miss = same_id & (ref["id"] == O.MISS_ID)
    res["miss_t_ok"] = bool(np.all(gpu["t"][miss] == O.FLT_MAX))
    res["bit_identical"] = bool(gpu.tobytes() == ref.tobytes())
    if hit.any():
        rel = np.abs(gpu["t"][hit].astype(np.float64) - ref["t"][hit].astype(np.float64)) / np.abs(ref["t"][hit].astype(np.float64))
        res["max_rel_t"] = float(rel.max())
    else:
        res["max_rel_t"] = 0.0
    return res
