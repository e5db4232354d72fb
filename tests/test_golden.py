"""Committed golden fixtures (tests/golden/): the oracle must keep reproducing them on the CPU, and the CUDA path must
reproduce the same bytes through the C ABI.  tests/golden/make_golden.py regenerates them."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
import scene_build as SB
from bvhtracer_b200 import _ffi, examples

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)
import make_golden as MG  # noqa: E402

F = np.float32
NTHREADS = max(1, O.max_threads())


def load_hits():
    z = np.load(os.path.join(GOLDEN, "hits.npz"))
    return {k: z[k].view(O.HIT) for k in z.files}


def load_structures():
    with open(os.path.join(GOLDEN, "structures.json")) as fh:
        return json.load(fh)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------------------------------ CPU: the oracle is pinned
def test_fixture_files_are_consistent():
    hits, st = load_hits(), load_structures()
    assert set(hits) == set(st["frames_sha256"]) and len(hits) == len(MG.FRAMES) + 1
    for k, v in hits.items():
        assert sha(v) == st["frames_sha256"][k], k
        assert (v["id"] != O.MISS_ID).sum() > 100, k


def test_oracle_reproduces_golden_frames():
    hits = load_hits()
    now = MG.frames()
    for k, v in hits.items():
        assert now[k].tobytes() == v.tobytes(), k


def test_oracle_reproduces_golden_structures():
    st = load_structures()
    now = MG.structures()
    assert now["assets"] == st["assets"]
    assert now["tlas"] == st["tlas"]


def test_reference_known_answers_from_fixture():
    with open(os.path.join(GOLDEN, "reference_kats.json")) as fh:
        kat = json.load(fh)
    n = kat["tri_mesh_normal"]
    v0, v1, v2 = (np.array(n[k], F) for k in ("v0", "v1", "v2"))
    got = O.normalize(O.cross(O.normalize(v2 - v0), O.normalize(v1 - v0)))
    assert got.tobytes() == np.array(n["expected_normal"], F).tobytes()
    s3 = np.sqrt(F(3))
    tri = np.array([0, F(1) / F(2), 0, -F(1) / s3, -F(1) / F(2), 0, F(1) / s3, -F(1) / F(2), 0], F)
    exprs = {"5": F(5), "sqrt(101/4)": np.sqrt(F(101) / F(4)), "sqrt(307/12)": np.sqrt(F(307) / F(12)), "sqrt(19/2)": np.sqrt(F(19) / F(2))}
    targets = {"centre": [0, 0, 0], "v0": tri[0:3], "v1": tri[3:6], "v2": tri[6:9]}
    for case in kat["triangle_t"]["cases"]:
        d = O.normalize(np.asarray(targets[case["target"]], F) - np.array(kat["triangle_t"]["origin"], F))
        r = O.triangle_intersect(tri, O.ray_new(kat["triangle_t"]["origin"], d))
        assert r is not None and F(r[0]).tobytes() == F(exprs[case["expected_t_f32_expr"]]).tobytes()
    a = kat["aabb_axis_hits"]
    for o in a["origins"]:
        o = np.array(o, F)
        t = O.aabb_intersect(a["box"][0], a["box"][1], O.ray_new(o, O.normalize(-o)))
        assert F(t).tobytes() == F(a["expected_t"]).tobytes()
    u = O.load_asset("unity.tri")
    assert u.shape[0] == kat["unity_tri_count"]["triangles"] and np.all(u[-1] == kat["unity_tri_count"]["last_triangle_all"])
    scene, _ = SB.oracle_scene(examples.cube())
    c = kat["cube_scene"]
    o = np.array(c["camera_position"], F)
    rays = np.array([list(o) + list(O.normalize(np.array(k["target"], F) - o)) + [O.FLT_MAX] for k in c["cases"]], F)
    hits = scene.trace_rays(rays)
    exp_t = float(exprs[c["expected_t_f32_expr"]])
    for h, k in zip(hits, c["cases"]):
        assert int(h["id"]) & 0xFFFFF == k["primitive_index"] and int(h["id"]) >> 20 == k["instance_index"]
        assert abs(float(h["t"]) - exp_t) <= exp_t * float(np.finfo(F).eps)
    one = O.Blas(tri.reshape(1, 9))
    lay = kat["bvh_one_triangle_layout"]
    assert one.nodes_used == lay["nodes_used"] and one.nodes["prim_count"][0] == lay["root_primitive_count"]
    assert one.nodes["left_first"][0] == lay["root_first_primitive"]


# ------------------------------------------------------------------------------------------ GPU: same bytes through the C ABI
@pytest.mark.gpu
@pytest.mark.parametrize("flags", [_ffi.FLAG_STRICT, _ffi.FLAG_STRICT | _ffi.FLAG_LEAF_ACCEL], ids=["strict-brute", "strict-accel"])
def test_cuda_path_reproduces_golden_frames(flags):
    from bvhtracer_b200 import Engine
    hits = load_hits()
    for name, make, w, h in MG.FRAMES:
        scene, cam = SB.oracle_scene(make())            # host-side scene preparation (BVH/TLAS build) is the oracle's, as in every GPU test
        with Engine(flags=flags) as eng:
            SB.upload_scene(eng, scene)
            got = eng.trace_primary(SB.to_ffi_camera(cam), w, h)
        assert got.tobytes() == hits[name].tobytes(), name


@pytest.mark.gpu
def test_cuda_build_and_refit_reproduce_golden_structures_and_frame():
    from bvhtracer_b200 import Engine
    st, hits = load_structures(), load_hits()
    with Engine(flags=_ffi.FLAG_STRICT | _ffi.FLAG_LEAF_ACCEL) as eng:
        for asset, exp in st["assets"].items():         # BvhBuilder::build_for on the device vs the committed hashes
            bid = eng.blas_build(O.load_asset(asset))
            n, used = eng.blas_info(bid)
            assert (n, used) == (exp["n_tris"], exp["nodes_used"]), asset
            assert sha(eng.blas_read_nodes(bid, used)) == exp["nodes_sha256"], asset
            assert sha(eng.blas_read_triangles(bid, n)) == exp["reordered_tris_sha256"], asset
    # big_ben_clock after two device-side animate + refit steps
    with Engine(flags=_ffi.FLAG_STRICT | _ffi.FLAG_LEAF_ACCEL) as eng:
        bid = eng.blas_build(O.load_asset("bigben.tri"))
        n, used = eng.blas_info(bid)
        anim = examples.BigBenAnimation(eng.blas_read_triangles(bid, n))
        _, cam = SB.oracle_scene(examples.big_ben_clock())
        scene = O.Scene([O.Blas(O.load_asset("bigben.tri"))], [(0, O.mat4_identity())], with_transform=False)
        SB.upload_scene(eng, scene, blas_ids=[bid])
        for _ in range(2):
            eng.blas_update_vertices(bid, anim.animate())
            eng.blas_refit(bid)
        got = eng.trace_primary(SB.to_ffi_camera(cam), 96, 54)
        assert got.tobytes() == hits["big_ben_clock_refit2_96x54"].tobytes()
