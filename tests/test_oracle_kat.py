"""Pin the CPU oracle against the reference's own known-answer tests.

Every test names the reference test it restates (paths relative to /root/reference/bvhtracer/).
All comparisons that are `assert_eq!` in the reference are bit-exact here.
"""
import numpy as np
import pytest

import oracle_lib as O

F = np.float32
S3 = np.sqrt(F(3))


def top_triangle():
    # tests/test_triangle_intersection.rs:11-17
    return np.array([0, F(1) / F(2), 0, -F(1) / S3, -F(1) / F(2), 0, F(1) / S3, -F(1) / F(2), 0], np.float32)


def ray_towards(origin, target):
    d = O.normalize(np.asarray(target, np.float32) - np.asarray(origin, np.float32))
    return O.ray_new(origin, d)


# ---------------------------------------------------------------- cglinalg pins
def test_normalize_and_cross_pinned_by_tri_mesh_normal():
    # tests/test_tri_mesh.rs:47-59 via mesh/decoders.rs:120-123
    v0 = np.array([-0.577350, -0.5, -0.1], np.float32)
    v1 = np.array([0.0, 0.5, -0.1], np.float32)
    v2 = np.array([0.0, 0.5, 0.1], np.float32)
    v0v2 = O.normalize(v2 - v0)
    v0v1 = O.normalize(v1 - v0)
    n = O.normalize(O.cross(v0v2, v0v1))
    expected = np.array([-0.86602545, 0.49999988, -1.7462564e-7], np.float32)
    assert n.tobytes() == expected.tobytes()


def test_tri_decoder_text_case():
    # tests/test_tri_mesh.rs:24-30, 62-72: backslash is whitespace, 2 triangles
    text = ("   \\\n"
            "    0.577350 -0.500000 -0.100000 0.577350 -0.500000  0.100000 -0.577350 -0.500000  0.100000  \\\n"
            "    -0.577350 -0.500000 -0.100000 0.000000  0.500000 -0.100000  0.000000  0.500000  0.100000 \\\n"
            "    ")
    tris = O.parse_tri(text)
    expected = np.array([[0.577350, -0.5, -0.1, 0.577350, -0.5, 0.1, -0.577350, -0.5, 0.1],
                         [-0.577350, -0.5, -0.1, 0.0, 0.5, -0.1, 0.0, 0.5, 0.1]], np.float32)
    assert tris.tobytes() == expected.tobytes()


def test_tri_loader_keeps_sentinel_unity_12583():
    # tri_loader/tests/test_lib.rs:9-15
    tris = O.load_asset("unity.tri")
    assert tris.shape == (12583, 9)
    assert np.all(tris[-1] == 999.0)


def test_tri_comment_and_blank_lines():
    # tri_loader/src/lexer.rs:62-67 (# comments), loader.rs:160-176 (blank lines skipped)
    tris = O.parse_tri("# header\n\n1 2 3 4 5 6 7 8 9\n\n# c\n9 8 7 6 5 4 3 2 1\n")
    assert tris.shape == (2, 9) and tris[1, 0] == 9


def test_obj_decoder_faces():
    tris = O.parse_obj("g q\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\nf 1/1/1 3/1/1 4/1/1\n")
    assert tris.shape == (2, 9)
    assert tris[1].tolist() == [0, 0, 0, 1, 1, 0, 0, 1, 0]


# ---------------------------------------------------------------- Triangle::intersect
def test_triangle_hits_center_and_vertices():
    # tests/test_triangle_intersection.rs:19-65
    tri = top_triangle()
    o = [0, 0, 5]
    assert O.triangle_intersect(tri, ray_towards(o, [0, 0, 0])) is not None
    for k in range(3):
        assert O.triangle_intersect(tri, ray_towards(o, tri[3 * k:3 * k + 3])) is not None


@pytest.mark.parametrize("target_idx,expected", [
    (None, F(5)),                       # :67-79  centre
    (0, np.sqrt(F(101) / F(4))),        # :81-94
    (1, np.sqrt(F(307) / F(12))),       # :96-109
    (2, np.sqrt(F(307) / F(12))),       # :111-124
])
def test_triangle_bit_exact_t(target_idx, expected):
    tri = top_triangle()
    target = [0, 0, 0] if target_idx is None else tri[3 * target_idx:3 * target_idx + 3]
    r = O.triangle_intersect(tri, ray_towards([0, 0, 5], target))
    assert r is not None
    assert F(r[0]).tobytes() == F(expected).tobytes()


def test_triangle_vertex_misses():
    # tests/test_triangle_intersection.rs:126-166
    tri = top_triangle()
    for k, disp in ((0, [0, 0.5, 0]), (1, [0, -0.5, 0]), (2, [0, -0.5, 0])):
        target = tri[3 * k:3 * k + 3] + np.array(disp, np.float32)
        assert O.triangle_intersect(tri, ray_towards([0, 0, 5], target)) is None


@pytest.mark.parametrize("ray_t", [np.finfo(np.float32).max, 100.0, 0.01])
def test_triangle_t_clamped_to_ray_t(ray_t):
    # tests/test_triangle_intersection.rs:168-207 : new_t = min(ray.t, t)
    tri = top_triangle()
    d = O.normalize(np.array([0, 0, -5], np.float32))
    r = O.triangle_intersect(tri, O.ray_new([0, 0, 5], d, ray_t))
    assert r is not None and r[0] <= F(ray_t)


# ---------------------------------------------------------------- Aabb::intersect
BOX = ([-1, -1, -1], [1, 1, 1])


@pytest.mark.parametrize("plane", ["xy", "yz", "zx"])
def test_aabb_rings_miss_and_hit(plane):
    # tests/test_aabb_intersection.rs:29-147 (FRAC_2_PI = 2/pi as in the reference)
    frac_2_pi = F(2) / F(np.pi)
    for i in range(64):
        ang = (F(i) / F(64)) * frac_2_pi
        c, s = F(10) * np.cos(ang, dtype=np.float32), F(10) * np.sin(ang, dtype=np.float32)
        o = {"xy": [c, s, 0], "yz": [0, c, s], "zx": [s, 0, c]}[plane]
        o = np.array(o, np.float32)
        away = O.ray_new(o, O.normalize(o))
        toward = O.ray_new(o, O.normalize(-o))
        assert O.aabb_intersect(*BOX, away) is None
        assert O.aabb_intersect(*BOX, toward) is not None


@pytest.mark.parametrize("origin", [[5, 0, 0], [0, 5, 0], [0, 0, 5], [-5, 0, 0], [0, -5, 0], [0, 0, -5]])
def test_aabb_axis_hits_bit_exact_4(origin):
    # tests/test_aabb_intersection.rs:149-188: zero direction components -> +-inf reciprocals, t == 4.0 exactly
    o = np.array(origin, np.float32)
    r = O.ray_new(o, O.normalize(-o))
    t = O.aabb_intersect(*BOX, r)
    assert t is not None and F(t).tobytes() == F(4).tobytes()


# ---------------------------------------------------------------- BLAS: 1 / 2 / 3 triangles
def test_bvh_one_triangle_layout():
    # src/model/bvh.rs:801-831: root is a leaf {aabb, count 1, first 0}; node 1 default; nodes_used 2
    b = O.Blas(top_triangle().reshape(1, 9))
    assert b.nodes_used == 2
    tri = top_triangle().reshape(3, 3)
    assert b.nodes["min"][0].tolist() == tri.min(axis=0).tolist()
    assert b.nodes["max"][0].tolist() == tri.max(axis=0).tolist()
    assert b.nodes["prim_count"][0] == 1 and b.nodes["left_first"][0] == 0
    assert b.nodes[1].tobytes() == bytes(32)


def test_bvh_one_triangle_kats():
    # tests/test_bvh_one_triangle.rs:34-180 (same KATs through ModelInstance::intersect)
    tri = top_triangle()
    b = O.Blas(tri.reshape(1, 9))
    o = [0, 0, 5]
    exp = [(None, F(5)), (0, np.sqrt(F(101) / F(4))), (1, np.sqrt(F(307) / F(12))), (2, np.sqrt(F(307) / F(12)))]
    for k, e in exp:
        target = [0, 0, 0] if k is None else tri[3 * k:3 * k + 3]
        h = b.intersect(ray_towards(o, target))
        assert h is not None and F(h["t"]).tobytes() == F(e).tobytes() and h["id"] == 0
    for k, disp in ((0, [0, 0.5, 0]), (1, [0, -0.5, 0]), (2, [0, -0.5, 0])):
        target = tri[3 * k:3 * k + 3] + np.array(disp, np.float32)
        assert b.intersect(ray_towards(o, target)) is None


def test_bvh_two_triangles():
    # tests/test_bvh_two_triangles.rs:15-85
    tri = top_triangle().reshape(3, 3)
    d = np.array([5, 0, 0], np.float32)
    b = O.Blas(np.stack([(tri - d).reshape(9), (tri + d).reshape(9)]))
    for k in range(2):
        c = O.triangle_centroid(b.tris[k])
        assert b.intersect(ray_towards([0, 0, 5], c)) is not None
    assert b.intersect(ray_towards([0, 0, 5], [0, 0, 0])) is None


def test_bvh_three_triangles():
    # tests/test_bvh_three_triangles.rs:16-62
    tri = top_triangle().reshape(3, 3)
    d = np.array([0, 10, 0], np.float32)
    b = O.Blas(np.stack([tri.reshape(9), (tri - d).reshape(9), (tri + d).reshape(9)]))
    for k in range(3):
        c = O.triangle_centroid(b.tris[k])
        r = ray_towards([0, 0, 10], c)
        assert O.triangle_intersect(b.tris[k], r) is not None
        assert b.intersect(r) is not None


# ---------------------------------------------------------------- closest hit
def stacked_scene():
    # tests/test_bvh_closest_intersection.rs:25-43
    top = top_triangle().reshape(3, 3)
    tris = [(top - np.array([0, 0, F(i)], np.float32)).reshape(9) for i in range(100)]
    return O.Blas(np.stack(tris))


def test_closest_hit_equals_top_triangle():
    # tests/test_bvh_closest_intersection.rs:61-77
    b = stacked_scene()
    top = top_triangle()
    r = ray_towards([0, 0, 5], O.triangle_centroid(top))
    h = b.intersect(r)
    e = O.triangle_intersect(top, r)
    assert h is not None and e is not None and F(h["t"]).tobytes() == F(e[0]).tobytes()


def test_closest_hit_is_brute_force_min():
    # tests/test_bvh_closest_intersection.rs:79-118
    b = stacked_scene()
    top = top_triangle()
    c = O.triangle_centroid(top)
    r = ray_towards([c[0], c[1], 5], c)
    all_hits = [O.triangle_intersect(t, r) for t in b.tris]
    assert all(h is not None for h in all_hits)
    tmin = min(F(h[0]) for h in all_hits)
    h = b.intersect(r)
    assert F(h["t"]).tobytes() == F(tmin).tobytes()
    assert F(tmin).tobytes() == F(O.triangle_intersect(top, r)[0]).tobytes()


# ---------------------------------------------------------------- refit + structure
def diagonal_bvh():
    # tests/test_bvh_refit.rs:11-33 == src/model/bvh.rs:553-574
    t0 = top_triangle().reshape(3, 3)
    dx = np.array([5, 0, 0], np.float32)
    dy = np.array([0, 5, 0], np.float32)
    tris = [((t0 + F(i) * dx) + F(i) * dy).reshape(9) for i in range(-100, 100)]
    return O.Blas(np.stack(tris).astype(np.float32))


def test_refit_identity_on_same_mesh():
    # tests/test_bvh_refit.rs:47-54
    b = diagonal_bvh()
    before = b.nodes.tobytes()
    b.refit()
    assert b.nodes.tobytes() == before


def test_refit_idempotent_after_animation():
    # tests/test_bvh_refit.rs:35-65
    b = diagonal_bvh()
    b.tris[:, 0] += F(0.3)
    b.tris[:, 4] += F(0.3)
    b.tris[:, 8] += F(0.3)
    b.refit()
    once = b.nodes.tobytes()
    b.refit()
    assert b.nodes.tobytes() == once


def test_bvh_structure_invariants():
    # src/model/bvh.rs:576-643, 720-723
    assert O.BVH_NODE.itemsize == 32
    b = diagonal_bvh()
    n = b.n_tris
    assert len(b.nodes) == 2 * n and b.nodes_used < 2 * n
    assert b.nodes[1].tobytes() == bytes(32)                       # dummy node
    for i in range(b.nodes_used):
        if i == 1:
            continue
        assert b.nodes[i].tobytes() != bytes(32)                   # used nodes are not default
        if b.nodes["prim_count"][i] == 0:
            l = int(b.nodes["left_first"][i])
            assert l > i and l + 1 > i                             # children after parents, left < right
    assert all(b.nodes[i].tobytes() == bytes(32) for i in range(b.nodes_used, 2 * n))
    before1 = b.nodes[1].tobytes()
    b.refit()
    assert b.nodes[1].tobytes() == before1                         # :635-642


# ---------------------------------------------------------------- camera
def test_camera_box_corner_rays_f32_analogue():
    # tests/test_camera.rs:14-174 (camera1) and :294-456 (camera2), restated in f32
    cam1 = O.camera_box(-4, 4, -3, 3, 1, [0, 0, -5], [0, 0, 1], [1, 0, 0], [0, -1, 0])
    cam2 = O.camera_box(-4, 4, -3, 3, 1, [0, 0, 5], [0, 0, -1], [1, 0, 0], [0, 1, 0])
    cases1 = {(0, 0): [-4, -3, -4], (0, 1): [-4, 3, -4], (1, 0): [4, -3, -4], (1, 1): [4, 3, -4]}
    cases2 = {(0, 0): [-4, 3, 4], (0, 1): [-4, -3, 4], (1, 0): [4, 3, 4], (1, 1): [4, -3, 4]}
    for cam, pos, cases, fwd in ((cam1, [0, 0, -5], cases1, [0, 0, 1]), (cam2, [0, 0, 5], cases2, [0, 0, -1])):
        assert cam["tl"][0].tolist() == [-4, 3, -1]
        assert cam["tr"][0].tolist() == [4, 3, -1]
        assert cam["bl"][0].tolist() == [-4, -3, -1]
        for (u, v), target in cases.items():
            got = O.camera_ray_world(cam, u, v)
            exp = ray_towards(pos, target)
            assert got.tobytes() == exp.tobytes()
        got = O.camera_ray_world(cam, 0.5, 0.5)
        exp = O.ray_new(pos, fwd)
        # -0.0 vs +0.0 components compare equal under assert_eq!; compare by value
        assert np.array_equal(got["o"], exp["o"]) and np.array_equal(got["d"], exp["d"])


# ---------------------------------------------------------------- full scenes
def cube_scene():
    # tests/test_scene_cube.rs:28-84
    pos = np.array([0, 4, 0], np.float32)
    fwd = O.normalize(-pos)
    cam = O.camera_box(-1, 1, -1, 1, 1, pos, fwd, [-1, 0, 0], [0, 0, 1])
    blas = O.Blas(O.load_asset("cube.obj"))
    m = O.transform_new_rot_xz([2, 2, 2], [-1, -1, -1], 0.0, 0.0)
    return O.Scene([blas], [(0, m)]), cam


@pytest.mark.parametrize("target,prim", [([0.5, 1.0, -0.5], 6), ([-0.5, 1.0, 0.5], 9)])
def test_scene_cube_ids_and_t(target, prim):
    # tests/test_scene_cube.rs:86-136 and :177-227
    scene, _ = cube_scene()
    h = scene.intersect(ray_towards([0, 4, 0], target))
    assert h is not None
    # approx::assert_relative_eq!(result, expected, epsilon = 1e-7): |a-b| <= eps OR <= max(|a|,|b|) * f32::EPSILON
    expected = np.sqrt(F(19) / F(2))
    diff = abs(float(h["t"]) - float(expected))
    assert diff <= 1e-7 or diff <= max(abs(float(h["t"])), float(expected)) * float(np.finfo(np.float32).eps)
    assert (int(h["id"]) >> 20) == 0
    assert (int(h["id"]) & 0xFFFFF) == prim


def quad_scene():
    # tests/test_scene_quad.rs:52-131
    cam = O.camera_box(-1, 1, -1, 1, 1, [0, 0, 2], [0, 0, -1], [1, 0, 0], [0, 1, 0])
    tris = np.array([[-1, -1, 0, 1, 1, 0, -1, 1, 0], [-1, -1, 0, 1, -1, 0, 1, 1, 0]], np.float32)
    return O.Scene([O.Blas(tris)], [(0, O.mat4_identity())]), cam


def quad_test_case():
    # tests/test_scene_quad.rs:133-197
    o = np.array([0, 0, 2], np.float32)
    tl, tr, bl = (np.array(p, np.float32) for p in ([-1, 1, 0], [1, 1, 0], [-1, -1, 0]))
    half = np.sqrt(F(6)) / F(2)

    def interp(target):
        r = ray_towards(o, target)
        return r["o"][0] + r["d"][0] * half

    tlc, trc, blc = interp(tl), interp(tr), interp(bl)
    dims = np.array([(trc[0] - tlc[0]) / F(2), (tlc[1] - blc[1]) / F(2)], np.float32)
    uv_tl = np.array([(tlc[0] - tl[0]) / F(2), (tl[1] - tlc[1]) / F(2)], np.float32)
    return uv_tl, dims


def test_scene_quad_bit_exact_depths():
    # tests/test_scene_quad.rs:200-261
    scene, cam = quad_scene()
    uv_tl, dims = quad_test_case()
    h = scene.intersect(O.camera_ray_world(cam, 0.5, 0.5))
    assert F(h["t"]).tobytes() == F(2).tobytes()
    s6 = np.sqrt(F(6))
    for u, v in ((uv_tl[0], uv_tl[1]), (uv_tl[0] + dims[0], uv_tl[1]), (uv_tl[0], uv_tl[1] + dims[1]),
                 (uv_tl[0] + dims[0], uv_tl[1] + dims[1])):
        h = scene.intersect(O.camera_ray_world(cam, u, v))
        assert h is not None and F(h["t"]).tobytes() == s6.tobytes()


def test_scene_quad_entire_viewport_mask():
    # tests/test_scene_quad.rs:355-377: 640 x 640 rays, hit iff (u, v) inside the quad's solid angle.
    #
    # 409,584 of 409,600 pixels reproduce the reference's expectation.  The other 16 lie exactly on the quad's top edge
    # (row 160, v == 0.25) or right edge (column 480, u == 0.75), where u+v == 1 in exact arithmetic and the two roundings
    # of f = 1/area and f * X decide (u+v comes out 1 ulp above 1.0 -> miss).  tests/test_quad_edge_variants.py enumerates
    # every f32 variant of cglinalg's dot / cross / normalize consistent with the reference's visible source: none of the 24
    # hits all 642 edge rays (16 or 12 misses), so the upstream expectation cannot be met by the reference's own
    # Triangle::intersect (triangle.rs:53-62).  We pin the variant the other KATs pin: exactly these 16.
    scene, cam = quad_scene()
    uv_tl, dims = quad_test_case()
    assert uv_tl.tolist() == [0.25, 0.25] and dims.tolist() == [0.5, 0.5]
    w = h = 640
    hits = scene.render(cam, w, h).reshape(h, w)
    u = (np.arange(w, dtype=np.float32) / F(w))[None, :]
    v = (np.arange(h, dtype=np.float32) / F(h))[:, None]
    inside = (u >= uv_tl[0]) & (u <= uv_tl[0] + dims[0]) & (v >= uv_tl[1]) & (v <= uv_tl[1] + dims[1])
    got = hits["id"] != O.MISS_ID
    assert np.array_equal(hits["t"] < O.FLT_MAX, got)
    bad = np.argwhere(got != inside)
    assert len(bad) == 16
    assert all((r == 160 or c == 480) for r, c in bad)          # only exact-edge rays
    assert not got[bad[:, 0], bad[:, 1]].any()                   # all are "expected hit, computed miss"
    interior = inside.copy()
    interior[160, :] = False
    interior[:, 480] = False
    assert got[interior].all() and not got[~inside].any()


def test_render_threads_and_regions_agree():
    scene, cam = cube_scene()
    a = scene.render(cam, 64, 48, threads=1)
    b = scene.render(cam, 64, 48, threads=4)
    assert a.tobytes() == b.tobytes()
    c = scene.render(cam, 64, 48, region=(0, 0, 64, 24))
    c = scene.render(cam, 64, 48, region=(0, 24, 64, 48), out=c)
    assert a.tobytes() == c.tobytes()


def test_tri_decoder_normals_kat():
    # tests/test_tri_mesh.rs:24-72: the decoder's per-vertex normals of the two-triangle case, bit-exact
    tris = np.array([[0.577350, -0.5, -0.1, 0.577350, -0.5, 0.1, -0.577350, -0.5, 0.1],
                     [-0.577350, -0.5, -0.1, 0.0, 0.5, -0.1, 0.0, 0.5, 0.1]], np.float32)
    n = O.tri_normals(tris)
    assert n[0].reshape(3, 3).tolist() == [[0.0, 1.0, 0.0]] * 3
    expected = np.array([-0.86602545, 0.49999988, -1.7462564e-7], np.float32)
    for v in range(3):
        assert n[1, 3 * v:3 * v + 3].tobytes() == expected.tobytes()


def test_texture_accumulator_restatement():
    # renderer.rs:297-333 + materials/material.rs:44-52, restated independently in numpy: barycentric mix of the three
    # coordinates, `as usize` (saturating; negative -> 0) then `% size`, texel * (1 / 256), RadianceToRgbShader.
    rng = np.random.default_rng(3)
    n_prims, tw, th = 5, 7, 4
    tc = rng.uniform(-1.5, 3.0, (n_prims, 6)).astype(np.float32)
    tc[4] = [1e25, 0.5, 1e25, 0.5, 1e25, 0.5]                     # uv.x * width >= 2^64: usize::MAX % width
    tex = rng.integers(0, 256, (th, tw, 3)).astype(np.uint8)
    n = 400
    hits = np.zeros(n, O.HIT)
    hits["u"] = rng.uniform(0, 1, n).astype(np.float32)
    hits["v"] = (rng.uniform(0, 1, n) * (1 - hits["u"])).astype(np.float32)
    hits["t"] = 1.0
    hits["id"] = rng.integers(0, n_prims, n).astype(np.uint32) | (rng.integers(0, 3, n).astype(np.uint32) << 20)
    hits["id"][::9] = O.MISS_ID
    hits["id"][5] = 77                                            # primitive index past object 0's mesh: black here, a panic there
    got = O.shade_texture(tc, tex, hits)
    F = np.float32
    for i in range(n):
        rgb = (0, 0, 0)
        prim = int(hits["id"][i]) & 0xFFFFF
        if hits["id"][i] != O.MISS_ID and prim < n_prims:
            u, v = F(hits["u"][i]), F(hits["v"][i])
            w0 = F(F(1) - u) - v
            idx = []
            for k, size in ((0, tw), (1, th)):
                with np.errstate(over="ignore"):
                    c = F(F(F(tc[prim, k] * w0) + F(tc[prim, 2 + k] * u)) + F(tc[prim, 4 + k] * v))
                    x = float(F(c * F(size)))
                xi = 0 if not x > 0 else (2 ** 64 - 1 if x >= 2.0 ** 64 else int(x))
                idx.append(xi % size)
            px = tex[idx[1], idx[0]]
            rgb = tuple(int(F(255) * F(F(p) * F(1 / 256))) for p in px)
        exp = rgb[0] | (rgb[1] << 8) | (rgb[2] << 16) | 0xFF000000
        assert got[i] == exp, (i, hex(got[i]), hex(exp))
    assert len(np.unique(got)) > 10
    # whole-texel identity: a quad with uv == barycentric position reads texel (x, y) back as floor(255 * p / 256)
    one = np.zeros(1, O.HIT); one["u"] = 0.0; one["v"] = 0.0; one["id"] = 0; one["t"] = 1.0
    tc0 = np.zeros((1, 6), np.float32); tc0[0, 0:2] = [(3 + 0.5) / tw, (2 + 0.5) / th]
    px = tex[2, 3].astype(int)
    exp = sum(int(255 * p / 256) << (8 * k) for k, p in enumerate(px)) | 0xFF000000
    assert O.shade_texture(tc0, tex, one)[0] == exp
