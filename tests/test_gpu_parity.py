"""GPU parity tests: the CUDA path (through the C ABI of include/bvht.h) against the CPU oracle.

Bar (BASELINE.json north_star): strict mode -- hit ids bit-exact, t within 1 ulp (we require bit-identical
records); fast mode -- >= 99.99 % identical ids, t within 1e-5 relative.
Sizes are chosen so the oracle finishes in seconds; full-size runs compare the leaf accelerator against the
brute-force leaves on the GPU itself (both strict), which is size-independent evidence of equivalence.
"""
import numpy as np
import pytest

import oracle_lib as O
import scene_build as SB
from bvhtracer_b200 import Engine, BvhtError, _ffi, examples
from bvhtracer_b200 import FLAG_FAST, FLAG_LEAF_ACCEL, FLAG_STRICT

pytestmark = pytest.mark.gpu

F = np.float32
NTHREADS = max(1, O.max_threads())

STRICT_MODES = [FLAG_STRICT, FLAG_STRICT | FLAG_LEAF_ACCEL]
MODE_IDS = ["strict-brute", "strict-accel"]


def render_both(spec, w, h, flags, tile=8):
    scene, cam = SB.oracle_scene(spec)
    ref = scene.render(cam, w, h, tile=tile, threads=NTHREADS)
    with Engine(flags=flags) as eng:
        SB.upload_scene(eng, scene)
        got = eng.trace_primary(SB.to_ffi_camera(cam), w, h, tile=tile)
    return got, ref


def assert_strict(got, ref):
    r = SB.compare_hits(got, ref)
    assert r["id_mismatch"] == 0, r
    assert r["max_ulp_t"] == 0 and r["max_ulp_u"] == 0 and r["max_ulp_v"] == 0, r
    assert r["bit_identical"], r


def assert_fast(got, ref):
    r = SB.compare_hits(got, ref)
    assert r["id_mismatch"] <= 1e-4 * r["n"], r
    assert r["max_rel_t"] <= 1e-5, r


# ------------------------------------------------------------------------------------------ C1 / quad
@pytest.mark.parametrize("flags", STRICT_MODES, ids=MODE_IDS)
def test_cube_640_bit_exact(flags):
    got, ref = render_both(examples.cube(), 640, 640, flags)
    assert (ref["id"] != O.MISS_ID).sum() > 1000
    assert_strict(got, ref)


@pytest.mark.parametrize("flags", STRICT_MODES, ids=MODE_IDS)
def test_cube_golden_ids_through_abi(flags):
    # bvhtracer/tests/test_scene_cube.rs:97-136, 188-227 through bvht_trace_rays
    scene, _ = SB.oracle_scene(examples.cube())
    o = np.array([0, 4, 0], F)
    rays = []
    for target in ([0.5, 1.0, -0.5], [-0.5, 1.0, 0.5]):
        d = O.normalize(np.array(target, F) - o)
        rays.append(list(o) + list(d) + [O.FLT_MAX])
    with Engine(flags=flags) as eng:
        SB.upload_scene(eng, scene)
        hits = eng.trace_rays(np.array(rays, F))
    assert [int(h) & 0xFFFFF for h in hits["id"]] == [6, 9]
    assert [int(h) >> 20 for h in hits["id"]] == [0, 0]
    exp = np.sqrt(F(19) / F(2))
    assert all(abs(float(t) - float(exp)) <= float(exp) * float(np.finfo(F).eps) for t in hits["t"])


@pytest.mark.parametrize("flags", STRICT_MODES, ids=MODE_IDS)
def test_quad_viewport_640(flags):
    # bvhtracer/tests/test_scene_quad.rs:355-377 (see tests/test_oracle_kat.py for the 16 exact-edge pixels)
    got, ref = render_both(examples.quad(), 640, 640, flags)
    assert_strict(got, ref)
    g = got.reshape(640, 640)
    assert F(g["t"][320, 320]).tobytes() == F(2).tobytes()          # test_scene_quad.rs:201-209
    inside = np.zeros((640, 640), bool)
    inside[161:481, 160:480] = True                                  # interior of the quad's solid angle
    assert (g["id"][inside] != O.MISS_ID).all()
    outside = np.ones((640, 640), bool)
    outside[160:481, 160:481] = False
    assert (g["id"][outside] == O.MISS_ID).all()
    # the two exact-edge lines (row 160: v == 0.25, column 480: u == 0.75), ray by ray: 16 of their 642 rays miss -- in f32 no
    # arithmetic consistent with the reference's Triangle::intersect hits them all (tests/test_quad_edge_variants.py)
    import test_quad_edge_variants as QV
    pinned = QV.hits_for_variant(QV.edge_rays(), "div", "lr", "plain", "f_times")
    got_edges = np.concatenate([g["id"][160, 160:481], g["id"][160:481, 480]]) != O.MISS_ID
    assert np.array_equal(got_edges, pinned) and int((~got_edges).sum()) == 16
    assert (g["id"][160:481, 160] != O.MISS_ID).all() and (g["id"][480, 160:481] != O.MISS_ID).all()     # left / bottom edges: all hit


def test_triangle_and_aabb_kats_through_abi():
    # bvhtracer/tests/test_bvh_one_triangle.rs:82-139 (bit-exact t), test_triangle_intersection.rs:126-166 (misses)
    s3 = np.sqrt(F(3))
    tri = np.array([[0, 0.5, 0, -F(1) / s3, -0.5, 0, F(1) / s3, -0.5, 0]], F)
    blas = O.Blas(tri)
    scene = O.Scene([blas], [(0, O.mat4_identity())])
    o = np.array([0, 0, 5], F)
    targets = [[0, 0, 0], tri[0, 0:3], tri[0, 3:6], tri[0, 6:9],
               tri[0, 0:3] + np.array([0, 0.5, 0], F), tri[0, 3:6] + np.array([0, -0.5, 0], F)]
    rays = np.array([list(o) + list(O.normalize(np.array(t, F) - o)) + [O.FLT_MAX] for t in targets], F)
    with Engine() as eng:
        SB.upload_scene(eng, scene)
        hits = eng.trace_rays(rays)
    exp = [F(5), np.sqrt(F(101) / F(4)), np.sqrt(F(307) / F(12)), np.sqrt(F(307) / F(12))]
    for h, e in zip(hits[:4], exp):
        assert F(h["t"]).tobytes() == F(e).tobytes() and h["id"] == 0
    assert all(h["id"] == O.MISS_ID and h["t"] == O.FLT_MAX for h in hits[4:])


# ------------------------------------------------------------------------------------------ C2
@pytest.mark.parametrize("flags", STRICT_MODES, ids=MODE_IDS)
@pytest.mark.parametrize("frame", ["canonical", "initial"])
def test_two_armadillos_bit_exact(frame, flags):
    # "initial" has both instances coincident: every hit is an exact tie between instances (order matters)
    got, ref = render_both(examples.two_armadillos(frame), 384, 216, flags)
    assert (ref["id"] != O.MISS_ID).mean() > 0.05
    assert_strict(got, ref)


# ------------------------------------------------------------------------------------------ C3 / C4
@pytest.mark.parametrize("flags", STRICT_MODES, ids=MODE_IDS)
@pytest.mark.parametrize("frame", [0, 1, 37])
def test_sixteen_armadillos_bit_exact(frame, flags):
    got, ref = render_both(examples.sixteen_armadillos(frame), 320, 180, flags)
    assert (ref["id"] != O.MISS_ID).mean() > 0.05
    assert_strict(got, ref)


@pytest.mark.parametrize("flags", STRICT_MODES, ids=MODE_IDS)
@pytest.mark.parametrize("frame", [0, 25])
def test_trippy_teapots_bit_exact(frame, flags):
    got, ref = render_both(examples.trippy_teapots(frame), 960, 540, flags)
    assert (ref["id"] != O.MISS_ID).mean() > 0.02
    assert_strict(got, ref)


# ------------------------------------------------------------------------------------------ C5: refit
@pytest.mark.parametrize("flags", STRICT_MODES, ids=MODE_IDS)
def test_big_ben_animate_refit_trace(flags):
    spec = examples.big_ben_clock()
    scene, cam = SB.oracle_scene(spec)
    blas = O.Blas(O.load_asset("bigben.tri"))              # private copy: vertices get animated
    scene = O.Scene([blas], [(0, O.mat4_identity())], with_transform=False)
    anim = examples.BigBenAnimation(blas.tris)
    with Engine(flags=flags) as eng:
        ids = SB.upload_scene(eng, scene)
        for frame in range(3):
            verts = anim.animate()
            blas.tris[:] = verts
            blas.refit()                                   # oracle: Bvh::refit
            scene.refresh_blas()
            eng.blas_update_vertices(ids[0], verts)
            eng.blas_refit(ids[0])                         # device: K2
            nodes = eng.blas_read_nodes(ids[0], blas.nodes_used)
            ref_nodes = blas.nodes[:blas.nodes_used]
            assert np.array_equal(nodes["aabb_min"], ref_nodes["min"])
            assert np.array_equal(nodes["aabb_max"], ref_nodes["max"])
            assert np.array_equal(nodes["prim_count"], ref_nodes["prim_count"])
            assert np.array_equal(nodes["left_first"], ref_nodes["left_first"])
            ref = scene.render(cam, 256, 144, threads=NTHREADS)
            got = eng.trace_primary(SB.to_ffi_camera(cam), 256, 144)
            assert_strict(got, ref)


def test_refit_identity_and_idempotent_on_device():
    # bvhtracer/tests/test_bvh_refit.rs:47-65 through the C ABI
    s3 = np.sqrt(F(3))
    t0 = np.array([[0, 0.5, 0], [-F(1) / s3, -0.5, 0], [F(1) / s3, -0.5, 0]], F)
    tris = np.stack([((t0 + F(i) * np.array([5, 0, 0], F)) + F(i) * np.array([0, 5, 0], F)).reshape(9) for i in range(-100, 100)])
    blas = O.Blas(tris.astype(F))
    with Engine() as eng:
        bid = eng.blas_create(blas.tris, blas.nodes.view(_ffi.BVH_NODE), blas.nodes_used)
        eng.blas_refit(bid)
        n1 = eng.blas_read_nodes(bid, blas.nodes_used)
        assert n1.tobytes() == blas.nodes[:blas.nodes_used].tobytes()       # unchanged on the same mesh
        moved = blas.tris.copy()
        moved[:, 0] += F(0.3); moved[:, 4] += F(0.3); moved[:, 8] += F(0.3)
        eng.blas_update_vertices(bid, moved)
        eng.blas_refit(bid)
        a = eng.blas_read_nodes(bid, blas.nodes_used)
        eng.blas_refit(bid)
        b = eng.blas_read_nodes(bid, blas.nodes_used)
        assert a.tobytes() == b.tobytes()                                    # idempotent
        blas.tris[:] = moved
        blas.refit()
        assert np.array_equal(a["aabb_min"], blas.nodes["min"][:blas.nodes_used])
        assert np.array_equal(a["aabb_max"], blas.nodes["max"][:blas.nodes_used])
        assert a[1].tobytes() == bytes(32)                                   # node 1 untouched (bvh.rs:635-642)


# ------------------------------------------------------------------------------------------ arbitrary rays
@pytest.mark.parametrize("flags", STRICT_MODES, ids=MODE_IDS)
def test_random_rays_sixteen_armadillos(flags):
    scene, _ = SB.oracle_scene(examples.sixteen_armadillos(12))
    rng = np.random.default_rng(1234)
    n = 20000
    o = rng.uniform(-8, 8, (n, 3)).astype(F)
    target = rng.uniform(-4, 4, (n, 3)).astype(F)
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(F)
    t = np.where(rng.random(n) < 0.2, rng.uniform(0.5, 12.0, n), O.FLT_MAX).astype(F)
    rays = np.concatenate([o, d.astype(F), t[:, None]], axis=1).astype(F)
    # axis-aligned directions: zero components -> +-inf reciprocals (aabb.rs KATs rely on this)
    rays[:60, 3:6] = 0
    for i in range(60):
        rays[i, 3 + (i % 3)] = 1.0 if (i // 3) % 2 == 0 else -1.0
    ref = scene.trace_rays(rays, threads=NTHREADS)
    with Engine(flags=flags) as eng:
        SB.upload_scene(eng, scene)
        got = eng.trace_rays(rays)
    assert (ref["id"] != O.MISS_ID).sum() > 500
    assert_strict(got, ref)


def test_ray_batches_rebake_the_leaf_accelerator():
    # Ray batches far outside the default |o| / |d| limits, with un-normalised directions: the accelerator is re-baked for
    # the batch (one reduction kernel), results stay bit-identical, and alternating with primary frames stays exact too.
    scene, cam = SB.oracle_scene(examples.sixteen_armadillos(5))
    rng = np.random.default_rng(99)
    n = 6000
    o = (rng.normal(size=(n, 3)) * 300.0).astype(F)
    target = rng.uniform(-4, 4, (n, 3)).astype(F)
    d = ((target - o) * rng.uniform(0.01, 40.0, (n, 1))).astype(F)
    rays = np.concatenate([o, d, np.full((n, 1), O.FLT_MAX, F)], axis=1).astype(F)
    near = rays.copy()
    near[:, :3] = rng.uniform(-6, 6, (n, 3)).astype(F)
    near[:, 3:6] = (target - near[:, :3]).astype(F)
    ref_far = scene.trace_rays(rays, threads=NTHREADS)
    ref_near = scene.trace_rays(near, threads=NTHREADS)
    ref_img = scene.render(cam, 96, 64, tile=8)
    assert (ref_far["id"] != O.MISS_ID).sum() > 100
    with Engine(flags=_ffi.FLAG_STRICT | _ffi.FLAG_LEAF_ACCEL) as eng:
        SB.upload_scene(eng, scene)
        for _ in range(2):
            assert_strict(eng.trace_rays(rays), ref_far)
            assert eng.trace_primary(SB.to_ffi_camera(cam), 96, 64, tile=8).tobytes() == ref_img.tobytes()
            assert_strict(eng.trace_rays(near), ref_near)
        # NaN / inf rays must not poison the bake for the finite ones in the same batch
        bad = near.copy()
        bad[0, 0] = np.inf; bad[1, 4] = np.nan
        got = eng.trace_rays(bad)
        ref_bad = scene.trace_rays(bad, threads=NTHREADS)
        assert np.array_equal(got["id"][2:], ref_bad["id"][2:])
        assert got["t"][2:].tobytes() == ref_bad["t"][2:].tobytes()


def test_empty_and_ragged_inputs():
    scene, cam = SB.oracle_scene(examples.cube())
    with Engine() as eng:
        SB.upload_scene(eng, scene)
        assert len(eng.trace_rays(np.zeros((0, 7), F))) == 0                # empty ray list
        # ragged image: not a multiple of the tile, odd tile, sub-rectangle region
        w, h = 101, 67
        ref = scene.render(cam, w, h, tile=8)
        for tile in (8, 5, 16):
            got = eng.trace_primary(SB.to_ffi_camera(cam), w, h, tile=tile)
            assert got.tobytes() == ref.tobytes()
        region = (13, 9, 77, 50)
        got = eng.trace_primary(SB.to_ffi_camera(cam), w, h, region=region)
        exp = np.zeros(w * h, _ffi.HIT); exp["t"] = O.FLT_MAX; exp["id"] = O.MISS_ID
        e2 = exp.reshape(h, w); r2 = ref.reshape(h, w)
        e2[9:50, 13:77] = r2[9:50, 13:77]
        assert got.tobytes() == exp.tobytes()
        # empty region is a no-op
        got = eng.trace_primary(SB.to_ffi_camera(cam), w, h, region=(5, 5, 5, 9))
        assert (got["id"] == O.MISS_ID).all()


def test_tile_bands_reassemble_full_frame():
    # multi-GPU sharding contract: disjoint regions written into one buffer == full frame
    scene, cam = SB.oracle_scene(examples.trippy_teapots(3))
    w, h = 640, 360
    with Engine(flags=FLAG_LEAF_ACCEL) as eng:
        SB.upload_scene(eng, scene)
        full = eng.trace_primary(SB.to_ffi_camera(cam), w, h)
        out = None
        for r in range(4):
            y0, y1 = r * 88, min(h, (r + 1) * 88)
            out = eng.trace_primary(SB.to_ffi_camera(cam), w, h, region=(0, y0, w, y1), out=out)
        out = eng.trace_primary(SB.to_ffi_camera(cam), w, h, region=(0, 352, w, h), out=out)
    assert out.tobytes() == full.tobytes()


# ------------------------------------------------------------------------------------------ accel == brute at full size
@pytest.mark.parametrize("name,frame,size", [
    ("two_armadillos", "canonical", (1920, 1080)),
    ("two_armadillos", "initial", (1920, 1080)),
    ("sixteen_armadillos", 0, (3840, 2160)),
    ("sixteen_armadillos", 45, (1920, 1080)),
    ("trippy_teapots", 10, (3840, 2160)),
    ("big_ben_clock", None, (1920, 1080)),
])
def test_leaf_accel_equals_brute_force_full_size(name, frame, size):
    spec = examples.CONFIGS[name]() if frame is None else examples.CONFIGS[name](frame)
    scene, cam = SB.oracle_scene(spec)
    w, h = size
    res = []
    for flags in STRICT_MODES:
        with Engine(flags=flags) as eng:
            SB.upload_scene(eng, scene)
            res.append(eng.trace_primary(SB.to_ffi_camera(cam), w, h))
    assert (res[0]["id"] != O.MISS_ID).sum() > 1000
    assert res[0].tobytes() == res[1].tobytes()


# ------------------------------------------------------------------------------------------ fast mode
@pytest.mark.parametrize("flags", [FLAG_FAST, FLAG_FAST | FLAG_LEAF_ACCEL], ids=["fast-brute", "fast-accel"])
@pytest.mark.parametrize("name,frame,size", [("two_armadillos", "canonical", (384, 216)),
                                             ("sixteen_armadillos", 5, (320, 180)),
                                             ("trippy_teapots", 8, (640, 360))])
def test_fast_mode_tolerance(name, frame, size, flags):
    got, ref = render_both(examples.CONFIGS[name](frame), size[0], size[1], flags)
    assert_fast(got, ref)


@pytest.mark.parametrize("flags", [FLAG_FAST, FLAG_FAST | FLAG_LEAF_ACCEL], ids=["fast-brute", "fast-accel"])
def test_fast_mode_on_exact_ties_cube_640(flags):
    # The axis-aligned cube view puts whole pixel columns exactly on the diagonals shared by two triangles (u + v == 1 in exact
    # arithmetic).  The fast build filters with FMA contraction but decides every candidate that survives the (widened) filter
    # in the reference's arithmetic, so even here the records are the strict ones.
    got, ref = render_both(examples.cube(), 640, 640, flags)
    r = SB.compare_hits(got, ref)
    assert r["id_mismatch"] <= 4 and r["max_rel_t"] <= 1e-6, r


# ------------------------------------------------------------------------------------------ error behaviour
def test_errors_are_codes_not_crashes():
    scene, cam = SB.oracle_scene(examples.cube())
    blas = scene.blases[0]
    with Engine() as eng:
        with pytest.raises(BvhtError) as e:                       # trace before tlas_set
            eng.trace_primary(SB.to_ffi_camera(cam), 8, 8)
        assert e.value.status == _ffi.ERR_NOT_READY
        bad = blas.nodes[:blas.nodes_used].copy().view(_ffi.BVH_NODE)
        bad["left_first"][0] = 1000                               # child out of range
        with pytest.raises(BvhtError) as e:
            eng.blas_create(blas.tris, bad, blas.nodes_used)
        assert e.value.status == _ffi.ERR_MALFORMED_BVH
        bad = blas.nodes[:blas.nodes_used].copy().view(_ffi.BVH_NODE)
        leaf = int(np.argmax(bad["prim_count"] > 0))
        bad["prim_count"][leaf] = 10 ** 6                         # leaf beyond the triangle buffer
        with pytest.raises(BvhtError) as e:
            eng.blas_create(blas.tris, bad, blas.nodes_used)
        assert e.value.status == _ffi.ERR_MALFORMED_BVH
        with pytest.raises(BvhtError) as e:                       # empty model
            eng.blas_create(np.zeros((0, 9), F), blas.nodes.view(_ffi.BVH_NODE), 2)
        assert e.value.status == _ffi.ERR_INVALID_ARG
        ids = SB.upload_scene(eng, scene)
        inst = np.zeros(1, _ffi.INSTANCE); inst["blas_id"] = 77  # unknown blas
        with pytest.raises(BvhtError) as e:
            eng.tlas_set(scene.tlas.view(_ffi.TLAS_NODE), scene.tlas_used, inst)
        assert e.value.status == _ffi.ERR_BAD_HANDLE
        with pytest.raises(BvhtError):
            eng.blas_update_vertices(ids[0], np.zeros((5, 9), F))  # wrong triangle count
        with pytest.raises(BvhtError) as e:                       # scene without objects
            eng.tlas_set(scene.tlas.view(_ffi.TLAS_NODE), scene.tlas_used, np.zeros(0, _ffi.INSTANCE))
        assert e.value.status == _ffi.ERR_INVALID_ARG
        cyc = scene.tlas[:scene.tlas_used].copy().view(_ffi.TLAS_NODE)
        cyc["left_right"][0] = (0 << 16) | 0 | 1                  # node 0 -> children (0, 1): a cycle through the root
        with pytest.raises(BvhtError) as e:
            eng.tlas_set(cyc, scene.tlas_used, np.zeros(1, _ffi.INSTANCE))
        assert e.value.status == _ffi.ERR_MALFORMED_BVH
        # the context is still usable after every error
        got = eng.trace_primary(SB.to_ffi_camera(cam), 64, 64)
        assert got.tobytes() == scene.render(cam, 64, 64).tobytes()


# ------------------------------------------------------------------------------------------ f1: fused shading
@pytest.mark.parametrize("kind", ["depth", "intersection", "uv"])
def test_render_frame_shaders_match_oracle(kind):
    # renderer.rs:116-245 accumulator + pixel shader pairs, fused into the trace kernel's epilogue
    scene, cam = SB.oracle_scene(examples.sixteen_armadillos(9))
    w, h = 320, 184
    ref_hits = scene.render(cam, w, h, threads=NTHREADS)
    with Engine(flags=FLAG_LEAF_ACCEL) as eng:
        SB.upload_scene(eng, scene)
        if kind == "depth":
            shade, ref = eng.shade_depth(80.0, 3.0), O.shade(1, ref_hits, 80.0, 3.0)
        elif kind == "intersection":
            shade, ref = eng.shade_intersection((255, 200, 10, 255), (1, 2, 3, 255)), \
                O.shade(2, ref_hits, hit_rgba=0xFF0AC8FF, miss_rgba=0xFF030201)
        else:
            shade, ref = eng.shade_uv(), O.shade(3, ref_hits)
        frame, hits = eng.render_frame(SB.to_ffi_camera(cam), w, h, shade, want_hits=True)
        assert hits.tobytes() == ref_hits.tobytes()
        assert frame.tobytes() == ref.tobytes()
        assert len(np.unique(frame)) >= 2
        # frame only (no hit records travel), pinned destination, sub-region
        pinned = eng.pinned_array(w * h, "<u4")
        pinned[:] = 0
        eng.render_frame(SB.to_ffi_camera(cam), w, h, shade, frame_out=pinned, region=(0, 8, w, 96))
        exp = np.zeros(w * h, "<u4")
        exp.reshape(h, w)[8:96] = ref.reshape(h, w)[8:96]
        assert pinned.tobytes() == exp.tobytes()
        eng.free_pinned(pinned)


def test_render_frame_bands_full_size_equal_single_launch():
    # 4K frame: the banded, copy-overlapped host path must equal the single-launch device path
    scene, cam = SB.oracle_scene(examples.sixteen_armadillos(2))
    w, h = 3840, 2160
    with Engine(flags=FLAG_LEAF_ACCEL) as eng:
        SB.upload_scene(eng, scene)
        shade = eng.shade_depth()
        frame, hits = eng.render_frame(SB.to_ffi_camera(cam), w, h, shade, want_hits=True)
        dh = eng.device_alloc(w * h * 16)
        df = eng.device_alloc(w * h * 4)
        eng.render_frame_device(SB.to_ffi_camera(cam), w, h, shade, 8, None, df, dh)
        eng.sync()
        h2 = eng.memcpy_d2h(np.zeros(w * h, _ffi.HIT), dh)
        f2 = eng.memcpy_d2h(np.zeros(w * h, "<u4"), df)
        eng.device_free(dh); eng.device_free(df)
    assert hits.tobytes() == h2.tobytes() and frame.tobytes() == f2.tobytes()
    assert np.array_equal(frame, O.shade(1, hits))


@pytest.mark.parametrize("scene_name,size", [("sixteen_armadillos", (1283, 717)), ("cube", (333, 251))])
def test_pipelined_host_frame_every_band_plan(scene_name, size):
    # bvht_render_frame traces a frame with ONE launch that pulls its pixel blocks band by band and raises a flag per finished band;
    # the device->host copies wait for the flags on other streams.  Every band count, pull order and number of copy streams must
    # deliver the same bytes as the plain device frame -- ragged sizes, sub-regions and tile-row shards included.  (A band whose
    # copy ran before its pixels were written would show up as stale data: the host buffers are poisoned before every call.)
    spec = examples.sixteen_armadillos(3) if scene_name == "sixteen_armadillos" else examples.cube()
    scene, cam = SB.oracle_scene(spec)
    w, h = size
    fcam = SB.to_ffi_camera(cam)
    ref = scene.render(cam, w, h, threads=NTHREADS)
    ref_frame = O.shade(1, ref)
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        SB.upload_scene(eng, scene)
        shade = eng.shade_depth()
        frame = np.zeros(w * h, "<u4")
        hits = np.zeros(w * h, _ffi.HIT)
        for bands, order, streams in [(1, 0, 1), (2, 1, 1), (3, 2, 2), (7, 3, 3), (16, 1, 2), (32, 1, 3), (32, 0, 1), (-1, -1, -1)]:
            eng.set_option(_ffi.OPT_BANDS, bands)
            eng.set_option(_ffi.OPT_BAND_ORDER, order)
            eng.set_option(_ffi.OPT_COPY_STREAMS, streams)
            for k0 in (0, 1):
                eng.set_option(_ffi.OPT_K0, k0)
                frame[:] = 0xDEADBEEF
                hits.view(np.uint8)[:] = 0xA5
                eng.render_frame(fcam, w, h, shade, frame_out=frame, hits_out=hits)
                assert hits.tobytes() == ref.tobytes(), (bands, order, streams, k0)
                assert frame.tobytes() == ref_frame.tobytes(), (bands, order, streams, k0)
        eng.set_option(_ffi.OPT_K0, -1)
        # a sub-region: only its pixels are written
        eng.set_option(_ffi.OPT_BANDS, 5)
        region = (40, 24, w - 33, h - 19)
        frame[:] = 0xDEADBEEF
        eng.render_frame(fcam, w, h, shade, region=region, frame_out=frame)
        inside = np.zeros((h, w), bool)
        inside[region[1]:region[3], region[0]:region[2]] = True
        inside = inside.reshape(-1)
        assert np.array_equal(frame[inside], ref_frame[inside]) and (frame[~inside] == 0xDEADBEEF).all()
        # three tile-row shards into ONE host frame, each with its own band plan
        frame[:] = 0xDEADBEEF
        for shard, bands in ((0, 4), (1, 1), (2, 9)):
            eng.set_shard(shard, 3)
            eng.set_option(_ffi.OPT_BANDS, bands)
            eng.render_frame(fcam, w, h, shade, frame_out=frame)
        eng.set_shard(0, 1)
        assert frame.tobytes() == ref_frame.tobytes()


# ------------------------------------------------------------------------------------------ through the host mirror
def test_host_mirror_renderer_animated_frames():
    # Renderer::new(Box::new(CudaPathTracer::new())) + AppState::update loop of sixteen_armadillos.rs:132-163
    from bvhtracer_b200 import host
    anim = examples.GridAnimation()
    scene, models = host.build_scene(examples.sixteen_armadillos(0))
    renderer = host.Renderer(flags=FLAG_STRICT | FLAG_LEAF_ACCEL)
    w, h = 320, 176
    state = host.RendererState(host.depth_pipeline(80.0, 3.0), w, h, keep_hits=True)
    for frame in range(4):
        if frame > 0:
            anim.update()
            for i, o in enumerate(anim.objects()):
                scene.set_transform(i, host.object_transform(o))
            scene.rebuild()
        assert renderer.render(state, scene) == w * h                 # rays traced, like PathTracer::evaluate
        ref_scene, ref_cam = SB.oracle_scene(examples.sixteen_armadillos(frame))
        ref = ref_scene.render(ref_cam, w, h, threads=NTHREADS)
        assert state.hits().tobytes() == ref.tobytes()
        assert state.frame_buffer().tobytes() == O.shade(1, ref).tobytes()


def test_two_frames_in_flight_deliver_render_frames_bytes():
    # Renderer::render_begin / render_end (bvht_render_frame_begin / _end): frame n's device->host copies run under frame n+1's
    # kernels, out of the other of two device staging frames.  Every frame of the animation must arrive byte for byte as the
    # oracle renders it, whatever was queued behind it (the next frame's scene update and kernels), and the begin/end pairing
    # rules hold: a third begin and an end with nothing in flight are refused, render() drains what is in flight.
    from bvhtracer_b200 import host
    anim = examples.GridAnimation()
    scene, models = host.build_scene(examples.sixteen_armadillos(0))
    renderer = host.Renderer(flags=FLAG_STRICT | FLAG_LEAF_ACCEL)
    w, h = 1283, 717                                                  # ragged size, several bands
    states = [host.RendererState(host.depth_pipeline(80.0, 3.0), w, h, keep_hits=(i == 0)) for i in range(2)]
    n_frames = 5
    refs = []
    for frame in range(n_frames):
        ref_scene, ref_cam = SB.oracle_scene(examples.sixteen_armadillos(frame))
        refs.append(ref_scene.render(ref_cam, w, h, threads=NTHREADS))
    with pytest.raises(host.HostError):
        renderer.render_end()                                         # nothing in flight
    for frame in range(n_frames + 1):
        if frame < n_frames:
            if frame > 0:
                anim.update()
                for i, o in enumerate(anim.objects()):
                    scene.set_transform(i, host.object_transform(o))
                scene.rebuild()
            assert renderer.render_begin(states[frame & 1], scene) == w * h
        if frame >= 1:
            renderer.render_end()                                     # completes frame - 1 while frame is being traced
            done = frame - 1
            assert states[done & 1].frame_buffer().tobytes() == O.shade(1, refs[done]).tobytes(), done
            if (done & 1) == 0:
                assert states[0].hits().tobytes() == refs[done].tobytes(), done
    with pytest.raises(host.HostError):
        renderer.render_end()
    # two begun, a third refused; render() completes both before it renders
    third = host.RendererState(host.depth_pipeline(80.0, 3.0), w, h, keep_hits=False)
    renderer.render_begin(states[0], scene)
    renderer.render_begin(states[1], scene)
    with pytest.raises(host.HostError):
        renderer.render_begin(third, scene)
    renderer.render(third, scene)
    for st in (states[0], states[1], third):
        assert st.frame_buffer().tobytes() == O.shade(1, refs[-1]).tobytes()
    with pytest.raises(host.HostError):
        renderer.render_end()


def test_frames_in_flight_with_vertex_updates_between_them():
    # big_ben_clock: the vertices of frame n+1 are uploaded and refitted while frame n's copies are still running
    _, cam = SB.oracle_scene(examples.big_ben_clock())
    blas = O.Blas(O.load_asset("bigben.tri"))              # private copy: vertices get animated
    scene = O.Scene([blas], [(0, O.mat4_identity())], with_transform=False)
    anim = examples.BigBenAnimation(blas.tris)
    fcam = SB.to_ffi_camera(cam)
    w, h = 640, 360
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        ids = SB.upload_scene(eng, scene)
        shade = eng.shade_depth()
        bufs = [(eng.pinned_array(w * h, "<u4"), eng.pinned_array(w * h, _ffi.HIT)) for _ in range(2)]
        want = []
        for frame in range(4):
            verts = anim.animate()
            blas.tris[:] = verts
            blas.refit()
            scene.refresh_blas()
            want.append(scene.render(cam, w, h, threads=NTHREADS))
            eng.blas_update_vertices(ids[0], verts)
            eng.blas_refit(ids[0])
            f, hh = bufs[frame & 1]
            if frame >= 2:
                eng.render_frame_end()                                # frame - 2 owned these buffers
                assert hh.tobytes() == want[frame - 2].tobytes(), frame - 2
            f[:] = 0xDEADBEEF
            eng.render_frame_begin(fcam, w, h, shade, f, hh)
        eng.sync()                                                    # completes the two still in flight
        for frame in (2, 3):
            f, hh = bufs[frame & 1]
            assert hh.tobytes() == want[frame].tobytes() and f.tobytes() == O.shade(1, want[frame]).tobytes(), frame
        with pytest.raises(BvhtError):
            eng.render_frame_end()
        for f, hh in bufs:
            eng.free_pinned(f); eng.free_pinned(hh)


def test_host_mirror_big_ben_refit_and_scene_intersect():
    # big_ben_clock.rs:67-103: animate() + ModelInstance::refit(), IntersectionAccumulator + IntersectionShader
    from bvhtracer_b200 import host
    scene, models = host.build_scene(examples.big_ben_clock())
    ref_blas = O.Blas(O.load_asset("bigben.tri"))
    ref_scene = O.Scene([ref_blas], [(0, O.mat4_identity())], with_transform=False)
    _, ref_cam = SB.oracle_scene(examples.big_ben_clock())
    anim = examples.BigBenAnimation(models[0].primitives())
    renderer = host.Renderer(flags=FLAG_STRICT | FLAG_LEAF_ACCEL)
    w, h = 256, 144
    state = host.RendererState(host.intersection_pipeline((255, 255, 255, 255), (0, 0, 0, 255)), w, h, keep_hits=True)
    for frame in range(3):
        verts = anim.animate()
        models[0].set_primitives(verts)
        models[0].refit()
        ref_blas.tris[:] = verts
        ref_blas.refit()
        ref_scene.refresh_blas()
        renderer.render(state, scene)
        nodes, used = models[0].nodes()                               # refitted boxes read back into the host Bvh
        assert np.array_equal(nodes["aabb_min"][:used], ref_blas.nodes["min"][:used])
        assert np.array_equal(nodes["aabb_max"][:used], ref_blas.nodes["max"][:used])
        ref = ref_scene.render(ref_cam, w, h, threads=NTHREADS)
        assert state.hits().tobytes() == ref.tobytes()
        assert state.frame_buffer().tobytes() == O.shade(2, ref, hit_rgba=0xFFFFFFFF, miss_rgba=0xFF000000).tobytes()
    # Scene::intersect(&Ray) through the integrator
    rays = np.array([[0, 2.75, -2.5, 0, 0, 1, O.FLT_MAX], [0, 2.75, -2.5, 0, 1, 0, O.FLT_MAX]], F)
    got = renderer.intersect(scene, rays)
    assert got.tobytes() == ref_scene.trace_rays(rays).tobytes()


def test_sharded_launches_assemble_in_one_buffer():
    # the multi-GPU contract on one device: n contexts, shard i of n each, all writing into ONE device buffer
    scene, cam = SB.oracle_scene(examples.sixteen_armadillos(4))
    w, h, n = 512, 296, 3
    with Engine(flags=FLAG_LEAF_ACCEL) as e0:
        SB.upload_scene(e0, scene)
        full = e0.trace_primary(SB.to_ffi_camera(cam), w, h)
        buf = e0.device_alloc(w * h * 16)
        e0.memcpy_h2d(buf, np.zeros(w * h, _ffi.HIT))
        for i in range(n):
            with Engine(flags=FLAG_LEAF_ACCEL) as ei:
                SB.upload_scene(ei, scene)
                ei.set_shard(i, n)
                ei.trace_primary_device(SB.to_ffi_camera(cam), w, h, 8, None, buf)
                ei.sync()
                if i == 0:                                            # after one shard only its rows are filled
                    part = e0.memcpy_d2h(np.zeros(w * h, _ffi.HIT), buf).reshape(h, w)
                    rows = _ffi.shard_tile_rows((0, 0, w, h), 8, 0, n)
                    own = np.zeros(h, bool)
                    for r in rows:
                        own[r * 8:(r + 1) * 8] = True
                    assert part[own].tobytes() == full.reshape(h, w)[own].tobytes()
                    assert not part[~own].view(np.uint8).any()
        got = e0.memcpy_d2h(np.zeros(w * h, _ffi.HIT), buf)
        e0.device_free(buf)
    assert got.tobytes() == full.tobytes()


@pytest.mark.parametrize("flags", STRICT_MODES, ids=MODE_IDS)
@pytest.mark.parametrize("name,arg", [("two_armadillos", "canonical"), ("sixteen_armadillos", 7), ("big_ben_clock", None)])
def test_axis_parallel_rays_zero_direction_components(name, arg, flags):
    # d has exact zeros -> +-inf reciprocals (Ray::new, ray.rs:23-31).  The reference slab test handles them
    # (aabb KATs); OUR conservative boxes (sub-BVH, tight TLAS boxes) must stay conservative for them too.
    spec = examples.CONFIGS[name]() if arg is None else examples.CONFIGS[name](arg)
    scene, _ = SB.oracle_scene(spec)
    g = np.linspace(-3.0, 3.0, 97, dtype=F)
    rays = []
    for axis, sign in ((2, 1.0), (2, -1.0), (0, 1.0), (1, -1.0)):
        a, b = np.meshgrid(g, g + F(1.0))
        o = np.zeros((a.size, 3), F)
        other = [k for k in range(3) if k != axis]
        o[:, other[0]] = a.ravel()
        o[:, other[1]] = b.ravel()
        o[:, axis] = -8.0 * sign
        d = np.zeros((a.size, 3), F)
        d[:, axis] = sign
        rays.append(np.concatenate([o, d, np.full((a.size, 1), O.FLT_MAX, F)], axis=1))
    rays = np.concatenate(rays).astype(F)
    ref = scene.trace_rays(rays, threads=NTHREADS)
    with Engine(flags=flags) as eng:
        SB.upload_scene(eng, scene)
        got = eng.trace_rays(rays)
    assert (ref["id"] != O.MISS_ID).sum() > 200
    assert_strict(got, ref)


def test_accel_rebake_across_camera_distances():
    # The conservative inflation of the sub-BVH is baked for the current camera (ensure_bake): move the camera far away,
    # back, and very close; every frame must stay bit-identical to the oracle, and arbitrary rays traced afterwards
    # (outside the baked limits -> brute-force leaves) must too.
    spec = examples.two_armadillos("canonical")
    scene, _ = SB.oracle_scene(spec)
    w, h = 192, 108
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        SB.upload_scene(eng, scene)
        for z, near in ((-2.5, 2.0), (-60.0, 60.0), (-2.5, 2.0), (-1.3, 0.5), (-400.0, 900.0), (-2.5, 2.0)):
            cam = O.camera_symmetric_fov(90.0, 1.0, near, [0, 1, z], [0, 0, 1], [1, 0, 0], [0, 1, 0])
            ref = scene.render(cam, w, h, threads=NTHREADS)
            got = eng.trace_primary(SB.to_ffi_camera(cam), w, h)
            assert_strict(got, ref)
        rng = np.random.default_rng(7)
        n = 3000
        o = (rng.normal(size=(n, 3)) * 300).astype(F)             # far outside any baked |o| limit
        d = (-o / np.linalg.norm(o, axis=1, keepdims=True)).astype(F)
        d += (rng.normal(size=(n, 3)) * 0.004).astype(F)
        rays = np.concatenate([o, d, np.full((n, 1), O.FLT_MAX, F)], axis=1).astype(F)
        ref = scene.trace_rays(rays, threads=NTHREADS)
        got = eng.trace_rays(rays)
        assert (ref["id"] != O.MISS_ID).sum() > 50
        assert_strict(got, ref)


@pytest.mark.parametrize("name,arg,size", [("trippy_teapots", 10, (3840, 2160)), ("sixteen_armadillos", 30, (1920, 1080)),
                                           ("big_ben_clock", None, (3840, 2160))])
def test_fast_mode_full_size_against_strict_gpu(name, arg, size):
    # fast (FMA) build vs the strict build on the GPU itself at full size: >= 99.99 % identical ids and, because the
    # winning triangle's (t, u, v) are re-evaluated without contraction, t within 1e-5 relative on identical ids
    spec = examples.CONFIGS[name]() if arg is None else examples.CONFIGS[name](arg)
    scene, cam = SB.oracle_scene(spec)
    w, h = size
    res = {}
    for label, flags in (("strict", FLAG_STRICT | FLAG_LEAF_ACCEL), ("fast", FLAG_FAST | FLAG_LEAF_ACCEL), ("fast-brute", FLAG_FAST)):
        if label == "fast-brute" and name != "trippy_teapots":
            continue
        with Engine(flags=flags) as eng:
            SB.upload_scene(eng, scene)
            res[label] = eng.trace_primary(SB.to_ffi_camera(cam), w, h)
    for label in res:
        if label != "strict":
            assert_fast(res[label], res["strict"])


# ------------------------------------------------------------------------------------------ full size vs the oracle, sampled
@pytest.mark.parametrize("name,arg,size", [("two_armadillos", "canonical", (1920, 1080)), ("sixteen_armadillos", 0, (3840, 2160)),
                                           ("sixteen_armadillos", 17, (3840, 2160)), ("trippy_teapots", 10, (3840, 2160)),
                                           ("big_ben_clock", None, (7680, 4320))])
def test_full_size_frames_match_the_oracle_on_sampled_tiles(name, arg, size):
    # BASELINE.json's own resolutions: the whole frame is traced on the GPU (strict, leaf accelerator), the oracle re-traces
    # 320 of its 8x8 tiles -- half drawn among tiles that contain hits, half anywhere -- and every record must be identical
    spec = examples.CONFIGS[name]() if arg is None else examples.CONFIGS[name](arg)
    scene, cam = SB.oracle_scene(spec)
    w, h = size
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        SB.upload_scene(eng, scene)
        got = eng.trace_primary(SB.to_ffi_camera(cam), w, h).reshape(h, w)
    rng = np.random.default_rng(2024)
    tw, th = w // 8, h // 8
    hit_tiles = np.argwhere((got["id"] != O.MISS_ID).reshape(th, 8, tw, 8).any(axis=(1, 3)))
    assert len(hit_tiles) > 200
    picks = [tuple(hit_tiles[i]) for i in rng.choice(len(hit_tiles), 160, replace=False)]
    picks += [(int(rng.integers(th)), int(rng.integers(tw))) for _ in range(160)]
    ref = np.zeros(w * h, O.HIT); ref["t"] = O.FLT_MAX; ref["id"] = O.MISS_ID
    n_hits = 0
    for ty, tx in picks:
        region = (tx * 8, ty * 8, tx * 8 + 8, ty * 8 + 8)
        scene.render(cam, w, h, region=region, out=ref)
        a = got[ty * 8:ty * 8 + 8, tx * 8:tx * 8 + 8]
        b = ref.reshape(h, w)[ty * 8:ty * 8 + 8, tx * 8:tx * 8 + 8]
        assert a.tobytes() == b.tobytes(), (name, arg, tx, ty)
        n_hits += int((b["id"] != O.MISS_ID).sum())
    assert n_hits > 2000


# ------------------------------------------------------------------------------------------ NormalMappingAccumulator
@pytest.mark.parametrize("name,arg,asset", [("trippy_teapots", 6, "teapot.obj"), ("cube", None, "cube.obj")])
def test_normal_mapping_shader_matches_oracle(name, arg, asset):
    # renderer.rs:256-286: normals[prim] of object 0's model (UN-reordered array, reordered index; instance always 0),
    # object 0's transform, normalize, (n + 1) / 2, then RadianceToRgbShader -- what cube.rs and trippy_teapots.rs display
    spec = examples.CONFIGS[name]() if arg is None else examples.CONFIGS[name](arg)
    scene, cam = SB.oracle_scene(spec)
    normals = O.load_asset_normals(asset)
    m0 = SB.object_matrix(spec.objects[0])
    w, h = 480, 272
    ref_hits = scene.render(cam, w, h, threads=NTHREADS)
    ref = O.shade_normal(normals, m0, ref_hits)
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        ids = SB.upload_scene(eng, scene)
        with pytest.raises(BvhtError):                               # normals not uploaded yet
            eng.render_frame(SB.to_ffi_camera(cam), w, h, eng.shade_normal(m0))
        eng.blas_set_normals(ids[0], normals)
        frame, hits = eng.render_frame(SB.to_ffi_camera(cam), w, h, eng.shade_normal(m0), want_hits=True)
    assert hits.tobytes() == ref_hits.tobytes()
    assert frame.tobytes() == ref.tobytes()
    assert len(np.unique(frame)) >= 2


def test_texture_material_shader_matches_oracle():
    # quad.rs main: TextureMaterialAccumulator + RadianceToRgbShader.  Nearest-texel lookup with Rust's saturating
    # `as usize` and `% size` (material.rs:44-52), instance index always 0 (renderer.rs:309-327).
    w, h = 320, 320
    tex = examples.brick_texture(96, 64)
    for frame in (0, 17):
        spec = examples.quad_example(frame)
        scene, cam = SB.oracle_scene(spec)
        ref_hits = scene.render(cam, w, h, threads=NTHREADS)
        for tc in (examples.QUAD_TEX_COORDS, (examples.QUAD_TEX_COORDS * F(3.7) - F(1.2)).astype(F)):   # also wrap + negative
            ref = O.shade_texture(tc, tex, ref_hits)
            with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
                ids = SB.upload_scene(eng, scene)
                with pytest.raises(BvhtError):                           # nothing uploaded yet
                    eng.render_frame(SB.to_ffi_camera(cam), w, h, eng.shade_texture())
                eng.blas_set_tex_coords(ids[0], tc)
                with pytest.raises(BvhtError):                           # coordinates but no texture
                    eng.render_frame(SB.to_ffi_camera(cam), w, h, eng.shade_texture())
                with pytest.raises(BvhtError):                           # `% 0` in the reference
                    eng.blas_set_texture(ids[0], np.zeros((0, 4, 3), np.uint8))
                eng.blas_set_texture(ids[0], tex)
                frame_px, hits = eng.render_frame(SB.to_ffi_camera(cam), w, h, eng.shade_texture(), want_hits=True)
            assert hits.tobytes() == ref_hits.tobytes()
            assert frame_px.tobytes() == ref.tobytes()
            assert len(np.unique(frame_px)) > 50


@pytest.mark.parametrize("flags", [FLAG_STRICT | FLAG_LEAF_ACCEL, FLAG_FAST | FLAG_LEAF_ACCEL], ids=["strict", "fast"])
def test_texture_material_many_instances_use_object0(flags):
    # sixteen teapots, random per-vertex coordinates incl. huge ones (usize saturation): every instance's hits index object 0's
    # coordinate array and texture.  Fast mode: pixels are identical wherever the hit record is.
    spec = examples.trippy_teapots(4)
    scene, cam = SB.oracle_scene(spec)
    rng = np.random.default_rng(11)
    n_tris = scene.blases[0].tris.shape[0]
    tc = rng.uniform(-2, 5, (n_tris, 6)).astype(F)
    tc[::37] *= F(1e24)
    tex = rng.integers(0, 256, (33, 129, 3)).astype(np.uint8)
    w, h = 400, 224
    ref_hits = scene.render(cam, w, h, threads=NTHREADS)
    with Engine(flags=flags) as eng:
        ids = SB.upload_scene(eng, scene)
        eng.blas_set_tex_coords(ids[0], tc)
        eng.blas_set_texture(ids[0], tex)
        frame_px, hits = eng.render_frame(SB.to_ffi_camera(cam), w, h, eng.shade_texture(), want_hits=True)
    if flags & FLAG_FAST:
        assert frame_px.tobytes() == O.shade_texture(tc, tex, hits).tobytes()        # shading is exact given the record
        assert (hits["id"] == ref_hits["id"]).mean() >= 0.9999
    else:
        assert hits.tobytes() == ref_hits.tobytes()
        assert frame_px.tobytes() == O.shade_texture(tc, tex, ref_hits).tobytes()
    assert (ref_hits["id"] != O.MISS_ID).mean() > 0.05


def test_host_mirror_textured_quad_example():
    # quad.rs: MeshBuilder::with_primitive(tri, tex_coords, normals), ModelBuilder::with_texture, TextureMaterialAccumulator +
    # RadianceToRgbShader through Renderer::render, several frames of the spinning quad
    from bvhtracer_b200 import host
    tex = examples.brick_texture(128, 128)
    mesh = host.Mesh.from_triangles(examples.QUAD_TRIS, examples.QUAD_NORMALS).set_tex_coords(examples.QUAD_TEX_COORDS)
    model = host.ModelBuilder().with_mesh(mesh).with_texture(tex).build()
    renderer = host.Renderer(flags=FLAG_STRICT | FLAG_LEAF_ACCEL)
    w, h = 256, 256
    state = host.RendererState(host.texture_pipeline(), w, h, keep_hits=True)
    scene = None
    for frame in (0, 3, 41):
        spec = examples.quad_example(frame)
        if scene is None:
            scene, _ = host.build_scene(spec, models=[model])
        else:
            scene.set_transform(0, host.object_transform(spec.objects[0]))
            scene.rebuild()
        renderer.render(state, scene)
        ref_scene, ref_cam = SB.oracle_scene(spec)
        ref_hits = ref_scene.render(ref_cam, w, h, threads=NTHREADS)
        assert state.hits().tobytes() == ref_hits.tobytes()
        assert state.frame_buffer().tobytes() == O.shade_texture(examples.QUAD_TEX_COORDS, tex, ref_hits).tobytes()
        assert (ref_hits["id"] != O.MISS_ID).mean() > 0.2


def test_host_mirror_normal_mapping_example():
    # trippy_teapots.rs main: NormalMappingAccumulator + RadianceToRgbShader through Renderer::render
    from bvhtracer_b200 import host
    spec = examples.trippy_teapots(3)
    scene, models = host.build_scene(spec)
    renderer = host.Renderer(flags=FLAG_STRICT | FLAG_LEAF_ACCEL)
    w, h = 400, 224
    state = host.RendererState(host.normal_pipeline(), w, h, keep_hits=True)
    renderer.render(state, scene)
    ref_scene, ref_cam = SB.oracle_scene(spec)
    ref_hits = ref_scene.render(ref_cam, w, h, threads=NTHREADS)
    assert state.hits().tobytes() == ref_hits.tobytes()
    ref = O.shade_normal(O.load_asset_normals("teapot.obj"), SB.object_matrix(spec.objects[0]), ref_hits)
    assert state.frame_buffer().tobytes() == ref.tobytes()


@pytest.mark.parametrize("flags", STRICT_MODES, ids=MODE_IDS)
def test_many_instances_deep_tlas(flags):
    # 150 randomly placed/rotated/scaled teapots + cubes: more than 32 instances (no tile masks), a deep agglomerative
    # TLAS, two BLASes, overlapping instances.  No reference test covers multi-instance scenes; parity is vs the oracle.
    rng = np.random.default_rng(99)
    blases = [SB.oracle_blas("teapot.obj"), SB.oracle_blas("cube.obj")]
    objs = []
    for i in range(150):
        s = float(rng.uniform(0.3, 1.2))
        m = O.transform_new_rot_xz([s, s, s], rng.uniform(-7, 7, 3), float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)))
        objs.append((i % 2, m))
    scene = O.Scene(blases, objs)
    cam = O.camera_symmetric_fov(90.0, 1.0, 1.0, [0, 0, -14], [0, 0, 1], [1, 0, 0], [0, 1, 0])
    w, h = 512, 288
    ref = scene.render(cam, w, h, threads=NTHREADS)
    assert (ref["id"] != O.MISS_ID).mean() > 0.1
    with Engine(flags=flags) as eng:
        SB.upload_scene(eng, scene)
        got = eng.trace_primary(SB.to_ffi_camera(cam), w, h)
    assert_strict(got, ref)


@pytest.mark.parametrize("flags", STRICT_MODES, ids=MODE_IDS)
def test_large_deformation_refit_multi_instance(flags):
    # A shared BLAS deformed far beyond its original bounds (the sub-BVH keeps its topology and is refitted on the
    # device; tight TLAS boxes and the bake follow), 16 instances, several steps; then back to the original shape.
    spec = examples.trippy_teapots(4)
    base = O.load_asset("teapot.obj")
    blas = O.Blas(base)
    objs = [(0, SB.object_matrix(o)) for o in spec.objects]
    scene = O.Scene([blas], objs)
    cam = SB.oracle_camera(spec.camera)
    orig = blas.tris.copy()
    w, h = 480, 272
    with Engine(flags=flags) as eng:
        ids = SB.upload_scene(eng, scene)
        for step, (sx, sy, shear, off) in enumerate([(1.6, 0.7, 0.5, 0.4), (0.5, 2.2, -0.8, -0.6), (1.0, 1.0, 0.0, 0.0)]):
            v = orig.reshape(-1, 3, 3).copy()
            v[:, :, 0] = orig.reshape(-1, 3, 3)[:, :, 0] * F(sx) + orig.reshape(-1, 3, 3)[:, :, 1] * F(shear)
            v[:, :, 1] = orig.reshape(-1, 3, 3)[:, :, 1] * F(sy) + F(off)
            blas.tris[:] = v.reshape(-1, 9)
            blas.refit()
            scene.rebuild()                                   # instance world bounds + TLAS from the refitted model bounds
            eng.blas_update_vertices(ids[0], blas.tris)
            eng.blas_refit(ids[0])
            nodes = eng.blas_read_nodes(ids[0], blas.nodes_used)
            assert np.array_equal(nodes["aabb_min"], blas.nodes["min"][:blas.nodes_used])
            assert np.array_equal(nodes["aabb_max"], blas.nodes["max"][:blas.nodes_used])
            SB.upload_scene(eng, scene, blas_ids=ids)
            ref = scene.render(cam, w, h, threads=NTHREADS)
            got = eng.trace_primary(SB.to_ffi_camera(cam), w, h)
            assert (ref["id"] != O.MISS_ID).mean() > 0.03
            assert_strict(got, ref)


def test_sharded_host_renders_fill_one_shared_frame():
    # multi-GPU e2e contract on one device: n contexts, shard i of n each, every one copying ONLY its tile rows into the
    # same page-locked host frame (ragged height: the last tile row is partial)
    scene, cam = SB.oracle_scene(examples.sixteen_armadillos(8))
    w, h, n = 448, 250, 3
    with Engine(flags=FLAG_LEAF_ACCEL) as e0:
        SB.upload_scene(e0, scene)
        shade = e0.shade_depth()
        full_frame, full_hits = e0.render_frame(SB.to_ffi_camera(cam), w, h, shade, want_hits=True)
        frame = e0.pinned_array(w * h, "<u4"); frame[:] = 0xDEADBEEF
        hits = np.zeros(w * h, _ffi.HIT)
        for i in range(n):
            with Engine(flags=FLAG_LEAF_ACCEL) as ei:
                SB.upload_scene(ei, scene)
                ei.set_shard(i, n)
                ei.render_frame(SB.to_ffi_camera(cam), w, h, shade, frame_out=frame, hits_out=hits)
                if i == 0:
                    rows = _ffi.shard_tile_rows((0, 0, w, h), 8, 0, n)
                    own = np.zeros(h, bool)
                    for r in rows:
                        own[r * 8:(r + 1) * 8] = True
                    f2 = frame.reshape(h, w)
                    assert np.array_equal(f2[own], full_frame.reshape(h, w)[own])
                    assert (f2[~own] == 0xDEADBEEF).all()              # rows of other shards are not touched
        assert frame.tobytes() == full_frame.tobytes()
        assert hits.tobytes() == full_hits.tobytes()
        e0.free_pinned(frame)


# ------------------------------------------------------------------------------------------ chain-skipping TLAS walk
def _random_instances(rng, n, spread):
    objs = []
    for i in range(n):
        s = float(rng.uniform(0.4, 1.3))
        m = O.transform_new_rot_xz([s, s, s], rng.uniform(-spread, spread, 3), float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3)))
        objs.append((i % 2, m))
    return objs


@pytest.mark.parametrize("n,seed", [(3, 1), (5, 2), (9, 3), (16, 4), (17, 5), (24, 6), (32, 7)])
def test_tlas_chain_skipping_random_scenes(n, seed):
    # <= 32 instances: tile-level candidate masks are on and, from 3 instances, the walk skips chains of nodes with a single
    # relevant child (trace_kernels.cuh).  Overlapping, rotated instances of two models; <= 16 instances take the shuffle
    # build of the skip table, 17..32 the shared-memory one.  Bit-identical to the oracle's plain ordered walk.
    rng = np.random.default_rng(seed)
    blases = [SB.oracle_blas("teapot.obj"), SB.oracle_blas("cube.obj")]
    scene = O.Scene(blases, _random_instances(rng, n, 2.0 + 0.07 * n))
    cam = O.camera_symmetric_fov(90.0, 1.0, 1.0, [0.3, 0.2, -6], [0, 0, 1], [1, 0, 0], [0, 1, 0])
    w, h = 400, 232
    ref = scene.render(cam, w, h, threads=NTHREADS)
    assert (ref["id"] != O.MISS_ID).mean() > 0.015
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        SB.upload_scene(eng, scene)
        got = eng.trace_primary(SB.to_ffi_camera(cam), w, h)
        assert_strict(got, ref)
        # the same scene through the region / tile variants of the launch (ragged blocks)
        got2 = eng.trace_primary(SB.to_ffi_camera(cam), w, h, tile=5)
        ref2 = scene.render(cam, w, h, tile=5, threads=NTHREADS)
        assert_strict(got2, ref2)


def test_tlas_with_boxes_that_are_not_nested_takes_the_plain_walk():
    # bvht_tlas_set accepts any node boxes.  Chain skipping is only valid for nested boxes, so a TLAS whose interior boxes
    # clip their children must fall back to the plain walk -- and still equal the oracle, which applies every box test.
    rng = np.random.default_rng(11)
    blases = [SB.oracle_blas("teapot.obj"), SB.oracle_blas("cube.obj")]
    scene = O.Scene(blases, _random_instances(rng, 12, 2.5))
    cam = O.camera_symmetric_fov(90.0, 1.0, 1.0, [0, 0, -6], [0, 0, 1], [1, 0, 0], [0, 1, 0])
    w, h = 400, 232
    nested = scene.render(cam, w, h, threads=NTHREADS)
    n = len(scene.objects)
    tl = scene.tlas
    for i in range(n + 1, scene.tlas_used):          # interior nodes: pull both corners towards the centre
        c = 0.5 * (tl["min"][i] + tl["max"][i])
        tl["min"][i] = c + 0.8 * (tl["min"][i] - c)
        tl["max"][i] = c + 0.8 * (tl["max"][i] - c)
    ref = scene.render(cam, w, h, threads=NTHREADS)
    assert (ref["id"] != O.MISS_ID).mean() > 0.05
    assert (ref["id"] != nested["id"]).sum() > 100    # the clipped boxes do change the reference's answer
    for flags in STRICT_MODES:
        with Engine(flags=flags) as eng:
            SB.upload_scene(eng, scene)
            got = eng.trace_primary(SB.to_ffi_camera(cam), w, h)
        assert_strict(got, ref)


def test_tlas_garbage_in_unreachable_slots_is_ignored():
    # Tlas::rebuild leaves a slot the walk never reaches (the last merged node is COPIED into node 0, tlas.rs:248), and a caller's
    # pool may hold anything there.  The per-node device tables (instance masks, chain-skip table) index through every slot, so
    # the upload neutralises unreachable ones: child indices far out of range there must neither fault nor change a record.
    rng = np.random.default_rng(12)
    blases = [SB.oracle_blas("teapot.obj"), SB.oracle_blas("cube.obj")]
    for n in (5, 24):                                   # shuffle build and shared-memory build of the skip table
        scene = O.Scene(blases, _random_instances(rng, n, 2.5))
        cam = O.camera_symmetric_fov(90.0, 1.0, 1.0, [0, 0, -6], [0, 0, 1], [1, 0, 0], [0, 1, 0])
        w, h = 400, 232
        ref = scene.render(cam, w, h, threads=NTHREADS)
        tl = scene.tlas.view(_ffi.TLAS_NODE).copy()
        reach, todo = set(), [0]
        while todo:
            i = todo.pop()
            reach.add(i)
            lr = int(tl["left_right"][i])
            if lr:
                todo += [lr >> 16, lr & 0xFFFF]
        dead = [i for i in range(scene.tlas_used) if i not in reach]
        assert dead, "expected at least one unreachable slot"
        for i in dead:
            tl["left_right"][i] = 0xFFFEFFFD             # children 65534 / 65533
            tl["blas"][i] = 0xFFFFFFFF
            tl["aabb_min"][i] = np.nan
        inst = np.zeros(len(scene.inst), _ffi.INSTANCE)
        for flags in STRICT_MODES:
            with Engine(flags=flags) as eng:
                ids = [eng.blas_create(b.tris, b.nodes.view(_ffi.BVH_NODE), b.nodes_used) for b in scene.blases]
                inst["transform_inv"] = scene.inst["inv"]
                inst["blas_id"] = [ids[int(b)] for b in scene.inst["blas_id"]]
                eng.tlas_set(tl, scene.tlas_used, inst)
                assert_strict(eng.trace_primary(SB.to_ffi_camera(cam), w, h), ref)
                frame, hits = eng.render_frame(SB.to_ffi_camera(cam), w, h, Engine.shade_depth(80.0, 3.0), want_hits=True)
                assert hits.tobytes() == ref.tobytes()


# ------------------------------------------------------------------------------------------ K0: classify + fill
@pytest.mark.parametrize("k0", [0, 1])
def test_k0_classify_fill_equals_single_kernel(k0):
    # K0 (classify_fill_kernel) writes the records of blocks no instance can be seen from and lists the others for K1.
    # Forced on and off (bvht_set_option): hit records AND shaded frames must equal the oracle either way -- ragged tile sizes,
    # sub-regions, tile-row shards into one buffer.
    rng = np.random.default_rng(21)
    blases = [SB.oracle_blas("teapot.obj"), SB.oracle_blas("cube.obj")]
    scene = O.Scene(blases, _random_instances(rng, 7, 2.5))
    cam = O.camera_symmetric_fov(90.0, 1.0, 1.0, [0.3, 0.2, -7], [0, 0, 1], [1, 0, 0], [0, 1, 0])
    w, h = 403, 229
    fcam = SB.to_ffi_camera(cam)
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        eng.set_option(_ffi.OPT_K0, k0)
        SB.upload_scene(eng, scene)
        for tile in (8, 5, 16):
            ref = scene.render(cam, w, h, tile=tile, threads=NTHREADS)
            assert 0.01 < (ref["id"] != O.MISS_ID).mean() < 0.5
            assert_strict(eng.trace_primary(fcam, w, h, tile=tile), ref)
        ref = scene.render(cam, w, h, threads=NTHREADS)
        # a sub-region: pixels outside keep the caller's initial records
        region = (37, 19, 301, 180)
        got = eng.trace_primary(fcam, w, h, region=region)
        inside = np.zeros((h, w), bool)
        inside[region[1]:region[3], region[0]:region[2]] = True
        inside = inside.reshape(-1)
        assert got[inside].tobytes() == ref[inside].tobytes()
        assert (got["id"][~inside] == O.MISS_ID).all()
        # fused shading through render_frame (bands) and the depth shader
        frame, hits = eng.render_frame(fcam, w, h, Engine.shade_depth(80.0, 3.0), want_hits=True)
        assert frame.tobytes() == O.shade(1, ref, 80.0, 3.0).tobytes()
        assert hits.tobytes() == ref.tobytes()


# ------------------------------------------------------------------------------------------ per-triangle block coverage
@pytest.mark.parametrize("cover", [0, 1])
@pytest.mark.parametrize("camera_z,n,seed", [(-7.0, 7, 31), (-2.2, 9, 32), (-0.4, 12, 33)])
def test_coverage_raster_equals_plain_trace(cover, camera_z, n, seed):
    # cover_kernels.cu marks, per 8x4 block, the instances whose (conservatively grown) triangle boxes project onto it; the trace
    # kernels drop the other instances from the block's candidates.  Forced on and off (bvht_set_option): records and shaded frames
    # equal the oracle either way -- also with the camera INSIDE the cloud of instances (boxes behind / across the eye plane
    # make an instance "visible everywhere"), with rotated instances, sub-regions and the banded host path.
    rng = np.random.default_rng(seed)
    blases = [SB.oracle_blas("teapot.obj"), SB.oracle_blas("cube.obj")]
    scene = O.Scene(blases, _random_instances(rng, n, 2.5))
    cam = O.camera_symmetric_fov(90.0, 1.0, 1.0, [0.2, 0.1, camera_z], [0, 0, 1], [1, 0, 0], [0, 1, 0])
    w, h = 416, 232
    fcam = SB.to_ffi_camera(cam)
    ref = scene.render(cam, w, h, threads=NTHREADS)
    assert (ref["id"] != O.MISS_ID).mean() > 0.01
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        eng.set_option(_ffi.OPT_COVER, cover)
        SB.upload_scene(eng, scene)
        for _ in range(2):
            assert_strict(eng.trace_primary(fcam, w, h), ref)
        region = (40, 24, 300, 200)
        got = eng.trace_primary(fcam, w, h, region=region)
        inside = np.zeros((h, w), bool)
        inside[region[1]:region[3], region[0]:region[2]] = True
        inside = inside.reshape(-1)
        assert got[inside].tobytes() == ref[inside].tobytes()
        frame, hits = eng.render_frame(fcam, w, h, Engine.shade_depth(80.0, 3.0), want_hits=True)
        assert frame.tobytes() == O.shade(1, ref, 80.0, 3.0).tobytes()
        assert hits.tobytes() == ref.tobytes()
