"""GPU: BvhBuilder::build_for on the device (bvht_blas_build / bvht_blas_rebuild) against the oracle's host build --
node pool, node numbering and the in-place reordering of the triangles, bit for bit."""
import numpy as np
import pytest

import oracle_lib as O
import scene_build as SB
from bvhtracer_b200 import Engine, BvhtError, _ffi, examples
from bvhtracer_b200 import FLAG_LEAF_ACCEL, FLAG_STRICT

pytestmark = pytest.mark.gpu
F = np.float32
NTHREADS = max(1, O.max_threads())


def assert_same_build(eng, bid, tris_in, prev_perm=None):
    ref = O.Blas(tris_in)                                   # orc_bvh_build: reorders a private copy
    n, used = eng.blas_info(bid)
    assert n == ref.n_tris and used == ref.nodes_used, (n, used, ref.nodes_used)
    nodes = eng.blas_read_nodes(bid, used)
    assert nodes.tobytes() == ref.nodes[:used].tobytes()
    got = eng.blas_read_triangles(bid, n)
    assert got.tobytes() == ref.tris.tobytes()
    perm = eng.blas_read_permutation(bid, n)
    assert sorted(perm.tolist()) == list(range(n))
    if prev_perm is not None:                               # rebuild: the permutation composes with the earlier one
        inv = np.empty(n, np.int64); inv[prev_perm] = np.arange(n)
        perm = inv[perm]
    assert np.asarray(tris_in, F).reshape(-1, 9)[perm].tobytes() == got.tobytes()
    return ref


@pytest.mark.parametrize("asset", ["cube.obj", "teapot.obj", "armadillo.tri", "bigben.tri", "unity.tri"])
def test_device_build_matches_host_build_on_assets(asset):
    tris = O.load_asset(asset)                              # file order, sentinel included
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        bid = eng.blas_build(tris)
        ref = assert_same_build(eng, bid, tris)
        st = eng.stats()
        assert st["last_build_levels"] >= 1 and st["last_build_ms"] > 0
    assert ref.nodes_used >= 2


@pytest.mark.parametrize("seed", range(6))
def test_device_build_random_soups(seed):
    # random soups WITHOUT the sentinel and not straddling the origin: deeper trees, many small nodes, ragged chunk edges
    rng = np.random.default_rng(seed)
    n = int(rng.choice([1, 2, 3, 7, 511, 512, 513, 1500, 4097, 20000]))
    c = rng.uniform(-1, 1, (n, 1, 3)) * rng.choice([1.0, 30.0]) + rng.choice([0.0, 50.0])
    tris = (c + rng.normal(size=(n, 3, 3)) * rng.choice([0.01, 0.3])).astype(F).reshape(n, 9)
    if seed % 2:
        tris[::5] = tris[0]                                  # duplicates: identical centroids
    with Engine(flags=FLAG_STRICT) as eng:
        assert_same_build(eng, eng.blas_build(tris), tris)


def test_device_build_degenerate_inputs():
    with Engine(flags=FLAG_STRICT) as eng:
        same = np.tile(np.array([[1, 2, 3, 2, 2, 3, 1, 3, 3]], F), (700, 1))          # every centroid equal: no axis to split
        assert_same_build(eng, eng.blas_build(same), same)
        line = np.zeros((1300, 9), F)                                                  # centroids on one axis only
        line[:, 0::3] = np.arange(1300, dtype=F)[:, None] + np.array([0, 1, 0], F)
        line[:, 1] = 1
        assert_same_build(eng, eng.blas_build(line), line)
        with pytest.raises(BvhtError):
            eng.blas_build(np.zeros((0, 9), F))


def test_device_built_scene_traces_like_the_oracle():
    # sixteen_armadillos with the BLAS built on the device from the file-order mesh: hit records bit-identical
    spec = examples.sixteen_armadillos(7)
    scene, cam = SB.oracle_scene(spec)
    ref = scene.render(cam, 320, 180, threads=NTHREADS)
    for flags in (FLAG_STRICT, FLAG_STRICT | FLAG_LEAF_ACCEL):
        with Engine(flags=flags) as eng:
            bid = eng.blas_build(O.load_asset("armadillo.tri"))
            SB.upload_scene(eng, scene, blas_ids=[bid])
            got = eng.trace_primary(SB.to_ffi_camera(cam), 320, 180)
        assert got.tobytes() == ref.tobytes()


@pytest.mark.parametrize("flags", [FLAG_STRICT, FLAG_STRICT | FLAG_LEAF_ACCEL], ids=["brute", "accel"])
def test_rebuild_after_animation_matches_host_rebuild(flags):
    # bench_bvh_refit_rebuild.rs's two arms: refit keeps the topology, rebuild runs BvhBuilder::build_for on the moved mesh
    spec = examples.big_ben_clock()
    _, cam = SB.oracle_scene(spec)
    blas = O.Blas(O.load_asset("bigben.tri"))
    anim = examples.BigBenAnimation(blas.tris)
    with Engine(flags=flags) as eng:
        bid = eng.blas_build(O.load_asset("bigben.tri"))
        scene = O.Scene([blas], [(0, O.mat4_identity())], with_transform=False)
        SB.upload_scene(eng, scene, blas_ids=[bid])
        for _ in range(2):
            verts = anim.animate()                           # new positions, in the CURRENT (reordered) order
            prev = eng.blas_read_permutation(bid, blas.n_tris)
            eng.blas_update_vertices(bid, verts)
            eng.blas_rebuild(bid)
            rebuilt = assert_same_build(eng, bid, verts, prev_perm=prev)     # host rebuild of the same moved mesh
            scene = O.Scene([rebuilt], [(0, O.mat4_identity())], with_transform=False)
            ref = scene.render(cam, 256, 144, threads=NTHREADS)
            got = eng.trace_primary(SB.to_ffi_camera(cam), 256, 144)
            assert got.tobytes() == ref.tobytes()
            anim = examples.BigBenAnimation(rebuilt.tris)    # keep animating the reordered mesh, like the example would


def test_host_mirror_builds_models_on_the_device():
    # ModelBuilder::build through CudaPathTracer::build_model: same Model as the host build (nodes, reordered primitives),
    # renders the same frames; rebuild_model after an animation step equals a host rebuild of the moved mesh.
    from bvhtracer_b200 import host
    spec = examples.sixteen_armadillos(3)
    renderer = host.Renderer(flags=FLAG_STRICT | FLAG_LEAF_ACCEL)
    mesh = host.load_asset_mesh("armadillo.tri")
    on_device = renderer.build_model(mesh)
    on_host = host.ModelBuilder().with_mesh(host.load_asset_mesh("armadillo.tri")).build()
    nd, ud = on_device.nodes(); nh, uh = on_host.nodes()
    assert ud == uh and nd[:ud].tobytes() == nh[:uh].tobytes()
    assert on_device.primitives().tobytes() == on_host.primitives().tobytes()
    scene, _ = host.build_scene(spec, models=[on_device])
    w, h = 320, 180
    state = host.RendererState(host.depth_pipeline(), w, h, keep_hits=True)
    renderer.render(state, scene)
    ref_scene, ref_cam = SB.oracle_scene(spec)
    assert state.hits().tobytes() == ref_scene.render(ref_cam, w, h, threads=NTHREADS).tobytes()

    # big_ben: animate, rebuild on the device, compare with the oracle's rebuild
    bb = renderer.build_model(host.load_asset_mesh("bigben.tri"))
    anim = examples.BigBenAnimation(bb.primitives())
    moved = anim.animate()
    bb.set_primitives(moved)
    renderer.rebuild_model(bb)
    ref = O.Blas(moved)
    nodes, used = bb.nodes()
    assert used == ref.nodes_used and nodes[:used].tobytes() == ref.nodes[:used].tobytes()
    assert bb.primitives().tobytes() == ref.tris.tobytes()
    spec5 = examples.big_ben_clock()
    scene5, _ = host.build_scene(spec5, models=[bb])
    state5 = host.RendererState(host.intersection_pipeline(), 256, 144, keep_hits=True)
    renderer.render(state5, scene5)
    _, cam5 = SB.oracle_scene(spec5)
    ref_scene5 = O.Scene([ref], [(0, O.mat4_identity())], with_transform=False)
    assert state5.hits().tobytes() == ref_scene5.render(cam5, 256, 144, threads=NTHREADS).tobytes()


def test_device_build_at_the_maximum_primitive_count():
    # 2^20 triangles: the 20-bit primitive index of InstancePrimitiveIndex is the limit of the boundary (bvht.h);
    # 2049 chunks at the root, a 258-node tree, the leaf accelerator over 1 M triangles on top of it
    rng = np.random.default_rng(1)
    n = 1 << 20
    c = rng.uniform(-30, 30, (n, 1, 3))
    tris = (c + rng.normal(size=(n, 3, 3)) * 0.05).astype(F).reshape(n, 9)
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        bid = eng.blas_build(tris)
        ref = assert_same_build(eng, bid, tris)
        assert ref.nodes_used > 100
        # and it traces: a handful of rays against the brute-force answer of the oracle
        scene = O.Scene([ref], [(0, O.mat4_identity())])
        rays = np.zeros((64, 7), F)
        rays[:, 0:3] = rng.uniform(-40, 40, (64, 3)); rays[:, 2] = -80
        rays[:, 5] = 1.0; rays[:, 6] = O.FLT_MAX
        SB.upload_scene(eng, scene, blas_ids=[bid])
        got = eng.trace_rays(rays)
        exp = scene.trace_rays(rays, threads=NTHREADS)
        assert got.tobytes() == exp.tobytes()
        with pytest.raises(BvhtError):
            eng.blas_build(np.zeros((n + 1, 9), F))          # one more than the index can address
