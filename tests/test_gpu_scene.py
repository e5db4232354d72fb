"""GPU: `SceneObject::set_transform` x n + `Tlas::rebuild` on the device (bvht_scene_set_transforms, K6) against the
oracle's restatement of scene_object.rs:60-75 and tlas.rs:179-250 -- inverse transforms, world bounds, TLAS node pool
(boxes, child packing, node numbering) bit for bit; then the traced frame against the host-fed path."""
import numpy as np
import pytest

import oracle_lib as O
import scene_build as SB
from bvhtracer_b200 import Engine, BvhtError, _ffi, examples
from bvhtracer_b200 import FLAG_LEAF_ACCEL, FLAG_STRICT

pytestmark = pytest.mark.gpu
F = np.float32
NTHREADS = max(1, O.max_threads())


def assert_same_scene(eng, scene, blas_ids):
    """Device rebuild of `scene`'s objects == the oracle's Scene.rebuild()."""
    mats = np.stack([m for _, m in scene.objects]).astype(F)
    ids = np.array([blas_ids[b] for b, _ in scene.objects], np.uint32)
    eng.scene_set_transforms(mats, ids)
    nodes, inst, bounds = eng.tlas_read(with_bounds=True)
    n = len(scene.objects)
    assert nodes.shape[0] == scene.tlas_used == 2 * n
    assert inst["transform_inv"].tobytes() == scene.inst["inv"].tobytes()
    assert bounds[:, :3].tobytes() == np.ascontiguousarray(scene.bounds["min"]).tobytes()
    assert bounds[:, 3:].tobytes() == np.ascontiguousarray(scene.bounds["max"]).tobytes()
    assert nodes.tobytes() == scene.tlas[:scene.tlas_used].tobytes()


@pytest.mark.parametrize("flags", [FLAG_STRICT, FLAG_STRICT | FLAG_LEAF_ACCEL])
@pytest.mark.parametrize("name,frames", [("cube", [0]), ("two_armadillos", ["initial", "canonical"]),
                                         ("sixteen_armadillos", [0, 1, 37]), ("trippy_teapots", [0, 5, 59])])
def test_device_scene_update_matches_host_on_examples(name, frames, flags):
    W = H = 160
    with Engine(flags=flags) as eng:
        blas_ids = None
        for fr in frames:
            if name == "cube":
                spec = getattr(examples, name)()
            else:
                spec = getattr(examples, name)(fr)
            scene, cam = SB.oracle_scene(spec)
            if blas_ids is None:
                blas_ids = [eng.blas_create(b.tris, b.nodes.view(_ffi.BVH_NODE), b.nodes_used) for b in scene.blases]
            assert_same_scene(eng, scene, blas_ids)
            got = eng.trace_primary(SB.to_ffi_camera(cam), W, H)
            ref = scene.render(cam, W, H, threads=NTHREADS)
            assert got.tobytes() == ref.tobytes()


def random_transforms(rng, n, spread):
    mats = np.zeros((n, 16), F)
    for i in range(n):
        s = rng.uniform(0.2, 2.0, 3) if rng.random() < 0.5 else [rng.uniform(0.2, 2.0)] * 3
        t = rng.uniform(-spread, spread, 3)
        mats[i] = O.transform_new_rot_xz(s, t, float(rng.uniform(-3.2, 3.2)), float(rng.uniform(-3.2, 3.2)))
    return mats


@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 31, 32, 33, 64, 65, 200, 257, 1000, 1025, 1500])
def test_device_tlas_rebuild_random_scenes(n):
    # one warp (n <= 64), 256 threads (n <= 1024) and 1024 threads: every block shape of the kernel
    rng = np.random.default_rng(n)
    cube = SB.oracle_blas("cube.obj")
    teapot = SB.oracle_blas("teapot.obj")
    mats = random_transforms(rng, n, spread=3.0 * n ** (1.0 / 3.0))
    models = rng.integers(0, 2, n)
    scene = O.Scene([cube, teapot], [(int(models[i]), mats[i]) for i in range(n)])
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        ids = [eng.blas_create(b.tris, b.nodes.view(_ffi.BVH_NODE), b.nodes_used) for b in scene.blases]
        assert_same_scene(eng, scene, ids)
        if n <= 257:
            cam = O.camera_symmetric_fov(90.0, 1.0, 1.0, [0, 0, -4.0 * n ** (1.0 / 3.0)], [0, 0, 1], [1, 0, 0], [0, 1, 0])
            got = eng.trace_primary(SB.to_ffi_camera(cam), 96, 96)
            ref = scene.render(cam, 96, 96, threads=NTHREADS)
            assert got.tobytes() == ref.tobytes()


def test_device_tlas_rebuild_ties_and_coincident_objects():
    # identical boxes everywhere: every arg-min is a tie, the FIRST candidate must win (tlas.rs:195 strict <)
    cube = SB.oracle_blas("cube.obj")
    for n in (2, 7, 40, 110):                                # (coincident piles cluster into chains: depth ~ n / 2 <= 64)
        mats = np.tile(O.mat4_identity(), (n, 1))
        mats[n // 2:, 12] = 5.0                              # two piles of coincident objects
        scene = O.Scene([cube], [(0, mats[i]) for i in range(n)])
        with Engine(flags=FLAG_STRICT) as eng:
            ids = [eng.blas_create(cube.tris, cube.nodes.view(_ffi.BVH_NODE), cube.nodes_used)]
            assert_same_scene(eng, scene, ids)


def test_device_tlas_rebuild_on_a_line():
    # objects on a line with growing gaps: deep, chain-like trees and the `a == last list position` case (tlas.rs:238-243)
    cube = SB.oracle_blas("cube.obj")
    for n, ratio in ((12, 1.7), (40, 1.2), (60, 1.05)):
        mats = np.tile(O.mat4_identity(), (n, 1))
        mats[:, 12] = np.cumsum(3.0 * ratio ** np.arange(n)).astype(F)
        for order in (np.arange(n), np.arange(n)[::-1], np.random.default_rng(n).permutation(n)):
            scene = O.Scene([cube], [(0, mats[i]) for i in order])
            with Engine(flags=FLAG_STRICT) as eng:
                ids = [eng.blas_create(cube.tris, cube.nodes.view(_ffi.BVH_NODE), cube.nodes_used)]
                assert_same_scene(eng, scene, ids)


def test_device_scene_update_follows_device_refit():
    # Model::bounds() is the root box: after animate + refit on the device the instance's world box must follow
    # (the example itself never calls set_transform: its object keeps Aabb::new_empty bounds, scene_object.rs:101-107)
    blas = O.Blas(O.load_asset("bigben.tri"))               # private copy: the cached one must stay pristine
    scene = O.Scene([blas], [(0, O.mat4_identity())])
    cam = SB.oracle_camera(examples.big_ben_clock().camera)
    rng = np.random.default_rng(3)
    with Engine(flags=FLAG_STRICT | FLAG_LEAF_ACCEL) as eng:
        ids = [eng.blas_create(blas.tris, blas.nodes.view(_ffi.BVH_NODE), blas.nodes_used)]
        moved = blas.tris.copy()
        real = np.abs(moved).max(axis=1) < 100.0             # keep the sentinel where it is
        moved[real] = (moved[real].reshape(-1, 3, 3) * F(1.5) + rng.normal(size=(int(real.sum()), 3, 3)).astype(F) * F(0.01)).reshape(-1, 9)
        blas.tris[:] = moved
        blas.refit()
        eng.blas_update_vertices(ids[0], moved)
        eng.blas_refit(ids[0])
        m = O.transform_new_rot_xz([1.5, 0.5, 1.0], [0.3, -0.2, 0.1], 0.4, -0.9)
        scene.set_transform(0, m)
        scene.rebuild()                                      # reads blas.bounds() = the refitted root box
        assert_same_scene(eng, scene, ids)
        got = eng.trace_primary(SB.to_ffi_camera(cam), 128, 128)
        ref = scene.render(cam, 128, 128, threads=NTHREADS)
        assert got.tobytes() == ref.tobytes()


def test_device_scene_update_errors():
    cube = SB.oracle_blas("cube.obj")
    with Engine(flags=FLAG_STRICT) as eng:
        bid = eng.blas_create(cube.tris, cube.nodes.view(_ffi.BVH_NODE), cube.nodes_used)
        ident = O.mat4_identity()
        with pytest.raises(BvhtError):                       # unknown model
            eng.scene_set_transforms([ident], [bid + 7])
        with pytest.raises(BvhtError):                       # empty scene
            eng.scene_set_transforms(np.zeros((0, 16), F), np.zeros(0, np.uint32))
        singular = ident.copy(); singular[0] = 0.0           # scale x = 0: det == 0, the reference unwraps None
        with pytest.raises(BvhtError, match="singular"):
            eng.scene_set_transforms([ident, singular], [bid, bid])
        cam = SB.to_ffi_camera(O.camera_symmetric_fov(90.0, 1.0, 1.0, [0, 0, -4], [0, 0, 1], [1, 0, 0], [0, 1, 0]))
        with pytest.raises(BvhtError):                       # a failed update leaves no traceable scene behind
            eng.trace_primary(cam, 32, 32)
        huge = ident.copy(); huge[12] = 3e38                 # bounds overflow to inf: no merge candidate (area not < MAX)
        with pytest.raises(BvhtError):
            eng.scene_set_transforms([ident, huge], [bid, bid])
        pile = np.tile(ident, (300, 1))                      # 300 coincident objects cluster into a chain deeper than the
        with pytest.raises(BvhtError, match="depth"):        # traversal stack: rejected exactly like bvht_tlas_set would
            eng.scene_set_transforms(pile, [bid] * 300)
        eng.scene_set_transforms([ident, ident], [bid, bid]) # and the context is still usable
        eng.trace_primary(cam, 32, 32)


def test_host_mirror_update_transforms_on_device():
    # CudaPathTracer::update_transforms: the host Scene adopts the device's inverses, bounds and TLAS; frames stay identical
    from bvhtracer_b200 import host
    anim = examples.GridAnimation()
    scene, models = host.build_scene(examples.sixteen_armadillos(0))
    renderer = host.Renderer(flags=FLAG_STRICT | FLAG_LEAF_ACCEL)
    w, h = 320, 176
    state = host.RendererState(host.depth_pipeline(80.0, 3.0), w, h, keep_hits=True)
    for frame in range(4):
        if frame > 0:
            anim.update()
            renderer.update_transforms(scene, [host.object_transform(o) for o in anim.objects()])
        assert renderer.render(state, scene) == w * h
        ref_scene, ref_cam = SB.oracle_scene(examples.sixteen_armadillos(frame))
        tl, used = scene.tlas()
        assert used == ref_scene.tlas_used and tl[:used].tobytes() == ref_scene.tlas[:used].tobytes()
        for i in range(len(scene)):
            inv, b = scene.instance(i)
            assert inv.tobytes() == ref_scene.inst["inv"][i].tobytes()
            assert b[:3].tobytes() == ref_scene.bounds["min"][i].tobytes() and b[3:].tobytes() == ref_scene.bounds["max"][i].tobytes()
        ref = ref_scene.render(ref_cam, w, h, threads=NTHREADS)
        assert state.hits().tobytes() == ref.tobytes()
        assert state.frame_buffer().tobytes() == O.shade(1, ref).tobytes()
