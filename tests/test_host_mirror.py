"""The C++ host mirror (bvhtracer_b200/host) must hand the device exactly what the reference's host code
would: compare every host-built structure bit-for-bit with the oracle's restatement (CPU only)."""
import numpy as np
import pytest

import oracle_lib as O
import scene_build as SB
from bvhtracer_b200 import examples, host


@pytest.mark.parametrize("asset", ["cube.obj", "teapot.obj", "armadillo.tri", "bigben.tri", "unity.tri"])
def test_bvh_builder_matches_oracle(asset):
    tris = O.load_asset(asset)
    ref = O.Blas(tris)
    model = host.ModelBuilder().with_mesh(host.Mesh.from_triangles(tris)).build()
    nodes, used = model.nodes()
    assert used == ref.nodes_used
    assert len(nodes) == 2 * len(tris)
    assert nodes.tobytes() == ref.nodes.tobytes()                  # same tree, same boxes, same unused tail
    assert model.primitives().tobytes() == ref.tris.tobytes()      # same in-place reorder


def test_decoders_match_oracle():
    text = "# c\n0.5 -0.5 -0.1 0.57735 -0.5 0.1\\\n -0.57735 -0.5 0.1\n\n1e-3 2 3 4 5 6 7 8 999\n"
    assert host.TriMeshDecoder(text).read_mesh().primitives().tobytes() == O.parse_tri(text).tobytes()
    obj = "g q\nv 0.1 0 0\nv 1 0 0.3333333333\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\nf 1/1/1 3/1/1 4/1/1\nf 1 2 3 4\n"
    assert host.ObjMeshDecoder(obj).read_mesh().primitives().tobytes() == O.parse_obj(obj).tobytes()
    with pytest.raises(host.HostError):
        host.TriMeshDecoder("1 2 3 x").read_mesh()
    with pytest.raises(host.HostError):
        host.TriMeshDecoder("1 2 3").read_mesh()                   # not a multiple of nine floats
    with pytest.raises(host.HostError):
        host.ModelBuilder().with_mesh(host.Mesh.from_triangles(np.zeros((0, 9), np.float32))).build()


@pytest.mark.parametrize("name,arg", [("cube", None), ("two_armadillos", "canonical"), ("two_armadillos", "initial"),
                                      ("sixteen_armadillos", 0), ("sixteen_armadillos", 1), ("sixteen_armadillos", 40),
                                      ("trippy_teapots", 17), ("big_ben_clock", None)])
def test_scene_state_matches_oracle(name, arg):
    spec = examples.CONFIGS[name]() if arg is None else examples.CONFIGS[name](arg)
    ref_scene, ref_cam = SB.oracle_scene(spec)
    scene, _ = host.build_scene(spec)
    tlas, used = scene.tlas()
    assert used == ref_scene.tlas_used
    assert tlas[:used].tobytes() == ref_scene.tlas[:used].tobytes()
    for i in range(len(scene)):
        inv, bounds = scene.instance(i)
        assert inv.tobytes() == ref_scene.inst["inv"][i].tobytes()
        assert bounds[:3].tobytes() == ref_scene.bounds["min"][i].tobytes()
        assert bounds[3:].tobytes() == ref_scene.bounds["max"][i].tobytes()
    cam = scene.camera()
    assert cam.tobytes() == SB.to_ffi_camera(ref_cam).tobytes()


def test_scene_update_and_rebuild_match_oracle():
    # sixteen_armadillos.rs:132-163: set_transform x16 then Scene::rebuild, several frames in a row
    anim = examples.GridAnimation()
    spec = examples.sixteen_armadillos(0)
    scene, models = host.build_scene(spec)
    ref_scene, _ = SB.oracle_scene(spec)
    for frame in range(1, 6):
        anim.update()
        objs = anim.objects()
        for i, o in enumerate(objs):
            scene.set_transform(i, host.object_transform(o))
            ref_scene.set_transform(i, SB.object_matrix(o))
        scene.rebuild()
        ref_scene.rebuild()
        tlas, used = scene.tlas()
        assert used == ref_scene.tlas_used and tlas[:used].tobytes() == ref_scene.tlas[:used].tobytes()
        for i in range(16):
            assert scene.instance(i)[0].tobytes() == ref_scene.inst["inv"][i].tobytes()


def test_batched_set_transforms_is_the_per_object_loop():
    # Scene.set_transforms (one binding call) = set_transform(i, t) for i in 0..n-1: same TLAS, same inverses after rebuild
    anim = examples.GridAnimation()
    spec = examples.sixteen_armadillos(0)
    one, _ = host.build_scene(spec)
    batched, _ = host.build_scene(spec)
    for frame in range(3):
        anim.update()
        transforms = [host.object_transform(o) for o in anim.objects()]
        for i, t in enumerate(transforms):
            one.set_transform(i, t)
        batched.set_transforms(np.stack([t.matrix for t in transforms]))
        one.rebuild(); batched.rebuild()
        (ta, ua), (tb, ub) = one.tlas(), batched.tlas()
        assert ua == ub and ta[:ua].tobytes() == tb[:ub].tobytes()
        for i in range(16):
            assert one.instance(i)[0].tobytes() == batched.instance(i)[0].tobytes()
    with pytest.raises(host.HostError):
        batched.set_transforms(np.zeros((17, 16), "<f4"))             # more transforms than objects


def test_transform_matches_oracle():
    rng = np.random.default_rng(5)
    for _ in range(50):
        s, t = rng.uniform(0.2, 3, 3), rng.uniform(-5, 5, 3)
        ax, az = rng.uniform(-3, 3, 2)
        a = host.Transform3.new(s, t, ax, az)
        b = O.transform_new_rot_xz(s, t, ax, az)
        assert a.matrix.tobytes() == b.tobytes()
        assert a.inverse().matrix.tobytes() == O.mat4_inverse(b).tobytes()


def test_decoder_normals_match_oracle():
    tris = O.load_asset("unity.tri")[:800]
    assert host.tri_face_normals(tris).tobytes() == O.tri_normals(tris).tobytes()
    obj = "g q\nv 0.1 0 0\nv 1 0 0.25\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nvn 0.6 0 0.8\nf 1//1 2//2 3//1\nf 1/1/2 3/1/1 4/1/2\nf 1 2 3 4\n"
    m = host.ObjMeshDecoder(obj).read_mesh()
    assert m.normals().tobytes() == O.parse_obj_normals(obj).tobytes()
    assert m.primitives().tobytes() == O.parse_obj(obj).tobytes()
    assert m.normals()[0].tolist() == [0, 0, 1, np.float32(0.6), 0, np.float32(0.8), 0, 0, 1]
    assert not m.normals()[2:].any()                     # faces without vn get zero normals (decoders.rs:176-181)


def test_decoder_texture_coordinates_and_mesh_builder():
    # decoders.rs:182-203: v/vt/vn and v/vt corners carry (vt.u, vt.v) as f32, the others Vector2::zero(); a fan keeps them
    obj = ("g q\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0.25 0.75\nvt 1 0.3333333333 0\nvt -2 5\nvn 0 0 1\n"
           "f 1/1/1 2/2/1 3/3/1\nf 1/3 3/1 4/2\nf 1//1 2//1 3//1\nf 1/1 2/2 3/3 4/-1\n")
    m = host.ObjMeshDecoder(obj).read_mesh()
    tc = m.tex_coords()
    third = np.float32(0.3333333333)
    assert tc.shape == (5, 6)
    assert tc[0].tolist() == [0.25, 0.75, 1, third, -2, 5]
    assert tc[1].tolist() == [-2, 5, 0.25, 0.75, 1, third]
    assert not tc[2].any()
    assert tc[3].tolist() == [0.25, 0.75, 1, third, -2, 5] and tc[4].tolist() == [0.25, 0.75, -2, 5, -2, 5]
    # MeshBuilder::with_primitive(triangle, tex_coords, normals) (mesh.rs:189-198) through the flat API
    mesh = host.Mesh.from_triangles(examples.QUAD_TRIS, examples.QUAD_NORMALS).set_tex_coords(examples.QUAD_TEX_COORDS)
    assert mesh.tex_coords().tobytes() == examples.QUAD_TEX_COORDS.tobytes()
    with pytest.raises(host.HostError):
        host.Mesh.from_triangles(examples.QUAD_TRIS).set_tex_coords(np.zeros((3, 6), np.float32))
    model = host.ModelBuilder().with_mesh(mesh).with_texture(examples.brick_texture(8, 4)).build()
    assert model.primitives().shape == (2, 9)
