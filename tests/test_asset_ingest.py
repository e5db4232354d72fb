"""Asset ingest of the product path (mesh/decoders.rs:102-216 through the C++ host mirror): the packed soups under assets/ ARE
the product decoders' output for the reference's .tri / .obj files (checked against the reference tree where it exists), text
assets placed in the asset directory are decoded directly, and the decoders survive an exact text round trip."""
import os
import subprocess
import sys

import numpy as np
import pytest

from bvhtracer_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not on this machine (GPU box)")
def test_packed_assets_are_the_product_decoders_output_for_the_reference_files():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pack_assets.py"), REF, "--check"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert r.stdout.count("identical") == 7 and "DIFFERS" not in r.stdout


def tri_text(tris):
    return "\n".join(" ".join(repr(float(np.float32(x))) for x in t) for t in np.asarray(tris, "<f4").reshape(-1, 9)) + "\n"


def test_text_asset_in_the_asset_directory_is_decoded_directly(tmp_path):
    # a .tri text next to (instead of) the packed soup: same mesh, normals derived by the decoder (decoders.rs:120-124)
    soup = np.fromfile(os.path.join(ROOT, "assets", "unity.tri.f32"), "<f4").reshape(-1, 9)[:500]
    (tmp_path / "part.tri").write_text(tri_text(soup))
    soup.tofile(tmp_path / "packed.tri.f32")
    a = host.load_asset_mesh("part.tri", str(tmp_path))
    b = host.load_asset_mesh("packed.tri", str(tmp_path))
    assert a.primitives().tobytes() == soup.tobytes() == b.primitives().tobytes()
    assert a.normals().tobytes() == b.normals().tobytes()
    with pytest.raises(host.HostError):
        host.read_mesh_file(str(tmp_path / "packed.tri.f32"))          # unknown extension


def test_obj_round_trip_with_shared_vertices_and_normals(tmp_path):
    # OBJ: f64 parse narrowed to f32 (decoders.rs:184-199), faces index shared vertices / normals
    cube = np.fromfile(os.path.join(ROOT, "assets", "cube.obj.f32"), "<f4").reshape(-1, 3, 3)
    nrm = np.fromfile(os.path.join(ROOT, "assets", "cube.obj.normals.f32"), "<f4").reshape(-1, 3, 3)
    verts, vidx, norms, nidx = [], {}, [], {}
    faces = []
    for t, n in zip(cube, nrm):
        face = []
        for p, q in zip(t, n):
            kp, kq = p.tobytes(), q.tobytes()
            if kp not in vidx:
                vidx[kp] = len(verts) + 1; verts.append(p)
            if kq not in nidx:
                nidx[kq] = len(norms) + 1; norms.append(q)
            face.append(f"{vidx[kp]}//{nidx[kq]}")
        faces.append("f " + " ".join(face))
    text = "o cube\n" + "\n".join("v " + " ".join(repr(float(x)) for x in v) for v in verts) + "\n" \
        + "\n".join("vn " + " ".join(repr(float(x)) for x in v) for v in norms) + "\n" + "\n".join(faces) + "\n"
    (tmp_path / "cube2.obj").write_text(text)
    m = host.read_mesh_file(str(tmp_path / "cube2.obj"))
    assert m.primitives().tobytes() == cube.tobytes()
    assert m.normals().tobytes() == nrm.tobytes()
