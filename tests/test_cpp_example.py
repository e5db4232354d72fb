"""The C++ example (examples/sixteen_armadillos.cpp, the reference's example written against the C++ host mirror)
compiles against include/bvht.h + host/bvhtracer.hpp; without a GPU it fails loudly, with one it reproduces the frame
the Python-driven path renders."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bvhtracer_b200", "lib")


def build_example(tmp_path):
    from bvhtracer_b200 import build as bvht_build
    bvht_build.build_all()
    exe = str(tmp_path / "sixteen_armadillos")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", os.path.join(ROOT, "examples", "sixteen_armadillos.cpp"),
                           "-L" + LIB, "-lbvht_cuda", "-Wl,-rpath," + LIB, "-o", exe])
    return exe


def test_cpp_example_compiles_and_has_no_cpu_fallback(tmp_path):
    from bvhtracer_b200 import _ffi
    exe = build_example(tmp_path)
    if _ffi.load().bvht_device_count() > 0:
        pytest.skip("a CUDA device is present")
    p = subprocess.run([exe, os.path.join(ROOT, "assets", "armadillo.tri.f32"), "1"], capture_output=True, text=True)
    assert p.returncode == 1 and "no CUDA device" in p.stderr


@pytest.mark.gpu
def test_cpp_example_matches_python_driven_frames(tmp_path):
    from bvhtracer_b200 import examples, host
    exe = build_example(tmp_path)
    frames = 5
    p = subprocess.run([exe, os.path.join(ROOT, "assets", "armadillo.tri.f32"), str(frames)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    checksum_cpp = int(p.stdout.strip().split()[-1], 16)
    anim = examples.GridAnimation()
    scene, models = host.build_scene(examples.sixteen_armadillos(0))
    renderer = host.Renderer()
    state = host.RendererState(host.depth_pipeline(80.0, 3.0), 640, 640)
    renderer.render(state, scene)
    for _ in range(frames):
        anim.update()
        for i, o in enumerate(anim.objects()):
            scene.set_transform(i, host.object_transform(o))
        scene.rebuild()
        renderer.render(state, scene)
    fb = state.frame_buffer().astype(np.uint64)
    idx = np.arange(fb.size, dtype=np.uint64) | np.uint64(1)
    checksum_py = int(np.bitwise_xor.reduce((fb * idx) & np.uint64(0xFFFFFFFF)))
    assert checksum_cpp == checksum_py
