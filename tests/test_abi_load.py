"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/bvht.h declares; without a device it reports BVHT_ERR_NO_DEVICE (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import bvhtracer_b200
from bvhtracer_b200 import _ffi, build as bvht_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bvht.h")).read()
    return sorted(set(re.findall(r"^BVHT_API\s+[\w\s\*]+?\b(bvht_\w+)\s*\(", text, flags=re.M)))


def test_library_builds_and_exports_every_declared_symbol():
    path = bvht_build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bvht.h but not exported"
    assert sorted(n for n, _, _ in _ffi.SYMBOLS) == names          # the Python binding covers the whole header


def test_abi_version_and_struct_sizes():
    lib = _ffi.load()
    assert lib.bvht_abi_version() == 5
    assert _ffi.BVH_NODE.itemsize == 32          # bvh.rs:720-723
    assert _ffi.TLAS_NODE.itemsize == 32
    assert _ffi.HIT.itemsize == 16               # intersection.rs:77-83
    assert _ffi.INSTANCE.itemsize == 68 and _ffi.CAMERA.itemsize == 100 and _ffi.RAY.itemsize == 28
    assert lib.bvht_status_string(-2).decode().startswith("no CUDA device")


def test_python_constants_match_the_header():
    # bvht_option, flags and status codes are spelled twice (include/bvht.h and bvhtracer_b200/_ffi.py): keep them equal
    text = open(os.path.join(ROOT, "include", "bvht.h")).read()
    enum = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"\b(BVHT_[A-Z0-9_]+)\s*=\s*(-?(?:0x[0-9a-fA-F]+|\d+))\b", text)}
    for name in ("COVER", "K0", "BANDS", "BAND_ORDER", "COPY_STREAMS", "TIMELINE"):
        assert getattr(_ffi, "OPT_" + name) == enum["BVHT_OPT_" + name], name
    assert len({enum[k] for k in enum if k.startswith("BVHT_OPT_")}) == 6          # no two options share a number
    assert _ffi.ERR_NO_DEVICE == enum["BVHT_ERR_NO_DEVICE"]
    for name in ("DEPTH", "UV", "NORMAL", "TEXTURE"):
        assert getattr(_ffi, "SHADE_" + name) == enum["BVHT_SHADE_" + name], name


def test_host_mirror_exports_the_frames_in_flight_calls():
    # Renderer::render_begin / render_end and the batched set_transform of the C++ mirror (no device needed to load it)
    from bvhtracer_b200 import host
    L = host.lib()
    for name in ("bvhx_renderer_render", "bvhx_renderer_render_begin", "bvhx_renderer_render_end", "bvhx_scene_set_transforms"):
        assert hasattr(L, name), name
    assert callable(host.Renderer.render_begin) and callable(host.Renderer.render_end)


def test_no_device_is_an_error_not_a_fallback():
    lib = _ffi.load()
    if lib.bvht_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(bvhtracer_b200.BvhtError) as e:
        bvhtracer_b200.Engine()
    assert e.value.status == _ffi.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under bvhtracer_b200/ may reference it
    pkg = os.path.join(ROOT, "bvhtracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".inc")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "bvht_oracle" not in text, f


def test_release_library_reads_no_environment():
    # the tuning / diagnostic knobs (Knobs in csrc/bvht_api.cu) exist only in -DBVHT_EXPERIMENT builds: the product library
    # must hold none of their environment names, so that no variable can change what a frame computes or how it is copied back
    blob = open(bvht_build.build(), "rb").read()
    src = open(os.path.join(ROOT, "bvhtracer_b200", "csrc", "bvht_api.cu")).read()
    names = sorted(set(re.findall(r'"(BVHT_[A-Z0-9_]+)"', src)))
    assert len(names) >= 15, names
    for n in names:
        assert n.encode() not in blob, f"{n} is compiled into the release library"
    for unit in ("bvht_api.cu", "cover_kernels.cu", "upload_kernels.cu", "refit_kernels.cu", "scene_kernels.cu", "build_kernels.cu",
                 "leaf_accel.cpp", "trace_kernels.cuh"):
        text = open(os.path.join(ROOT, "bvhtracer_b200", "csrc", unit)).read()
        if unit != "bvht_api.cu":
            assert "getenv" not in text, unit
        else:
            body = text[text.index("static Knobs read_knobs()"):]
            assert "getenv" not in body[body.index("\n}\n"):], "getenv outside read_knobs()"
