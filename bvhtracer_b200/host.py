"""Python face of the C++ host mirror (bvhtracer_b200/host/bvhtracer.hpp via libbvht_host.so).

Same names and call shapes as the reference's Rust API (ModelBuilder, SceneObjectBuilder, SceneBuilder,
Camera, Renderer, ...), so that examples/ and tests/ read like the reference's own callers.  All logic is in
C++; the BVH / TLAS builds run on the host exactly as in the reference, the traversal runs on the GPU.
"""
import ctypes as C
import os

import numpy as np

from . import _ffi
from ._ffi import BVH_NODE, CAMERA, HIT, RAY, TLAS_NODE

HOST_LIB = os.path.join(_ffi.PKG, "lib", "libbvht_host.so")
_lib = None
_P = C.c_void_p


class HostError(RuntimeError):
    pass


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _release(obj, free_name):
    """Handle destructor; at interpreter shutdown module globals may already be gone, then the OS reclaims everything."""
    h = getattr(obj, "_h", None)
    if not h or _lib is None:
        return
    try:
        getattr(_lib, free_name)(h)
    except Exception:
        pass
    obj._h = None


def lib():
    global _lib
    if _lib is None:
        _ffi.load()                                      # libbvht_cuda.so first (the host mirror links against it)
        if not os.path.exists(HOST_LIB):
            raise ImportError(f"{HOST_LIB} is missing: build it with `python -m bvhtracer_b200.build`")
        L = C.CDLL(HOST_LIB)
        sig = {
            "bvhx_last_error": (C.c_char_p, []),
            "bvhx_mesh_from_triangles": (_P, [_P, _P, C.c_uint32]),
            "bvhx_mesh_normals": (_P, [_P]),
            "bvhx_mesh_set_tex_coords": (C.c_int, [_P, _P, C.c_uint32]),
            "bvhx_mesh_tex_coords": (_P, [_P, C.POINTER(C.c_uint32)]),
            "bvhx_model_build_textured": (_P, [_P, _P, C.c_uint32, C.c_uint32]),
            "bvhx_mesh_from_tri_text": (_P, [C.c_char_p, C.c_size_t]),
            "bvhx_mesh_from_obj_text": (_P, [C.c_char_p, C.c_size_t]),
            "bvhx_mesh_len": (C.c_uint32, [_P]),
            "bvhx_mesh_data": (_P, [_P]),
            "bvhx_mesh_free": (None, [_P]),
            "bvhx_model_build": (_P, [_P]),
            "bvhx_model_nodes": (_P, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
            "bvhx_model_tris": (_P, [_P, C.POINTER(C.c_uint32)]),
            "bvhx_model_set_vertices": (None, [_P, _P]),
            "bvhx_model_refit": (None, [_P]),
            "bvhx_model_free": (None, [_P]),
            "bvhx_transform_new": (None, [_P, _P, C.c_float, C.c_float, _P]),
            "bvhx_transform_from_scale_translation": (None, [_P, _P, _P]),
            "bvhx_transform_inverse": (C.c_int, [_P, _P]),
            "bvhx_camera_symmetric_fov": (_P, [C.c_float, C.c_float, C.c_float, C.c_float, _P, _P, _P, _P]),
            "bvhx_camera_box": (_P, [C.c_float] * 6 + [_P, _P, _P, _P]),
            "bvhx_camera_to_ffi": (None, [_P, _P]),
            "bvhx_camera_free": (None, [_P]),
            "bvhx_scene_builder_new": (_P, [_P]),
            "bvhx_scene_builder_add": (C.c_int, [_P, _P, _P]),
            "bvhx_scene_build": (_P, [_P]),
            "bvhx_scene_len": (C.c_uint32, [_P]),
            "bvhx_scene_set_transform": (C.c_int, [_P, C.c_uint32, _P]),
            "bvhx_scene_set_transforms": (C.c_int, [_P, _P, C.c_uint32]),
            "bvhx_scene_rebuild": (None, [_P]),
            "bvhx_scene_tlas": (_P, [_P, C.POINTER(C.c_uint32)]),
            "bvhx_scene_instance": (None, [_P, C.c_uint32, _P, _P]),
            "bvhx_scene_camera": (None, [_P, _P]),
            "bvhx_scene_free": (None, [_P]),
            "bvhx_renderer_new": (_P, [C.c_uint32, C.c_int, C.c_uint32]),
            "bvhx_renderer_ctx": (_P, [_P]),
            "bvhx_renderer_free": (None, [_P]),
            "bvhx_state_new": (_P, [C.c_uint32, C.c_float, C.c_float, _P, _P, C.c_uint32, C.c_uint32, C.c_int]),
            "bvhx_state_new_external": (_P, [C.c_uint32, C.c_float, C.c_float, _P, _P, C.c_uint32, C.c_uint32, _P]),
            "bvhx_state_free": (None, [_P]),
            "bvhx_state_frame": (_P, [_P]),
            "bvhx_state_hits": (_P, [_P]),
            "bvhx_renderer_render": (C.c_int64, [_P, _P, _P]),
            "bvhx_renderer_render_begin": (C.c_int64, [_P, _P, _P]),
            "bvhx_renderer_render_end": (C.c_int, [_P]),
            "bvhx_renderer_sync_scene": (C.c_int, [_P, _P]),
            "bvhx_renderer_update_transforms": (C.c_int, [_P, _P, _P, C.c_uint32]),
            "bvhx_renderer_build_model": (_P, [_P, _P]),
            "bvhx_renderer_rebuild_model": (C.c_int, [_P, _P]),
            "bvhx_renderer_intersect": (C.c_int, [_P, _P, _P, C.c_uint64, _P]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _err():
    return HostError(lib().bvhx_last_error().decode(errors="replace"))


def _nn(p):
    if not p:
        raise _err()
    return p


# ------------------------------------------------------------------ meshes / models
class Mesh:
    """mesh.rs Mesh<f32> (positions only)."""

    def __init__(self, handle):
        self._h = _nn(handle)

    @classmethod
    def from_triangles(cls, tris, normals=None):
        tris = np.ascontiguousarray(np.asarray(tris, "<f4").reshape(-1, 9))
        if normals is not None:
            normals = np.ascontiguousarray(np.asarray(normals, "<f4").reshape(-1, 9))
            assert normals.shape == tris.shape
        return cls(lib().bvhx_mesh_from_triangles(_ffi.ptr(tris), _ffi.ptr(normals) if normals is not None else None, tris.shape[0]))

    def normals(self):
        n = lib().bvhx_mesh_len(self._h)
        p = lib().bvhx_mesh_normals(self._h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(max(n, 1), 9))[:n].copy()

    def set_tex_coords(self, tex_coords):
        """TextureCoordinates<f32, 3> per primitive (mesh.rs:8-51): n x 6 floats"""
        tc = np.ascontiguousarray(np.asarray(tex_coords, "<f4").reshape(-1, 6))
        if lib().bvhx_mesh_set_tex_coords(self._h, _ffi.ptr(tc), tc.shape[0]) != 0:
            raise _err()
        return self

    def tex_coords(self):
        n = C.c_uint32()
        p = lib().bvhx_mesh_tex_coords(self._h, C.byref(n))
        if n.value == 0:
            return np.zeros((0, 6), "<f4")
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n.value, 6)).copy()

    def primitives(self):
        n = lib().bvhx_mesh_len(self._h)
        p = lib().bvhx_mesh_data(self._h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(max(n, 1), 9))[:n].copy()

    def __del__(self):
        try:
            _release(self, "bvhx_mesh_free")
        except Exception:          # interpreter shutdown: module globals already cleared
            pass


class TriMeshDecoder:
    """mesh/decoders.rs:84-134"""

    def __init__(self, text):
        self.text = text.encode() if isinstance(text, str) else bytes(text)

    def read_mesh(self):
        return Mesh(lib().bvhx_mesh_from_tri_text(self.text, len(self.text)))


class ObjMeshDecoder:
    """mesh/decoders.rs:137-216"""

    def __init__(self, text):
        self.text = text.encode() if isinstance(text, str) else bytes(text)

    def read_mesh(self):
        return Mesh(lib().bvhx_mesh_from_obj_text(self.text, len(self.text)))


class ModelInstance:
    """model.rs:16-60.  bvh nodes / reordered primitives are exposed for upload and parity checks."""

    def __init__(self, handle):
        self._h = _nn(handle)

    def nodes(self):
        used, total = C.c_uint32(), C.c_uint32()
        p = lib().bvhx_model_nodes(self._h, C.byref(used), C.byref(total))
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(total.value * 32,)).view(BVH_NODE)
        return arr.copy(), int(used.value)

    def primitives(self):
        n = C.c_uint32()
        p = lib().bvhx_model_tris(self._h, C.byref(n))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n.value, 9)).copy()

    def set_primitives(self, tris):
        """`model.borrow_mut().primitives_mut()[i] = ...` for every i (big_ben_clock.rs:97-102)."""
        tris = np.ascontiguousarray(np.asarray(tris, "<f4").reshape(-1, 9))
        lib().bvhx_model_set_vertices(self._h, _ffi.ptr(tris))

    def refit(self):
        lib().bvhx_model_refit(self._h)

    def __del__(self):
        try:
            _release(self, "bvhx_model_free")
        except Exception:          # interpreter shutdown: module globals already cleared
            pass


class ModelBuilder:
    """model.rs:117-146"""

    def __init__(self):
        self._mesh = None
        self._texture = None

    def with_mesh(self, mesh):
        self._mesh = mesh
        return self

    def with_texture(self, texels):
        """TextureMaterial::new(texture) (materials/material.rs:22-24): decoded Rgb<u8> texels[height, width, 3]"""
        texels = np.ascontiguousarray(np.asarray(texels, np.uint8))
        assert texels.ndim == 3 and texels.shape[2] == 3
        self._texture = texels
        return self

    def build(self):
        if self._texture is not None:
            t = self._texture
            return ModelInstance(lib().bvhx_model_build_textured(self._mesh._h, _ffi.ptr(t), t.shape[1], t.shape[0]))
        return ModelInstance(lib().bvhx_model_build(self._mesh._h))


# ------------------------------------------------------------------ transforms
class Transform3:
    """transform.rs:12-256 (column-major 16 floats)."""

    def __init__(self, m):
        self.matrix = np.asarray(m, "<f4").reshape(16).copy()

    @classmethod
    def identity(cls):
        return cls(np.eye(4, dtype="<f4"))

    @classmethod
    def new(cls, scale, translation, angle_x=0.0, angle_z=0.0):
        """Transform3::new(&scale, &translation, Rotation3::from_angle_x(ax) * Rotation3::from_angle_z(az))"""
        out = np.zeros(16, "<f4")
        lib().bvhx_transform_new(_f3(scale), _f3(translation), float(angle_x), float(angle_z), _ffi.ptr(out))
        return cls(out)

    @classmethod
    def from_scale_translation(cls, scale, translation):
        out = np.zeros(16, "<f4")
        lib().bvhx_transform_from_scale_translation(_f3(scale), _f3(translation), _ffi.ptr(out))
        return cls(out)

    def inverse(self):
        out = np.zeros(16, "<f4")
        if lib().bvhx_transform_inverse(_ffi.ptr(self.matrix), _ffi.ptr(out)) != 0:
            raise HostError("singular transform")
        return Transform3(out)


# ------------------------------------------------------------------ camera
class Camera:
    """camera.rs Camera<f32, PerspectiveProjection<f32>> from SymmetricFovSpec / BoxSpec + CameraAttitudeSpec."""

    def __init__(self, handle):
        self._h = _nn(handle)

    @classmethod
    def symmetric_fov(cls, fovy_deg, aspect, near, far, position, forward, right, up):
        return cls(lib().bvhx_camera_symmetric_fov(fovy_deg, aspect, near, far, _f3(position), _f3(forward), _f3(right), _f3(up)))

    @classmethod
    def box(cls, left, right_, bottom, top, near, far, position, forward, right, up):
        return cls(lib().bvhx_camera_box(left, right_, bottom, top, near, far, _f3(position), _f3(forward), _f3(right), _f3(up)))

    @classmethod
    def from_spec(cls, c):
        if c.box is not None:
            l, r, b, t = c.box
            return cls.box(l, r, b, t, c.near, 100.0, c.position, c.forward, c.right, c.up)
        return cls.symmetric_fov(c.fovy_deg, c.aspect, c.near, 10000.0, c.position, c.forward, c.right, c.up)

    def to_ffi(self):
        out = np.zeros(1, CAMERA)
        lib().bvhx_camera_to_ffi(self._h, _ffi.ptr(out))
        return out

    def __del__(self):
        try:
            _release(self, "bvhx_camera_free")
        except Exception:          # interpreter shutdown: module globals already cleared
            pass


# ------------------------------------------------------------------ scene
class Scene:
    """scene.rs:8-51"""

    def __init__(self, handle, models):
        self._h = _nn(handle)
        self._models = models                 # keep the ModelInstance handles alive

    def __len__(self):
        return int(lib().bvhx_scene_len(self._h))

    def set_transform(self, i, transform):
        """scene.get_mut_unchecked(i).set_transform(&t) (sixteen_armadillos.rs:143)"""
        if lib().bvhx_scene_set_transform(self._h, int(i), _ffi.ptr(transform.matrix)) != 0:
            raise _err()

    def set_transforms(self, matrices):
        """set_transform for objects 0..n-1 in one call; matrices: n x 16 f32, column-major"""
        m = np.ascontiguousarray(matrices, "<f4").reshape(-1, 16)
        if lib().bvhx_scene_set_transforms(self._h, _ffi.ptr(m), m.shape[0]) != 0:
            raise _err()

    def rebuild(self):
        lib().bvhx_scene_rebuild(self._h)

    def tlas(self):
        used = C.c_uint32()
        p = lib().bvhx_scene_tlas(self._h, C.byref(used))
        n = max(2 * len(self), 2)
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n * 32,)).view(TLAS_NODE)
        return arr.copy(), int(used.value)

    def instance(self, i):
        inv = np.zeros(16, "<f4")
        b = np.zeros(6, "<f4")
        lib().bvhx_scene_instance(self._h, int(i), _ffi.ptr(inv), _ffi.ptr(b))
        return inv, b

    def camera(self):
        out = np.zeros(1, CAMERA)
        lib().bvhx_scene_camera(self._h, _ffi.ptr(out))
        return out

    def __del__(self):
        try:
            _release(self, "bvhx_scene_free")
        except Exception:          # interpreter shutdown: module globals already cleared
            pass


class SceneBuilder:
    """scene.rs:53-97 with SceneObjectBuilder (scene_object.rs:92-137) folded in."""

    def __init__(self, camera):
        self._h = _nn(lib().bvhx_scene_builder_new(camera._h))
        self._models = []

    def with_object(self, model, transform=None):
        """SceneObjectBuilder::new(model, ..).with_transform(&t).build(); transform=None skips with_transform."""
        m = _ffi.ptr(transform.matrix) if transform is not None else None
        if lib().bvhx_scene_builder_add(self._h, model._h, m) != 0:
            raise _err()
        self._models.append(model)
        return self

    def build(self):
        h, self._h = self._h, None
        return Scene(lib().bvhx_scene_build(h), self._models)


# ------------------------------------------------------------------ renderer
class RendererState:
    """renderer.rs:76-102 with the accumulator + pixel shader pair given as a device shading pipeline."""

    def __init__(self, shading, width, height, keep_hits=False, frame=None):
        """frame: a caller-owned page-locked uint32 array of width * height pixels to render into (e.g. one frame in POSIX shared
        memory that several single-GPU processes fill, each the tile rows its integrator is sharded to)."""
        kind, scale, offset, hit, miss = shading
        self.width, self.height, self.keep_hits = int(width), int(height), bool(keep_hits)
        self._frame = frame
        if frame is not None:
            assert frame.dtype == np.uint32 and frame.size == self.width * self.height and frame.flags["C_CONTIGUOUS"]
            self._h = _nn(lib().bvhx_state_new_external(kind, scale, offset, (C.c_uint8 * 4)(*hit), (C.c_uint8 * 4)(*miss),
                                                        self.width, self.height, _ffi.ptr(frame)))
            return
        self._h = _nn(lib().bvhx_state_new(kind, scale, offset, (C.c_uint8 * 4)(*hit), (C.c_uint8 * 4)(*miss),
                                           self.width, self.height, int(keep_hits)))

    def frame_buffer(self):
        p = lib().bvhx_state_frame(self._h)
        if not p:
            raise HostError("frame buffer is allocated by the first render")
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(self.width * self.height,))

    def hits(self):
        p = lib().bvhx_state_hits(self._h)
        if not p:
            raise HostError("hit records were not requested (keep_hits=False) or nothing was rendered yet")
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(self.width * self.height * 16,)).view(HIT)

    def __del__(self):
        try:
            _release(self, "bvhx_state_free")
        except Exception:          # interpreter shutdown: module globals already cleared
            pass


def depth_pipeline(scale=80.0, offset=3.0):
    """DepthAccumulator::new() + DepthMappingShader::new(scale, offset)"""
    return (_ffi.SHADE_DEPTH, float(scale), float(offset), (0, 0, 0, 0), (0, 0, 0, 0))


def intersection_pipeline(hit=(255, 255, 255, 255), miss=(0, 0, 0, 255)):
    """IntersectionAccumulator + IntersectionShader::new(hit, miss)"""
    return (_ffi.SHADE_INTERSECTION, 0.0, 0.0, tuple(hit), tuple(miss))


def normal_pipeline():
    """NormalMappingAccumulator + RadianceToRgbShader (cube.rs, trippy_teapots.rs)"""
    return (_ffi.SHADE_NORMAL, 0.0, 0.0, (0, 0, 0, 0), (0, 0, 0, 0))


def texture_pipeline():
    """TextureMaterialAccumulator + RadianceToRgbShader (quad.rs)"""
    return (_ffi.SHADE_TEXTURE, 0.0, 0.0, (0, 0, 0, 0), (0, 0, 0, 0))


def uv_pipeline():
    """UvMappingAccumulator + RadianceToRgbShader"""
    return (_ffi.SHADE_UV, 0.0, 0.0, (0, 0, 0, 0), (0, 0, 0, 0))


class Renderer:
    """Renderer::new(Box::new(CudaPathTracer::new(flags))) (renderer.rs:388-400)."""

    def __init__(self, flags=_ffi.FLAG_STRICT | _ffi.FLAG_LEAF_ACCEL, device=0, tile=8):
        self._h = _nn(lib().bvhx_renderer_new(int(flags), int(device), int(tile)))
        self.flags = int(flags)

    def render(self, state, scene):
        """-> rays traced (usize), like Integrator::evaluate"""
        n = lib().bvhx_renderer_render(self._h, state._h, scene._h)
        if n < 0:
            raise _err()
        return int(n)

    def render_begin(self, state, scene):
        """render() without the wait: up to two frames in flight, each into its own RendererState (bvht_render_frame_begin)."""
        n = lib().bvhx_renderer_render_begin(self._h, state._h, scene._h)
        if n < 0:
            raise _err()
        return int(n)

    def render_end(self):
        """Wait for the oldest begun frame; its RendererState then holds the pixels (bvht_render_frame_end)."""
        if lib().bvhx_renderer_render_end(self._h) != 0:
            raise _err()

    def sync_scene(self, scene):
        if lib().bvhx_renderer_sync_scene(self._h, scene._h) != 0:
            raise _err()

    def update_transforms(self, scene, transforms):
        """set_transform for every object + scene.rebuild(), computed on the device; the host scene adopts the results."""
        m = np.ascontiguousarray(np.stack([np.asarray(t.matrix, "<f4").reshape(16) for t in transforms]))
        if lib().bvhx_renderer_update_transforms(self._h, scene._h, _ffi.ptr(m), m.shape[0]) != 0:
            raise _err()

    def build_model(self, mesh):
        """ModelBuilder::new().with_mesh(mesh).build() with BvhBuilder::build_for running on the device; the model is resident."""
        return ModelInstance(lib().bvhx_renderer_build_model(self._h, mesh._h))

    def rebuild_model(self, model):
        """Rebuild the model's BVH from its current vertices on the device (the alternative to refit)."""
        if lib().bvhx_renderer_rebuild_model(self._h, model._h) != 0:
            raise _err()

    def intersect(self, scene, rays):
        """Scene::intersect(&Ray) for a batch of rays (o, d, t)"""
        rays = np.ascontiguousarray(np.asarray(rays, "<f4").reshape(-1, 7))
        out = np.zeros(rays.shape[0], HIT)
        if lib().bvhx_renderer_intersect(self._h, scene._h, _ffi.ptr(rays), rays.shape[0], _ffi.ptr(out)) != 0:
            raise _err()
        return out

    def ctx_handle(self):
        return lib().bvhx_renderer_ctx(self._h)

    def engine(self):
        """The integrator's device context as an Engine (non-owning): resident buffers, sharding, IPC."""
        from .engine import Engine
        return Engine.from_handle(self.ctx_handle(), flags=self.flags)

    def stats(self):
        s = _ffi.Stats()
        _ffi.load().bvht_get_stats(C.c_void_p(self.ctx_handle()), C.byref(s))
        return s.as_dict()

    def set_stream(self, cuda_stream_handle):
        rc = _ffi.load().bvht_set_stream(C.c_void_p(self.ctx_handle()), C.c_void_p(cuda_stream_handle or 0))
        if rc != 0:
            raise HostError(f"bvht_set_stream failed: {rc}")

    def __del__(self):
        try:
            _release(self, "bvhx_renderer_free")
        except Exception:          # interpreter shutdown: module globals already cleared
            pass


# ------------------------------------------------------------------ example scenes through the mirror
def read_mesh_file(path):
    """`TriMeshDecoder::new(File::open(path)).read_mesh()` / `ObjMeshDecoder` by extension (mesh/decoders.rs:102-216): the text
    asset decoded by the C++ mirror's decoders -- the ingest path of the product."""
    text = open(path, "rb").read()
    if path.lower().endswith(".tri"):
        return TriMeshDecoder(text).read_mesh()
    if path.lower().endswith(".obj"):
        return ObjMeshDecoder(text).read_mesh()
    raise HostError(f"unknown mesh format: {path}")


def load_asset_mesh(name, asset_dir=None):
    """The example asset `name` (armadillo.tri, teapot.obj, ...): the text file itself when it lies in the asset directory
    (decoded by read_mesh_file), else its packed triangle soup assets/<name>.f32 -- the same decoders' output, written by
    tools/pack_assets.py where the reference tree exists (the .tri / .obj texts are the reference's files and are not copied
    into this repository; tests/test_asset_ingest.py keeps the two byte-identical)."""
    asset_dir = asset_dir or os.path.join(os.path.dirname(_ffi.PKG), "assets")
    if os.path.exists(os.path.join(asset_dir, name)):
        return read_mesh_file(os.path.join(asset_dir, name))
    tris = np.fromfile(os.path.join(asset_dir, name + ".f32"), dtype="<f4").reshape(-1, 9)
    npath = os.path.join(asset_dir, name + ".normals.f32")
    if os.path.exists(npath):                      # OBJ models: the `vn` normals of each face corner
        return Mesh.from_triangles(tris, np.fromfile(npath, dtype="<f4").reshape(-1, 9))
    if name.endswith(".tri"):                      # .tri models: face normals derived by the decoder (mesh/decoders.rs:120-124)
        return Mesh.from_triangles(tris, tri_face_normals(tris))
    return Mesh.from_triangles(tris)


def tri_face_normals(tris):
    """TriMeshDecoder's normals through the C++ decoder itself (text round trip is exact for f32 -> %.9g -> f32)."""
    text = "\n".join(" ".join(repr(float(np.float32(x))) for x in t) for t in np.asarray(tris, "<f4").reshape(-1, 9))
    return TriMeshDecoder(text).read_mesh().normals()


def object_transform(o):
    return Transform3.new(o.scale, o.translation, o.angle_x, o.angle_z)


def build_scene(spec, models=None):
    """SceneSpec (bvhtracer_b200/examples.py) -> (Scene, [ModelInstance]) the way the example's `new()` does."""
    from . import examples
    if models is None:
        models = []
        for a in spec.meshes:
            mesh = Mesh.from_triangles(examples.QUAD_TRIS) if a == "<quad>" else load_asset_mesh(a)
            models.append(ModelBuilder().with_mesh(mesh).build())
    sb = SceneBuilder(Camera.from_spec(spec.camera))
    for o in spec.objects:
        sb.with_object(models[o.model], object_transform(o) if o.with_transform else None)
    return sb.build(), models
