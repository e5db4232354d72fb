"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (bvhtracer_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ASSET_DIR = os.path.join(ROOT, "assets")

FLT_MAX = np.finfo(np.float32).max
MISS_ID = 0xFFFFFFFF

AABB = np.dtype([("min", "<f4", 3), ("max", "<f4", 3)])
BVH_NODE = np.dtype([("min", "<f4", 3), ("max", "<f4", 3), ("prim_count", "<u4"), ("left_first", "<u4")])
TLAS_NODE = np.dtype([("min", "<f4", 3), ("max", "<f4", 3), ("left_right", "<u4"), ("blas", "<u4")])
HIT = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("id", "<u4")])
INSTANCE = np.dtype([("inv", "<f4", 16), ("blas_id", "<u4")])
CAMERA = np.dtype([("tl", "<f4", 3), ("tr", "<f4", 3), ("bl", "<f4", 3), ("view_inv", "<f4", 16)])
RAY = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("rd", "<f4", 3), ("t", "<f4")])
assert BVH_NODE.itemsize == 32 and TLAS_NODE.itemsize == 32 and HIT.itemsize == 16


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("rays", "hits", "blas_nodes", "tlas_nodes", "inst", "tri_area", "tri_u", "tri_v", "tri_t", "box_tests")] + \
               [("max_blas_stack", C.c_uint32), ("max_tlas_stack", C.c_uint32)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class _Blas(C.Structure):
    _fields_ = [("tris", C.c_void_p), ("n_tris", C.c_uint32), ("nodes", C.c_void_p), ("nodes_used", C.c_uint32)]


class _Scene(C.Structure):
    _fields_ = [("tlas", C.c_void_p), ("tlas_nodes_used", C.c_uint32), ("inst", C.c_void_p), ("n_inst", C.c_uint32),
                ("blas", C.c_void_p), ("n_blas", C.c_uint32)]


_lib = None


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(ORACLE_DIR, "bvht_oracle.c")):
            build()
        _lib = C.CDLL(so)
        _lib.orc_bvh_build.restype = C.c_uint32
        _lib.orc_tlas_build.restype = C.c_uint32
        _lib.orc_parse_tri.restype = C.c_int64
        _lib.orc_parse_obj.restype = C.c_int64
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def f3(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(3))


# ------------------------------------------------------------------ primitives
def ray_new(o, d, t=FLT_MAX):
    r = np.zeros(1, RAY)
    lib().orc_ray_new(_p(f3(o)), _p(f3(d)), C.c_float(t), _p(r))
    return r


def normalize(v):
    out = np.zeros(3, np.float32)
    lib().orc_vec3_normalize(_p(f3(v)), _p(out))
    return out


def cross(a, b):
    out = np.zeros(3, np.float32)
    lib().orc_vec3_cross(_p(f3(a)), _p(f3(b)), _p(out))
    return out


def aabb_intersect(bmin, bmax, ray):
    box = np.zeros(1, AABB)
    box["min"][0] = bmin
    box["max"][0] = bmax
    t = C.c_float(0)
    ok = lib().orc_aabb_intersect(_p(box), _p(ray), C.byref(t))
    return np.float32(t.value) if ok else None


def triangle_intersect(tri9, ray):
    tri = np.ascontiguousarray(np.asarray(tri9, np.float32).reshape(9))
    out = np.zeros(3, np.float32)
    ok = lib().orc_triangle_intersect(_p(tri), _p(ray), _p(out))
    return out if ok else None


def triangle_centroid(tri9):
    tri = np.ascontiguousarray(np.asarray(tri9, np.float32).reshape(9))
    out = np.zeros(3, np.float32)
    lib().orc_triangle_centroid(_p(tri), _p(out))
    return out


# ------------------------------------------------------------------ assets
def load_asset(name):
    """Packed asset (assets/<name>.f32, N x 9 f32, file order; see tools/pack_assets.py)."""
    return np.fromfile(os.path.join(ASSET_DIR, name + ".f32"), dtype="<f4").reshape(-1, 9).copy()


def parse_tri(text):
    b = text.encode() if isinstance(text, str) else text
    p = C.POINTER(C.c_float)()
    n = lib().orc_parse_tri(b, C.c_size_t(len(b)), C.byref(p))
    if n < 0:
        raise ValueError(f"orc_parse_tri failed: {n}")
    arr = np.ctypeslib.as_array(p, shape=(max(n, 1), 9))[:n].copy() if n else np.zeros((0, 9), np.float32)
    lib().orc_free(p)
    return arr


def parse_obj(text):
    b = text.encode() if isinstance(text, str) else text
    p = C.POINTER(C.c_float)()
    n = lib().orc_parse_obj(b, C.c_size_t(len(b)), C.byref(p))
    if n < 0:
        raise ValueError(f"orc_parse_obj failed: {n}")
    arr = np.ctypeslib.as_array(p, shape=(max(n, 1), 9))[:n].copy() if n else np.zeros((0, 9), np.float32)
    lib().orc_free(p)
    return arr


# ------------------------------------------------------------------ BLAS
class Blas:
    """ModelBuilder::build (model.rs:140-144): builds the BVH and reorders the triangles in place."""

    def __init__(self, tris):
        self.tris = np.ascontiguousarray(np.asarray(tris, np.float32).reshape(-1, 9)).copy()
        n = self.tris.shape[0]
        self.nodes = np.zeros(max(2 * n, 2), BVH_NODE)
        self.nodes_used = int(lib().orc_bvh_build(_p(self.tris), C.c_uint32(n), _p(self.nodes)))

    @property
    def n_tris(self):
        return self.tris.shape[0]

    def refit(self):
        lib().orc_bvh_refit(_p(self.tris), _p(self.nodes), C.c_uint32(self.nodes_used))

    def bounds(self):
        b = np.zeros(1, AABB)
        b["min"][0] = self.nodes["min"][0]
        b["max"][0] = self.nodes["max"][0]
        return b

    def _c(self):
        return _Blas(self.tris.ctypes.data, self.n_tris, self.nodes.ctypes.data, self.nodes_used)

    def intersect(self, ray, counters=None):
        hit = np.zeros(1, HIT)
        cb = self._c()
        ok = lib().orc_bvh_intersect(C.byref(cb), _p(ray), _p(hit), C.byref(counters) if counters is not None else None)
        return hit[0] if ok else None


# ------------------------------------------------------------------ transforms
def mat4_identity():
    return np.eye(4, dtype=np.float32).reshape(16).copy()


def mat4_inverse(m):
    m = np.ascontiguousarray(np.asarray(m, np.float32).reshape(16))
    out = np.zeros(16, np.float32)
    if not lib().orc_mat4_inverse(_p(m), _p(out)):
        raise ValueError("singular matrix")
    return out


def transform_new_rot_xz(scale, trans, angle_x, angle_z):
    out = np.zeros(16, np.float32)
    lib().orc_transform_new_rot_xz(_p(f3(scale)), _p(f3(trans)), C.c_float(angle_x), C.c_float(angle_z), _p(out))
    return out


def transform_from_scale_translation(scale, trans):
    out = np.zeros(16, np.float32)
    lib().orc_transform_from_scale_translation(_p(f3(scale)), _p(f3(trans)), _p(out))
    return out


def transform_from_translation(trans):
    return transform_from_scale_translation([1, 1, 1], trans)


def transform_point(m, p):
    m = np.ascontiguousarray(np.asarray(m, np.float32).reshape(16))
    out = np.zeros(3, np.float32)
    lib().orc_transform_point(_p(m), _p(f3(p)), _p(out))
    return out


def instance_bounds(m, model_bounds):
    m = np.ascontiguousarray(np.asarray(m, np.float32).reshape(16))
    out = np.zeros(1, AABB)
    lib().orc_instance_bounds(_p(m), _p(model_bounds), _p(out))
    return out


def tlas_build(bounds):
    """bounds: AABB array of n world-space instance boxes -> (nodes[2n], nodes_used)."""
    bounds = np.ascontiguousarray(bounds)
    n = bounds.shape[0]
    nodes = np.zeros(max(2 * n, 2), TLAS_NODE)
    used = int(lib().orc_tlas_build(_p(bounds), C.c_uint32(n), _p(nodes)))
    return nodes, used


# ------------------------------------------------------------------ camera
def camera_symmetric_fov(fovy_deg, aspect, near, pos, fwd, right, up):
    cam = np.zeros(1, CAMERA)
    lib().orc_camera_symmetric_fov(C.c_float(fovy_deg), C.c_float(aspect), C.c_float(near),
                                   _p(f3(pos)), _p(f3(fwd)), _p(f3(right)), _p(f3(up)), _p(cam))
    return cam


def camera_box(left, right_, bottom, top, near, pos, fwd, right, up):
    cam = np.zeros(1, CAMERA)
    lib().orc_camera_box(C.c_float(left), C.c_float(right_), C.c_float(bottom), C.c_float(top), C.c_float(near),
                         _p(f3(pos)), _p(f3(fwd)), _p(f3(right)), _p(f3(up)), _p(cam))
    return cam


def camera_ray_world(cam, u, v):
    r = np.zeros(1, RAY)
    lib().orc_camera_ray_world(_p(cam), C.c_float(u), C.c_float(v), _p(r))
    return r


# ------------------------------------------------------------------ scene
class Scene:
    """SceneBuilder::build (scene.rs:86-96): objects = (blas index, transform matrix) pairs."""

    def __init__(self, blases, objects, with_transform=True):
        self.blases = list(blases)
        self.objects = [(int(b), np.asarray(m, np.float32).reshape(16).copy()) for b, m in objects]
        self.with_transform = with_transform
        self.rebuild()

    def set_transform(self, i, m):
        self.objects[i] = (self.objects[i][0], np.asarray(m, np.float32).reshape(16).copy())

    def rebuild(self):
        n = len(self.objects)
        self.inst = np.zeros(n, INSTANCE)
        bounds = np.zeros(n, AABB)
        for i, (b, m) in enumerate(self.objects):
            self.inst["inv"][i] = mat4_inverse(m)
            self.inst["blas_id"][i] = b
            if self.with_transform:
                bounds[i] = instance_bounds(m, self.blases[b].bounds())[0]
            else:  # SceneObjectBuilder::new without with_transform: bounds stay new_empty (scene_object.rs:101-107)
                bounds["min"][i] = FLT_MAX
                bounds["max"][i] = -FLT_MAX
        self.bounds = bounds
        self.tlas, self.tlas_used = tlas_build(bounds)
        self._blas_c = (_Blas * len(self.blases))(*[b._c() for b in self.blases])
        self._scene_c = _Scene(self.tlas.ctypes.data, self.tlas_used, self.inst.ctypes.data, n,
                               C.cast(self._blas_c, C.c_void_p), len(self.blases))

    def refresh_blas(self):
        self._blas_c = (_Blas * len(self.blases))(*[b._c() for b in self.blases])
        self._scene_c.blas = C.cast(self._blas_c, C.c_void_p)

    def intersect(self, ray, counters=None):
        hit = np.zeros(1, HIT)
        ok = lib().orc_scene_intersect(C.byref(self._scene_c), _p(ray), _p(hit),
                                       C.byref(counters) if counters is not None else None)
        return hit[0] if ok else None

    def render(self, cam, width, height, tile=8, region=None, threads=1, counters=None, out=None):
        hits = out if out is not None else np.zeros(width * height, HIT)
        if out is None:
            hits["t"] = FLT_MAX
            hits["id"] = MISS_ID
        x0, y0, x1, y1 = region if region is not None else (0, 0, width, height)
        lib().orc_render(C.byref(self._scene_c), _p(cam), C.c_uint32(width), C.c_uint32(height), C.c_uint32(tile),
                         C.c_uint32(x0), C.c_uint32(y0), C.c_uint32(x1), C.c_uint32(y1), _p(hits),
                         C.byref(counters) if counters is not None else None, C.c_int(threads))
        return hits

    def trace_rays(self, odt, threads=1, counters=None):
        odt = np.ascontiguousarray(np.asarray(odt, np.float32).reshape(-1, 7))
        hits = np.zeros(odt.shape[0], HIT)
        lib().orc_trace_rays(C.byref(self._scene_c), _p(odt), C.c_uint64(odt.shape[0]), _p(hits),
                             C.byref(counters) if counters is not None else None, C.c_int(threads))
        return hits


def max_threads():
    return int(lib().orc_max_threads())


def shade(kind, hits, scale=80.0, offset=3.0, hit_rgba=0xFFFFFFFF, miss_rgba=0xFF000000):
    """Accumulator + PixelShader pair applied to hit records -> u32 Rgba<u8> pixels (renderer.rs:116-245)."""
    hits = np.ascontiguousarray(hits)
    out = np.zeros(hits.size, "<u4")
    lib().orc_shade(C.c_uint32(kind), C.c_float(scale), C.c_float(offset), C.c_uint32(hit_rgba), C.c_uint32(miss_rgba),
                    _p(hits), C.c_uint64(hits.size), _p(out))
    return out


def tri_normals(tris):
    """TriMeshDecoder normals (mesh/decoders.rs:120-124) for file-order triangles -> n x 9."""
    tris = np.ascontiguousarray(np.asarray(tris, np.float32).reshape(-1, 9))
    out = np.zeros_like(tris)
    lib().orc_tri_normals(_p(tris), C.c_uint64(tris.shape[0]), _p(out))
    return out


def parse_obj_normals(text):
    b = text.encode() if isinstance(text, str) else text
    p = C.POINTER(C.c_float)()
    lib().orc_parse_obj_normals.restype = C.c_int64
    n = lib().orc_parse_obj_normals(b, C.c_size_t(len(b)), C.byref(p))
    if n < 0:
        raise ValueError(f"orc_parse_obj_normals failed: {n}")
    arr = np.ctypeslib.as_array(p, shape=(max(n, 1), 9))[:n].copy() if n else np.zeros((0, 9), np.float32)
    lib().orc_free(p)
    return arr


def load_asset_normals(name):
    """Packed per-vertex normals (assets/<name>.normals.f32, file order)."""
    return np.fromfile(os.path.join(ASSET_DIR, name + ".normals.f32"), dtype="<f4").reshape(-1, 9).copy()


def shade_normal(normals, object0_transform, hits):
    """NormalMappingAccumulator + RadianceToRgbShader (renderer.rs:256-286, 124-132)."""
    normals = np.ascontiguousarray(np.asarray(normals, np.float32).reshape(-1, 9))
    m = np.ascontiguousarray(np.asarray(object0_transform, np.float32).reshape(16))
    hits = np.ascontiguousarray(hits)
    out = np.zeros(hits.size, "<u4")
    lib().orc_shade_normal(_p(normals), C.c_uint64(normals.shape[0]), _p(m), _p(hits), C.c_uint64(hits.size), _p(out))
    return out


def shade_texture(tex_coords, texels, hits):
    """TextureMaterialAccumulator + RadianceToRgbShader (renderer.rs:289-334, 124-132); texels[height, width, 3] u8."""
    tex_coords = np.ascontiguousarray(np.asarray(tex_coords, np.float32).reshape(-1, 6))
    texels = np.ascontiguousarray(np.asarray(texels, np.uint8))
    hits = np.ascontiguousarray(hits)
    out = np.zeros(hits.size, "<u4")
    lib().orc_shade_texture(_p(tex_coords), C.c_uint64(tex_coords.shape[0]), _p(texels), C.c_uint32(texels.shape[1]),
                            C.c_uint32(texels.shape[0]), _p(hits), C.c_uint64(hits.size), _p(out))
    return out
