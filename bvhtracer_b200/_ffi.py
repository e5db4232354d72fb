"""ctypes binding of include/bvht.h (libbvht_cuda.so).  Thin plumbing: no algorithm lives here.

The library is the product; if it is missing or cannot be loaded this module raises -- there is no
CPU fallback anywhere in the package.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libbvht_cuda.so")

# status codes / flags (include/bvht.h)
OK = 0
ERR_INVALID_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_OOM, ERR_BAD_HANDLE, ERR_MALFORMED_BVH, ERR_NOT_READY = -1, -2, -3, -4, -5, -6, -7
FLAG_STRICT, FLAG_FAST, FLAG_LEAF_ACCEL, FLAG_STAMP_INSTANCE = 0x0, 0x1, 0x2, 0x4

FLT_MAX = np.finfo(np.float32).max
MISS_ID = 0xFFFFFFFF

# POD layouts (include/bvht.h)
BVH_NODE = np.dtype([("aabb_min", "<f4", 3), ("aabb_max", "<f4", 3), ("prim_count", "<u4"), ("left_first", "<u4")])
TLAS_NODE = np.dtype([("aabb_min", "<f4", 3), ("aabb_max", "<f4", 3), ("left_right", "<u4"), ("blas", "<u4")])
INSTANCE = np.dtype([("transform_inv", "<f4", 16), ("blas_id", "<u4")])
CAMERA = np.dtype([("top_left_eye", "<f4", 3), ("top_right_eye", "<f4", 3), ("bottom_left_eye", "<f4", 3),
                   ("view_matrix_inv", "<f4", 16)])
RAY = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("t", "<f4")])
HIT = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("id", "<u4")])
assert BVH_NODE.itemsize == 32 and TLAS_NODE.itemsize == 32 and INSTANCE.itemsize == 68
assert CAMERA.itemsize == 100 and RAY.itemsize == 28 and HIT.itemsize == 16


class Rect(C.Structure):
    _fields_ = [("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32)]


class ShadeParams(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("depth_scale", C.c_float), ("depth_offset", C.c_float),
                ("hit_rgba", C.c_uint8 * 4), ("miss_rgba", C.c_uint8 * 4), ("object0_transform", C.c_float * 16)]


OPT_COVER, OPT_K0, OPT_BANDS, OPT_BAND_ORDER, OPT_COPY_STREAMS, OPT_TIMELINE = 1, 2, 3, 4, 5, 6

SHADE_NONE, SHADE_DEPTH, SHADE_INTERSECTION, SHADE_UV, SHADE_NORMAL, SHADE_TEXTURE = 0, 1, 2, 3, 4, 5


class Stats(C.Structure):
    _fields_ = [("last_trace_ms", C.c_float), ("last_refit_ms", C.c_float), ("last_upload_ms", C.c_float),
                ("last_trace_rays", C.c_uint64), ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("sm_count", C.c_uint32), ("trace_grid", C.c_uint32),
                ("trace_block", C.c_uint32), ("flags", C.c_uint32), ("last_build_ms", C.c_float), ("last_build_levels", C.c_uint32),
                ("last_k1_ms", C.c_float), ("rebakes", C.c_uint32), ("bake_d_max", C.c_float), ("bake_o_max", C.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/bvht.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("bvht_abi_version", C.c_int, []),
    ("bvht_device_count", C.c_int, []),
    ("bvht_create", C.c_int, [C.c_int, C.c_uint32, C.POINTER(_P)]),
    ("bvht_destroy", None, [_P]),
    ("bvht_last_error", C.c_char_p, [_P]),
    ("bvht_status_string", C.c_char_p, [C.c_int]),
    ("bvht_set_stream", C.c_int, [_P, _P]),
    ("bvht_sync", C.c_int, [_P]),
    ("bvht_set_option", C.c_int, [_P, C.c_uint32, C.c_int32]),
    ("bvht_blas_create", C.c_int, [_P, _P, C.c_uint32, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("bvht_blas_destroy", C.c_int, [_P, C.c_uint32]),
    ("bvht_blas_build", C.c_int, [_P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("bvht_blas_rebuild", C.c_int, [_P, C.c_uint32]),
    ("bvht_blas_info", C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("bvht_blas_read_triangles", C.c_int, [_P, C.c_uint32, _P, C.c_uint32]),
    ("bvht_blas_read_permutation", C.c_int, [_P, C.c_uint32, _P, C.c_uint32]),
    ("bvht_blas_set_normals", C.c_int, [_P, C.c_uint32, _P, C.c_uint32]),
    ("bvht_blas_set_tex_coords", C.c_int, [_P, C.c_uint32, _P, C.c_uint32]),
    ("bvht_blas_set_texture", C.c_int, [_P, C.c_uint32, _P, C.c_uint32, C.c_uint32]),
    ("bvht_blas_update_vertices", C.c_int, [_P, C.c_uint32, _P, C.c_uint32]),
    ("bvht_blas_refit", C.c_int, [_P, C.c_uint32]),
    ("bvht_blas_read_nodes", C.c_int, [_P, C.c_uint32, _P, C.c_uint32]),
    ("bvht_tlas_set", C.c_int, [_P, _P, C.c_uint32, _P, C.c_uint32]),
    ("bvht_scene_set_transforms", C.c_int, [_P, _P, _P, C.c_uint32]),
    ("bvht_tlas_read", C.c_int, [_P, _P, C.c_uint32, C.POINTER(C.c_uint32), _P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("bvht_trace_primary", C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, Rect, _P]),
    ("bvht_trace_primary_device", C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, Rect, _P]),
    ("bvht_render_frame", C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, Rect, C.POINTER(ShadeParams), _P, _P]),
    ("bvht_render_frame_begin", C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, Rect, C.POINTER(ShadeParams), _P, _P]),
    ("bvht_render_frame_end", C.c_int, [_P]),
    ("bvht_render_frame_device", C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, Rect, C.POINTER(ShadeParams), _P, _P]),
    ("bvht_set_shard", C.c_int, [_P, C.c_uint32, C.c_uint32]),
    ("bvht_shard_tile_rows", C.c_int, [Rect, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("bvht_trace_rays", C.c_int, [_P, _P, C.c_uint64, _P]),
    ("bvht_trace_rays_device", C.c_int, [_P, _P, C.c_uint64, _P]),
    ("bvht_device_alloc", C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    ("bvht_device_free", C.c_int, [_P, _P]),
    ("bvht_host_alloc", C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    ("bvht_host_free", C.c_int, [_P, _P]),
    ("bvht_host_register", C.c_int, [_P, _P, C.c_size_t]),
    ("bvht_host_unregister", C.c_int, [_P, _P]),
    ("bvht_memcpy_h2d", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("bvht_memcpy_d2h", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("bvht_ipc_export", C.c_int, [_P, _P, _P]),
    ("bvht_ipc_open", C.c_int, [_P, _P, C.POINTER(_P)]),
    ("bvht_ipc_close", C.c_int, [_P, _P]),
    ("bvht_get_stats", C.c_int, [_P, C.POINTER(Stats)]),
    ("bvht_debug_read_bandwidth", C.c_int, [_P, C.c_size_t, C.c_uint32, C.POINTER(C.c_double)]),
    ("bvht_debug_frame_timeline", C.c_int, [_P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("bvht_debug_trace_stats", C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, Rect, _P]),
]

_lib = None


class BvhtError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"bvht error {status}: {message}")
        self.status = status


def load():
    """dlopen libbvht_cuda.so and type every entry point.  Raises if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m bvhtracer_b200.build` "
                              "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(lib, name)        # AttributeError if a declared symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def shard_tile_rows(region, tile, index, count):
    """bvht_shard_tile_rows: -> list of the tile rows shard `index` of `count` owns inside `region` (x0, y0, x1, y1)."""
    first, n = C.c_uint32(), C.c_uint32()
    rc = load().bvht_shard_tile_rows(Rect(*region), int(tile), int(index), int(count), C.byref(first), C.byref(n))
    if rc != OK:
        raise BvhtError(rc, "bvht_shard_tile_rows")
    return [first.value + k * int(count) for k in range(n.value)]
