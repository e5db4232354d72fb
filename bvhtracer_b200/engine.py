"""Engine: a thin object wrapper over one bvht_ctx (include/bvht.h).  All work happens in libbvht_cuda.so."""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import BVH_NODE, CAMERA, HIT, INSTANCE, RAY, TLAS_NODE, BvhtError, Rect, ShadeParams, Stats, ptr


class Engine:
    def __init__(self, device=0, flags=_ffi.FLAG_STRICT):
        self._lib = _ffi.load()
        self._ctx = C.c_void_p()
        rc = self._lib.bvht_create(int(device), int(flags), C.byref(self._ctx))
        if rc != _ffi.OK:
            raise BvhtError(rc, self._lib.bvht_status_string(rc).decode())
        self.flags = int(flags)
        self.device = int(device)

    @classmethod
    def from_handle(cls, ctx_handle, flags=0, device=0):
        """Non-owning view of a bvht_ctx created elsewhere (e.g. by the C++ CudaPathTracer)."""
        self = cls.__new__(cls)
        self._lib = _ffi.load()
        self._ctx = C.c_void_p(ctx_handle)
        self._borrowed = True
        self.flags, self.device = int(flags), int(device)
        return self

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            if not getattr(self, "_borrowed", False):
                self._lib.bvht_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != _ffi.OK:
            raise BvhtError(rc, self._lib.bvht_last_error(self._ctx).decode(errors="replace"))

    def set_stream(self, cuda_stream_handle):
        self._check(self._lib.bvht_set_stream(self._ctx, C.c_void_p(cuda_stream_handle or 0)))

    def set_shard(self, index, count):
        """Only tile rows r with r % count == index are traced by the *_device entry points (multi-GPU)."""
        self._check(self._lib.bvht_set_shard(self._ctx, int(index), int(count)))

    def set_option(self, option, value):
        """bvht_set_option: scheduling overrides (coverage raster, K0, bands); value < 0 = the library's own rule."""
        self._check(self._lib.bvht_set_option(self._ctx, int(option), int(value)))

    def sync(self):
        self._check(self._lib.bvht_sync(self._ctx))

    def stats(self):
        s = Stats()
        self._check(self._lib.bvht_get_stats(self._ctx, C.byref(s)))
        return s.as_dict()

    # ------------------------------------------------------------------ scene state
    def blas_create(self, tris, nodes, nodes_used=None):
        tris = np.ascontiguousarray(np.asarray(tris, dtype="<f4").reshape(-1, 9))
        nodes = np.ascontiguousarray(nodes)
        assert nodes.dtype.itemsize == 32
        if nodes_used is None:
            nodes_used = len(nodes)
        out = C.c_uint32()
        self._check(self._lib.bvht_blas_create(self._ctx, ptr(tris), tris.shape[0], ptr(nodes), int(nodes_used), C.byref(out)))
        return int(out.value)

    def blas_destroy(self, blas_id):
        self._check(self._lib.bvht_blas_destroy(self._ctx, int(blas_id)))

    def blas_build(self, tris):
        """BvhBuilder::build_for on the device: tris in the mesh's own order -> blas id (tree + reordering as the reference's)."""
        tris = np.ascontiguousarray(np.asarray(tris, dtype="<f4").reshape(-1, 9))
        out = C.c_uint32()
        self._check(self._lib.bvht_blas_build(self._ctx, ptr(tris), tris.shape[0], C.byref(out)))
        return int(out.value)

    def blas_rebuild(self, blas_id):
        """Rebuild from the current vertices (the alternative to blas_refit)."""
        self._check(self._lib.bvht_blas_rebuild(self._ctx, int(blas_id)))

    def blas_info(self, blas_id):
        """-> (n_tris, nodes_used)"""
        n, u = C.c_uint32(), C.c_uint32()
        self._check(self._lib.bvht_blas_info(self._ctx, int(blas_id), C.byref(n), C.byref(u)))
        return int(n.value), int(u.value)

    def blas_read_triangles(self, blas_id, n_tris):
        out = np.zeros((int(n_tris), 9), "<f4")
        self._check(self._lib.bvht_blas_read_triangles(self._ctx, int(blas_id), ptr(out), int(n_tris)))
        return out

    def blas_read_permutation(self, blas_id, n_tris):
        out = np.zeros(int(n_tris), "<u4")
        self._check(self._lib.bvht_blas_read_permutation(self._ctx, int(blas_id), ptr(out), int(n_tris)))
        return out

    def blas_set_normals(self, blas_id, normals):
        normals = np.ascontiguousarray(np.asarray(normals, dtype="<f4").reshape(-1, 9))
        self._check(self._lib.bvht_blas_set_normals(self._ctx, int(blas_id), ptr(normals), normals.shape[0]))

    def blas_set_tex_coords(self, blas_id, tex_coords):
        """Mesh::tex_coords() (mesh.rs:146-154): n_tris x 6 f32, original primitive order."""
        tex_coords = np.ascontiguousarray(np.asarray(tex_coords, "<f4").reshape(-1, 6))
        self._check(self._lib.bvht_blas_set_tex_coords(self._ctx, int(blas_id), ptr(tex_coords), tex_coords.shape[0]))

    def blas_set_texture(self, blas_id, texels):
        """TextureMaterial<Rgb<u8>> (materials/material.rs:14-53): texels[height, width, 3] u8, already decoded."""
        texels = np.ascontiguousarray(np.asarray(texels, np.uint8))
        assert texels.ndim == 3 and texels.shape[2] == 3
        self._check(self._lib.bvht_blas_set_texture(self._ctx, int(blas_id), ptr(texels), texels.shape[1], texels.shape[0]))

    def blas_update_vertices(self, blas_id, tris):
        tris = np.ascontiguousarray(np.asarray(tris, dtype="<f4").reshape(-1, 9))
        self._check(self._lib.bvht_blas_update_vertices(self._ctx, int(blas_id), ptr(tris), tris.shape[0]))

    def blas_refit(self, blas_id):
        self._check(self._lib.bvht_blas_refit(self._ctx, int(blas_id)))

    def blas_read_nodes(self, blas_id, n):
        out = np.zeros(int(n), BVH_NODE)
        self._check(self._lib.bvht_blas_read_nodes(self._ctx, int(blas_id), ptr(out), int(n)))
        return out

    def tlas_set(self, nodes, nodes_used, instances):
        nodes = np.ascontiguousarray(nodes)
        instances = np.ascontiguousarray(instances)
        assert nodes.dtype.itemsize == 32 and instances.dtype.itemsize == 68
        self._check(self._lib.bvht_tlas_set(self._ctx, ptr(nodes), int(nodes_used), ptr(instances), len(instances)))

    def scene_set_transforms(self, transforms, blas_ids):
        """SceneObject::set_transform for every object + Tlas::rebuild on the device (scene_object.rs:60-75, tlas.rs:204-250).
        transforms: n x 16 f32 column-major forward matrices; blas_ids: n model ids."""
        transforms = np.ascontiguousarray(np.asarray(transforms, "<f4").reshape(-1, 16))
        blas_ids = np.ascontiguousarray(np.asarray(blas_ids, "<u4").reshape(-1))
        assert transforms.shape[0] == blas_ids.shape[0]
        self._check(self._lib.bvht_scene_set_transforms(self._ctx, ptr(transforms), ptr(blas_ids), transforms.shape[0]))

    def tlas_read(self, with_bounds=False):
        """Current TLAS nodes (reference layout), instances and, after scene_set_transforms, the objects' world bounds."""
        used, n = C.c_uint32(0), C.c_uint32(0)
        self._check(self._lib.bvht_tlas_read(self._ctx, None, 0, C.byref(used), None, None, 0, C.byref(n)))
        nodes = np.zeros(used.value, TLAS_NODE)
        inst = np.zeros(n.value, INSTANCE)
        bounds = np.zeros((n.value, 6), "<f4") if with_bounds else None
        self._check(self._lib.bvht_tlas_read(self._ctx, ptr(nodes), used.value, None, ptr(inst),
                                             ptr(bounds) if with_bounds else None, n.value, None))
        return (nodes, inst, bounds) if with_bounds else (nodes, inst)

    # ------------------------------------------------------------------ tracing
    def trace_primary(self, camera, width, height, tile=8, region=None, out=None):
        camera = np.ascontiguousarray(camera)
        assert camera.dtype.itemsize == 100
        if out is None:
            out = np.zeros(width * height, HIT)
            out["t"] = _ffi.FLT_MAX
            out["id"] = _ffi.MISS_ID
        x0, y0, x1, y1 = region if region is not None else (0, 0, width, height)
        self._check(self._lib.bvht_trace_primary(self._ctx, ptr(camera), int(width), int(height), int(tile),
                                                 Rect(x0, y0, x1, y1), ptr(out)))
        return out

    def trace_primary_device(self, camera, width, height, tile, region, out_device_ptr):
        camera = np.ascontiguousarray(camera)
        x0, y0, x1, y1 = region if region is not None else (0, 0, width, height)
        self._check(self._lib.bvht_trace_primary_device(self._ctx, ptr(camera), int(width), int(height), int(tile),
                                                        Rect(x0, y0, x1, y1), C.c_void_p(out_device_ptr)))

    @staticmethod
    def shade_depth(scale=80.0, offset=3.0):
        """DepthAccumulator + DepthMappingShader::new(scale, offset) (two/sixteen_armadillos.rs main)."""
        return ShadeParams(_ffi.SHADE_DEPTH, scale, offset, (C.c_uint8 * 4)(0, 0, 0, 0), (C.c_uint8 * 4)(0, 0, 0, 0))

    @staticmethod
    def shade_normal(object0_transform):
        """NormalMappingAccumulator + RadianceToRgbShader (cube.rs / trippy_teapots.rs main); needs blas_set_normals."""
        m = np.asarray(object0_transform, "<f4").reshape(16)
        return ShadeParams(_ffi.SHADE_NORMAL, 0.0, 0.0, (C.c_uint8 * 4)(0, 0, 0, 0), (C.c_uint8 * 4)(0, 0, 0, 0), (C.c_float * 16)(*m))

    @staticmethod
    def shade_texture():
        """TextureMaterialAccumulator + RadianceToRgbShader (quad.rs main); needs blas_set_tex_coords + blas_set_texture."""
        return ShadeParams(_ffi.SHADE_TEXTURE, 0.0, 0.0, (C.c_uint8 * 4)(0, 0, 0, 0), (C.c_uint8 * 4)(0, 0, 0, 0))

    @staticmethod
    def shade_intersection(hit=(255, 255, 255, 255), miss=(0, 0, 0, 255)):
        """IntersectionAccumulator + IntersectionShader::new(hit, miss) (big_ben_clock.rs main)."""
        return ShadeParams(_ffi.SHADE_INTERSECTION, 0.0, 0.0, (C.c_uint8 * 4)(*hit), (C.c_uint8 * 4)(*miss))

    @staticmethod
    def shade_uv():
        """UvMappingAccumulator + RadianceToRgbShader."""
        return ShadeParams(_ffi.SHADE_UV, 0.0, 0.0, (C.c_uint8 * 4)(0, 0, 0, 0), (C.c_uint8 * 4)(0, 0, 0, 0))

    def render_frame(self, camera, width, height, shade, tile=8, region=None, frame_out=None, hits_out=None, want_hits=False):
        """Integrator::evaluate: -> (frame u32[w*h] Rgba<u8>, hits or None); host buffers (pinned ones overlap copies)."""
        camera = np.ascontiguousarray(camera)
        if frame_out is None and shade is not None:
            frame_out = np.zeros(width * height, "<u4")
        if hits_out is None and want_hits:
            hits_out = np.zeros(width * height, HIT)
        x0, y0, x1, y1 = region if region is not None else (0, 0, width, height)
        self._check(self._lib.bvht_render_frame(
            self._ctx, ptr(camera), int(width), int(height), int(tile), Rect(x0, y0, x1, y1),
            C.byref(shade) if shade is not None else None,
            ptr(frame_out) if frame_out is not None else None, ptr(hits_out) if hits_out is not None else None))
        return frame_out, hits_out

    def render_frame_begin(self, camera, width, height, shade, frame_out, hits_out=None, tile=8, region=None):
        """render_frame without the wait (at most two frames in flight); frame_out / hits_out: page-locked arrays the caller keeps
        alive and untouched until the matching render_frame_end()."""
        camera = np.ascontiguousarray(camera)
        x0, y0, x1, y1 = region if region is not None else (0, 0, width, height)
        self._check(self._lib.bvht_render_frame_begin(
            self._ctx, ptr(camera), int(width), int(height), int(tile), Rect(x0, y0, x1, y1),
            C.byref(shade) if shade is not None else None,
            ptr(frame_out) if frame_out is not None else None, ptr(hits_out) if hits_out is not None else None))

    def render_frame_end(self):
        """Wait for the oldest begun frame."""
        self._check(self._lib.bvht_render_frame_end(self._ctx))

    def render_frame_device(self, camera, width, height, shade, tile, region, frame_dptr, hits_dptr):
        camera = np.ascontiguousarray(camera)
        x0, y0, x1, y1 = region if region is not None else (0, 0, width, height)
        self._check(self._lib.bvht_render_frame_device(
            self._ctx, ptr(camera), int(width), int(height), int(tile), Rect(x0, y0, x1, y1),
            C.byref(shade) if shade is not None else None, C.c_void_p(frame_dptr or 0), C.c_void_p(hits_dptr or 0)))

    def trace_rays(self, rays):
        rays = np.ascontiguousarray(rays)
        if rays.dtype != RAY:
            rays = np.ascontiguousarray(np.asarray(rays, "<f4").reshape(-1, 7)).view(RAY).reshape(-1)
        out = np.zeros(len(rays), HIT)
        self._check(self._lib.bvht_trace_rays(self._ctx, ptr(rays), len(rays), ptr(out)))
        return out

    def trace_rays_device(self, rays_device_ptr, n, out_device_ptr):
        self._check(self._lib.bvht_trace_rays_device(self._ctx, C.c_void_p(rays_device_ptr), int(n), C.c_void_p(out_device_ptr)))

    COUNTER_NAMES = ["rays_traced", "tlas_pairs", "instance_entries", "blas_pairs", "ref_leaves", "brute_tris", "sub_pairs", "sub_tris",
                     "accel_fallbacks", "hits", "mt_finishes", "tlas_chain_heads", "rays_in_empty_blocks", "blocks_pulled", "skip_tables",
                     "rays"]

    def debug_read_bandwidth(self, nbytes, passes=20):
        """GB/s of the library's streaming read kernel over an nbytes buffer (fits in L2 -> L2 bandwidth; >> L2 -> HBM)."""
        g = C.c_double()
        self._check(self._lib.bvht_debug_read_bandwidth(self._ctx, int(nbytes), int(passes), C.byref(g)))
        return float(g.value)

    def debug_frame_timeline(self):
        """Device timeline of the last render_frame (ms since its first device operation): dict with the coverage raster's end,
        the frame's end and, per band in launch order, (kernel end, copy end, tile rows)."""
        out = np.zeros(2 + 3 * 16, "<f4")
        n = C.c_uint32()
        self._check(self._lib.bvht_debug_frame_timeline(self._ctx, ptr(out), out.size, C.byref(n)))
        bands = [(float(out[2 + 3 * i]), float(out[3 + 3 * i]), int(out[4 + 3 * i])) for i in range((n.value - 2) // 3)]
        return {"cover_done_ms": float(out[0]), "frame_done_ms": float(out[1]), "bands": bands}

    def debug_trace_stats(self, camera, width, height, tile=8, region=None):
        """Per-frame work counters of the instrumented strict build, launched exactly like render_frame_device (coverage raster,
        K0, kernel flavour).  "rays" = rays of the region; "rays_traced" = those K1 generated a ray for.  Not a product path."""
        camera = np.ascontiguousarray(camera)
        out = np.zeros(16, "<u8")
        x0, y0, x1, y1 = region if region is not None else (0, 0, width, height)
        self._check(self._lib.bvht_debug_trace_stats(self._ctx, ptr(camera), int(width), int(height), int(tile),
                                                     Rect(x0, y0, x1, y1), ptr(out)))
        return {n: int(out[i]) for i, n in enumerate(self.COUNTER_NAMES)}

    # ------------------------------------------------------------------ device memory
    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self._lib.bvht_device_alloc(self._ctx, int(nbytes), C.byref(p)))
        return int(p.value)

    def device_free(self, dptr):
        self._check(self._lib.bvht_device_free(self._ctx, C.c_void_p(dptr)))

    def pinned_array(self, shape, dtype):
        """numpy array over page-locked host memory (bvht_host_alloc); freed with free_pinned(array)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._check(self._lib.bvht_host_alloc(self._ctx, max(n, 16), C.byref(p)))
        buf = (C.c_uint8 * max(n, 16)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8, count=n).view(dtype).reshape(shape)
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[arr.ctypes.data] = p.value
        return arr

    def free_pinned(self, arr):
        p = self._pinned.pop(arr.ctypes.data)
        self._check(self._lib.bvht_host_free(self._ctx, C.c_void_p(p)))

    def host_register(self, arr):
        """Page-lock an existing numpy buffer (e.g. over multiprocessing.shared_memory)."""
        self._check(self._lib.bvht_host_register(self._ctx, ptr(arr), arr.nbytes))

    def host_unregister(self, arr):
        self._check(self._lib.bvht_host_unregister(self._ctx, ptr(arr)))

    def memcpy_h2d(self, dptr, host_array):
        host_array = np.ascontiguousarray(host_array)
        self._check(self._lib.bvht_memcpy_h2d(self._ctx, C.c_void_p(dptr), ptr(host_array), host_array.nbytes))

    def memcpy_d2h(self, host_array, dptr, nbytes=None):
        assert host_array.flags["C_CONTIGUOUS"]
        self._check(self._lib.bvht_memcpy_d2h(self._ctx, ptr(host_array), C.c_void_p(dptr), int(nbytes or host_array.nbytes)))
        return host_array

    def ipc_export(self, dptr):
        h = (C.c_uint8 * 64)()
        self._check(self._lib.bvht_ipc_export(self._ctx, C.c_void_p(dptr), h))
        return bytes(h)

    def ipc_open(self, handle_bytes):
        h = (C.c_uint8 * 64).from_buffer_copy(handle_bytes)
        p = C.c_void_p()
        self._check(self._lib.bvht_ipc_open(self._ctx, h, C.byref(p)))
        return int(p.value)

    def ipc_close(self, dptr):
        self._check(self._lib.bvht_ipc_close(self._ctx, C.c_void_p(dptr)))
