/*
 * bvht.h -- C ABI of the B200-native closest-hit engine for bvhtracer's hot path.
 *
 * The reference (lambdaxymox/bvhtracer, pure Rust) has no FFI today; its plugin seam is the trait
 * `Integrator::evaluate(&mut self, &mut RendererState, &Scene) -> usize` (bvhtracer/src/renderer.rs:104-106)
 * plus `Scene::intersect(&self, &Ray<f32>) -> Option<Intersection<f32>>` (bvhtracer/src/scene/scene.rs:32-34).
 * This header is what a `CudaPathTracer: Integrator` inside the Rust crate binds with `extern "C"`
 * (INTEGRATION.md shows the binding).  Every entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all structs are POD with the layouts documented here
 *   - return value: BVHT_OK (0) or a negative bvht_status; bvht_last_error(ctx) gives the text.
 *     Nothing aborts, throws or unwinds across this boundary.  There is NO CPU fallback: with no
 *     CUDA device bvht_create fails with BVHT_ERR_NO_DEVICE.
 *   - the caller owns host buffers; the library copies during the call and never retains host pointers
 *   - one ctx <-> one device <-> one host thread at a time.  Calls are synchronous at return unless the
 *     name ends in _async / _device (then ordered on the ctx stream; bvht_sync waits).
 *   - miss is encoded in-band: t = FLT_MAX, u = v = 0, id = 0xFFFFFFFF
 */
#ifndef BVHT_H
#define BVHT_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVHT_ABI_VERSION 5

#if defined(__GNUC__)
#define BVHT_API __attribute__((visibility("default")))
#else
#define BVHT_API
#endif

typedef enum {
    BVHT_OK = 0,
    BVHT_ERR_INVALID_ARG = -1,
    BVHT_ERR_NO_DEVICE = -2,
    BVHT_ERR_CUDA = -3,
    BVHT_ERR_OUT_OF_MEMORY = -4,
    BVHT_ERR_BAD_HANDLE = -5,
    BVHT_ERR_MALFORMED_BVH = -6,     /* child index out of range, cycle, stack bound exceeded, ... */
    BVHT_ERR_NOT_READY = -7          /* trace before tlas_set, refit before vertices, ... */
} bvht_status;

/* bvht_create flags */
#define BVHT_FLAG_STRICT       0x0u  /* default: kernels built with --fmad=false, IEEE div/sqrt, no FTZ:
                                        hit ids bit-exact, t/u/v bit-exact vs the reference CPU path */
#define BVHT_FLAG_FAST         0x1u  /* FMA-contracted kernels: >= 99.99 % identical ids, t within 1e-5 rel. */
#define BVHT_FLAG_LEAF_ACCEL   0x2u  /* conservative sub-BVH inside each oversized reference leaf:
                                        same results, far fewer triangle tests */
#define BVHT_FLAG_STAMP_INSTANCE 0x4u /* EXTENSION (off by default): put the instance index in id bits 31..20.
                                        The reference always reports instance 0 (bvh.rs:296, intersection.rs:64-66) */

typedef struct bvht_ctx bvht_ctx;

/* model/bvh.rs:88-134  BvhLeafNode / BvhBranchNode (both #[repr(C)]): 32 bytes.
 * prim_count > 0  => leaf, left_first = first primitive index
 * prim_count == 0 => branch, children at left_first and left_first + 1.  Node 1 is the unused dummy. */
typedef struct {
    float    aabb_min[3];
    float    aabb_max[3];
    uint32_t prim_count;
    uint32_t left_first;
} bvht_bvh_node;

/* scene/tlas.rs:11-46  TlasNode: 32 bytes.  left_right == 0 => leaf (blas = instance index).
 * left_right keeps the reference's packing (LeftRightIndex::new(l, r) = l + (r << 16), tlas.rs:18-22)
 * and the engine reads it back exactly like the reference accessors do (upper half first, tlas.rs:25-32). */
typedef struct {
    float    aabb_min[3];
    float    aabb_max[3];
    uint32_t left_right;
    uint32_t blas;
} bvht_tlas_node;

/* scene/scene_object.rs:16-21 + transform_component.rs:8-11: cached INVERSE transform, column-major
 * (cglinalg Matrix4x4 storage), and the BLAS this instance refers to. */
typedef struct {
    float    transform_inv[16];
    uint32_t blas_id;
} bvht_instance;

/* camera/camera.rs:199-216 (eye-space corner points of the near plane) and :871-873 (view_matrix_inv,
 * column-major).  Ray for (u, v): camera.rs:994-1010. */
typedef struct {
    float top_left_eye[3];
    float top_right_eye[3];
    float bottom_left_eye[3];
    float view_matrix_inv[16];
} bvht_camera;

/* query/ray.rs:9-17 without the cached reciprocal (Ray::new recomputes it, ray.rs:23-31). */
typedef struct {
    float origin[3];
    float direction[3];
    float t;                         /* initial closest distance (f32::MAX for Ray::from_origin_dir) */
} bvht_ray;

/* query/intersection.rs:10-17, 33-67, 77-92: the 16-byte record {t, u, v, InstancePrimitiveIndex}. */
typedef struct {
    float    t;
    float    u;
    float    v;
    uint32_t id;                     /* ((instance & 0xFFF) << 20) | (primitive & 0xFFFFF); miss: 0xFFFFFFFF */
} bvht_hit;

typedef struct { uint32_t x0, y0, x1, y1; } bvht_rect;   /* pixels [x0,x1) x [y0,y1) */

/* Accumulator + PixelShader pairs of the reference (renderer.rs:116-245), evaluated on the device right
 * after the closest hit is known, producing `Rgba<u8>` pixels (r, g, b, a bytes; 4 B/pixel). */
typedef enum {
    BVHT_SHADE_NONE = 0,
    BVHT_SHADE_DEPTH = 1,          /* DepthAccumulator (:184-194) + DepthMappingShader::new(scale, offset) (:207-222) */
    BVHT_SHADE_INTERSECTION = 2,   /* IntersectionAccumulator (:145-153) + IntersectionShader::new(hit, miss) (:166-174) */
    BVHT_SHADE_UV = 3,             /* UvMappingAccumulator (:233-245) + RadianceToRgbShader (:124-132) */
    BVHT_SHADE_NORMAL = 4,         /* NormalMappingAccumulator (:256-286) + RadianceToRgbShader; needs bvht_blas_set_normals for
                                      the model of scene object 0 and object0_transform (the reference's instance index is
                                      always 0, so it always looks up object 0's model and transform, :258-276) */
    BVHT_SHADE_TEXTURE = 5         /* TextureMaterialAccumulator (:289-334) + RadianceToRgbShader; needs bvht_blas_set_tex_coords
                                      and bvht_blas_set_texture for the model of scene object 0 (instance index always 0) */
} bvht_shade_kind;

typedef struct {
    uint32_t kind;                 /* bvht_shade_kind */
    float    depth_scale;          /* DepthMappingShader.scale  (80 in the armadillo examples) */
    float    depth_offset;         /* DepthMappingShader.offset ( 3 in the armadillo examples) */
    uint8_t  hit_rgba[4];          /* IntersectionShader.hit_value  */
    uint8_t  miss_rgba[4];         /* IntersectionShader.miss_value */
    float    object0_transform[16];/* BVHT_SHADE_NORMAL: forward transform of scene object 0, column-major */
} bvht_shade_params;

typedef struct {
    float    last_trace_ms;          /* device time of the last trace kernel (CUDA events on the ctx stream) */
    float    last_refit_ms;
    float    last_upload_ms;
    uint64_t last_trace_rays;
    uint64_t kernel_launches;        /* total launches of this library's own kernels since bvht_create */
    uint64_t h2d_bytes;              /* totals since bvht_create */
    uint64_t d2h_bytes;
    uint32_t sm_count;
    uint32_t trace_grid;             /* CTAs of the last persistent trace launch */
    uint32_t trace_block;
    uint32_t flags;
    float    last_build_ms;          /* bvht_blas_build / bvht_blas_rebuild: upload + device build + read-back */
    uint32_t last_build_levels;      /* levels of the level-synchronous device build */
    float    last_k1_ms;             /* the trace kernel (K1) alone of the last frame: what bench.py's roofline divides by */
    uint32_t rebakes;                /* how often the leaf accelerator's boxes were re-inflated for new ray limits */
    float    bake_d_max, bake_o_max; /* ray limits of the last bake (model space |d|, |o|) */
} bvht_stats;

/* ---------------------------------------------------------------------------------------------- */
BVHT_API int         bvht_abi_version(void);
BVHT_API int         bvht_device_count(void);

/* Create a context on `device`.  Replaces nothing in the reference (it has no device). */
BVHT_API int         bvht_create(int device, uint32_t flags, bvht_ctx** out);
BVHT_API void        bvht_destroy(bvht_ctx* ctx);
BVHT_API const char* bvht_last_error(const bvht_ctx* ctx);
BVHT_API const char* bvht_status_string(int status);

/* Use an existing CUDA stream (cudaStream_t as void*) for all work of this ctx; NULL = the ctx's own. */
BVHT_API int         bvht_set_stream(bvht_ctx* ctx, void* cuda_stream);
BVHT_API int         bvht_sync(bvht_ctx* ctx);

/* Execution options.  None of them changes a result (tests/ force each both ways and compare with the oracle); they only
 * override how a frame is scheduled on the device.  value < 0 restores the library's own rule.  No reference counterpart. */
typedef enum bvht_option {
    BVHT_OPT_COVER = 1,   /* per-triangle block coverage raster before the trace (K7): 0 off, 1 on */
    BVHT_OPT_K0    = 2,   /* classify-and-fill pass for empty pixel blocks (K0): 0 off, 1 on */
    BVHT_OPT_BANDS = 3,   /* number of tile-row bands bvht_render_frame pipelines its device->host copies in (1..32) */
    BVHT_OPT_COPY_STREAMS = 5, /* streams the band copies of bvht_render_frame rotate over (1..3) */
    BVHT_OPT_BAND_ORDER = 4, /* order in which the bands are traced and copied: 0 image order, 1 cheapest first, 2 cheap ascending
                             * then expensive descending, 3 descending cost */
    BVHT_OPT_TIMELINE = 6   /* 1: bvht_render_frame records a timing event behind every band copy for bvht_debug_frame_timeline
                             * (about 3 us per band on the copy streams; off by default) */
} bvht_option;
BVHT_API int         bvht_set_option(bvht_ctx* ctx, uint32_t option, int32_t value);

/* Upload one model: the BVH-reordered triangle buffer `Mesh::primitives()` (mesh.rs:126-134; n_tris x 9 f32,
 * 36-byte stride) and the used prefix of `Bvh.nodes` (bvh.rs:228-233).  Replaces what
 * ModelBuilder::build leaves in memory for Model::intersect (model.rs:66-68, 140-144).  On upload the
 * node pool is flattened to 2 x float4 per node and the triangles are repacked SoA (v0, e1, e2). */
BVHT_API int         bvht_blas_create(bvht_ctx* ctx, const float* tris, uint32_t n_tris,
                             const bvht_bvh_node* nodes, uint32_t nodes_used, uint32_t* out_blas_id);
BVHT_API int         bvht_blas_destroy(bvht_ctx* ctx, uint32_t blas_id);

/* `BvhBuilder::build_for(&mut mesh)` (bvh.rs:524-541, subdivide :396-467, find_best_split_plane :333-394) ON THE DEVICE:
 * `tris` is the mesh in ITS OWN order (n_tris x 9 f32); the binned-SAH tree, the in-place reordering of the triangles
 * and the node numbering come out bit-identical to the reference's host build (tests compare them with the oracle).
 * Read the results back with bvht_blas_read_nodes / bvht_blas_read_triangles -- the reference reorders `Mesh.vertices`
 * in place, so its host copy must be replaced by the reordered triangles. */
BVHT_API int         bvht_blas_build(bvht_ctx* ctx, const float* tris, uint32_t n_tris, uint32_t* out_blas_id);

/* Rebuild the tree of an existing model from its CURRENT vertices (after bvht_blas_update_vertices): the "rebuild" arm of
 * bvhtracer/benches/bench_bvh_refit_rebuild.rs, where bvht_blas_refit is the "refit" arm.  Triangles are reordered again. */
BVHT_API int         bvht_blas_rebuild(bvht_ctx* ctx, uint32_t blas_id);

/* The model's triangles in their current (BVH) order, n_tris x 9 f32; and, for device-built models, the permutation
 * that produced it: out[i] = index in the array given to bvht_blas_build of the triangle now at position i. */
BVHT_API int         bvht_blas_info(bvht_ctx* ctx, uint32_t blas_id, uint32_t* n_tris, uint32_t* nodes_used);   /* Bvh.nodes_used (bvh.rs:232) */
BVHT_API int         bvht_blas_read_triangles(bvht_ctx* ctx, uint32_t blas_id, float* out, uint32_t n_tris);
BVHT_API int         bvht_blas_read_permutation(bvht_ctx* ctx, uint32_t blas_id, uint32_t* out, uint32_t n_tris);

/* Per-vertex normals of a model, `Mesh::normals()` (mesh.rs:158-166): n_tris x 9 f32 in the mesh's ORIGINAL primitive
 * order -- the reference reorders only the positions when it builds the BVH (bvh.rs:426), and indexes this array with
 * the reordered primitive index (renderer.rs:259-266); that behaviour is kept. */
BVHT_API int         bvht_blas_set_normals(bvht_ctx* ctx, uint32_t blas_id, const float* normals, uint32_t n_tris);

/* Per-vertex texture coordinates, `Mesh::tex_coords()` (mesh.rs:146-154): n_tris x 6 f32 (three Vector2), ORIGINAL
 * primitive order like the normals (renderer.rs:311-316 indexes them with the reordered primitive index). */
BVHT_API int         bvht_blas_set_tex_coords(bvht_ctx* ctx, uint32_t blas_id, const float* tex_coords, uint32_t n_tris);

/* The model's `TextureMaterial<Rgb<u8>>` (materials/material.rs:14-53): width x height texels, 3 bytes each, row-major
 * with texel (x, y) at (y * width + x) * 3 (texture_buffer.rs:211), already decoded (PNG decoding stays on the host).
 * Sampling is the reference's nearest-texel rule: ((uv.x * width) as usize) % width, likewise for y (material.rs:44-52). */
BVHT_API int         bvht_blas_set_texture(bvht_ctx* ctx, uint32_t blas_id, const uint8_t* rgb, uint32_t width, uint32_t height);

/* New vertex positions for an existing model (examples/big_ben_clock.rs:67-96 `animate`); same count. */
BVHT_API int         bvht_blas_update_vertices(bvht_ctx* ctx, uint32_t blas_id, const float* tris, uint32_t n_tris);

/* Bvh::refit (bvh.rs:469-493) + update_node_bounds (bvh.rs:317-330), bottom-up on the device. */
BVHT_API int         bvht_blas_refit(bvht_ctx* ctx, uint32_t blas_id);

/* Read the device node pool back in reference layout (parity checks of refit). */
BVHT_API int         bvht_blas_read_nodes(bvht_ctx* ctx, uint32_t blas_id, bvht_bvh_node* out, uint32_t max_nodes);

/* Per-frame scene state: `Tlas.nodes[0..nodes_used]` after Tlas::rebuild (tlas.rs:204-250) and one
 * bvht_instance per SceneObject, in `Scene.objects` order (scene.rs:10-15). */
BVHT_API int         bvht_tlas_set(bvht_ctx* ctx, const bvht_tlas_node* nodes, uint32_t nodes_used,
                          const bvht_instance* instances, uint32_t n_instances);

/* The same per-frame scene state, COMPUTED ON THE DEVICE from the objects' forward transforms: for every scene object
 * `SceneObject::set_transform` (scene_object.rs:60-75: the world AABB of the 8 transformed corners of `Model::bounds()` --
 * the model's CURRENT root box on the device, so it follows bvht_blas_refit / bvht_blas_rebuild without a read-back -- and
 * the cached inverse, transform_component.rs:17-27), then `Tlas::rebuild` (tlas.rs:204-250, find_best_match :179-202) with
 * the reference's clustering order, child packing and node numbering (always 2 * n nodes).  `transforms`: n x 16 f32,
 * column-major `Transform3` matrices, in `Scene.objects` order; `blas_ids`: the model of each object.  Replaces the host
 * loop `for o in objects { o.set_transform(..) }; tlas.rebuild(objects)` + bvht_tlas_set; results are bit-identical to it.
 * A singular transform, or bounds for which the clustering finds no candidate, are errors (the reference panics). */
BVHT_API int         bvht_scene_set_transforms(bvht_ctx* ctx, const float* transforms, const uint32_t* blas_ids, uint32_t n_instances);

/* The current TLAS in reference layout (`Tlas.nodes[0..nodes_used]`), the instances (cached inverses) and -- after
 * bvht_scene_set_transforms -- `SceneObject::bounds()` of every object (6 f32 each: min, max).  Any output may be NULL;
 * call once with NULL buffers to get the counts. */
BVHT_API int         bvht_tlas_read(bvht_ctx* ctx, bvht_tlas_node* nodes_out, uint32_t max_nodes, uint32_t* nodes_used_out,
                            bvht_instance* instances_out, float* bounds_out, uint32_t max_instances, uint32_t* n_instances_out);

/* PathTracer::evaluate, first loop (renderer.rs:345-368): one primary ray per pixel of `region` of a
 * width x height image through Camera::get_ray_world, Scene::intersect for each, tiles of `tile` x `tile`
 * pixels (8 in the reference).  out_host: width*height records, row-major (pixel_address =
 * x + y * width, renderer.rs:362); only the region's pixels are written.  Returns when out_host is filled. */
BVHT_API int         bvht_trace_primary(bvht_ctx* ctx, const bvht_camera* camera, uint32_t width, uint32_t height,
                               uint32_t tile, bvht_rect region, bvht_hit* out_host);

/* Same, result left in device memory (a device pointer obtained from bvht_device_alloc, from another
 * library, or a peer-mapped pointer of another GPU: the kernel stores straight into it).  Asynchronous. */
BVHT_API int         bvht_trace_primary_device(bvht_ctx* ctx, const bvht_camera* camera, uint32_t width, uint32_t height,
                                      uint32_t tile, bvht_rect region, void* out_device);

/* Integrator::evaluate as a whole (renderer.rs:345-384): trace the region AND run the accumulator + pixel
 * shader pair on the device, returning the `FrameBuffer<Rgba<u8>>` pixels (frame_out_host: width*height
 * u32, row-major, NOT vertically flipped -- the flip for GL is done by the caller, bvhtracer_demos lib.rs:114)
 * and, when hits_out_host is not NULL, the 16-byte hit records as well.  The region is cut into bands of
 * tile rows; each band's device->host copy overlaps the tracing of the next band (give pinned host memory,
 * e.g. from bvht_host_alloc, for the overlap to be real).  Returns when the host buffers are filled. */
BVHT_API int         bvht_render_frame(bvht_ctx* ctx, const bvht_camera* camera, uint32_t width, uint32_t height,
                               uint32_t tile, bvht_rect region, const bvht_shade_params* shade,
                               uint32_t* frame_out_host, bvht_hit* hits_out_host);
/* The same frame with up to TWO frames in flight (no reference counterpart: the reference's Renderer::render returns with the
 * frame, renderer.rs:394-399; an application that presents frame n while frame n+1 is traced calls these instead).
 * bvht_render_frame_begin queues the frame and returns; bvht_render_frame_end waits for the OLDEST begun frame, whose host
 * buffers (page-locked, and not to be touched in between) are then filled.  The device->host copies of frame n run under the
 * kernels of frame n+1: two device staging frames alternate, each band's copy is issued by the host when the band's flag comes
 * up in page-locked memory (from _begin, _end and every wait the library does in between: call one of them to keep copies
 * moving), and nothing makes the next frame's kernels wait for them.  Scene updates between
 * two begins (bvht_tlas_set, bvht_scene_set_transforms, bvht_blas_update_vertices / _refit) are ordered behind the kernels of
 * the frame already begun.  A third begin without an end is BVHT_ERR_NOT_READY, as is an end with nothing in flight;
 * bvht_render_frame and bvht_sync complete whatever is in flight first.  Results are those of bvht_render_frame, bit for bit. */
BVHT_API int         bvht_render_frame_begin(bvht_ctx* ctx, const bvht_camera* camera, uint32_t width, uint32_t height,
                               uint32_t tile, bvht_rect region, const bvht_shade_params* shade,
                               uint32_t* frame_out_host, bvht_hit* hits_out_host);
BVHT_API int         bvht_render_frame_end(bvht_ctx* ctx);
/* Same, everything left in device memory (either output may be NULL).  Asynchronous, single launch. */
BVHT_API int         bvht_render_frame_device(bvht_ctx* ctx, const bvht_camera* camera, uint32_t width, uint32_t height,
                                      uint32_t tile, bvht_rect region, const bvht_shade_params* shade,
                                      void* frame_out_device, void* hits_out_device);

/* Multi-GPU sharding (scene replicated per GPU, image tiles sharded): after bvht_set_shard(ctx, i, n) the
 * trace / render entry points only trace the tile rows r with r % n == i (interleaved for load balance) and
 * leave every other pixel untouched, so n contexts on n GPUs writing into ONE frame buffer (rank 0's, mapped
 * on the peers with bvht_ipc_open) assemble the frame over NVLink with no copy; the host-output calls copy
 * back (and write) only the owned rows, so n ranks can likewise fill ONE shared pinned host frame buffer,
 * each over its own PCIe link.  (0, 1) restores the default. */
BVHT_API int         bvht_set_shard(bvht_ctx* ctx, uint32_t shard_index, uint32_t shard_count);
/* Pure host helper (no device needed): which tile rows of `region` shard i of n owns -- rows
 * first_row, first_row + n, ... (n_rows of them).  The same function the launches use. */
BVHT_API int         bvht_shard_tile_rows(bvht_rect region, uint32_t tile, uint32_t shard_index, uint32_t shard_count,
                                  uint32_t* first_row, uint32_t* n_rows);

/* Scene::intersect(&Ray) for n arbitrary rays (scene.rs:32-34). */
BVHT_API int         bvht_trace_rays(bvht_ctx* ctx, const bvht_ray* rays, uint64_t n, bvht_hit* out_host);
BVHT_API int         bvht_trace_rays_device(bvht_ctx* ctx, const void* rays_device, uint64_t n, void* out_device);

/* Device memory helpers (multi-GPU gather, resident benchmarks). */
BVHT_API int         bvht_device_alloc(bvht_ctx* ctx, size_t bytes, void** out_device);
BVHT_API int         bvht_device_free(bvht_ctx* ctx, void* device_ptr);
/* Page-locked host memory (so that copies overlap with kernels).  ctx may be NULL for both calls. */
BVHT_API int         bvht_host_alloc(bvht_ctx* ctx, size_t bytes, void** out_host);
BVHT_API int         bvht_host_free(bvht_ctx* ctx, void* host_ptr);
/* Page-lock an existing host range (e.g. a POSIX shared-memory frame buffer that several ranks fill). */
BVHT_API int         bvht_host_register(bvht_ctx* ctx, void* host_ptr, size_t bytes);
BVHT_API int         bvht_host_unregister(bvht_ctx* ctx, void* host_ptr);
BVHT_API int         bvht_memcpy_h2d(bvht_ctx* ctx, void* dst_device, const void* src_host, size_t bytes);
BVHT_API int         bvht_memcpy_d2h(bvht_ctx* ctx, void* dst_host, const void* src_device, size_t bytes);
/* CUDA IPC: export a device allocation so another process (one rank per GPU) can map it over NVLink P2P */
BVHT_API int         bvht_ipc_export(bvht_ctx* ctx, void* device_ptr, uint8_t handle_out[64]);
BVHT_API int         bvht_ipc_open(bvht_ctx* ctx, const uint8_t handle[64], void** out_device);
BVHT_API int         bvht_ipc_close(bvht_ctx* ctx, void* device_ptr);

BVHT_API int         bvht_get_stats(const bvht_ctx* ctx, bvht_stats* out);

/* Debug: trace `region` once with an instrumented strict kernel and return per-frame work counters:
 * [0] rays, [1] TLAS pair tests, [2] instance entries, [3] reference BLAS pair tests, [4] reference leaves
 * visited, [5] brute-force triangle tests, [6] sub-BVH pair tests, [7] sub-BVH triangle tests,
 * [8] accel fallbacks, [9] hits, [10] Moeller-Trumbore evaluations that went past the filter, [11] TLAS chains resolved by the
 * skip table, [12] rays of blocks K1 found empty (no ray generated), [13] pixel blocks K1 pulled, [14] skip tables built,
 * [15] rays of the region (this shard's).  The frame is launched exactly as bvht_render_frame_device launches it (coverage
 * raster, K0, kernel flavour).  Not a product path. */
/* Measurement helper: read bandwidth (GB/s) of `passes` sweeps over a `bytes`-sized device buffer with this library's own
 * streaming kernel -- L2 -> SM delivery when the buffer fits in L2 (e.g. 32 MiB), HBM when it is much larger (e.g. 2 GiB).
 * bench.py records it next to MEASURED_PEAKS.json's HBM figure. */
BVHT_API int         bvht_debug_read_bandwidth(bvht_ctx* ctx, size_t bytes, uint32_t passes, double* gbs_out);
/* Debug: device timeline of the last bvht_render_frame made with BVHT_OPT_TIMELINE on (BVHT_ERR_NOT_READY otherwise), milliseconds
 * since its first device operation: [0] coverage raster done, [1] frame complete (last copy), then per band in the order the
 * copies were issued (the first 16): all kernels done, copy done, tile rows of the band. */
BVHT_API int         bvht_debug_frame_timeline(bvht_ctx* ctx, float* ms_out, uint32_t capacity, uint32_t* n_out);
BVHT_API int         bvht_debug_trace_stats(bvht_ctx* ctx, const bvht_camera* camera, uint32_t width, uint32_t height,
                                    uint32_t tile, bvht_rect region, uint64_t counters_out[16]);

#ifdef __cplusplus
}
#endif
#endif /* BVHT_H */
