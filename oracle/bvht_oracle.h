/*
 * bvht_oracle.h -- CPU ORACLE for the bvhtracer primary closest-hit path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * (lambdaxymox/bvhtracer) CPU algorithm, used as the checker in tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under bvhtracer_b200/ may include, link or call it.
 *
 * PARITY STATUS: the reference is Rust and cannot be compiled in this image
 * (no cargo/rustc), so there is no oracle/_ref.  The restatement is pinned
 * against every known-answer test the reference holds for this path
 * (tests/test_oracle_kat.py lists them with file:line); arithmetic that lives
 * in the un-vendored dependency `cglinalg` (req 0.16.10, Cargo.lock not
 * committed) is restated from its published semantics:
 *   - normalize  = v / sqrt(dot(v,v))       PINNED  (test_tri_mesh.rs:57-59)
 *   - cross      = textbook                 PINNED  (same KAT + MT t KATs)
 *   - dot        = (x*x + y*y) + z*z        parity unpinned (order), KATs pass
 *   - Mat4*Vec4  = ((c0*x + c1*y) + c2*z) + c3*w     parity unpinned
 *   - Mat4 inverse = adjugate / det         parity unpinned (host-side input)
 * Multi-instance TLAS scenes have no reference test: parity unpinned beyond
 * the restated source for those.
 *
 * Build:  make -C oracle      (gcc -O2 -ffp-contract=off -fno-fast-math)
 */
#ifndef BVHT_ORACLE_H
#define BVHT_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float min[3]; float max[3]; } orc_aabb;

/* model/bvh.rs:88-134  BvhLeafNode/BvhBranchNode: aabb, primitive_count, left_node|first_primitive_index */
typedef struct {
    float    min[3];
    float    max[3];
    uint32_t prim_count;   /* >0  => leaf */
    uint32_t left_first;   /* branch: left child (right = left+1); leaf: first primitive */
} orc_bvh_node;

/* scene/tlas.rs:41-46  TlasNode: aabb, left_right (u16|u16), blas */
typedef struct {
    float    min[3];
    float    max[3];
    uint32_t left_right;   /* 0 => leaf */
    uint32_t blas;
} orc_tlas_node;

/* query/ray.rs:9-17 */
typedef struct { float o[3]; float d[3]; float rd[3]; float t; } orc_ray;

/* query/intersection.rs:77-92: 16-byte record {t,u,v,instance_primitive}; miss: t=FLT_MAX,id=0xFFFFFFFF */
typedef struct { float t; float u; float v; uint32_t id; } orc_hit;

/* camera/camera.rs:199-216 corner points (eye space) + cached view_matrix_inv (column-major) */
typedef struct { float tl[3]; float tr[3]; float bl[3]; float view_inv[16]; } orc_camera;

/* scene/scene_object.rs:16-21: cached inverse transform (column-major) + which model */
typedef struct { float inv[16]; uint32_t blas_id; } orc_instance;

typedef struct {
    const float*        tris;        /* n_tris x 9 floats, 36 B stride, BVH-reordered (mesh.rs:126-134) */
    uint32_t            n_tris;
    const orc_bvh_node* nodes;
    uint32_t            nodes_used;
} orc_blas;

typedef struct {
    const orc_tlas_node* tlas;
    uint32_t             tlas_nodes_used;
    const orc_instance*  inst;
    uint32_t             n_inst;
    const orc_blas*      blas;
    uint32_t             n_blas;
} orc_scene;

/* Per-frame work counters under the reference's exact traversal order. */
typedef struct {
    uint64_t rays;
    uint64_t hits;
    uint64_t blas_nodes;     /* BLAS nodes whose record is fetched: 1 root per BLAS entry + 2 per interior step */
    uint64_t tlas_nodes;     /* TLAS nodes fetched: 1 root + 2 per interior step */
    uint64_t inst;           /* SceneObject::intersect calls */
    uint64_t tri_area;       /* MT tests leaving at |area| < 1e-4           (20 ops) */
    uint64_t tri_u;          /* ... leaving at the u test                   (30 ops) */
    uint64_t tri_v;          /* ... leaving at the v / u+v test             (46 ops) */
    uint64_t tri_t;          /* ... reaching the t test (accept or reject)  (53 ops) */
    uint64_t box_tests;      /* Aabb::intersect calls                       (22 ops) */
    uint32_t max_blas_stack;
    uint32_t max_tlas_stack;
} orc_counters;

/* ---- primitives (geometry/aabb.rs:65-84, geometry/triangle.rs:41-72, query/ray.rs:23-35) ---- */
void  orc_ray_new(const float o[3], const float d[3], float t, orc_ray* out);
int   orc_aabb_intersect(const orc_aabb* box, const orc_ray* ray, float* t_out);
int   orc_triangle_intersect(const float tri[9], const orc_ray* ray, float tuv_out[3]);
void  orc_vec3_normalize(const float v[3], float out[3]);
void  orc_vec3_cross(const float a[3], const float b[3], float out[3]);
void  orc_triangle_centroid(const float tri[9], float out[3]);

/* ---- asset decode (tri_loader/src/{lexer,loader}.rs, mesh/decoders.rs:108-133, :157-215) ---- */
/* returns triangle count or <0; *out is malloc'd n*9 floats, free with orc_free */
int64_t orc_parse_tri(const char* text, size_t len, float** out);
int64_t orc_parse_obj(const char* text, size_t len, float** out);
int64_t orc_load_mesh_file(const char* path, float** out);   /* by extension .tri/.obj */
/* per-vertex normals in FILE order (n x 9 floats), as the decoders build them:
 *  .tri: normal = normalize(cross(normalize(v2-v0), normalize(v1-v0))) for all three vertices (mesh/decoders.rs:120-124)
 *  .obj: the `vn` referenced by each face corner (f64 -> f32), zero when the corner has none (mesh/decoders.rs:174-203) */
void    orc_tri_normals(const float* tris, uint64_t n, float* normals_out);
int64_t orc_parse_obj_normals(const char* text, size_t len, float** out);
void    orc_free(void* p);

/* ---- BLAS build / refit (model/bvh.rs:317-541) ---- */
/* tris reordered in place; nodes must hold 2*n entries (zero-filled by callee); returns nodes_used */
uint32_t orc_bvh_build(float* tris, uint32_t n_tris, orc_bvh_node* nodes);
void     orc_bvh_refit(const float* tris, orc_bvh_node* nodes, uint32_t nodes_used);
int      orc_bvh_intersect(const orc_blas* blas, const orc_ray* ray, orc_hit* hit, orc_counters* c);

/* ---- transforms (transform.rs:24-43, 83-96, 219-234; cglinalg) ---- */
void orc_mat4_identity(float m[16]);
void orc_mat4_mul(const float a[16], const float b[16], float out[16]);
int  orc_mat4_inverse(const float m[16], float out[16]);
void orc_mat4_mul_vec4(const float m[16], const float v[4], float out[4]);
void orc_transform_point(const float m[16], const float p[3], float out[3]);
void orc_transform_vector(const float m[16], const float v[3], float out[3]);
/* Transform3::new(scale, translation, Rx(ax) * Rz(az))  (examples/sixteen_armadillos.rs:104-106,135-141) */
void orc_transform_new_rot_xz(const float scale[3], const float trans[3], float angle_x, float angle_z, float out[16]);
void orc_transform_from_scale_translation(const float scale[3], const float trans[3], float out[16]);
/* scene_object.rs:60-75 / :118-131 : world AABB of 8 transformed corners */
void orc_instance_bounds(const float m[16], const orc_aabb* model_bounds, orc_aabb* out);

/* ---- TLAS (scene/tlas.rs:179-280) ---- */
/* nodes must hold 2*n entries; returns nodes_used */
uint32_t orc_tlas_build(const orc_aabb* bounds, uint32_t n, orc_tlas_node* nodes);

/* ---- camera (camera/camera.rs:199-251, 343-367, 809-835, 994-1010) ---- */
void orc_camera_symmetric_fov(float fovy_deg, float aspect, float near_,
                              const float pos[3], const float fwd[3], const float right[3], const float up[3],
                              orc_camera* out);
void orc_camera_box(float left, float right_, float bottom, float top, float near_,
                    const float pos[3], const float fwd[3], const float right[3], const float up[3],
                    orc_camera* out);
void orc_camera_ray_world(const orc_camera* cam, float u, float v, orc_ray* out);

/* ---- scene traversal (scene/scene.rs:32-34, tlas.rs:123-177, scene_object.rs:78-89) ---- */
int  orc_scene_intersect(const orc_scene* s, const orc_ray* ray, orc_hit* hit, orc_counters* c);

/* ---- frame (renderer.rs:345-368): pixels [x0,x1) x [y0,y1) of a W x H image, tile x tile order.
 * hits is the full W*H row-major buffer (only the region is written).
 * n_threads<=1: serial tile loop as in the reference; >1: OpenMP over tiles (our parallelisation). */
void orc_render(const orc_scene* s, const orc_camera* cam, uint32_t width, uint32_t height, uint32_t tile,
                uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                orc_hit* hits, orc_counters* counters, int n_threads);
/* N arbitrary rays (Scene::intersect(&Ray)); rays given as o,d,t (rd recomputed by Ray::new) */
void orc_trace_rays(const orc_scene* s, const float* o_d_t /* n x 7 */, uint64_t n, orc_hit* hits,
                    orc_counters* counters, int n_threads);

/* ---- accumulator + pixel shader pairs (renderer.rs:116-245) applied to hit records.
 * kind 1: DepthAccumulator + DepthMappingShader(scale, offset); 2: IntersectionAccumulator + IntersectionShader(hit, miss);
 * 3: UvMappingAccumulator + RadianceToRgbShader.  rgba_out: r | g<<8 | b<<16 | a<<24 (Rgba<u8> byte order). */
void orc_shade(uint32_t kind, float scale, float offset, uint32_t hit_rgba, uint32_t miss_rgba,
               const orc_hit* hits, uint64_t n, uint32_t* rgba_out);
/* kind 4: NormalMappingAccumulator (renderer.rs:256-286) + RadianceToRgbShader (:124-132).  normals = the UN-reordered
 * per-vertex normals of scene object 0's model indexed with the (BVH-reordered) primitive index, object0_transform =
 * scene object 0's forward transform -- both exactly what the reference looks up (its instance index is always 0). */
void orc_shade_normal(const float* normals, uint64_t n_prims, const float object0_transform[16],
                      const orc_hit* hits, uint64_t n, uint32_t* rgba_out);

/* kind 5: TextureMaterialAccumulator (renderer.rs:289-334) + TextureMaterial::evaluate (materials/material.rs:44-52) +
 * RadianceToRgbShader (:124-132).  tex_coords = UN-reordered per-vertex coordinates of scene object 0's model (n_prims x 6
 * floats), texels = that model's Rgb<u8> texture (texel (x, y) at (y * width + x) * 3, texture_buffer.rs:211). */
void orc_shade_texture(const float* tex_coords, uint64_t n_prims, const uint8_t* texels, uint32_t tex_w, uint32_t tex_h,
                       const orc_hit* hits, uint64_t n, uint32_t* rgba_out);

int  orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
