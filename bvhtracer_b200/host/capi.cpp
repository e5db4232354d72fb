// capi.cpp -- flat C handles over bvhtracer.hpp so that Python (ctypes) tests, bench.py and the example
// scripts can drive the C++ host mirror.  Errors: functions return NULL / negative and bvhx_last_error()
// holds the message (the reference would panic at the same points).
#include "bvhtracer.hpp"

using namespace bvhtracer;

namespace {
thread_local std::string g_err;
template <typename F> auto guard(F&& f, decltype(f()) on_error) -> decltype(f()) {
    try { return f(); } catch (const std::exception& e) { g_err = e.what(); return on_error; }
}
struct SceneBuilderHandle { Camera camera; std::vector<SceneObject> objects; };
}

#define BVHX_API extern "C" __attribute__((visibility("default")))

BVHX_API const char* bvhx_last_error() { return g_err.c_str(); }

// ---- meshes
BVHX_API void* bvhx_mesh_from_triangles(const float* tris, const float* normals_or_null, uint32_t n) {
    return guard([&]() -> void* {
        Mesh* m = new Mesh();
        m->primitives.resize(n);
        m->normals.resize(n);
        if (n) std::memcpy((void*)m->primitives.data(), tris, (size_t)n * sizeof(Triangle));
        if (n && normals_or_null) std::memcpy((void*)m->normals.data(), normals_or_null, (size_t)n * sizeof(Normals));
        return m;
    }, nullptr);
}
BVHX_API const float* bvhx_mesh_normals(void* mesh) { return (const float*)((Mesh*)mesh)->normals.data(); }
BVHX_API int bvhx_mesh_set_tex_coords(void* mesh, const float* tex_coords, uint32_t n) {
    return guard([&]() -> int {
        Mesh* m = (Mesh*)mesh;
        if (n != m->primitives.size()) throw std::invalid_argument("texture coordinate count differs from the primitive count");
        m->tex_coords.resize(n);
        if (n) std::memcpy((void*)m->tex_coords.data(), tex_coords, (size_t)n * sizeof(TextureCoordinates));
        return 0;
    }, -1);
}
BVHX_API const float* bvhx_mesh_tex_coords(void* mesh, uint32_t* n) {
    Mesh* m = (Mesh*)mesh;
    if (n) *n = (uint32_t)m->tex_coords.size();
    return (const float*)m->tex_coords.data();
}
BVHX_API void* bvhx_mesh_from_tri_text(const char* text, size_t len) {
    return guard([&]() -> void* { return new Mesh(TriMeshDecoder::read_mesh(text, len)); }, nullptr);
}
BVHX_API void* bvhx_mesh_from_obj_text(const char* text, size_t len) {
    return guard([&]() -> void* { return new Mesh(ObjMeshDecoder::read_mesh(text, len)); }, nullptr);
}
BVHX_API uint32_t bvhx_mesh_len(void* mesh) { return (uint32_t)((Mesh*)mesh)->primitives.size(); }
BVHX_API const float* bvhx_mesh_data(void* mesh) { return (const float*)((Mesh*)mesh)->primitives.data(); }
BVHX_API void bvhx_mesh_free(void* mesh) { delete (Mesh*)mesh; }

// ---- models (ModelBuilder::build: BVH build + in-place reorder)
BVHX_API void* bvhx_model_build(void* mesh) {
    return guard([&]() -> void* { return new ModelInstance(ModelBuilder().with_mesh(*(Mesh*)mesh).build()); }, nullptr);
}
BVHX_API void* bvhx_model_build_textured(void* mesh, const uint8_t* rgb, uint32_t width, uint32_t height) {
    return guard([&]() -> void* {
        TextureMaterial t(width, height, std::vector<uint8_t>(rgb, rgb + (size_t)width * height * 3));
        return new ModelInstance(ModelBuilder().with_mesh(*(Mesh*)mesh).with_texture(std::move(t)).build());
    }, nullptr);
}
BVHX_API const void* bvhx_model_nodes(void* model, uint32_t* nodes_used, uint32_t* nodes_len) {
    Model* m = ((ModelInstance*)model)->get();
    if (nodes_used) *nodes_used = m->bvh.nodes_used;
    if (nodes_len) *nodes_len = (uint32_t)m->bvh.nodes.size();
    return m->bvh.nodes.data();
}
BVHX_API const float* bvhx_model_tris(void* model, uint32_t* n_tris) {
    Model* m = ((ModelInstance*)model)->get();
    if (n_tris) *n_tris = (uint32_t)m->primitives().size();
    return (const float*)m->primitives().data();
}
BVHX_API void bvhx_model_set_vertices(void* model, const float* tris) {
    Model* m = ((ModelInstance*)model)->get();
    auto& p = m->primitives_mut();
    std::memcpy((void*)p.data(), tris, p.size() * sizeof(Triangle));
}
BVHX_API void bvhx_model_refit(void* model) { ((ModelInstance*)model)->get()->refit(); }
BVHX_API void bvhx_model_free(void* model) { delete (ModelInstance*)model; }

// ---- transforms
BVHX_API void bvhx_transform_new(const float scale[3], const float trans[3], float angle_x, float angle_z, float out[16]) {
    Transform3 t = Transform3::new_(Vector3(scale[0], scale[1], scale[2]), Vector3(trans[0], trans[1], trans[2]),
                                    Rotation3::from_angle_x(angle_x) * Rotation3::from_angle_z(angle_z));
    std::memcpy(out, t.matrix.m, 64);
}
BVHX_API void bvhx_transform_from_scale_translation(const float scale[3], const float trans[3], float out[16]) {
    Transform3 t = Transform3::from_scale_translation(Vector3(scale[0], scale[1], scale[2]), Vector3(trans[0], trans[1], trans[2]));
    std::memcpy(out, t.matrix.m, 64);
}
BVHX_API int bvhx_transform_inverse(const float m[16], float out[16]) {
    Matrix4x4 a, b; std::memcpy(a.m, m, 64);
    if (!a.inverse(b)) return -1;
    std::memcpy(out, b.m, 64);
    return 0;
}

// ---- camera
static CameraAttitudeSpec attitude(const float* pos, const float* fwd, const float* right, const float* up) {
    CameraAttitudeSpec a;
    a.position = Vector3(pos[0], pos[1], pos[2]); a.forward = Vector3(fwd[0], fwd[1], fwd[2]);
    a.right = Vector3(right[0], right[1], right[2]); a.up = Vector3(up[0], up[1], up[2]); a.axis = a.forward;
    return a;
}
BVHX_API void* bvhx_camera_symmetric_fov(float fovy_deg, float aspect, float near_, float far_, const float* pos, const float* fwd,
                                         const float* right, const float* up) {
    return guard([&]() -> void* { return new Camera(SymmetricFovSpec{ fovy_deg, aspect, near_, far_ }, attitude(pos, fwd, right, up)); }, nullptr);
}
BVHX_API void* bvhx_camera_box(float left, float right_, float bottom, float top, float near_, float far_, const float* pos,
                               const float* fwd, const float* right, const float* up) {
    return guard([&]() -> void* { return new Camera(BoxSpec{ left, right_, bottom, top, near_, far_ }, attitude(pos, fwd, right, up)); }, nullptr);
}
BVHX_API void bvhx_camera_to_ffi(void* cam, bvht_camera* out) { *out = ((Camera*)cam)->to_ffi(); }
BVHX_API void bvhx_camera_free(void* cam) { delete (Camera*)cam; }

// ---- scene
BVHX_API void* bvhx_scene_builder_new(void* camera) { return new SceneBuilderHandle{ *(Camera*)camera, {} }; }
BVHX_API int bvhx_scene_builder_add(void* sb, void* model, const float* transform16_or_null) {
    return guard([&]() -> int {
        SceneObjectBuilder b(*(ModelInstance*)model);
        if (transform16_or_null) { Transform3 t; std::memcpy(t.matrix.m, transform16_or_null, 64); b.with_transform(t); }
        ((SceneBuilderHandle*)sb)->objects.push_back(b.build());
        return 0;
    }, -1);
}
BVHX_API void* bvhx_scene_build(void* sb) {
    SceneBuilderHandle* h = (SceneBuilderHandle*)sb;
    void* s = guard([&]() -> void* { return new Scene(SceneBuilder(h->camera).with_objects(std::move(h->objects)).build()); }, nullptr);
    delete h;
    return s;
}
BVHX_API uint32_t bvhx_scene_len(void* scene) { return (uint32_t)((Scene*)scene)->objects().size(); }
BVHX_API int bvhx_scene_set_transform(void* scene, uint32_t i, const float m[16]) {
    return guard([&]() -> int {
        Transform3 t; std::memcpy(t.matrix.m, m, 64);
        ((Scene*)scene)->get_mut_unchecked(i).set_transform(t);
        return 0;
    }, -1);
}
// `for (i, t) in transforms { scene.get_mut_unchecked(i).set_transform(&t) }` in one call (a binding with a per-call cost pays it once)
BVHX_API int bvhx_scene_set_transforms(void* scene, const float* m, uint32_t n) {
    return guard([&]() -> int {
        Scene& s = *(Scene*)scene;
        if (n > s.objects().size()) throw std::runtime_error("set_transforms: more transforms than scene objects");
        for (uint32_t i = 0; i < n; ++i) {
            Transform3 t; std::memcpy(t.matrix.m, m + 16 * (size_t)i, 64);
            s.get_mut_unchecked(i).set_transform(t);
        }
        return 0;
    }, -1);
}
BVHX_API void bvhx_scene_rebuild(void* scene) { ((Scene*)scene)->rebuild(); }
BVHX_API const void* bvhx_scene_tlas(void* scene, uint32_t* nodes_used) {
    const Tlas& t = ((Scene*)scene)->tlas();
    if (nodes_used) *nodes_used = t.nodes_used;
    return t.nodes.data();
}
BVHX_API void bvhx_scene_instance(void* scene, uint32_t i, float inv16[16], float bounds6[6]) {
    const SceneObject& o = ((Scene*)scene)->get_unchecked(i);
    if (inv16) std::memcpy(inv16, o.get_transform_inv().matrix.m, 64);
    if (bounds6) {
        Aabb b = o.bounds();
        bounds6[0] = b.bounds_min.x; bounds6[1] = b.bounds_min.y; bounds6[2] = b.bounds_min.z;
        bounds6[3] = b.bounds_max.x; bounds6[4] = b.bounds_max.y; bounds6[5] = b.bounds_max.z;
    }
}
BVHX_API void bvhx_scene_camera(void* scene, bvht_camera* out) { *out = ((Scene*)scene)->active_camera().to_ffi(); }
BVHX_API void bvhx_scene_free(void* scene) { delete (Scene*)scene; }

// ---- renderer (Renderer::new(Box::new(CudaPathTracer::new())))
BVHX_API void* bvhx_renderer_new(uint32_t flags, int device, uint32_t tile) {
    return guard([&]() -> void* { return new Renderer(std::make_unique<CudaPathTracer>(flags, device, tile)); }, nullptr);
}
BVHX_API void* bvhx_renderer_ctx(void* renderer) { return ((CudaPathTracer*)((Renderer*)renderer)->integrator())->context(); }
BVHX_API void bvhx_renderer_free(void* renderer) { delete (Renderer*)renderer; }
BVHX_API void* bvhx_state_new(uint32_t kind, float scale, float offset, const uint8_t hit[4], const uint8_t miss[4], uint32_t width,
                              uint32_t height, int keep_hits) {
    return guard([&]() -> void* {
        ShadingPipeline s = kind == BVHT_SHADE_DEPTH ? ShadingPipeline::depth(scale, offset)
                          : kind == BVHT_SHADE_INTERSECTION ? ShadingPipeline::intersection(hit, miss)
                          : kind == BVHT_SHADE_NORMAL ? ShadingPipeline::normal()
                          : kind == BVHT_SHADE_TEXTURE ? ShadingPipeline::texture() : ShadingPipeline::uv();
        return new RendererState(s, width, height, keep_hits != 0);
    }, nullptr);
}
BVHX_API void* bvhx_state_new_external(uint32_t kind, float scale, float offset, const uint8_t hit[4], const uint8_t miss[4], uint32_t width,
                                       uint32_t height, uint32_t* frame) {
    return guard([&]() -> void* {
        ShadingPipeline s = kind == BVHT_SHADE_DEPTH ? ShadingPipeline::depth(scale, offset)
                          : kind == BVHT_SHADE_INTERSECTION ? ShadingPipeline::intersection(hit, miss)
                          : kind == BVHT_SHADE_NORMAL ? ShadingPipeline::normal()
                          : kind == BVHT_SHADE_TEXTURE ? ShadingPipeline::texture() : ShadingPipeline::uv();
        if (!frame) throw std::runtime_error("null external frame buffer");
        return new RendererState(s, width, height, frame);
    }, nullptr);
}
BVHX_API void bvhx_state_free(void* state) { delete (RendererState*)state; }
BVHX_API const uint32_t* bvhx_state_frame(void* state) { return ((RendererState*)state)->frame_buffer(); }
BVHX_API const bvht_hit* bvhx_state_hits(void* state) { return ((RendererState*)state)->hits(); }
BVHX_API int64_t bvhx_renderer_render(void* renderer, void* state, void* scene) {
    return guard([&]() -> int64_t { return (int64_t)((Renderer*)renderer)->render(*(RendererState*)state, *(Scene*)scene); }, (int64_t)-1);
}
BVHX_API int64_t bvhx_renderer_render_begin(void* renderer, void* state, void* scene) {
    return guard([&]() -> int64_t { return (int64_t)((Renderer*)renderer)->render_begin(*(RendererState*)state, *(Scene*)scene); }, (int64_t)-1);
}
BVHX_API int bvhx_renderer_render_end(void* renderer) {
    return guard([&]() -> int { ((Renderer*)renderer)->render_end(); return 0; }, -1);
}
BVHX_API int bvhx_renderer_sync_scene(void* renderer, void* scene) {
    return guard([&]() -> int { ((CudaPathTracer*)((Renderer*)renderer)->integrator())->sync_scene(*(Scene*)scene); return 0; }, -1);
}
BVHX_API int bvhx_renderer_update_transforms(void* renderer, void* scene, const float* transforms16, uint32_t n) {
    return guard([&]() -> int {
        std::vector<Transform3> t(n);
        for (uint32_t i = 0; i < n; ++i) std::memcpy(t[i].matrix.m, transforms16 + (size_t)i * 16, 64);
        ((CudaPathTracer*)((Renderer*)renderer)->integrator())->update_transforms(*(Scene*)scene, t);
        return 0;
    }, -1);
}
BVHX_API void* bvhx_renderer_build_model(void* renderer, void* mesh) {
    return guard([&]() -> void* {
        return new ModelInstance(((CudaPathTracer*)((Renderer*)renderer)->integrator())->build_model(*(Mesh*)mesh));
    }, nullptr);
}
BVHX_API int bvhx_renderer_rebuild_model(void* renderer, void* model) {
    return guard([&]() -> int { ((CudaPathTracer*)((Renderer*)renderer)->integrator())->rebuild_model(**(ModelInstance*)model); return 0; }, -1);
}
BVHX_API int bvhx_renderer_intersect(void* renderer, void* scene, const bvht_ray* rays, uint64_t n, bvht_hit* out) {
    return guard([&]() -> int { ((CudaPathTracer*)((Renderer*)renderer)->integrator())->intersect(*(Scene*)scene, rays, n, out); return 0; }, -1);
}
