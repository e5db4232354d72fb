// bvhtracer.hpp -- C++ host-side mirror of the reference's scene / BVH / TLAS / renderer interface.
//
// The reference's host code is Rust; there is no Rust toolchain in this image, so the host side above the
// C ABI (include/bvht.h) is written in C++ with the reference's names, argument meaning and error behaviour.
// It builds exactly the data the Rust crate would hand to the device (BVH-reordered triangles, node pools,
// TLAS nodes, inverse transforms, camera corner points) and renders through `CudaPathTracer`, a second
// `Integrator` next to the reference's `PathTracer` (renderer.rs:104-106, 337-385).
//
// There is deliberately NO CPU traversal in here: Scene::intersect and Integrator::evaluate run on the GPU
// through libbvht_cuda.so or fail.  (The CPU restatement of the traversal lives in oracle/ and is test-only.)
//
// Arithmetic notes: all f32, evaluated left to right with no FMA contraction (compile with
// -ffp-contract=off), mirroring rustc.  cglinalg (un-vendored) semantics are restated as in DESIGN.md:
// dot = (x*x + y*y) + z*z, normalize = v / |v|, Mat4*Vec4 = ((c0*x + c1*y) + c2*z) + c3*w, inverse = adjugate/det.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/bvht.h"

namespace bvhtracer {

// ------------------------------------------------------------------------------------------ cglinalg subset
struct Vector3 {
    float x = 0, y = 0, z = 0;
    Vector3() = default;
    Vector3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    static Vector3 from_fill(float v) { return Vector3(v, v, v); }
    static Vector3 zero() { return Vector3(0, 0, 0); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    Vector3 operator+(const Vector3& o) const { return Vector3(x + o.x, y + o.y, z + o.z); }
    Vector3 operator-(const Vector3& o) const { return Vector3(x - o.x, y - o.y, z - o.z); }
    Vector3 operator-() const { return Vector3(-x, -y, -z); }
    Vector3 operator*(float s) const { return Vector3(x * s, y * s, z * s); }
    Vector3 operator/(float s) const { return Vector3(x / s, y / s, z / s); }
    float dot(const Vector3& o) const { return (x * o.x + y * o.y) + z * o.z; }
    Vector3 cross(const Vector3& o) const { return Vector3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
    float magnitude() const { return std::sqrt(dot(*this)); }
    Vector3 normalize() const { return *this / magnitude(); }           // pinned: test_tri_mesh.rs:57-59
    static Vector3 component_min(const Vector3& a, const Vector3& b) { return Vector3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)); }
    static Vector3 component_max(const Vector3& a, const Vector3& b) { return Vector3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)); }
    static Vector3 unit_x() { return Vector3(1, 0, 0); }
    static Vector3 unit_y() { return Vector3(0, 1, 0); }
    static Vector3 unit_z() { return Vector3(0, 0, 1); }
};

// column-major 4x4 (cglinalg Matrix4x4: m[c][r])
struct Matrix4x4 {
    float m[16];
    static Matrix4x4 identity() { Matrix4x4 r; std::memset(r.m, 0, sizeof r.m); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
    float& at(int c, int r) { return m[c * 4 + r]; }
    float at(int c, int r) const { return m[c * 4 + r]; }
    void mul_vec4(const float v[4], float out[4]) const {
        float t[4];
        for (int r = 0; r < 4; ++r) t[r] = ((m[r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r] * v[3];
        std::memcpy(out, t, sizeof t);
    }
    Matrix4x4 operator*(const Matrix4x4& b) const {
        Matrix4x4 r;
        for (int c = 0; c < 4; ++c) mul_vec4(b.m + 4 * c, r.m + 4 * c);
        return r;
    }
    bool inverse(Matrix4x4& out) const {
        auto A = [&](int r, int c) { return m[c * 4 + r]; };
        float s0 = A(0,0) * A(1,1) - A(1,0) * A(0,1), s1 = A(0,0) * A(1,2) - A(1,0) * A(0,2);
        float s2 = A(0,0) * A(1,3) - A(1,0) * A(0,3), s3 = A(0,1) * A(1,2) - A(1,1) * A(0,2);
        float s4 = A(0,1) * A(1,3) - A(1,1) * A(0,3), s5 = A(0,2) * A(1,3) - A(1,2) * A(0,3);
        float c5 = A(2,2) * A(3,3) - A(3,2) * A(2,3), c4 = A(2,1) * A(3,3) - A(3,1) * A(2,3);
        float c3 = A(2,1) * A(3,2) - A(3,1) * A(2,2), c2 = A(2,0) * A(3,3) - A(3,0) * A(2,3);
        float c1 = A(2,0) * A(3,2) - A(3,0) * A(2,2), c0 = A(2,0) * A(3,1) - A(3,0) * A(2,1);
        float det = ((((s0 * c5 - s1 * c4) + s2 * c3) + s3 * c2) - s4 * c1) + s5 * c0;
        if (det == 0.0f) return false;
        float inv = 1.0f / det;
        float b[4][4];
        b[0][0] = ((A(1,1) * c5 - A(1,2) * c4) + A(1,3) * c3) * inv;
        b[0][1] = ((-A(0,1) * c5 + A(0,2) * c4) - A(0,3) * c3) * inv;
        b[0][2] = ((A(3,1) * s5 - A(3,2) * s4) + A(3,3) * s3) * inv;
        b[0][3] = ((-A(2,1) * s5 + A(2,2) * s4) - A(2,3) * s3) * inv;
        b[1][0] = ((-A(1,0) * c5 + A(1,2) * c2) - A(1,3) * c1) * inv;
        b[1][1] = ((A(0,0) * c5 - A(0,2) * c2) + A(0,3) * c1) * inv;
        b[1][2] = ((-A(3,0) * s5 + A(3,2) * s2) - A(3,3) * s1) * inv;
        b[1][3] = ((A(2,0) * s5 - A(2,2) * s2) + A(2,3) * s1) * inv;
        b[2][0] = ((A(1,0) * c4 - A(1,1) * c2) + A(1,3) * c0) * inv;
        b[2][1] = ((-A(0,0) * c4 + A(0,1) * c2) - A(0,3) * c0) * inv;
        b[2][2] = ((A(3,0) * s4 - A(3,1) * s2) + A(3,3) * s0) * inv;
        b[2][3] = ((-A(2,0) * s4 + A(2,1) * s2) - A(2,3) * s0) * inv;
        b[3][0] = ((-A(1,0) * c3 + A(1,1) * c1) - A(1,2) * c0) * inv;
        b[3][1] = ((A(0,0) * c3 - A(0,1) * c1) + A(0,2) * c0) * inv;
        b[3][2] = ((-A(3,0) * s3 + A(3,1) * s1) - A(3,2) * s0) * inv;
        b[3][3] = ((A(2,0) * s3 - A(2,1) * s1) + A(2,2) * s0) * inv;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out.m[c * 4 + r] = b[r][c];
        return true;
    }
};

// cglinalg Rotation3 (3x3, row-major here), enough for the examples: from_angle_x / from_angle_z / product
struct Rotation3 {
    float r[3][3];
    static Rotation3 identity() { Rotation3 o; std::memset(o.r, 0, sizeof o.r); o.r[0][0] = o.r[1][1] = o.r[2][2] = 1.0f; return o; }
    static Rotation3 from_angle_x(float a) {
        float c = std::cos(a), s = std::sin(a);
        Rotation3 o = identity(); o.r[1][1] = c; o.r[1][2] = -s; o.r[2][1] = s; o.r[2][2] = c; return o;
    }
    static Rotation3 from_angle_z(float a) {
        float c = std::cos(a), s = std::sin(a);
        Rotation3 o = identity(); o.r[0][0] = c; o.r[0][1] = -s; o.r[1][0] = s; o.r[1][1] = c; return o;
    }
    Rotation3 operator*(const Rotation3& b) const {
        Rotation3 o;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.r[i][j] = (r[i][0] * b.r[0][j] + r[i][1] * b.r[1][j]) + r[i][2] * b.r[2][j];
        return o;
    }
};

// ------------------------------------------------------------------------------------------ geometry / query
// geometry/aabb.rs:11-63 (Aabb::intersect is device-only here)
struct Aabb {
    Vector3 bounds_min, bounds_max;                                    // Default = point box at the origin (aabb.rs:12)
    static Aabb new_empty() { Aabb a; a.bounds_min = Vector3::from_fill(FLT_MAX); a.bounds_max = Vector3::from_fill(-FLT_MAX); return a; }
    void grow(const Vector3& p) { bounds_min = Vector3::component_min(bounds_min, p); bounds_max = Vector3::component_max(bounds_max, p); }
    void grow_aabb(const Aabb& o) { if (o.bounds_min.x != FLT_MAX) { grow(o.bounds_min); grow(o.bounds_max); } }
    float area() const { Vector3 e = bounds_max - bounds_min; return (e.x * e.y + e.y * e.z) + e.z * e.x; }
};

// geometry/triangle.rs:9-39 (Triangle::intersect is device-only here)
struct Triangle {
    Vector3 vertices[3];
    Vector3 centroid() const { const float one_third = 1.0f / 3.0f; return ((vertices[0] + vertices[1]) + vertices[2]) * one_third; }
};
static_assert(sizeof(Triangle) == 36, "Triangle must be 36 bytes (mesh.rs:126-134)");

// query/ray.rs:9-39
struct Ray {
    Vector3 origin, direction, recip_direction; float t;
    Ray() : t(FLT_MAX) {}
    Ray(const Vector3& o, const Vector3& d, float t_) : origin(o), direction(d), recip_direction(1.0f / d.x, 1.0f / d.y, 1.0f / d.z), t(t_) {}
    static Ray from_origin_dir(const Vector3& o, const Vector3& d) { return Ray(o, d, FLT_MAX); }
};

// query/intersection.rs:33-104 (16-byte record)
struct Intersection {
    float t, u, v; uint32_t instance_primitive;
    uint32_t instance_index() const { return (instance_primitive & 0xFFF00000u) >> 20; }
    uint32_t primitive_index() const { return instance_primitive & 0x000FFFFFu; }
};

// ------------------------------------------------------------------------------------------ mesh + decoders
// mesh/mesh.rs:100-206 (positions only: tex coords / normals are not on the traced path)
// Normals<f32, 3> (mesh.rs:55-96) has the layout of a Triangle: three Vector3
using Normals = Triangle;

// TextureCoordinates<f32, 3> (mesh.rs:8-51): three Vector2
struct Vector2 { float x = 0, y = 0; };
struct TextureCoordinates { Vector2 uv[3]; };

struct Mesh {
    std::vector<Triangle> primitives;
    std::vector<Normals> normals;             // per-vertex normals, NEVER reordered by the BVH build (bvh.rs:426 swaps positions only)
    std::vector<TextureCoordinates> tex_coords;   // likewise never reordered
    size_t len_primitives() const { return primitives.size(); }
};

struct MeshBuilder {
    Mesh mesh;
    MeshBuilder& with_primitive(const Triangle& t) { return with_primitive(t, TextureCoordinates(), Normals()); }
    MeshBuilder& with_primitive(const Triangle& t, const Normals& n) { return with_primitive(t, TextureCoordinates(), n); }
    MeshBuilder& with_primitive(const Triangle& t, const TextureCoordinates& tc, const Normals& n) {     // mesh.rs:189-198
        mesh.primitives.push_back(t); mesh.tex_coords.push_back(tc); mesh.normals.push_back(n); return *this;
    }
    Mesh build() { return std::move(mesh); }
};

// materials/material.rs:14-53 over texture_buffer.rs: a decoded Rgb<u8> image, texel (x, y) at (y * width + x) * 3.
// Decoding (PNG/JPEG, materials/decoders.rs) stays with the caller; evaluation runs on the device (BVHT_SHADE_TEXTURE).
struct TextureMaterial {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgb;
    TextureMaterial() = default;
    TextureMaterial(uint32_t w, uint32_t h, std::vector<uint8_t> texels) : width(w), height(h), rgb(std::move(texels)) {
        if (rgb.size() != (size_t)w * h * 3) throw std::invalid_argument("TextureMaterial: texel buffer does not match width x height x 3");
    }
    bool empty() const { return width == 0 || height == 0; }
};

// tri_loader/src/{lexer,loader}.rs + mesh/decoders.rs:102-134: nine f32 per triangle, ' ' '\\' '\t' are
// whitespace, '#' comments, every triangle kept (including the 999 sentinel that ends the shipped assets)
struct TriMeshDecoder {
    static Mesh read_mesh(const char* text, size_t len) {
        std::vector<float> vals;
        size_t i = 0;
        while (i < len) {
            char ch = text[i];
            if (ch == ' ' || ch == '\\' || ch == '\t' || ch == '\n' || ch == '\r') { ++i; continue; }
            if (ch == '#') { while (i < len && text[i] != '\n' && text[i] != '\r') ++i; continue; }
            size_t s = i;
            while (i < len && !(text[i] == ' ' || text[i] == '\\' || text[i] == '\t' || text[i] == '\n' || text[i] == '\r')) ++i;
            std::string tok(text + s, i - s);
            char* end = nullptr;
            float v = std::strtof(tok.c_str(), &end);
            if (end == tok.c_str() || *end != 0) throw std::runtime_error("Expected a floating point number but got `" + tok + "` instead.");
            vals.push_back(v);
        }
        if (vals.size() % 9 != 0) throw std::runtime_error("Reached the end of the input in the process of getting the next token.");
        MeshBuilder b;
        for (size_t k = 0; k + 8 < vals.size(); k += 9) {
            Triangle t;
            for (int v = 0; v < 3; ++v) t.vertices[v] = Vector3(vals[k + 3 * v], vals[k + 3 * v + 1], vals[k + 3 * v + 2]);
            // mesh/decoders.rs:120-124: one face normal on all three vertices
            Vector3 v0v2 = (t.vertices[2] - t.vertices[0]).normalize();
            Vector3 v0v1 = (t.vertices[1] - t.vertices[0]).normalize();
            Vector3 normal = v0v2.cross(v0v1).normalize();
            Normals n; n.vertices[0] = n.vertices[1] = n.vertices[2] = normal;
            b.with_primitive(t, n);
        }
        return b.build();
    }
};

// mesh/decoders.rs:150-216 over cgwavefront_obj: positions f64 -> f32, faces of the first object, fan-triangulated
struct ObjMeshDecoder {
    static Mesh read_mesh(const char* text, size_t len) {
        std::vector<Vector3> pos, nrm;
        std::vector<Vector2> tex;
        MeshBuilder b;
        size_t i = 0; int objects = 0;
        while (i < len) {
            size_t ls = i; while (i < len && text[i] != '\n') ++i;
            std::string line(text + ls, i - ls); if (i < len) ++i;
            size_t p = line.find_first_not_of(" \t");
            if (p == std::string::npos) continue;
            line = line.substr(p);
            if (line.size() < 2) continue;
            if (line[0] == 'o' && (line[1] == ' ' || line[1] == '\t')) { if (++objects > 1 && !b.mesh.primitives.empty()) break; continue; }
            bool is_v = line[0] == 'v' && (line[1] == ' ' || line[1] == '\t');
            bool is_vn = line.size() > 2 && line[0] == 'v' && line[1] == 'n' && (line[2] == ' ' || line[2] == '\t');
            bool is_vt = line.size() > 2 && line[0] == 'v' && line[1] == 't' && (line[2] == ' ' || line[2] == '\t');
            if (is_vt) {                                  // vt u v [w]: (u, v) kept, f64 -> f32 (decoders.rs:190)
                const char* c = line.c_str() + 2; char* e = nullptr; Vector2 t;
                double u = std::strtod(c, &e); if (e == c) throw std::runtime_error("bad texture vertex"); c = e;
                double v = std::strtod(c, &e); if (e == c) v = 0.0;
                t.x = (float)u; t.y = (float)v; tex.push_back(t);
            } else if (is_v || is_vn) {
                const char* c = line.c_str() + (is_v ? 1 : 2); char* e = nullptr; float v[3];
                for (int k = 0; k < 3; ++k) { double d = std::strtod(c, &e); if (e == c) throw std::runtime_error("bad vertex"); v[k] = (float)d; c = e; }
                (is_v ? pos : nrm).push_back(Vector3(v[0], v[1], v[2]));
            } else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
                std::vector<long> idx, nidx, tidx; const char* c = line.c_str() + 1;
                while (*c) {
                    while (*c == ' ' || *c == '\t' || *c == '\r') ++c;
                    if (!*c) break;
                    char* e = nullptr; long vi = std::strtol(c, &e, 10); if (e == c) break;
                    if (vi < 0) vi = (long)pos.size() + vi + 1;
                    c = e;
                    long ni = 0, ti = 0; bool have = false, have_t = false;
                    if (*c == '/') {                      // v/vt/vn, v//vn or v/vt
                        ++c;
                        if (*c != '/') { char* e1 = nullptr; ti = std::strtol(c, &e1, 10); if (e1 != c) { have_t = true; c = e1; } }
                        if (*c == '/') { ++c; char* e2 = nullptr; ni = std::strtol(c, &e2, 10); if (e2 != c) { have = true; c = e2; } }
                    }
                    if (have && ni < 0) ni = (long)nrm.size() + ni + 1;
                    if (have_t && ti < 0) ti = (long)tex.size() + ti + 1;
                    idx.push_back(vi - 1); nidx.push_back(have ? ni - 1 : -1); tidx.push_back(have_t ? ti - 1 : -1);
                    while (*c && *c != ' ' && *c != '\t') ++c;
                }
                for (size_t k = 1; k + 1 < idx.size(); ++k) {
                    size_t tri[3] = { 0, k, k + 1 };
                    Triangle t; Normals n; TextureCoordinates tc;
                    for (int v = 0; v < 3; ++v) {
                        long pi = idx[tri[v]], ni = nidx[tri[v]], ti = tidx[tri[v]];
                        if (ti >= 0 && (size_t)ti < tex.size()) tc.uv[v] = tex[ti];
                        if (pi < 0 || (size_t)pi >= pos.size()) throw std::runtime_error("face index out of range");
                        t.vertices[v] = pos[pi];
                        n.vertices[v] = (ni >= 0 && (size_t)ni < nrm.size()) ? nrm[ni] : Vector3::zero();   // decoders.rs:182-203
                    }
                    b.with_primitive(t, tc, n);
                }
            }
        }
        return b.build();
    }
};

// ------------------------------------------------------------------------------------------ BLAS (host build)
// model/bvh.rs:88-134 -- same 32-byte layout as bvht_bvh_node
struct BvhNode {
    Aabb aabb; uint32_t primitive_count = 0; uint32_t left_first = 0;
    bool is_leaf() const { return primitive_count > 0; }
};
static_assert(sizeof(BvhNode) == 32 && sizeof(BvhNode) == sizeof(bvht_bvh_node), "BvhNode must be 32 bytes (bvh.rs:720-723)");

struct Bvh {
    std::vector<BvhNode> nodes;      // 2N allocated (bvh.rs:530)
    uint32_t root_node_index = 0;
    uint32_t nodes_used = 2;         // node 1 is the alignment dummy (bvh.rs:505-522)
    Aabb bounds() const { return nodes[0].aabb; }
};

class BvhBuilder {
public:
    // model/bvh.rs:524-541; reorders `mesh` in place (bvh.rs:419-430)
    Bvh build_for(std::vector<Triangle>& mesh) {
        Bvh bvh;
        bvh.nodes.assign(2 * mesh.size(), BvhNode());
        if (mesh.empty()) throw std::runtime_error("BvhBuilder::build_for: empty mesh (the reference indexes nodes[0] and panics)");
        bvh.nodes[0].left_first = 0;
        bvh.nodes[0].primitive_count = (uint32_t)mesh.size();
        update_node_bounds(bvh, mesh, 0);
        subdivide(bvh, mesh, 0);
        return bvh;
    }
    // model/bvh.rs:317-330
    static void update_node_bounds(Bvh& bvh, const std::vector<Triangle>& mesh, uint32_t ni) {
        BvhNode& n = bvh.nodes[ni];
        Aabb a = Aabb::new_empty();
        for (uint32_t i = 0; i < n.primitive_count; ++i) {
            const Triangle& t = mesh[n.left_first + i];
            for (int v = 0; v < 3; ++v) a.bounds_min = Vector3::component_min(a.bounds_min, t.vertices[v]);
            for (int v = 0; v < 3; ++v) a.bounds_max = Vector3::component_max(a.bounds_max, t.vertices[v]);
        }
        n.aabb = a;
    }
private:
    static constexpr int BIN_COUNT = 8;
    // model/bvh.rs:333-394, quirks kept: bounds_max starts at 1e-30; bins and sweep boxes start as Aabb::default()
    // (the point box at the origin, NOT new_empty()); bin index = saturating float->usize cast then min(7, .)
    static void find_best_split_plane(const Bvh& bvh, const std::vector<Triangle>& mesh, const BvhNode& node,
                                      int& best_axis, float& best_position, float& best_cost) {
        best_axis = -1; best_position = 0.0f; best_cost = FLT_MAX;
        for (int axis = 0; axis < 3; ++axis) {
            float bounds_min = 1e30f, bounds_max = 1e-30f;
            for (uint32_t i = 0; i < node.primitive_count; ++i) {
                float c = mesh[node.left_first + i].centroid()[axis];
                bounds_min = std::fmin(bounds_min, c); bounds_max = std::fmax(bounds_max, c);
            }
            if (bounds_min == bounds_max) continue;
            Aabb bin_box[BIN_COUNT]; uint32_t bin_cnt[BIN_COUNT] = {};
            float bin_scale = (float)BIN_COUNT / (bounds_max - bounds_min);
            for (uint32_t i = 0; i < node.primitive_count; ++i) {
                const Triangle& t = mesh[node.left_first + i];
                float f = (t.centroid()[axis] - bounds_min) * bin_scale;
                int bi = f >= (float)BIN_COUNT ? BIN_COUNT - 1 : (f > 0.0f ? (int)f : 0);
                bin_cnt[bi] += 1;
                bin_box[bi].grow(t.vertices[0]); bin_box[bi].grow(t.vertices[1]); bin_box[bi].grow(t.vertices[2]);
            }
            float left_area[BIN_COUNT - 1], right_area[BIN_COUNT - 1];
            uint32_t left_count[BIN_COUNT - 1], right_count[BIN_COUNT - 1];
            Aabb left_box, right_box; uint32_t left_sum = 0, right_sum = 0;
            for (int i = 0; i < BIN_COUNT - 1; ++i) {
                left_sum += bin_cnt[i]; left_count[i] = left_sum;
                left_box.grow_aabb(bin_box[i]); left_area[i] = left_box.area();
                right_sum += bin_cnt[BIN_COUNT - 1 - i]; right_count[BIN_COUNT - 2 - i] = right_sum;
                right_box.grow_aabb(bin_box[BIN_COUNT - 1 - i]); right_area[BIN_COUNT - 2 - i] = right_box.area();
            }
            float scale = (bounds_max - bounds_min) / (float)BIN_COUNT;
            for (int i = 0; i < BIN_COUNT - 1; ++i) {
                float plane_cost = (float)left_count[i] * left_area[i] + (float)right_count[i] * right_area[i];
                if (plane_cost < best_cost) { best_axis = axis; best_position = bounds_min + scale * (float)(i + 1); best_cost = plane_cost; }
            }
        }
        (void)bvh;
    }
    // model/bvh.rs:396-467
    void subdivide(Bvh& bvh, std::vector<Triangle>& mesh, uint32_t ni) {
        int axis; float split, cost;
        find_best_split_plane(bvh, mesh, bvh.nodes[ni], axis, split, cost);
        BvhNode& node = bvh.nodes[ni];
        float no_split_cost = (float)node.primitive_count * node.aabb.area();
        if (cost >= no_split_cost || axis < 0) return;
        int64_t i = node.left_first, j = i + (int64_t)node.primitive_count - 1;
        while (i <= j) {
            if (mesh[(size_t)i].centroid()[axis] < split) i += 1;
            else { std::swap(mesh[(size_t)i], mesh[(size_t)j]); j -= 1; }
        }
        uint32_t left_count = (uint32_t)(i - (int64_t)node.left_first);
        if (left_count == 0 || left_count == node.primitive_count) return;
        uint32_t l = bvh.nodes_used++, r = bvh.nodes_used++;
        bvh.nodes[l].left_first = node.left_first; bvh.nodes[l].primitive_count = left_count;
        bvh.nodes[r].left_first = (uint32_t)i;     bvh.nodes[r].primitive_count = node.primitive_count - left_count;
        node.left_first = l; node.primitive_count = 0;
        update_node_bounds(bvh, mesh, l); update_node_bounds(bvh, mesh, r);
        subdivide(bvh, mesh, l); subdivide(bvh, mesh, r);
    }
};

// model/model.rs:16-146.  `refit()` marks the model: the refit itself (bvh.rs:469-493) runs on the device the
// next time an integrator renders, and the refitted node boxes are read back into `bvh.nodes`.
struct Model {
    Mesh mesh; Bvh bvh; TextureMaterial texture_;
    const TextureMaterial& texture() const { return texture_; }
    uint64_t geometry_version = 1;     // bumped when vertices change
    bool refit_requested = false;
    Aabb bounds() const { return bvh.bounds(); }
    std::vector<Triangle>& primitives_mut() { geometry_version++; return mesh.primitives; }
    const std::vector<Triangle>& primitives() const { return mesh.primitives; }
    const std::vector<Normals>& normals() const { return mesh.normals; }
    const std::vector<TextureCoordinates>& tex_coords() const { return mesh.tex_coords; }
    void refit() { refit_requested = true; }
};
using ModelInstance = std::shared_ptr<Model>;

struct ModelBuilder {
    Mesh mesh; TextureMaterial texture;
    ModelBuilder& with_mesh(Mesh m) { mesh = std::move(m); return *this; }
    ModelBuilder& with_texture(TextureMaterial t) { texture = std::move(t); return *this; }   // model.rs:128-132
    ModelInstance build() {                                             // model.rs:140-144
        auto model = std::make_shared<Model>();
        model->mesh = std::move(mesh);
        model->texture_ = std::move(texture);
        model->bvh = BvhBuilder().build_for(model->mesh.primitives);
        return model;
    }
};

// ------------------------------------------------------------------------------------------ transforms
// transform.rs:12-256
struct Transform3 {
    Matrix4x4 matrix = Matrix4x4::identity();
    static Transform3 identity() { return Transform3(); }
    static Transform3 new_(const Vector3& scale, const Vector3& translation, const Rotation3& rotation) {   // :24-43
        Transform3 t;
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) t.matrix.at(c, r) = rotation.r[r][c];
        t.matrix.at(3, 0) = translation.x; t.matrix.at(3, 1) = translation.y; t.matrix.at(3, 2) = translation.z;
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) t.matrix.at(c, r) *= scale[c];
        return t;
    }
    static Transform3 from_scale_translation(const Vector3& scale, const Vector3& translation) {            // :83-96
        Transform3 t;
        t.matrix.at(3, 0) = translation.x; t.matrix.at(3, 1) = translation.y; t.matrix.at(3, 2) = translation.z;
        t.matrix.at(0, 0) = scale.x; t.matrix.at(1, 1) = scale.y; t.matrix.at(2, 2) = scale.z;
        return t;
    }
    static Transform3 from_translation(const Vector3& translation) { return from_scale_translation(Vector3(1, 1, 1), translation); }
    Vector3 transform_point(const Vector3& p) const { float v[4] = { p.x, p.y, p.z, 1.0f }, r[4]; matrix.mul_vec4(v, r); return Vector3(r[0], r[1], r[2]); }
    Vector3 transform_vector(const Vector3& p) const { float v[4] = { p.x, p.y, p.z, 0.0f }, r[4]; matrix.mul_vec4(v, r); return Vector3(r[0], r[1], r[2]); }
    Transform3 inverse() const {
        Transform3 t;
        if (!matrix.inverse(t.matrix)) throw std::runtime_error("called `Option::unwrap()` on a `None` value (singular transform)");
        return t;
    }
};

// ------------------------------------------------------------------------------------------ camera
struct SymmetricFovSpec { float fovy_degrees, aspect, near_, far_; };
struct BoxSpec { float left, right, bottom, top, near_, far_; };
struct CameraAttitudeSpec { Vector3 position, forward, right, up, axis; };

// camera/camera.rs:157-251, 343-367, 809-835, 888-1011 (ray generation itself is device-only)
class Camera {
public:
    Camera(const SymmetricFovSpec& s, const CameraAttitudeSpec& a) {
        float fovy_over_two = s.fovy_degrees / 2.0f;
        float tan_half = std::tan(fovy_over_two * ((float)M_PI / 180.0f));   // Degrees::tan (cglinalg): parity unpinned
        float top = s.near_ * tan_half, bottom = -top, left = -s.aspect * top, right = s.aspect * top;
        set_frustum(left, top, right - left, top - bottom, s.near_);
        set_attitude(a);
    }
    Camera(const BoxSpec& s, const CameraAttitudeSpec& a) {
        set_frustum(s.left, s.top, s.right - s.left, s.top - s.bottom, s.near_);
        set_attitude(a);
    }
    Vector3 top_left_eye() const { return tl_; }
    Vector3 top_right_eye() const { return tr_; }
    Vector3 bottom_left_eye() const { return bl_; }
    const Matrix4x4& view_matrix() const { return view_; }
    const Matrix4x4& view_matrix_inv() const { return view_inv_; }
    Vector3 position() const { return position_; }
    bvht_camera to_ffi() const {
        bvht_camera c;
        const Vector3* src[3] = { &tl_, &tr_, &bl_ };
        float* dst[3] = { c.top_left_eye, c.top_right_eye, c.bottom_left_eye };
        for (int k = 0; k < 3; ++k) { dst[k][0] = src[k]->x; dst[k][1] = src[k]->y; dst[k][2] = src[k]->z; }
        std::memcpy(c.view_matrix_inv, view_inv_.m, sizeof c.view_matrix_inv);
        return c;
    }
private:
    void set_frustum(float left, float top, float ext_x, float ext_y, float near_) {
        tl_ = Vector3(left, top, -near_); tr_ = Vector3(left + ext_x, top, -near_); bl_ = Vector3(left, top - ext_y, -near_);
    }
    void set_attitude(const CameraAttitudeSpec& a) {
        position_ = a.position;
        Matrix4x4 tr = Matrix4x4::identity();
        tr.at(3, 0) = -a.position.x; tr.at(3, 1) = -a.position.y; tr.at(3, 2) = -a.position.z;
        Matrix4x4 rot = Matrix4x4::identity();
        const float cols[16] = { a.right.x, a.up.x, -a.forward.x, 0.0f, a.right.y, a.up.y, -a.forward.y, 0.0f,
                                 a.right.z, a.up.z, -a.forward.z, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f };
        std::memcpy(rot.m, cols, sizeof cols);
        view_ = rot * tr;
        if (!view_.inverse(view_inv_)) throw std::runtime_error("singular view matrix");
    }
    Vector3 tl_, tr_, bl_, position_;
    Matrix4x4 view_, view_inv_;
};

// ------------------------------------------------------------------------------------------ scene
// scene/scene_object.rs:16-137
class SceneObject {
public:
    SceneObject(ModelInstance model, const Transform3& transform, const Aabb& bounds)
        : model_(std::move(model)), transform_init_(transform), transform_(transform), transform_inv_(transform.inverse()), bounds_(bounds) {}
    const Transform3& get_transform() const { return transform_; }
    const Transform3& get_transform_inv() const { return transform_inv_; }
    Aabb bounds() const { return bounds_; }
    ModelInstance model() const { return model_; }
    void set_transform(const Transform3& t) {                           // :60-75
        bounds_ = world_bounds(model_->bounds(), t);
        transform_ = t; transform_inv_ = t.inverse();
    }
    // state computed by the device for this object (CudaPathTracer::update_transforms): same values set_transform yields
    void adopt(const Transform3& t, const Transform3& t_inv, const Aabb& world) { transform_ = t; transform_inv_ = t_inv; bounds_ = world; }
    static Aabb world_bounds(const Aabb& ob, const Transform3& t) {
        Aabb nb = Aabb::new_empty();
        for (int i = 0; i < 8; ++i) {
            Vector3 p((i & 1) ? ob.bounds_max.x : ob.bounds_min.x, (i & 2) ? ob.bounds_max.y : ob.bounds_min.y,
                      (i & 4) ? ob.bounds_max.z : ob.bounds_min.z);
            nb.grow(t.transform_point(p));
        }
        return nb;
    }
private:
    ModelInstance model_;
    Transform3 transform_init_, transform_, transform_inv_;
    Aabb bounds_;
};

struct SceneObjectBuilder {                                             // :92-137
    ModelInstance model; Transform3 transform = Transform3::identity(); Aabb bounds = Aabb::new_empty();
    explicit SceneObjectBuilder(ModelInstance m) : model(std::move(m)) {}
    SceneObjectBuilder& with_transform(const Transform3& t) { bounds = SceneObject::world_bounds(model->bounds(), t); transform = t; return *this; }
    SceneObject build() const { return SceneObject(model, transform, bounds); }
};

// scene/tlas.rs:11-280 -- same 32-byte layout as bvht_tlas_node
struct TlasNode { Aabb aabb = Aabb::new_empty(); uint32_t left_right = 0; uint32_t blas = 0; };
static_assert(sizeof(TlasNode) == 32 && sizeof(TlasNode) == sizeof(bvht_tlas_node), "TlasNode must be 32 bytes");

class Tlas {
public:
    std::vector<TlasNode> nodes; uint32_t nodes_used = 2;
    void rebuild(const std::vector<SceneObject>& blas) {               // :204-250
        size_t n = blas.size();
        if (nodes.size() < 2 * std::max<size_t>(n, 1)) nodes.resize(2 * std::max<size_t>(n, 1));
        if (n == 0) { nodes_used = 2; return; }
        std::vector<int32_t> idx(n);
        int32_t count = (int32_t)n, used = 1;
        for (size_t i = 0; i < n; ++i) {
            idx[i] = used;
            nodes[used].aabb = blas[i].bounds(); nodes[used].blas = (uint32_t)i; nodes[used].left_right = 0;
            used += 1;
        }
        int32_t a = 0, b = find_best_match(idx, count, a);
        while (count > 1) {
            int32_t c = find_best_match(idx, count, b);
            if (a == c) {
                int32_t ia = idx[a], ib = idx[b];
                TlasNode na = nodes[ia], nb = nodes[ib];
                TlasNode& nn = nodes[used];
                nn.left_right = (uint32_t)ia + ((uint32_t)ib << 16);   // LeftRightIndex::new(left = ia, right = ib), :18-22
                nn.aabb.bounds_min = Vector3::component_min(na.aabb.bounds_min, nb.aabb.bounds_min);
                nn.aabb.bounds_max = Vector3::component_max(na.aabb.bounds_max, nb.aabb.bounds_max);
                idx[a] = used; used += 1;
                idx[b] = idx[count - 1]; count -= 1;
                b = find_best_match(idx, count, a);
            } else { a = b; b = c; }
        }
        nodes[0] = nodes[idx[a]];
        nodes_used = (uint32_t)used;
    }
private:
    int32_t find_best_match(const std::vector<int32_t>& list, int32_t n, int32_t a) const {   // :179-202
        float smallest = FLT_MAX; int32_t best = -1;
        for (int32_t b = 0; b < n; ++b) {
            if (b == a) continue;
            Vector3 mx = Vector3::component_max(nodes[list[a]].aabb.bounds_max, nodes[list[b]].aabb.bounds_max);
            Vector3 mn = Vector3::component_min(nodes[list[a]].aabb.bounds_min, nodes[list[b]].aabb.bounds_min);
            Vector3 e = mx - mn;
            float area = (e.x * e.y + e.y * e.z) + e.z * e.x;
            if (area < smallest) { smallest = area; best = b; }
        }
        return best;
    }
};

struct TlasBuilder {                                                    // :253-280
    static Tlas build_for(const std::vector<SceneObject>& blas) {
        Tlas t; t.nodes.assign(2 * std::max<size_t>(blas.size(), 1), TlasNode()); t.rebuild(blas); return t;
    }
};

// scene/scene.rs:8-97
class Scene {
public:
    Scene(std::vector<SceneObject> objects, Camera camera) : objects_(std::move(objects)), camera_(std::move(camera)), tlas_(TlasBuilder::build_for(objects_)) {}
    const SceneObject& get_unchecked(size_t i) const { return objects_[i]; }
    SceneObject& get_mut_unchecked(size_t i) { version_++; return objects_[i]; }
    const std::vector<SceneObject>& objects() const { return objects_; }
    const Camera& active_camera() const { return camera_; }
    Camera& active_camera_mut() { return camera_; }
    const Tlas& tlas() const { return tlas_; }
    void rebuild() { tlas_.rebuild(objects_); version_++; }
    uint64_t version() const { return version_; }
    // the device ran set_transform + Tlas::rebuild (CudaPathTracer::update_transforms): take its results
    Tlas& tlas_mut() { version_++; return tlas_; }
private:
    std::vector<SceneObject> objects_;
    Camera camera_;
    Tlas tlas_;
    uint64_t version_ = 1;
};

struct SceneBuilder {
    Camera camera; std::vector<SceneObject> objects;
    explicit SceneBuilder(Camera c) : camera(std::move(c)) {}
    SceneBuilder& with_object(SceneObject o) { objects.push_back(std::move(o)); return *this; }
    SceneBuilder& with_objects(std::vector<SceneObject> o) { objects = std::move(o); return *this; }
    Scene build() { return Scene(std::move(objects), camera); }
};

// ------------------------------------------------------------------------------------------ renderer
// Accumulator + PixelShader pairs (renderer.rs:116-245) that the device evaluates; see bvht_shade_params.
struct ShadingPipeline {
    bvht_shade_params params;
    static ShadingPipeline depth(float scale, float offset) { ShadingPipeline s{}; s.params.kind = BVHT_SHADE_DEPTH; s.params.depth_scale = scale; s.params.depth_offset = offset; return s; }
    static ShadingPipeline intersection(const uint8_t hit[4], const uint8_t miss[4]) {
        ShadingPipeline s{}; s.params.kind = BVHT_SHADE_INTERSECTION; std::memcpy(s.params.hit_rgba, hit, 4); std::memcpy(s.params.miss_rgba, miss, 4); return s;
    }
    static ShadingPipeline uv() { ShadingPipeline s{}; s.params.kind = BVHT_SHADE_UV; return s; }
    static ShadingPipeline normal() { ShadingPipeline s{}; s.params.kind = BVHT_SHADE_NORMAL; return s; }   // object0_transform is filled per frame
    static ShadingPipeline texture() { ShadingPipeline s{}; s.params.kind = BVHT_SHADE_TEXTURE; return s; }  // TextureMaterialAccumulator + RadianceToRgbShader
};

// renderer.rs:76-102: frame buffer (Rgba<u8>, 4 B/px) + the per-pixel hit records (the GPU's "accumulation buffer")
class RendererState {
public:
    RendererState(const ShadingPipeline& shading, size_t width, size_t height, bool keep_hits)
        : shading_(shading), width_(width), height_(height), keep_hits_(keep_hits) {}
    // A frame buffer the caller owns (page-locked, width * height Rgba<u8>): several processes, one per GPU, can then share ONE
    // frame in POSIX shared memory and each fill the tile rows its integrator is sharded to (bvht_set_shard).
    RendererState(const ShadingPipeline& shading, size_t width, size_t height, uint32_t* external_frame)
        : shading_(shading), width_(width), height_(height), keep_hits_(false), frame_(external_frame), external_(true) {}
    ~RendererState() { release(); }
    RendererState(const RendererState&) = delete;
    RendererState& operator=(const RendererState&) = delete;
    size_t width() const { return width_; }
    size_t height() const { return height_; }
    const uint32_t* frame_buffer() const { return frame_; }
    const bvht_hit* hits() const { return hits_; }
    const ShadingPipeline& shading() const { return shading_; }
    bool keep_hits() const { return keep_hits_; }
    // buffers are page-locked host memory (not tied to the integrator's context: a state may outlive it)
    void bind(bvht_ctx*) {
        if (frame_) return;
        void* p = nullptr;
        if (bvht_host_alloc(nullptr, width_ * height_ * 4, &p) != BVHT_OK) throw std::runtime_error("bvht_host_alloc failed (frame buffer)");
        frame_ = (uint32_t*)p;
        for (size_t i = 0; i < width_ * height_; ++i) frame_[i] = 0xFF000000u;        // Rgba::from([0, 0, 0, 255]), renderer.rs:87-91
        if (keep_hits_) {
            if (bvht_host_alloc(nullptr, width_ * height_ * sizeof(bvht_hit), &p) != BVHT_OK) throw std::runtime_error("bvht_host_alloc failed (hit records)");
            hits_ = (bvht_hit*)p;
        }
    }
    uint32_t* frame_mut() { return frame_; }
    bvht_hit* hits_mut() { return hits_; }
private:
    void release() {
        if (frame_ && !external_) bvht_host_free(nullptr, frame_);
        if (hits_) bvht_host_free(nullptr, hits_);
        frame_ = nullptr; hits_ = nullptr;
    }
    ShadingPipeline shading_;
    size_t width_, height_;
    bool keep_hits_;
    uint32_t* frame_ = nullptr;
    bvht_hit* hits_ = nullptr;
    bool external_ = false;
};

// renderer.rs:104-106
class Integrator {
public:
    virtual ~Integrator() = default;
    virtual size_t evaluate(RendererState& state, Scene& scene) = 0;
};

// The drop-in: a second Integrator next to the reference's PathTracer (renderer.rs:337-385).  evaluate() =
// (1) upload every model once (bvht_blas_create) / after vertex changes (bvht_blas_update_vertices + bvht_blas_refit),
// (2) upload this frame's TLAS + inverse transforms + camera (bvht_tlas_set), (3) ONE call that traces all
// primary rays and shades them on the device (bvht_render_frame).  Returns rays traced, like PathTracer.
class CudaPathTracer : public Integrator {
public:
    explicit CudaPathTracer(uint32_t flags = BVHT_FLAG_STRICT | BVHT_FLAG_LEAF_ACCEL, int device = 0, uint32_t tile = 8) : tile_(tile) {
        int rc = bvht_create(device, flags, &ctx_);
        if (rc != BVHT_OK) throw std::runtime_error(std::string("bvht_create: ") + bvht_status_string(rc));
    }
    ~CudaPathTracer() override { if (ctx_) bvht_destroy(ctx_); }
    CudaPathTracer(const CudaPathTracer&) = delete;
    CudaPathTracer& operator=(const CudaPathTracer&) = delete;
    bvht_ctx* context() const { return ctx_; }

    // ModelBuilder::build (model.rs:140-144) with BvhBuilder::build_for running ON THE DEVICE (bvht_blas_build): the model
    // comes back with the reference's node pool and its primitives reordered in place, already resident for rendering.
    ModelInstance build_model(Mesh mesh, TextureMaterial texture = TextureMaterial()) {
        if (mesh.primitives.empty()) throw std::runtime_error("BvhBuilder::build_for: empty mesh (the reference indexes nodes[0] and panics)");
        auto model = std::make_shared<Model>();
        model->mesh = std::move(mesh);
        model->texture_ = std::move(texture);
        const uint32_t n = (uint32_t)model->mesh.primitives.size();
        uint32_t id = 0, n_tris = 0, used = 0;
        check(bvht_blas_build(ctx_, (const float*)model->mesh.primitives.data(), n, &id));
        check(bvht_blas_info(ctx_, id, &n_tris, &used));
        model->bvh.nodes.assign(2 * (size_t)n, BvhNode());
        model->bvh.nodes_used = used;
        check(bvht_blas_read_nodes(ctx_, id, (bvht_bvh_node*)model->bvh.nodes.data(), used));
        check(bvht_blas_read_triangles(ctx_, id, (float*)model->mesh.primitives.data(), n));
        upload_attributes(*model, id);
        uploaded_.push_back(Uploaded{ model.get(), id, model->geometry_version });
        return model;
    }

    // The "rebuild" alternative to ModelInstance::refit (benches/bench_bvh_refit_rebuild.rs) for a model whose vertices moved
    void rebuild_model(Model& m) {
        Uploaded* u = find(&m);
        if (!u) throw std::runtime_error("rebuild_model: the model is not resident on this integrator");
        if (u->version != m.geometry_version)
            check(bvht_blas_update_vertices(ctx_, u->blas_id, (const float*)m.primitives().data(), (uint32_t)m.primitives().size()));
        check(bvht_blas_rebuild(ctx_, u->blas_id));
        uint32_t n_tris = 0, used = 0;
        check(bvht_blas_info(ctx_, u->blas_id, &n_tris, &used));
        m.bvh.nodes.assign(2 * (size_t)n_tris, BvhNode());
        m.bvh.nodes_used = used;
        check(bvht_blas_read_nodes(ctx_, u->blas_id, (bvht_bvh_node*)m.bvh.nodes.data(), used));
        check(bvht_blas_read_triangles(ctx_, u->blas_id, (float*)m.mesh.primitives.data(), n_tris));
        m.refit_requested = false;
        u->version = m.geometry_version;
        scene_ = nullptr;                                                   // bounds changed: the next frame re-sends the TLAS
    }

    // upload / refresh every model the scene uses; returns the blas id of each scene object
    std::vector<uint32_t> sync_models(Scene& scene) {
        std::vector<uint32_t> blas_ids(scene.objects().size());
        for (size_t i = 0; i < scene.objects().size(); ++i) {
            Model* m = scene.objects()[i].model().get();
            Uploaded* u = find(m);
            if (!u) {
                uploaded_.push_back(Uploaded{ m, 0, 0 });
                u = &uploaded_.back();
                check(bvht_blas_create(ctx_, (const float*)m->primitives().data(), (uint32_t)m->primitives().size(),
                                       (const bvht_bvh_node*)m->bvh.nodes.data(), m->bvh.nodes_used, &u->blas_id));
                u->version = m->geometry_version;
                upload_attributes(*m, u->blas_id);
            } else if (u->version != m->geometry_version || m->refit_requested) {
                if (u->version != m->geometry_version)
                    check(bvht_blas_update_vertices(ctx_, u->blas_id, (const float*)m->primitives().data(), (uint32_t)m->primitives().size()));
                u->version = m->geometry_version;
            }
            if (m->refit_requested) {                                   // ModelInstance::refit (model.rs:31-33)
                check(bvht_blas_refit(ctx_, u->blas_id));
                check(bvht_blas_read_nodes(ctx_, u->blas_id, (bvht_bvh_node*)m->bvh.nodes.data(), m->bvh.nodes_used));
                m->refit_requested = false;
            }
            blas_ids[i] = u->blas_id;
        }
        return blas_ids;
    }

    void sync_scene(Scene& scene) {
        std::vector<uint32_t> blas_ids = sync_models(scene);       // (1) models
        // (2) per-frame state
        if (scene_ != &scene || scene_version_ != scene.version() || instances_.size() != scene.objects().size()) {
            instances_.resize(scene.objects().size());
            for (size_t i = 0; i < instances_.size(); ++i) {
                std::memcpy(instances_[i].transform_inv, scene.objects()[i].get_transform_inv().matrix.m, 64);
                instances_[i].blas_id = blas_ids[i];
            }
            check(bvht_tlas_set(ctx_, (const bvht_tlas_node*)scene.tlas().nodes.data(), scene.tlas().nodes_used, instances_.data(),
                                (uint32_t)instances_.size()));
            scene_ = &scene; scene_version_ = scene.version();
        }
    }

    // `for (i, t) in transforms { scene.get_mut_unchecked(i).set_transform(&t) }; scene.rebuild()` (sixteen_armadillos.rs:132-163)
    // ON THE DEVICE (bvht_scene_set_transforms): inverses, world bounds and the TLAS are computed there -- from the models'
    // CURRENT root boxes, i.e. after any pending refit -- and copied back into the host scene, which ends up in exactly the
    // state the host loop would have produced.  The next evaluate() sends nothing but the camera.
    void update_transforms(Scene& scene, const std::vector<Transform3>& transforms) {
        const size_t n = scene.objects().size();
        if (transforms.size() != n) throw std::runtime_error("update_transforms: one transform per scene object");
        std::vector<uint32_t> blas_ids = sync_models(scene);
        std::vector<float> fwd(n * 16);
        for (size_t i = 0; i < n; ++i) std::memcpy(&fwd[i * 16], transforms[i].matrix.m, 64);
        check(bvht_scene_set_transforms(ctx_, fwd.data(), blas_ids.data(), (uint32_t)n));
        instances_.resize(n);
        std::vector<float> bounds(n * 6);
        Tlas& tlas = scene.tlas_mut();
        if (tlas.nodes.size() < 2 * n) tlas.nodes.resize(2 * n);
        uint32_t used = 0;
        check(bvht_tlas_read(ctx_, (bvht_tlas_node*)tlas.nodes.data(), (uint32_t)tlas.nodes.size(), &used, instances_.data(), bounds.data(),
                             (uint32_t)n, nullptr));
        tlas.nodes_used = used;
        for (size_t i = 0; i < n; ++i) {
            Transform3 inv; std::memcpy(inv.matrix.m, instances_[i].transform_inv, 64);
            Aabb b; b.bounds_min = Vector3(bounds[i * 6 + 0], bounds[i * 6 + 1], bounds[i * 6 + 2]);
            b.bounds_max = Vector3(bounds[i * 6 + 3], bounds[i * 6 + 4], bounds[i * 6 + 5]);
            scene.get_mut_unchecked(i).adopt(transforms[i], inv, b);
        }
        scene_ = &scene; scene_version_ = scene.version();
    }

    size_t evaluate(RendererState& state, Scene& scene) override { return submit(state, scene, false); }

    // Two frames in flight (bvht_render_frame_begin / _end): evaluate_begin queues the frame into `state` and returns;
    // evaluate_end waits for the oldest begun frame, whose state then holds its pixels.  Alternate two RendererStates: the
    // copies of one frame run under the tracing of the next.  `state` must stay alive and untouched until its evaluate_end.
    size_t evaluate_begin(RendererState& state, Scene& scene) { return submit(state, scene, true); }
    void evaluate_end() { check(bvht_render_frame_end(ctx_)); }

    // Scene::intersect(&Ray) (scene.rs:32-34) for a batch of rays, on the device
    void intersect(Scene& scene, const bvht_ray* rays, uint64_t n, bvht_hit* out) {
        sync_scene(scene);
        check(bvht_trace_rays(ctx_, rays, n, out));
    }
private:
    size_t submit(RendererState& state, Scene& scene, bool in_flight) {
        sync_scene(scene);
        state.bind(ctx_);
        bvht_camera cam = scene.active_camera().to_ffi();
        bvht_rect region = { 0, 0, (uint32_t)state.width(), (uint32_t)state.height() };
        bvht_shade_params shade = state.shading().params;
        if (shade.kind == BVHT_SHADE_NORMAL && !scene.objects().empty())     // scene.get_unchecked(0).get_transform(), renderer.rs:275-278
            std::memcpy(shade.object0_transform, scene.objects()[0].get_transform().matrix.m, 64);
        check((in_flight ? bvht_render_frame_begin : bvht_render_frame)(ctx_, &cam, (uint32_t)state.width(), (uint32_t)state.height(), tile_,
                                                                       region, &shade, state.frame_mut(),
                                                                       state.keep_hits() ? state.hits_mut() : nullptr));
        return state.width() * state.height();
    }
    struct Uploaded { Model* model; uint32_t blas_id; uint64_t version; };
    void upload_attributes(const Model& m, uint32_t blas_id) {
        if (m.normals().size() == m.primitives().size())
            check(bvht_blas_set_normals(ctx_, blas_id, (const float*)m.normals().data(), (uint32_t)m.normals().size()));
        if (m.tex_coords().size() == m.primitives().size())
            check(bvht_blas_set_tex_coords(ctx_, blas_id, (const float*)m.tex_coords().data(), (uint32_t)m.tex_coords().size()));
        if (!m.texture().empty())
            check(bvht_blas_set_texture(ctx_, blas_id, m.texture().rgb.data(), m.texture().width, m.texture().height));
    }
    Uploaded* find(Model* m) { for (auto& u : uploaded_) if (u.model == m) return &u; return nullptr; }
    void check(int rc) { if (rc != BVHT_OK) throw std::runtime_error(std::string(bvht_status_string(rc)) + ": " + bvht_last_error(ctx_)); }
    bvht_ctx* ctx_ = nullptr;
    uint32_t tile_;
    std::vector<Uploaded> uploaded_;
    std::vector<bvht_instance> instances_;
    Scene* scene_ = nullptr;
    uint64_t scene_version_ = 0;
};

// renderer.rs:388-400
class Renderer {
public:
    explicit Renderer(std::unique_ptr<Integrator> pipeline) : integrator_(std::move(pipeline)) {}
    size_t render(RendererState& state, Scene& scene) { return integrator_->evaluate(state, scene); }
    // render() split in two for integrators that keep frames in flight (CudaPathTracer): see evaluate_begin / evaluate_end
    size_t render_begin(RendererState& state, Scene& scene) { return cuda().evaluate_begin(state, scene); }
    void render_end() { cuda().evaluate_end(); }
    Integrator* integrator() { return integrator_.get(); }
private:
    CudaPathTracer& cuda() {
        auto* c = dynamic_cast<CudaPathTracer*>(integrator_.get());
        if (!c) throw std::runtime_error("render_begin / render_end: the integrator keeps no frames in flight");
        return *c;
    }
    std::unique_ptr<Integrator> integrator_;
};

} // namespace bvhtracer
