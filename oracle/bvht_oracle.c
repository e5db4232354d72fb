/*
 * bvht_oracle.c -- CPU ORACLE (test infrastructure; see bvht_oracle.h header note).
 *
 * Plain-C restatement of the reference's primary closest-hit path.  Every
 * function cites the reference file:line (paths relative to the reference
 * repository root, bvhtracer/src/...) it follows.  Scalar IEEE f32, no FMA
 * contraction (-ffp-contract=off), IEEE div/sqrt: the arithmetic model of
 * rustc on x86-64.
 */
#include "bvht_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* f32::min / f32::max: if one operand is NaN the other is returned (== C fminf/fmaxf). */
static inline float fmin_(float a, float b) { return fminf(a, b); }
static inline float fmax_(float a, float b) { return fmaxf(a, b); }

/* ------------------------------------------------------------------------------------------
 * cglinalg Vector3 arithmetic (un-vendored dependency; see header for pinned/unpinned status)
 * ------------------------------------------------------------------------------------------ */
static inline float dot3(const float a[3], const float b[3]) {
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
static inline void cross3(const float a[3], const float b[3], float o[3]) {
    float x = a[1] * b[2] - a[2] * b[1];
    float y = a[2] * b[0] - a[0] * b[2];
    float z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline void sub3(const float a[3], const float b[3], float o[3]) {
    o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2];
}
void orc_vec3_cross(const float a[3], const float b[3], float out[3]) { cross3(a, b, out); }

/* Pinned by bvhtracer/tests/test_tri_mesh.rs:57-59: component-wise divide by the magnitude. */
void orc_vec3_normalize(const float v[3], float out[3]) {
    float m = sqrtf(dot3(v, v));
    out[0] = v[0] / m; out[1] = v[1] / m; out[2] = v[2] / m;
}

/* geometry/triangle.rs:33-39 */
void orc_triangle_centroid(const float tri[9], float out[3]) {
    const float one = 1.0f;
    const float three = one + one + one;
    const float one_third = one / three;
    for (int k = 0; k < 3; ++k) out[k] = ((tri[k] + tri[3 + k]) + tri[6 + k]) * one_third;
}
static inline float centroid_axis(const float* tri, int axis) {
    const float one_third = 1.0f / 3.0f;
    return ((tri[axis] + tri[3 + axis]) + tri[6 + axis]) * one_third;
}

/* query/ray.rs:23-31  Ray::new : three IEEE divides, zero components give +-inf */
void orc_ray_new(const float o[3], const float d[3], float t, orc_ray* out) {
    for (int k = 0; k < 3; ++k) { out->o[k] = o[k]; out->d[k] = d[k]; out->rd[k] = 1.0f / d[k]; }
    out->t = t;
}

/* geometry/aabb.rs:65-84  Aabb::intersect (slab test, returns entry distance) */
static inline int aabb_intersect(const float bmin[3], const float bmax[3], const orc_ray* ray, float* t_out) {
    float t_x1 = (bmin[0] - ray->o[0]) * ray->rd[0];
    float t_x2 = (bmax[0] - ray->o[0]) * ray->rd[0];
    float t_min = fmin_(t_x1, t_x2);
    float t_max = fmax_(t_x1, t_x2);
    float t_y1 = (bmin[1] - ray->o[1]) * ray->rd[1];
    float t_y2 = (bmax[1] - ray->o[1]) * ray->rd[1];
    t_min = fmax_(t_min, fmin_(t_y1, t_y2));
    t_max = fmin_(t_max, fmax_(t_y1, t_y2));
    float t_z1 = (bmin[2] - ray->o[2]) * ray->rd[2];
    float t_z2 = (bmax[2] - ray->o[2]) * ray->rd[2];
    t_min = fmax_(t_min, fmin_(t_z1, t_z2));
    t_max = fmin_(t_max, fmax_(t_z1, t_z2));
    if ((t_max >= t_min) && (t_min < ray->t) && (t_max > 0.0f)) { *t_out = t_min; return 1; }
    return 0;
}
int orc_aabb_intersect(const orc_aabb* box, const orc_ray* ray, float* t_out) {
    return aabb_intersect(box->min, box->max, ray, t_out);
}

/* geometry/triangle.rs:41-72  Triangle::intersect (Moeller-Trumbore, no culling).
 * stage (optional): 0 = left at |area|, 1 = at u, 2 = at v/u+v, 3 = reached t. */
static inline int triangle_intersect(const float* tri, const orc_ray* ray, float tuv[3], int* stage) {
    const float threshold = (float)0.0001; /* num_traits::cast(0.0001_f64) */
    float edge1[3], edge2[3], normal[3], s[3], q[3];
    sub3(tri + 3, tri, edge1);
    sub3(tri + 6, tri, edge2);
    cross3(ray->d, edge2, normal);
    float area = dot3(edge1, normal);
    if (fabsf(area) < threshold) { if (stage) *stage = 0; return 0; }
    float f = 1.0f / area;
    sub3(ray->o, tri, s);
    float u = f * dot3(s, normal);
    if (u < 0.0f || u > 1.0f) { if (stage) *stage = 1; return 0; }
    cross3(s, edge1, q);
    float v = f * dot3(ray->d, q);
    if (v < 0.0f || u + v > 1.0f) { if (stage) *stage = 2; return 0; }
    float t = f * dot3(edge2, q);
    if (stage) *stage = 3;
    if (t > threshold) {
        tuv[0] = fmin_(ray->t, t); tuv[1] = u; tuv[2] = v;
        return 1;
    }
    return 0;
}
int orc_triangle_intersect(const float tri[9], const orc_ray* ray, float tuv_out[3]) {
    return triangle_intersect(tri, ray, tuv_out, NULL);
}

/* ------------------------------------------------------------------------------------------
 * asset decode
 * ------------------------------------------------------------------------------------------ */
void orc_free(void* p) { free(p); }

typedef struct { float* v; size_t n, cap; } fvec;
static int fvec_push(fvec* a, float x) {
    if (a->n == a->cap) {
        size_t nc = a->cap ? a->cap * 2 : 4096;
        float* nv = (float*)realloc(a->v, nc * sizeof(float));
        if (!nv) return -1;
        a->v = nv; a->cap = nc;
    }
    a->v[a->n++] = x;
    return 0;
}

/* tri_loader/src/lexer.rs:12-24, 62-103 and loader.rs:119-180.
 * whitespace = ' ', '\\', '\t'; newline = '\n', '\r'; '#' starts a comment when it begins a token.
 * Nine f32 per triangle (str::parse::<f32> is correctly rounded == strtof); every triangle is kept,
 * including the 999-sentinel that ends the shipped assets (tri_loader/tests/test_lib.rs:9-15). */
int64_t orc_parse_tri(const char* text, size_t len, float** out) {
    fvec a = {0, 0, 0};
    size_t i = 0;
    char buf[128];
    while (i < len) {
        char ch = text[i];
        if (ch == ' ' || ch == '\\' || ch == '\t' || ch == '\n' || ch == '\r') { ++i; continue; }
        if (ch == '#') { while (i < len && text[i] != '\n' && text[i] != '\r') ++i; continue; }
        size_t s = i;
        while (i < len) {
            char c = text[i];
            if (c == ' ' || c == '\\' || c == '\t' || c == '\n' || c == '\r') break;
            ++i;
        }
        size_t tl = i - s;
        if (tl >= sizeof buf) { free(a.v); return -2; }
        memcpy(buf, text + s, tl); buf[tl] = 0;
        char* end = NULL;
        float val = strtof(buf, &end);
        if (end == buf || *end != 0) { free(a.v); return -3; }
        if (fvec_push(&a, val)) { free(a.v); return -4; }
    }
    if (a.n % 9 != 0) { free(a.v); return -5; }
    *out = a.v;
    return (int64_t)(a.n / 9);
}

/* mesh/decoders.rs:157-215 over cgwavefront_obj 1.0.4 (un-vendored): positions are parsed as f64 and
 * narrowed with `as f32`; only Face elements of the first object are used; polygons fan-triangulated. */
int64_t orc_parse_obj(const char* text, size_t len, float** out) {
    fvec pos = {0, 0, 0};
    fvec tri = {0, 0, 0};
    size_t i = 0;
    int objects_seen = 0;
    while (i < len) {
        size_t ls = i;
        while (i < len && text[i] != '\n') ++i;
        size_t le = i; if (i < len) ++i;
        while (ls < le && (text[ls] == ' ' || text[ls] == '\t')) ++ls;
        if (ls >= le) continue;
        char line[512];
        size_t ll = le - ls; if (ll >= sizeof line) ll = sizeof line - 1;
        memcpy(line, text + ls, ll); line[ll] = 0;
        if (line[0] == 'o' && (line[1] == ' ' || line[1] == '\t')) {
            if (++objects_seen > 1 && tri.n > 0) break;
            continue;
        }
        if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
            char* p = line + 1;
            for (int k = 0; k < 3; ++k) {
                char* e = NULL;
                double d = strtod(p, &e);
                if (e == p) { free(pos.v); free(tri.v); return -3; }
                p = e;
                if (fvec_push(&pos, (float)d)) { free(pos.v); free(tri.v); return -4; }
            }
        } else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
            long idx[64]; int nv = 0;
            char* p = line + 1;
            while (*p && nv < 64) {
                while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
                if (!*p) break;
                char* e = NULL;
                long vi = strtol(p, &e, 10);
                if (e == p) break;
                long nverts = (long)(pos.n / 3);
                if (vi < 0) vi = nverts + vi + 1;
                idx[nv++] = vi - 1;
                p = e;
                while (*p && *p != ' ' && *p != '\t') ++p; /* skip /vt/vn */
            }
            for (int k = 1; k + 1 < nv; ++k) {
                long tri_idx[3] = { idx[0], idx[k], idx[k + 1] };
                for (int c = 0; c < 3; ++c) {
                    long vi = tri_idx[c];
                    if (vi < 0 || (size_t)vi * 3 + 2 >= pos.n) { free(pos.v); free(tri.v); return -6; }
                    for (int d = 0; d < 3; ++d)
                        if (fvec_push(&tri, pos.v[vi * 3 + d])) { free(pos.v); free(tri.v); return -4; }
                }
            }
        }
    }
    free(pos.v);
    *out = tri.v;
    return (int64_t)(tri.n / 9);
}

int64_t orc_load_mesh_file(const char* path, float** out) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* buf = (char*)malloc((size_t)sz + 1);
    if (!buf) { fclose(f); return -4; }
    size_t rd = fread(buf, 1, (size_t)sz, f);
    fclose(f);
    buf[rd] = 0;
    size_t pl = strlen(path);
    int64_t n;
    if (pl >= 4 && strcmp(path + pl - 4, ".obj") == 0) n = orc_parse_obj(buf, rd, out);
    else n = orc_parse_tri(buf, rd, out);
    free(buf);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * BLAS build (model/bvh.rs:317-467, 524-541)
 * ------------------------------------------------------------------------------------------ */
/* geometry/aabb.rs:36-39 */
static inline void aabb_grow(orc_aabb* b, const float p[3]) {
    for (int k = 0; k < 3; ++k) { b->min[k] = fmin_(b->min[k], p[k]); b->max[k] = fmax_(b->max[k], p[k]); }
}
/* geometry/aabb.rs:41-46 */
static inline void aabb_grow_aabb(orc_aabb* b, const orc_aabb* o) {
    if (o->min[0] != FLT_MAX) { aabb_grow(b, o->min); aabb_grow(b, o->max); }
}
/* geometry/aabb.rs:48-57 */
static inline float aabb_area(const float bmin[3], const float bmax[3]) {
    float ex = bmax[0] - bmin[0], ey = bmax[1] - bmin[1], ez = bmax[2] - bmin[2];
    return (ex * ey + ey * ez) + ez * ex;
}

/* model/bvh.rs:317-330 update_node_bounds (node_indices is the identity: bvh.rs:526-528) */
static void update_node_bounds(const float* tris, orc_bvh_node* node) {
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
    float mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    uint32_t first = node->left_first;
    for (uint32_t i = 0; i < node->prim_count; ++i) {
        const float* t = tris + (size_t)(first + i) * 9;
        for (int vtx = 0; vtx < 3; ++vtx)
            for (int k = 0; k < 3; ++k) mn[k] = fmin_(mn[k], t[vtx * 3 + k]);
        for (int vtx = 0; vtx < 3; ++vtx)
            for (int k = 0; k < 3; ++k) mx[k] = fmax_(mx[k], t[vtx * 3 + k]);
    }
    memcpy(node->min, mn, sizeof mn);
    memcpy(node->max, mx, sizeof mx);
}

#define ORC_BINS 8

/* model/bvh.rs:333-394 find_best_split_plane.  Quirks kept on purpose:
 *  - bounds_max starts at 1e-30 (not -1e30)                      bvh.rs:340
 *  - bins and the sweep boxes start as Aabb::default() = the point box at the origin
 *    (bvh.rs:47-51, 349, 365-366; aabb.rs:12), NOT new_empty() -- this is what makes the
 *    reference's trees shallow with giant leaves
 *  - bin index = float -> usize saturating cast, then min(7, .)  bvh.rs:352-353 */
static void find_best_split_plane(const float* tris, const orc_bvh_node* node,
                                  int* best_axis_o, float* best_pos_o, float* best_cost_o) {
    int best_axis = -1;
    float best_position = 0.0f;
    float best_cost = FLT_MAX;
    uint32_t first = node->left_first, count = node->prim_count;
    for (int axis = 0; axis < 3; ++axis) {
        float bounds_min = 1e30f;
        float bounds_max = 1e-30f;
        for (uint32_t i = 0; i < count; ++i) {
            float c = centroid_axis(tris + (size_t)(first + i) * 9, axis);
            bounds_min = fmin_(bounds_min, c);
            bounds_max = fmax_(bounds_max, c);
        }
        if (bounds_min == bounds_max) continue;

        orc_aabb bin_box[ORC_BINS];
        uint32_t bin_count[ORC_BINS];
        memset(bin_box, 0, sizeof bin_box);
        memset(bin_count, 0, sizeof bin_count);
        float bin_scale = (float)ORC_BINS / (bounds_max - bounds_min);
        for (uint32_t i = 0; i < count; ++i) {
            const float* t = tris + (size_t)(first + i) * 9;
            float f = (centroid_axis(t, axis) - bounds_min) * bin_scale;
            int idx;
            if (f >= (float)ORC_BINS) idx = ORC_BINS - 1;       /* saturating `as usize` then min(7,.) */
            else if (f > 0.0f) idx = (int)f;                    /* truncation toward zero */
            else idx = 0;                                       /* negative and NaN -> 0 */
            bin_count[idx] += 1;
            aabb_grow(&bin_box[idx], t);
            aabb_grow(&bin_box[idx], t + 3);
            aabb_grow(&bin_box[idx], t + 6);
        }

        float left_area[ORC_BINS - 1], right_area[ORC_BINS - 1];
        uint32_t left_count[ORC_BINS - 1], right_count[ORC_BINS - 1];
        orc_aabb left_box, right_box;
        memset(&left_box, 0, sizeof left_box);
        memset(&right_box, 0, sizeof right_box);
        uint32_t left_sum = 0, right_sum = 0;
        for (int i = 0; i < ORC_BINS - 1; ++i) {
            left_sum += bin_count[i];
            left_count[i] = left_sum;
            aabb_grow_aabb(&left_box, &bin_box[i]);
            left_area[i] = aabb_area(left_box.min, left_box.max);

            right_sum += bin_count[ORC_BINS - 1 - i];
            right_count[ORC_BINS - 2 - i] = right_sum;
            aabb_grow_aabb(&right_box, &bin_box[ORC_BINS - 1 - i]);
            right_area[ORC_BINS - 2 - i] = aabb_area(right_box.min, right_box.max);
        }

        float scale = (bounds_max - bounds_min) / (float)ORC_BINS;
        for (int i = 0; i < ORC_BINS - 1; ++i) {
            float plane_cost = (float)left_count[i] * left_area[i] + (float)right_count[i] * right_area[i];
            if (plane_cost < best_cost) {
                best_axis = axis;
                best_position = bounds_min + scale * (float)(i + 1);
                best_cost = plane_cost;
            }
        }
    }
    *best_axis_o = best_axis; *best_pos_o = best_position; *best_cost_o = best_cost;
}

static void swap_tri(float* a, float* b) {
    float tmp[9];
    memcpy(tmp, a, sizeof tmp); memcpy(a, b, sizeof tmp); memcpy(b, tmp, sizeof tmp);
}

/* model/bvh.rs:396-467 subdivide (recursion order: left then right) */
static void subdivide(float* tris, orc_bvh_node* nodes, uint32_t* nodes_used, uint32_t node_index) {
    int best_axis; float best_position, best_cost;
    find_best_split_plane(tris, &nodes[node_index], &best_axis, &best_position, &best_cost);

    orc_bvh_node* node = &nodes[node_index];
    float no_split_cost = (float)node->prim_count * aabb_area(node->min, node->max);
    if (best_cost >= no_split_cost) return;
    if (best_axis < 0) return; /* unreachable for finite inputs (the reference would panic) */

    /* in-place partition (bvh.rs:419-430).  The reference uses u32 i/j; j can only step below
     * zero when first_primitive_index == 0 and everything goes right, which the reference cannot
     * survive either -- signed arithmetic turns that case into the "empty side" abort below. */
    int64_t i = node->left_first;
    int64_t j = i + (int64_t)node->prim_count - 1;
    while (i <= j) {
        if (centroid_axis(tris + (size_t)i * 9, best_axis) < best_position) {
            i += 1;
        } else {
            swap_tri(tris + (size_t)i * 9, tris + (size_t)j * 9);
            j -= 1;
        }
    }
    uint32_t left_count = (uint32_t)(i - (int64_t)node->left_first);
    if (left_count == 0 || left_count == node->prim_count) return;

    uint32_t left_child = (*nodes_used)++;
    uint32_t right_child = (*nodes_used)++;
    nodes[left_child].left_first = node->left_first;
    nodes[left_child].prim_count = left_count;
    nodes[right_child].left_first = (uint32_t)i;
    nodes[right_child].prim_count = node->prim_count - left_count;
    node->left_first = left_child;
    node->prim_count = 0;

    update_node_bounds(tris, &nodes[left_child]);
    update_node_bounds(tris, &nodes[right_child]);
    subdivide(tris, nodes, nodes_used, left_child);
    subdivide(tris, nodes, nodes_used, right_child);
}

/* model/bvh.rs:505-541 BvhBuilder::new + build_for (node 1 is the alignment dummy) */
uint32_t orc_bvh_build(float* tris, uint32_t n_tris, orc_bvh_node* nodes) {
    memset(nodes, 0, sizeof(orc_bvh_node) * 2 * (size_t)n_tris);
    uint32_t nodes_used = 2;
    if (n_tris == 0) return nodes_used;
    nodes[0].left_first = 0;
    nodes[0].prim_count = n_tris;
    update_node_bounds(tris, &nodes[0]);
    subdivide(tris, nodes, &nodes_used, 0);
    return nodes_used;
}

/* model/bvh.rs:469-493 refit: reverse index sweep, node 1 skipped */
void orc_bvh_refit(const float* tris, orc_bvh_node* nodes, uint32_t nodes_used) {
    for (int64_t ni = (int64_t)nodes_used - 1; ni >= 0; --ni) {
        if (ni == 1) continue;
        orc_bvh_node* node = &nodes[ni];
        if (node->prim_count > 0) { update_node_bounds(tris, node); continue; }
        const orc_bvh_node* l = &nodes[node->left_first];
        const orc_bvh_node* r = &nodes[node->left_first + 1];
        for (int k = 0; k < 3; ++k) {
            node->min[k] = fmin_(l->min[k], r->min[k]);
            node->max[k] = fmax_(l->max[k], r->max[k]);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * BLAS traversal (model/bvh.rs:242-305)
 * ------------------------------------------------------------------------------------------ */
#define ORC_STACK 256

static int bvh_intersect(const orc_blas* blas, const orc_ray* ray, orc_ray* closest_out, orc_hit* hit, orc_counters* c) {
    const orc_bvh_node* nodes = blas->nodes;
    const orc_bvh_node* current = &nodes[0];            /* root AABB is never tested (bvh.rs:243) */
    const orc_bvh_node* stack[ORC_STACK];
    uint32_t sp = 0;
    orc_ray closest_ray = *ray;
    int have = 0;
    float ct = 0, cu = 0, cv = 0;
    uint32_t closest_prim = 0;
    if (c) c->blas_nodes += 1;
    for (;;) {
        if (current->prim_count > 0) {
            uint32_t base = current->left_first;       /* node_indices[] is the identity (bvh.rs:236-240) */
            for (uint32_t k = 0; k < current->prim_count; ++k) {
                uint32_t pi = base + k;
                float tuv[3]; int stage;
                /* NOTE: tested against the ENTRY ray, not closest_ray (bvh.rs:251) */
                int ok = triangle_intersect(blas->tris + (size_t)pi * 9, ray, tuv, &stage);
                if (c) { if (stage == 0) c->tri_area++; else if (stage == 1) c->tri_u++; else if (stage == 2) c->tri_v++; else c->tri_t++; }
                if (ok && tuv[0] < closest_ray.t) {
                    closest_ray.t = tuv[0];
                    ct = tuv[0]; cu = tuv[1]; cv = tuv[2];
                    have = 1;
                    closest_prim = pi;
                }
            }
            if (sp > 0) current = stack[--sp]; else break;
        } else {
            const orc_bvh_node* left = &nodes[current->left_first];
            const orc_bvh_node* right = &nodes[current->left_first + 1];
            float ld = 0, rdist = 0;
            int lh = aabb_intersect(left->min, left->max, &closest_ray, &ld);
            int rh = aabb_intersect(right->min, right->max, &closest_ray, &rdist);
            if (c) { c->blas_nodes += 2; c->box_tests += 2; }
            const orc_bvh_node *near_n, *far_n; int near_h, far_h;
            if ((lh ? ld : FLT_MAX) < (rh ? rdist : FLT_MAX)) { near_n = left; near_h = lh; far_n = right; far_h = rh; }
            else { near_n = right; near_h = rh; far_n = left; far_h = lh; }   /* ties and double-miss: right first */
            if (near_h) {
                current = near_n;
                if (far_h) {
                    if (sp >= ORC_STACK) { fprintf(stderr, "oracle: BLAS stack overflow\n"); abort(); }
                    stack[sp++] = far_n;
                    if (c && sp > c->max_blas_stack) c->max_blas_stack = sp;
                }
                continue;
            }
            if (sp > 0) current = stack[--sp]; else break;
        }
    }
    if ((closest_ray.t < FLT_MAX) && have) {
        hit->t = ct; hit->u = cu; hit->v = cv;
        hit->id = closest_prim & 0x000FFFFFu;   /* InstancePrimitiveIndex::from_primitive: instance bits 0 (intersection.rs:64-66) */
        if (closest_out) *closest_out = closest_ray;
        return 1;
    }
    return 0;
}

int orc_bvh_intersect(const orc_blas* blas, const orc_ray* ray, orc_hit* hit, orc_counters* c) {
    orc_ray cr;
    int ok = bvh_intersect(blas, ray, &cr, hit, c);
    if (!ok) { hit->t = FLT_MAX; hit->u = 0; hit->v = 0; hit->id = 0xFFFFFFFFu; }
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * transforms
 * ------------------------------------------------------------------------------------------ */
void orc_mat4_identity(float m[16]) {
    memset(m, 0, 16 * sizeof(float));
    m[0] = m[5] = m[10] = m[15] = 1.0f;
}

/* column-major m[c*4+r]; cglinalg Matrix4x4 * Vector4 (order of the sum: parity unpinned) */
void orc_mat4_mul_vec4(const float m[16], const float v[4], float out[4]) {
    float r[4];
    for (int row = 0; row < 4; ++row)
        r[row] = ((m[0 + row] * v[0] + m[4 + row] * v[1]) + m[8 + row] * v[2]) + m[12 + row] * v[3];
    memcpy(out, r, sizeof r);
}
/* transform.rs:219-223 */
void orc_transform_point(const float m[16], const float p[3], float out[3]) {
    float v[4] = { p[0], p[1], p[2], 1.0f }, r[4];
    orc_mat4_mul_vec4(m, v, r);
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
/* transform.rs:230-234 */
void orc_transform_vector(const float m[16], const float v3[3], float out[3]) {
    float v[4] = { v3[0], v3[1], v3[2], 0.0f }, r[4];
    orc_mat4_mul_vec4(m, v, r);
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
/* cglinalg Matrix4x4 * Matrix4x4: column c of the product = a * (column c of b) */
void orc_mat4_mul(const float a[16], const float b[16], float out[16]) {
    float r[16];
    for (int c = 0; c < 4; ++c) orc_mat4_mul_vec4(a, b + 4 * c, r + 4 * c);
    memcpy(out, r, sizeof r);
}

/* cglinalg Matrix4x4::inverse (parity unpinned): adjugate over determinant, cofactors by 2x2 minors. */
int orc_mat4_inverse(const float m[16], float out[16]) {
    /* a(r,c) = m[c*4+r] */
#define A(r, c) m[(c) * 4 + (r)]
    float s0 = A(0,0) * A(1,1) - A(1,0) * A(0,1);
    float s1 = A(0,0) * A(1,2) - A(1,0) * A(0,2);
    float s2 = A(0,0) * A(1,3) - A(1,0) * A(0,3);
    float s3 = A(0,1) * A(1,2) - A(1,1) * A(0,2);
    float s4 = A(0,1) * A(1,3) - A(1,1) * A(0,3);
    float s5 = A(0,2) * A(1,3) - A(1,2) * A(0,3);
    float c5 = A(2,2) * A(3,3) - A(3,2) * A(2,3);
    float c4 = A(2,1) * A(3,3) - A(3,1) * A(2,3);
    float c3 = A(2,1) * A(3,2) - A(3,1) * A(2,2);
    float c2 = A(2,0) * A(3,3) - A(3,0) * A(2,3);
    float c1 = A(2,0) * A(3,2) - A(3,0) * A(2,2);
    float c0 = A(2,0) * A(3,1) - A(3,0) * A(2,1);
    float det = ((((s0 * c5 - s1 * c4) + s2 * c3) + s3 * c2) - s4 * c1) + s5 * c0;
    if (det == 0.0f) return 0;
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
#undef A
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[c * 4 + r] = b[r][c];
    return 1;
}

/* cglinalg Rotation3::from_angle_x/z, Rotation3 * Rotation3, to_affine_matrix; then Transform3::new
 * (transform.rs:24-43): translation into column 3, column c scaled by scale[c]. */
void orc_transform_new_rot_xz(const float scale[3], const float trans[3], float angle_x, float angle_z, float out[16]) {
    float cx = cosf(angle_x), sx = sinf(angle_x);
    float cz = cosf(angle_z), sz = sinf(angle_z);
    /* row-major 3x3 */
    float rx[3][3] = { {1, 0, 0}, {0, cx, -sx}, {0, sx, cx} };
    float rz[3][3] = { {cz, -sz, 0}, {sz, cz, 0}, {0, 0, 1} };
    float r[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r[i][j] = (rx[i][0] * rz[0][j] + rx[i][1] * rz[1][j]) + rx[i][2] * rz[2][j];
    orc_mat4_identity(out);
    for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) out[c * 4 + rr] = r[rr][c];
    out[12] = trans[0]; out[13] = trans[1]; out[14] = trans[2];
    for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) out[c * 4 + rr] *= scale[c];
}
/* transform.rs:83-96 */
void orc_transform_from_scale_translation(const float scale[3], const float trans[3], float out[16]) {
    orc_mat4_identity(out);
    out[12] = trans[0]; out[13] = trans[1]; out[14] = trans[2];
    out[0] = scale[0]; out[5] = scale[1]; out[10] = scale[2];
}

/* scene_object.rs:60-75 (= :118-131): Aabb::new_empty grown by the 8 transformed corners */
void orc_instance_bounds(const float m[16], const orc_aabb* ob, orc_aabb* out) {
    orc_aabb nb;
    for (int k = 0; k < 3; ++k) { nb.min[k] = FLT_MAX; nb.max[k] = -FLT_MAX; }
    for (int i = 0; i < 8; ++i) {
        float p[3] = {
            (i & 1) ? ob->max[0] : ob->min[0],
            (i & 2) ? ob->max[1] : ob->min[1],
            (i & 4) ? ob->max[2] : ob->min[2],
        };
        float q[3];
        orc_transform_point(m, p, q);
        aabb_grow(&nb, q);
    }
    *out = nb;
}

/* ------------------------------------------------------------------------------------------
 * TLAS build (scene/tlas.rs:179-250, 253-280)
 * ------------------------------------------------------------------------------------------ */
static int find_best_match(const orc_tlas_node* nodes, const int32_t* list, int32_t n, int32_t a) {
    float smallest = FLT_MAX;
    int32_t best_b = -1;
    for (int32_t b = 0; b < n; ++b) {
        if (b == a) continue;
        const orc_tlas_node* na = &nodes[list[a]];
        const orc_tlas_node* nb = &nodes[list[b]];
        float ex[3];
        for (int k = 0; k < 3; ++k) ex[k] = fmax_(na->max[k], nb->max[k]) - fmin_(na->min[k], nb->min[k]);
        float area = (ex[0] * ex[1] + ex[1] * ex[2]) + ex[2] * ex[0];
        if (area < smallest) { smallest = area; best_b = b; }
    }
    return best_b;
}

uint32_t orc_tlas_build(const orc_aabb* bounds, uint32_t n, orc_tlas_node* nodes) {
    /* TlasBuilder::build_for: 2n default nodes (aabb = new_empty) */
    for (uint32_t i = 0; i < 2 * n; ++i) {
        for (int k = 0; k < 3; ++k) { nodes[i].min[k] = FLT_MAX; nodes[i].max[k] = -FLT_MAX; }
        nodes[i].left_right = 0; nodes[i].blas = 0;
    }
    if (n == 0) return 2;
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * n);
    int32_t count = (int32_t)n;
    int32_t nodes_used = 1;
    for (uint32_t i = 0; i < n; ++i) {
        idx[i] = nodes_used;
        memcpy(nodes[nodes_used].min, bounds[i].min, sizeof(float) * 3);
        memcpy(nodes[nodes_used].max, bounds[i].max, sizeof(float) * 3);
        nodes[nodes_used].blas = i;
        nodes[nodes_used].left_right = 0;
        nodes_used += 1;
    }
    int32_t a = 0;
    int32_t b = find_best_match(nodes, idx, count, a);
    while (count > 1) {
        int32_t c = find_best_match(nodes, idx, count, b);
        if (a == c) {
            int32_t ia = idx[a], ib = idx[b];
            orc_tlas_node na = nodes[ia], nb = nodes[ib];
            orc_tlas_node* nn = &nodes[nodes_used];
            /* LeftRightIndex::new(left=ia, right=ib) stores left + (right << 16)  (tlas.rs:18-22);
             * the accessors read it back swapped (tlas.rs:25-32) -- handled in the traversal. */
            nn->left_right = (uint32_t)ia + ((uint32_t)ib << 16);
            for (int k = 0; k < 3; ++k) {
                nn->min[k] = fmin_(na.min[k], nb.min[k]);
                nn->max[k] = fmax_(na.max[k], nb.max[k]);
            }
            idx[a] = nodes_used;
            nodes_used += 1;
            idx[b] = idx[count - 1];
            count -= 1;
            b = find_best_match(nodes, idx, count, a);
        } else {
            a = b;
            b = c;
        }
    }
    nodes[0] = nodes[idx[a]];
    free(idx);
    return (uint32_t)nodes_used;
}

/* ------------------------------------------------------------------------------------------
 * camera
 * ------------------------------------------------------------------------------------------ */
/* camera.rs:809-835 CameraAttitude::from_spec: view = rotation * translation(-pos); inverse cached */
static void camera_attitude(const float pos[3], const float fwd[3], const float right[3], const float up[3],
                            float view_inv[16]) {
    float tr[16], rot[16], view[16];
    orc_mat4_identity(tr);
    tr[12] = -pos[0]; tr[13] = -pos[1]; tr[14] = -pos[2];
    /* Matrix4x4::new takes columns: (right.x, up.x, -fwd.x, 0), (right.y, up.y, -fwd.y, 0), ... */
    float r[16] = {
        right[0], up[0], -fwd[0], 0.0f,
        right[1], up[1], -fwd[1], 0.0f,
        right[2], up[2], -fwd[2], 0.0f,
        0.0f,     0.0f,  0.0f,    1.0f };
    memcpy(rot, r, sizeof r);
    orc_mat4_mul(rot, tr, view);
    if (!orc_mat4_inverse(view, view_inv)) orc_mat4_identity(view_inv);
}

static void camera_corners(float left, float top, float ext_x, float ext_y, float near_, orc_camera* out) {
    /* camera.rs:199-211 */
    out->tl[0] = left;         out->tl[1] = top;         out->tl[2] = -near_;
    out->tr[0] = left + ext_x; out->tr[1] = top;         out->tr[2] = -near_;
    out->bl[0] = left;         out->bl[1] = top - ext_y; out->bl[2] = -near_;
}

/* camera.rs:223-251 From<SymmetricFovSpec> for Frustum.  Degrees::tan (cglinalg, un-vendored) is taken as
 * tanf(deg * (pi/180)) in f32: parity unpinned; for the examples' 90 degrees it yields tan = 1.0. */
void orc_camera_symmetric_fov(float fovy_deg, float aspect, float near_,
                              const float pos[3], const float fwd[3], const float right[3], const float up[3],
                              orc_camera* out) {
    float fovy_over_two = fovy_deg / 2.0f;
    float tan_half = tanf(fovy_over_two * ((float)M_PI / 180.0f));
    float top = near_ * tan_half;
    float bottom = -top;
    float left = -aspect * top;
    float right_ = aspect * top;
    camera_corners(left, top, right_ - left, top - bottom, near_, out);
    camera_attitude(pos, fwd, right, up, out->view_inv);
}

/* camera.rs:343-367 From<BoxSpec> for Frustum */
void orc_camera_box(float left, float right_, float bottom, float top, float near_,
                    const float pos[3], const float fwd[3], const float right[3], const float up[3],
                    orc_camera* out) {
    camera_corners(left, top, right_ - left, top - bottom, near_, out);
    camera_attitude(pos, fwd, right, up, out->view_inv);
}

/* camera.rs:994-1010 get_ray_eye + get_ray_world */
void orc_camera_ray_world(const orc_camera* cam, float u, float v, orc_ray* out) {
    float pix[3], dir[3];
    for (int k = 0; k < 3; ++k) {
        float origin = 0.0f;
        pix[k] = ((origin + cam->tl[k]) + (cam->tr[k] - cam->tl[k]) * u) + (cam->bl[k] - cam->tl[k]) * v;
        pix[k] = pix[k] - origin;
    }
    orc_vec3_normalize(pix, dir);
    float o4[4] = { 0.0f, 0.0f, 0.0f, 1.0f }, d4[4] = { dir[0], dir[1], dir[2], 0.0f }, ow[4], dw[4];
    orc_mat4_mul_vec4(cam->view_inv, o4, ow);
    orc_mat4_mul_vec4(cam->view_inv, d4, dw);
    orc_ray_new(ow, dw, FLT_MAX, out);
}

/* ------------------------------------------------------------------------------------------
 * scene traversal
 * ------------------------------------------------------------------------------------------ */
/* scene_object.rs:78-89 SceneObject::intersect */
static int instance_intersect(const orc_scene* s, uint32_t inst_index, const orc_ray* ray, orc_ray* model_closest,
                              orc_hit* hit, orc_counters* c) {
    const orc_instance* inst = &s->inst[inst_index];
    float o[3], d[3];
    orc_transform_point(inst->inv, ray->o, o);
    orc_transform_vector(inst->inv, ray->d, d);
    orc_ray mray;
    orc_ray_new(o, d, ray->t, &mray);
    if (c) c->inst += 1;
    return bvh_intersect(&s->blas[inst->blas_id], &mray, model_closest, hit, c);
}

/* scene/tlas.rs:123-177 Tlas::intersect.  left_blas() = upper 16 bits, right_blas() = lower 16 bits
 * (tlas.rs:25-32, 70-77) although new(l, r) stored l in the LOWER half: kept as is. */
int orc_scene_intersect(const orc_scene* s, const orc_ray* ray, orc_hit* hit, orc_counters* c) {
    const orc_tlas_node* nodes = s->tlas;
    const orc_tlas_node* current = &nodes[0];
    const orc_tlas_node* stack[ORC_STACK];
    uint32_t sp = 0;
    orc_ray closest_ray = *ray;
    int have = 0;
    orc_hit best = { FLT_MAX, 0.0f, 0.0f, 0xFFFFFFFFu };
    if (c) { c->rays += 1; c->tlas_nodes += 1; }
    for (;;) {
        if (current->left_right == 0) {
            orc_hit h; orc_ray mc;
            if (instance_intersect(s, current->blas, &closest_ray, &mc, &h, c)) {
                if (mc.t < closest_ray.t) {
                    closest_ray.t = mc.t;
                    best = h;
                    have = 1;
                }
            }
            if (sp > 0) current = stack[--sp]; else break;
        } else {
            const orc_tlas_node* left = &nodes[(current->left_right & 0xFFFF0000u) >> 16];
            const orc_tlas_node* right = &nodes[current->left_right & 0x0000FFFFu];
            float ld = 0, rdist = 0;
            int lh = aabb_intersect(left->min, left->max, &closest_ray, &ld);
            int rh = aabb_intersect(right->min, right->max, &closest_ray, &rdist);
            if (c) { c->tlas_nodes += 2; c->box_tests += 2; }
            const orc_tlas_node *near_n, *far_n; int near_h, far_h;
            if ((lh ? ld : FLT_MAX) < (rh ? rdist : FLT_MAX)) { near_n = left; near_h = lh; far_n = right; far_h = rh; }
            else { near_n = right; near_h = rh; far_n = left; far_h = lh; }
            if (near_h) {
                current = near_n;
                if (far_h) {
                    if (sp >= ORC_STACK) { fprintf(stderr, "oracle: TLAS stack overflow\n"); abort(); }
                    stack[sp++] = far_n;
                    if (c && sp > c->max_tlas_stack) c->max_tlas_stack = sp;
                }
                continue;
            }
            if (sp > 0) current = stack[--sp]; else break;
        }
    }
    if ((closest_ray.t < FLT_MAX) && have) { *hit = best; if (c) c->hits += 1; return 1; }
    hit->t = FLT_MAX; hit->u = 0.0f; hit->v = 0.0f; hit->id = 0xFFFFFFFFu;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * frame loop (renderer.rs:345-368)
 * ------------------------------------------------------------------------------------------ */
static void counters_add(orc_counters* a, const orc_counters* b) {
    a->rays += b->rays; a->hits += b->hits; a->blas_nodes += b->blas_nodes; a->tlas_nodes += b->tlas_nodes;
    a->inst += b->inst; a->tri_area += b->tri_area; a->tri_u += b->tri_u; a->tri_v += b->tri_v; a->tri_t += b->tri_t;
    a->box_tests += b->box_tests;
    if (b->max_blas_stack > a->max_blas_stack) a->max_blas_stack = b->max_blas_stack;
    if (b->max_tlas_stack > a->max_tlas_stack) a->max_tlas_stack = b->max_tlas_stack;
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* A tiny work-pulling pool: the reference is single-threaded (renderer.rs:353-368), so n_threads == 1
 * is the faithful configuration; n_threads > 1 distributes chunks of tiles/rays over pthreads and is
 * OUR parallelisation, used only for the "all host cores" CPU baseline. */
typedef struct {
    const orc_scene* s; const orc_camera* cam;
    uint32_t width, height, tile, x0, y0, x1, y1, tx0, ty0;
    int64_t ntx, n_items, chunk;
    const float* odt;
    orc_hit* hits;
    int want_counters;
    int64_t next;                 /* atomic cursor */
    pthread_mutex_t lock;
    orc_counters total;
} orc_job;

static void job_run_tile(const orc_job* j, int64_t tidx, orc_counters* lc) {
    uint32_t tx = j->tx0 + (uint32_t)(tidx % j->ntx);
    uint32_t ty = j->ty0 + (uint32_t)(tidx / j->ntx);
    for (uint32_t v = 0; v < j->tile; ++v) {
        for (uint32_t u = 0; u < j->tile; ++u) {
            uint32_t px = j->tile * tx + u, py = j->tile * ty + v;
            if (px < j->x0 || px >= j->x1 || py < j->y0 || py >= j->y1) continue;
            orc_ray ray;
            /* renderer.rs:358-361: usize as f32 / usize as f32 */
            orc_camera_ray_world(j->cam, (float)px / (float)j->width, (float)py / (float)j->height, &ray);
            orc_scene_intersect(j->s, &ray, &j->hits[(size_t)py * j->width + px], lc);
        }
    }
}

static void* job_worker(void* arg) {
    orc_job* j = (orc_job*)arg;
    orc_counters local; memset(&local, 0, sizeof local);
    orc_counters* lc = j->want_counters ? &local : NULL;
    for (;;) {
        int64_t b = __atomic_fetch_add(&j->next, j->chunk, __ATOMIC_RELAXED);
        if (b >= j->n_items) break;
        int64_t e = b + j->chunk; if (e > j->n_items) e = j->n_items;
        for (int64_t i = b; i < e; ++i) {
            if (j->odt) {
                orc_ray ray;
                orc_ray_new(j->odt + i * 7, j->odt + i * 7 + 3, j->odt[i * 7 + 6], &ray);
                orc_scene_intersect(j->s, &ray, &j->hits[i], lc);
            } else {
                job_run_tile(j, i, lc);
            }
        }
    }
    if (j->want_counters) {
        pthread_mutex_lock(&j->lock);
        counters_add(&j->total, &local);
        pthread_mutex_unlock(&j->lock);
    }
    return NULL;
}

static void job_execute(orc_job* j, int n_threads, orc_counters* counters) {
    j->next = 0;
    j->want_counters = counters != NULL;
    memset(&j->total, 0, sizeof j->total);
    pthread_mutex_init(&j->lock, NULL);
    if (n_threads <= 1) {
        job_worker(j);
    } else {
        if (n_threads > 1024) n_threads = 1024;
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
        int started = 0;
        for (int i = 0; i < n_threads; ++i) if (pthread_create(&th[started], NULL, job_worker, j) == 0) ++started;
        if (started == 0) job_worker(j);
        for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&j->lock);
    if (counters) *counters = j->total;
}

void orc_render(const orc_scene* s, const orc_camera* cam, uint32_t width, uint32_t height, uint32_t tile,
                uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                orc_hit* hits, orc_counters* counters, int n_threads) {
    if (tile == 0) tile = 8;
    if (x1 > width) x1 = width;
    if (y1 > height) y1 = height;
    orc_job j; memset(&j, 0, sizeof j);
    j.s = s; j.cam = cam; j.width = width; j.height = height; j.tile = tile;
    j.x0 = x0; j.y0 = y0; j.x1 = x1; j.y1 = y1;
    j.tx0 = x0 / tile; j.ty0 = y0 / tile;
    uint32_t tx1 = (x1 + tile - 1) / tile, ty1 = (y1 + tile - 1) / tile;
    j.ntx = (int64_t)tx1 - j.tx0;
    int64_t nty = (int64_t)ty1 - j.ty0;
    j.n_items = (j.ntx > 0 && nty > 0) ? j.ntx * nty : 0;
    j.chunk = 8;
    j.odt = NULL; j.hits = hits;
    job_execute(&j, n_threads, counters);
}

void orc_trace_rays(const orc_scene* s, const float* odt, uint64_t n, orc_hit* hits, orc_counters* counters, int n_threads) {
    orc_job j; memset(&j, 0, sizeof j);
    j.s = s; j.odt = odt; j.hits = hits; j.n_items = (int64_t)n; j.chunk = 256;
    job_execute(&j, n_threads, counters);
}

/* ------------------------------------------------------------------------------------------
 * accumulators + pixel shaders (renderer.rs:116-245)
 * ------------------------------------------------------------------------------------------ */
/* Rust `f32 as i32`: saturating, NaN -> 0 */
static int32_t f32_as_i32(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}
/* Rust `f32 as u8`: saturating, NaN -> 0 */
static uint32_t f32_as_u8(float x) {
    if (x != x) return 0;
    if (x >= 255.0f) return 255;
    if (x <= 0.0f) return 0;
    return (uint32_t)x;
}

void orc_shade(uint32_t kind, float scale, float offset, uint32_t hit_rgba, uint32_t miss_rgba,
               const orc_hit* hits, uint64_t n, uint32_t* rgba_out) {
    for (uint64_t i = 0; i < n; ++i) {
        const orc_hit* h = &hits[i];
        int hit = h->id != 0xFFFFFFFFu;
        uint32_t px = 0;
        if (kind == 1) {
            /* DepthAccumulator: radiance = from_fill(t or f32::MAX) (:184-194); DepthMappingShader (:207-222) */
            float nearest_t = hit ? h->t : FLT_MAX;
            if (nearest_t < FLT_MAX) {
                uint32_t color = 255u - (uint32_t)f32_as_i32((nearest_t - offset) * scale);   /* wrapping u32 */
                uint32_t c = color * 0x010101u;
                uint32_t r = (c & 0x00FF0000u) >> 16, g = (c & 0x0000FF00u) >> 8, b = c & 0x000000FFu;
                px = r | (g << 8) | (b << 16) | 0xFF000000u;
            } else {
                px = 0xFF000000u;
            }
        } else if (kind == 2) {
            /* IntersectionAccumulator (:145-153) + IntersectionShader (:166-174), hit/miss radiance non-zero/zero */
            px = hit ? hit_rgba : miss_rgba;
        } else if (kind == 3) {
            /* UvMappingAccumulator (:233-245) + RadianceToRgbShader (:124-132) */
            float rx = hit ? h->u : 0.0f, ry = hit ? h->v : 0.0f, rz = hit ? 1.0f - (h->u + h->v) : 0.0f;
            uint32_t r = f32_as_u8(255.0f * rx), g = f32_as_u8(255.0f * ry), b = f32_as_u8(255.0f * rz);
            if (r > 255) r = 255; if (g > 255) g = 255; if (b > 255) b = 255;
            px = r | (g << 8) | (b << 16) | 0xFF000000u;
        }
        rgba_out[i] = px;
    }
}

/* mesh/decoders.rs:120-124 (TriMeshDecoder): one face normal replicated on the three vertices */
void orc_tri_normals(const float* tris, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const float* t = tris + i * 9;
        float a[3], b[3], an[3], bn[3], c[3], nn[3];
        sub3(t + 6, t, a);              /* v0v2 = (v2 - v0).normalize() */
        sub3(t + 3, t, b);              /* v0v1 = (v1 - v0).normalize() */
        orc_vec3_normalize(a, an);
        orc_vec3_normalize(b, bn);
        cross3(an, bn, c);
        orc_vec3_normalize(c, nn);
        for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) out[i * 9 + v * 3 + k] = nn[k];
    }
}

/* mesh/decoders.rs:174-203: normals[i] = vn of the face corner (f64 -> f32), zero when absent; fan-triangulated like the faces */
int64_t orc_parse_obj_normals(const char* text, size_t len, float** out) {
    fvec vn = {0, 0, 0};
    fvec res = {0, 0, 0};
    size_t i = 0;
    int objects_seen = 0;
    size_t n_pos = 0;
    while (i < len) {
        size_t ls = i;
        while (i < len && text[i] != '\n') ++i;
        size_t le = i; if (i < len) ++i;
        while (ls < le && (text[ls] == ' ' || text[ls] == '\t')) ++ls;
        if (ls >= le) continue;
        char line[512];
        size_t ll = le - ls; if (ll >= sizeof line) ll = sizeof line - 1;
        memcpy(line, text + ls, ll); line[ll] = 0;
        if (line[0] == 'o' && (line[1] == ' ' || line[1] == '\t')) { if (++objects_seen > 1 && res.n > 0) break; continue; }
        if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) { n_pos++; continue; }
        if (line[0] == 'v' && line[1] == 'n' && (line[2] == ' ' || line[2] == '\t')) {
            char* p = line + 2;
            for (int k = 0; k < 3; ++k) {
                char* e = NULL; double d = strtod(p, &e);
                if (e == p) { free(vn.v); free(res.v); return -3; }
                p = e;
                if (fvec_push(&vn, (float)d)) { free(vn.v); free(res.v); return -4; }
            }
        } else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
            long nidx[64]; int nv = 0;
            char* p = line + 1;
            while (*p && nv < 64) {
                while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
                if (!*p) break;
                char* e = NULL;
                long vi = strtol(p, &e, 10);
                if (e == p) break;
                (void)vi;
                p = e;
                long ni = 0; int have = 0;
                if (*p == '/') {               /* v/vt/vn or v//vn or v/vt */
                    ++p;
                    if (*p != '/') { strtol(p, &e, 10); p = e; }
                    if (*p == '/') { ++p; char* e2 = NULL; ni = strtol(p, &e2, 10); if (e2 != p) { have = 1; p = e2; } }
                }
                long count = (long)(vn.n / 3);
                if (have && ni < 0) ni = count + ni + 1;
                nidx[nv++] = have ? ni - 1 : -1;
                while (*p && *p != ' ' && *p != '\t') ++p;
            }
            for (int k = 1; k + 1 < nv; ++k) {
                long tri_idx[3] = { nidx[0], nidx[k], nidx[k + 1] };
                for (int c = 0; c < 3; ++c) {
                    long ni = tri_idx[c];
                    for (int d = 0; d < 3; ++d) {
                        float val = (ni >= 0 && (size_t)ni * 3 + 2 < vn.n) ? vn.v[ni * 3 + d] : 0.0f;
                        if (fvec_push(&res, val)) { free(vn.v); free(res.v); return -4; }
                    }
                }
            }
        }
    }
    (void)n_pos;
    free(vn.v);
    *out = res.v;
    return (int64_t)(res.n / 9);
}

/* Rust `f32 as usize` on a 64-bit target: saturating, NaN -> 0 */
static uint64_t f32_as_usize(float x) {
    if (!(x > 0.0f)) return 0u;
    if (x >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)x;
}

void orc_shade_texture(const float* tex_coords, uint64_t n_prims, const uint8_t* texels, uint32_t tex_w, uint32_t tex_h,
                       const orc_hit* hits, uint64_t n, uint32_t* rgba_out) {
    for (uint64_t i = 0; i < n; ++i) {
        const orc_hit* h = &hits[i];
        float r[3] = { 0.0f, 0.0f, 0.0f };                  /* miss: Vector3::zero() (renderer.rs:331) */
        if (h->id != 0xFFFFFFFFu) {
            uint32_t prim = h->id & 0x000FFFFFu;            /* instance index is always 0: object 0's model (renderer.rs:309-316) */
            if (prim < n_prims && tex_w && tex_h) {
                const float* tc = tex_coords + (size_t)prim * 6;
                float w0 = (1.0f - h->u) - h->v;
                float uvx = (tc[0] * w0 + tc[2] * h->u) + tc[4] * h->v;      /* renderer.rs:320 */
                float uvy = (tc[1] * w0 + tc[3] * h->u) + tc[5] * h->v;
                uint64_t iu = f32_as_usize(uvx * (float)tex_w) % tex_w;      /* material.rs:45-48 */
                uint64_t iv = f32_as_usize(uvy * (float)tex_h) % tex_h;
                const uint8_t* px = texels + ((size_t)iv * tex_w + iu) * 3;
                const float s = 1.0f / 256.0f;                               /* renderer.rs:299-306 */
                for (int k = 0; k < 3; ++k) r[k] = (float)px[k] * s;
            }
        }
        uint32_t c[3];
        for (int k = 0; k < 3; ++k) { c[k] = f32_as_u8(255.0f * r[k]); if (c[k] > 255) c[k] = 255; }
        rgba_out[i] = c[0] | (c[1] << 8) | (c[2] << 16) | 0xFF000000u;
    }
}

void orc_shade_normal(const float* normals, uint64_t n_prims, const float m[16], const orc_hit* hits, uint64_t n, uint32_t* rgba_out) {
    for (uint64_t i = 0; i < n; ++i) {
        const orc_hit* h = &hits[i];
        float r[3] = { 0.0f, 0.0f, 0.0f };
        if (h->id != 0xFFFFFFFFu) {
            uint32_t prim = h->id & 0x000FFFFFu;            /* instance index is always 0 */
            if (prim < n_prims) {
                const float* nm = normals + (size_t)prim * 9;
                float w0 = (1.0f - h->u) - h->v;             /* 1_f32 - u - v */
                float ms[3], ws[3], nn[3];
                for (int k = 0; k < 3; ++k) ms[k] = (nm[k] * w0 + nm[3 + k] * h->u) + nm[6 + k] * h->v;
                orc_transform_vector(m, ms, ws);
                orc_vec3_normalize(ws, nn);
                for (int k = 0; k < 3; ++k) r[k] = (nn[k] + 1.0f) * 0.5f;
            }
        }
        uint32_t c[3];
        for (int k = 0; k < 3; ++k) { c[k] = f32_as_u8(255.0f * r[k]); if (c[k] > 255) c[k] = 255; }
        rgba_out[i] = c[0] | (c[1] << 8) | (c[2] << 16) | 0xFF000000u;
    }
}
