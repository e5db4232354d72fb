// trace_kernels.cuh -- the closest-hit kernels (K1 primary rays, K1b arbitrary rays).
//
// This file is compiled TWICE (trace_strict.cu / trace_fast.cu) with different floating-point flags:
//   strict: --fmad=false -prec-div=true -prec-sqrt=true -ftz=false   (rustc's arithmetic model: no contraction)
//   fast  : --fmad=true  (FMA contraction; results within the fast-mode tolerance of BASELINE.json)
// BVHT_MODE_NS names the namespace of the instantiation.
//
// What it replaces in the reference (bvhtracer/src/...):
//   renderer.rs:345-368   PathTracer::evaluate, first loop (tiles of 8x8 pixels, one primary ray per pixel)
//   camera.rs:994-1010    Camera::get_ray_eye / get_ray_world
//   query/ray.rs:23-35    Ray::new / from_origin_dir (three IEEE divides)
//   scene/tlas.rs:123-177 Tlas::intersect (ordered near-first walk, root box never tested, swapped accessors)
//   scene/scene_object.rs:78-89  SceneObject::intersect (world -> model ray, t shared)
//   model/bvh.rs:242-305  Bvh::intersect_subtree (leaf triangles against the ENTRY ray, boxes against the
//                         shrinking ray, ties / double-miss descend RIGHT first)
//   geometry/aabb.rs:65-84       Aabb::intersect (slab)
//   geometry/triangle.rs:41-72   Triangle::intersect (Moeller-Trumbore, thresholds 1e-4, no culling)
//
// Design (DESIGN.md "Kernels"): persistent CTAs, one per SM slot; each WARP pulls a 32-pixel slice of an
// 8x8 tile with one atomicAdd (lane 0) + shuffle broadcast, generates its 32 primary rays in registers,
// walks TLAS -> instance -> BLAS with two short per-thread stacks, and writes one 16-byte record per pixel.
#pragma once
#include <cfloat>
#include "device_types.cuh"

#ifndef BVHT_MODE_NS
#error "define BVHT_MODE_NS (strict|fast) before including trace_kernels.cuh"
#endif

#ifndef BVHT_MIN_BLOCKS
#define BVHT_MIN_BLOCKS 8      // CTAs of 128 threads per SM the register allocation must allow: 64 registers (swept 5-10 on B200)
#endif

namespace bvht {
namespace BVHT_MODE_NS {

struct RayM {            // a ray in some space with its cached reciprocal (query/ray.rs:9-17)
    float ox, oy, oz;
    float dx, dy, dz;
    float rdx, rdy, rdz;
    // For the FMA-form slab test of OUR conservative boxes only (never reference boxes): reciprocal clamped to
    // +-1e30 (so that a zero / denormal direction component cannot produce inf - inf = NaN) and -(o * that).
    float fx, fy, fz;
    float nx, ny, nz;
};

__device__ __forceinline__ void ray_prepare_fma(RayM& r) {
    r.fx = fminf(fmaxf(r.rdx, -1e30f), 1e30f);
    r.fy = fminf(fmaxf(r.rdy, -1e30f), 1e30f);
    r.fz = fminf(fmaxf(r.rdz, -1e30f), 1e30f);
    r.nx = -(r.ox * r.fx); r.ny = -(r.oy * r.fy); r.nz = -(r.oz * r.fz);
}

struct HitRec { float t, u, v; uint32_t id; };

// Result stores.  BVHT_STORE_CS: 1 = streaming stores (evict-first) for the records / pixels K0 writes, 2 = for K1's as well
// (measured on C3 4K and C5 8K, resident and end to end: no difference either way, so plain stores stay the default).
#ifndef BVHT_STORE_CS
#define BVHT_STORE_CS 0
#endif
__device__ __forceinline__ void store_hit_k0(uint4* p, uint4 v) {
#if BVHT_STORE_CS >= 1
    __stcs(p, v);
#else
    *p = v;
#endif
}
__device__ __forceinline__ void store_px_k0(uint32_t* p, uint32_t v) {
#if BVHT_STORE_CS >= 1
    __stcs(p, v);
#else
    *p = v;
#endif
}
__device__ __forceinline__ void store_hit_k1(uint4* p, uint4 v) {
#if BVHT_STORE_CS >= 2
    __stcs(p, v);
#else
    *p = v;
#endif
}
__device__ __forceinline__ void store_px_k1(uint32_t* p, uint32_t v) {
#if BVHT_STORE_CS >= 2
    __stcs(p, v);
#else
    *p = v;
#endif
}

// Per-thread work counters, compiled in only for the debug entry point (BVHT_STATS); otherwise an empty type
// whose calls vanish.  Slots: 0 rays, 1 tlas pair tests, 2 instance entries, 3 reference BLAS pair tests,
// 4 reference leaves visited, 5 brute-force triangle tests, 6 sub-BVH pair tests, 7 sub-BVH triangle tests,
// 8 accel fallbacks (ray outside the inflation limits), 9 hits, 10 mt_finish evaluations, 11 chain heads resolved by the
// skip table, 12 rays whose block saw no instance (no ray generated), 13 pixel blocks pulled by K1, 14 skip tables built
#ifdef BVHT_STATS
constexpr int kStatSlots = 15;
struct Stat {
    unsigned long long c[kStatSlots];
    __device__ __forceinline__ Stat() { for (int i = 0; i < kStatSlots; ++i) c[i] = 0; }
    __device__ __forceinline__ void add(int i, unsigned n = 1) { c[i] += n; }
};
#else
struct Stat {
    __device__ __forceinline__ void add(int, unsigned = 1) {}
};
#endif

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// cglinalg Matrix4x4 * Vector4, one row: ((c0*x + c1*y) + c2*z) + c3*w with round-to-nearest intrinsics (never contracted)
#define BVHT_MV4(cx0, cx1, cx2, cx3, x, y, z, wv) \
    __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx0, x), __fmul_rn(cx1, y)), __fmul_rn(cx2, z)), __fmul_rn(cx3, wv))

// geometry/aabb.rs:65-84.  `tcl` is the ray's current t.
__device__ __forceinline__ bool slab_test(const float4 lo, const float4 hi, const RayM& r, float tcl, float& tmin_out) {
    float t_x1 = (lo.x - r.ox) * r.rdx;
    float t_x2 = (hi.x - r.ox) * r.rdx;
    float t_min = fminf(t_x1, t_x2);
    float t_max = fmaxf(t_x1, t_x2);
    float t_y1 = (lo.y - r.oy) * r.rdy;
    float t_y2 = (hi.y - r.oy) * r.rdy;
    t_min = fmaxf(t_min, fminf(t_y1, t_y2));
    t_max = fminf(t_max, fmaxf(t_y1, t_y2));
    float t_z1 = (lo.z - r.oz) * r.rdz;
    float t_z2 = (hi.z - r.oz) * r.rdz;
    t_min = fmaxf(t_min, fminf(t_z1, t_z2));
    t_max = fminf(t_max, fmaxf(t_z1, t_z2));
    tmin_out = t_min;
    return (t_max >= t_min) && (t_min < tcl) && (t_max > 0.0f);
}

// Slab test for OUR conservative boxes (sub-BVH nodes, tight TLAS boxes) -- not a reference function.  One FMA per
// plane, t = lo * rd - o * rd (explicit __fmaf_rn, so identical in the strict and fast builds), inclusive on every
// bound.  Its rounding differs from (lo - o) * rd by at most ~eps * |o * rd|, i.e. a plane shift of eps * |o|, which
// the box inflation (leaf_accel.cpp delta_abs) covers.  The reciprocal is clamped to +-1e30 (ray_prepare_fma): an axis
// with d == 0 then yields t = +-(huge) with the right signs (inside the slab: [-huge, +huge]; outside: both on one side).
__device__ __forceinline__ bool slab_test_sub(const float4 lo, const float4 hi, const RayM& r, float tcl, float& tmin_out) {
    float t_x1 = __fmaf_rn(lo.x, r.fx, r.nx);
    float t_x2 = __fmaf_rn(hi.x, r.fx, r.nx);
    float t_y1 = __fmaf_rn(lo.y, r.fy, r.ny);
    float t_y2 = __fmaf_rn(hi.y, r.fy, r.ny);
    float t_z1 = __fmaf_rn(lo.z, r.fz, r.nz);
    float t_z2 = __fmaf_rn(hi.z, r.fz, r.nz);
    float t_min = fmaxf(fmaxf(fminf(t_x1, t_x2), fminf(t_y1, t_y2)), fminf(t_z1, t_z2));
    float t_max = fminf(fminf(fmaxf(t_x1, t_x2), fmaxf(t_y1, t_y2)), fmaxf(t_z1, t_z2));
    tmin_out = t_min;
    return (t_max >= t_min) && (t_min <= tcl) && (t_max >= 0.0f);
}

// The same test on a child box stored as CENTRE + HALF EXTENT (device_types.cuh BVHT_SUB_CH; bake_sub_nodes_kernel
// writes that form): t_centre = c * f - o * f, t_near/far = t_centre -/+ h * |f|.  No per-axis min/max: the work moves
// from the ALU pipe (FMNMX, the busiest pipe of this kernel) to the FMA pipes.  Two more roundings than the lo/hi form
// (eps * |t_centre| and eps * |h f|), which the bake adds to h (upload_kernels.cu).
__device__ __forceinline__ bool slab_test_ch(const float4 c, const float4 h, const RayM& r, float tcl, float& tmin_out) {
    float tcx = __fmaf_rn(c.x, r.fx, r.nx);
    float tcy = __fmaf_rn(c.y, r.fy, r.ny);
    float tcz = __fmaf_rn(c.z, r.fz, r.nz);
#if BVHT_SUB_CH == 2
    float ax = fabsf(r.fx), ay = fabsf(r.fy), az = fabsf(r.fz);
    float t_min = fmaxf(fmaxf(__fmaf_rn(-h.x, ax, tcx), __fmaf_rn(-h.y, ay, tcy)), __fmaf_rn(-h.z, az, tcz));
    float t_max = fminf(fminf(__fmaf_rn(h.x, ax, tcx), __fmaf_rn(h.y, ay, tcy)), __fmaf_rn(h.z, az, tcz));
#else
    float px = fabsf(__fmul_rn(h.x, r.fx)), py = fabsf(__fmul_rn(h.y, r.fy)), pz = fabsf(__fmul_rn(h.z, r.fz));
    float t_min = fmaxf(fmaxf(__fsub_rn(tcx, px), __fsub_rn(tcy, py)), __fsub_rn(tcz, pz));
    float t_max = fminf(fminf(__fadd_rn(tcx, px), __fadd_rn(tcy, py)), __fadd_rn(tcz, pz));
#endif
    tmin_out = t_min;
    return (t_max >= t_min) && (t_min <= tcl) && (t_max >= 0.0f);
}

// geometry/triangle.rs:41-72 on the repacked triangle (v0, e1 = v1 - v0, e2 = v2 - v0), split in two so that the
// hot loop is straight-line code with ONE rarely taken branch:
//
//  mt_filter  evaluates, with exactly the reference's operations, normal = d x e2, area = e1 . normal, s = o - v0,
//             X = s . normal, and decides -- exactly -- whether the reference would already have returned None at
//             `|area| < 1e-4` or at `u = (1 / area) * X; u < 0 || u > 1`, WITHOUT the IEEE divide (98 % of all tests
//             end here).  With a = |area| in [1e-4, 2^59] and x = |X|:
//              * signs differ and x >= 2^-60: f = fl(1/area) has |f| >= 2^-60, so fl(f * X) is a negative NORMAL
//                number: u < 0.  (A smaller x could round to -0.0, which the reference does not reject: undecided.)
//              * signs equal and x > fl(a * (1 + 2^-20)): then x > a (1 + 2^-21), fl(1/a) >= (1 - 2^-24) / a, so
//                fl(1/a) * x > 1 + 2^-22 and its rounding is still > 1: u > 1.
//             Anything undecided (including NaNs: every comparison is false) goes to mt_finish.
//  mt_finish  is the rest of Triangle::intersect verbatim (f, u, the u test again, q, v, t, thresholds).
struct MtPartial { float nx, ny, nz, area, sx, sy, sz, X; };

#ifdef BVHT_FAST_MODE
// FAST build.  The filter runs with FMA contraction (that is the mode's speed-up on the 98 % of tests that end here) and, since
// its inputs are then no longer the reference's values, with WIDENED thresholds: it only rejects what is outside by a margin
// (|area| < 0.99e-4, u < -1e-4, u > 1 + 1e-4).  What passes is evaluated by mt_finish in the reference's arithmetic
// (round-to-nearest intrinsics, never contracted), thresholds included -- so acceptance, (t, u, v) and every comparison
// between candidates are the strict build's.  Ids can differ from the reference only where contraction moves u by more
// than the margin (|area| tiny against its terms) or, with one sub-BVH per model, where two triangles tie exactly in t.
__device__ __forceinline__ bool mt_filter(const float4 v0, const float4 e1, const float4 e2, const RayM& r, MtPartial& p) {
    float nx = r.dy * e2.z - r.dz * e2.y;
    float ny = r.dz * e2.x - r.dx * e2.z;
    float nz = r.dx * e2.y - r.dy * e2.x;
    float area = (e1.x * nx + e1.y * ny) + e1.z * nz;
    p.sx = __fsub_rn(r.ox, v0.x); p.sy = __fsub_rn(r.oy, v0.y); p.sz = __fsub_rn(r.oz, v0.z);
    float X = (p.sx * nx + p.sy * ny) + p.sz * nz;
    float a = fabsf(area), x = fabsf(X);
    bool opposite = (__float_as_int(area) ^ __float_as_int(X)) < 0;
    bool rej_area = a < 0.000099f;
    bool rej_neg = opposite & (x > a * 0.0001f);
    bool rej_big = (!opposite) & (x > a * 1.0001f);
    return !(rej_area | rej_neg | rej_big);
}

__device__ __forceinline__ bool mt_finish(const float4 e1, const float4 e2, const RayM& r, const MtPartial& p, float entry_t,
                                          float& t_out, float& u_out, float& v_out) {
    // geometry/triangle.rs:41-72 verbatim, one IEEE operation per source operation
    const float threshold = 0.0001f;
    float nx = __fsub_rn(__fmul_rn(r.dy, e2.z), __fmul_rn(r.dz, e2.y));
    float ny = __fsub_rn(__fmul_rn(r.dz, e2.x), __fmul_rn(r.dx, e2.z));
    float nz = __fsub_rn(__fmul_rn(r.dx, e2.y), __fmul_rn(r.dy, e2.x));
    float area = __fadd_rn(__fadd_rn(__fmul_rn(e1.x, nx), __fmul_rn(e1.y, ny)), __fmul_rn(e1.z, nz));
    if (fabsf(area) < threshold) return false;
    float f = __fdiv_rn(1.0f, area);
    float u = __fmul_rn(f, __fadd_rn(__fadd_rn(__fmul_rn(p.sx, nx), __fmul_rn(p.sy, ny)), __fmul_rn(p.sz, nz)));
    if (u < 0.0f || u > 1.0f) return false;
    float qx = __fsub_rn(__fmul_rn(p.sy, e1.z), __fmul_rn(p.sz, e1.y));
    float qy = __fsub_rn(__fmul_rn(p.sz, e1.x), __fmul_rn(p.sx, e1.z));
    float qz = __fsub_rn(__fmul_rn(p.sx, e1.y), __fmul_rn(p.sy, e1.x));
    float v = __fmul_rn(f, __fadd_rn(__fadd_rn(__fmul_rn(r.dx, qx), __fmul_rn(r.dy, qy)), __fmul_rn(r.dz, qz)));
    if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return false;
    float t = __fmul_rn(f, __fadd_rn(__fadd_rn(__fmul_rn(e2.x, qx), __fmul_rn(e2.y, qy)), __fmul_rn(e2.z, qz)));
    if (!(t > threshold)) return false;
    t_out = fminf(entry_t, t);
    u_out = u;
    v_out = v;
    return true;
}
#else
__device__ __forceinline__ bool mt_filter(const float4 v0, const float4 e1, const float4 e2, const RayM& r, MtPartial& p) {
    p.nx = r.dy * e2.z - r.dz * e2.y;
    p.ny = r.dz * e2.x - r.dx * e2.z;
    p.nz = r.dx * e2.y - r.dy * e2.x;
    p.area = (e1.x * p.nx + e1.y * p.ny) + e1.z * p.nz;
    p.sx = r.ox - v0.x; p.sy = r.oy - v0.y; p.sz = r.oz - v0.z;
    p.X = (p.sx * p.nx + p.sy * p.ny) + p.sz * p.nz;
    float a = fabsf(p.area), x = fabsf(p.X);
    bool opposite = (__float_as_int(p.area) ^ __float_as_int(p.X)) < 0;
    bool rej_area = a < 0.0001f;
    bool rej_neg = opposite & (x >= 8.673617379884035e-19f) & (a <= 5.764607523034235e17f);
    bool rej_big = (!opposite) & (x > a * 1.00000095367431640625f);
    return !(rej_area | rej_neg | rej_big);
}

__device__ __forceinline__ bool mt_finish(const float4 e1, const float4 e2, const RayM& r, const MtPartial& p, float entry_t,
                                          float& t_out, float& u_out, float& v_out) {
    const float threshold = 0.0001f;
    if (fabsf(p.area) < threshold) return false;
    float f = 1.0f / p.area;
    float u = f * p.X;
    if (u < 0.0f || u > 1.0f) return false;
    float qx = p.sy * e1.z - p.sz * e1.y;
    float qy = p.sz * e1.x - p.sx * e1.z;
    float qz = p.sx * e1.y - p.sy * e1.x;
    float v = f * ((r.dx * qx + r.dy * qy) + r.dz * qz);
    if (v < 0.0f || u + v > 1.0f) return false;
    float t = f * ((e2.x * qx + e2.y * qy) + e2.z * qz);
    if (!(t > threshold)) return false;
    t_out = fminf(entry_t, t);
    u_out = u;
    v_out = v;
    return true;
}

#endif   // BVHT_FAST_MODE (filter / finish)

// Brute-force leaf: primitives base .. base+count ascending, strict '<' against the shrinking closest t
// (bvh.rs:250-258).  Triangles are tested against the ENTRY ray.
__device__ __forceinline__ void leaf_brute(const BlasDesc& B, uint32_t base, uint32_t count, const RayM& r, float entry_t,
                                           float& best_t, float& best_u, float& best_v, uint32_t& best_prim, bool& found, Stat& st) {
    st.add(5, count);
    const float4* tp = B.tri + 3 * (size_t)base;
    for (uint32_t k = 0; k < count; ++k, tp += 3) {
        uint32_t pi = base + k;
        float4 v0 = ldg4(tp + 0);
        float4 e1 = ldg4(tp + 1);
        float4 e2 = ldg4(tp + 2);
        MtPartial mp;
        if (mt_filter(v0, e1, e2, r, mp)) {
            float t, u, v;
            st.add(10);
            if (mt_finish(e1, e2, r, mp, entry_t, t, u, v)) {
                if (t < best_t) { best_t = t; best_u = u; best_v = v; best_prim = pi; found = true; }
            }
        }
    }
}

// Leaf accelerator: conservative sub-BVH built by us over the triangles of one oversized reference leaf
// (leaf_accel.hpp).  The reference's result for a leaf is the lexicographic minimum (t, primitive index)
// over the triangles that Triangle::intersect accepts with t < closest-at-entry, because the leaf loop
// tests against the ENTRY ray (bvh.rs:251) and accepts with strict '<' in ascending index order.  Any
// visiting order gives that same answer as long as no accepting triangle is skipped, which the
// pre-inflated boxes guarantee under the bound checked by the caller (DESIGN.md "Leaf accelerator").
__device__ __forceinline__ void leaf_accel(const BlasDesc& B, uint32_t sub_root, const RayM& r, float entry_t,
                                           float& best_t, float& best_u, float& best_v, uint32_t& best_prim, bool& found, Stat& st) {
    // within this leaf: lt/lu/lv/lp = lexicographic-min candidate; ties with the entry value are rejected
    // because lp starts at 0 (no index is < 0) while lt starts at the closest-at-entry value.
    float lt = best_t, lu = 0.0f, lv = 0.0f;
    uint32_t lp = 0u;
    bool lfound = false;
    uint32_t stack[kSubStack];
    int sp = 0;
    uint32_t ref = sub_root;
    for (;;) {
        if (ref & 0x80000000u) {
            // leaf ref: [30:28] = count-1, [27:0] = first triangle in sub order
            uint32_t first = ref & 0x0FFFFFFFu;
            uint32_t cnt = ((ref >> 28) & 7u) + 1u;
            st.add(7, cnt);
            const float4* tp = B.stri + 3 * (size_t)first;
            for (uint32_t k = 0; k < cnt; ++k, tp += 3) {
                float4 v0 = ldg4(tp + 0);
                float4 e1 = ldg4(tp + 1);
                float4 e2 = ldg4(tp + 2);
                MtPartial mp;
                if (mt_filter(v0, e1, e2, r, mp)) {
                    float t, u, v;
                    st.add(10);
                    if (mt_finish(e1, e2, r, mp, entry_t, t, u, v)) {
                        uint32_t pi = __float_as_uint(v0.w);
                        if (t < lt || (t == lt && pi < lp)) { lt = t; lu = u; lv = v; lp = pi; lfound = true; }
                    }
                }
            }
            if (sp == 0) break;
            ref = stack[--sp];
        } else {
            st.add(6);
            const float4* n = B.sub_nodes + (size_t)ref * 4;
            float4 a = ldg4(n + 0);   // child0 lo.xyz, child0 ref
            float4 b = ldg4(n + 1);   // child0 hi.xyz, child1 ref
            float4 c = ldg4(n + 2);   // child1 lo.xyz
            float4 d = ldg4(n + 3);   // child1 hi.xyz
            float t0, t1;
            // inclusive comparisons: a node holding a triangle with t == lt but a lower index must be visited
#if BVHT_SUB_CH
            bool h0 = slab_test_ch(a, b, r, lt, t0);
            bool h1 = slab_test_ch(c, d, r, lt, t1);
#else
            bool h0 = slab_test_sub(a, b, r, lt, t0);
            bool h1 = slab_test_sub(c, d, r, lt, t1);
#endif
            uint32_t r0 = __float_as_uint(a.w), r1 = __float_as_uint(b.w);
            if (h0 && h1) {
                bool swap = t1 < t0;
                uint32_t nearr = swap ? r1 : r0, farr = swap ? r0 : r1;
                stack[sp++] = farr;
                ref = nearr;
            } else if (h0) {
                ref = r0;
            } else if (h1) {
                ref = r1;
            } else {
                if (sp == 0) break;
                ref = stack[--sp];
            }
        }
    }
    if (lfound) { best_t = lt; best_u = lu; best_v = lv; best_prim = lp; found = true; }
}

// model/bvh.rs:242-305 for one instance.  `r` is the MODEL-space ray, entry_t its t at entry (= the world
// ray's current closest, scene_object.rs:87).  On return `found` says whether a strictly closer hit exists.
template <bool ACCEL>
__device__ __forceinline__ void blas_intersect(const BlasDesc& B, const RayM& r, float entry_t, bool use_accel,
                                               float& best_t, float& best_u, float& best_v, uint32_t& best_prim, bool& found, Stat& st) {
    best_t = entry_t;
    found = false;
#ifdef BVHT_FAST_MODE
    if (ACCEL && use_accel && B.leaf_sub_root == nullptr) {
        // fast mode: one sub-BVH over the whole model (sub node 0), closest hit = lexicographic min (t, primitive index)
        st.add(4);
        leaf_accel(B, 0u, r, entry_t, best_t, best_u, best_v, best_prim, found, st);
        return;
    }
#endif
    uint32_t stack[kBlasStack];
    int sp = 0;
    uint32_t ni = 0;                       // root: its AABB is never tested (bvh.rs:243)
    float4 n0 = ldg4(B.nodes + 0);
    float4 n1 = ldg4(B.nodes + 1);
    for (;;) {
        uint32_t count = __float_as_uint(n1.w);
        uint32_t lf = __float_as_uint(n0.w);
        if (count > 0) {
            bool done = false;
            st.add(4);
            if (ACCEL) {
                if (use_accel) {
                    uint32_t sr = __ldg(B.leaf_sub_root + ni);
                    if (sr != 0xFFFFFFFFu) { leaf_accel(B, sr, r, entry_t, best_t, best_u, best_v, best_prim, found, st); done = true; }
                }
            }
            if (!done) leaf_brute(B, lf, count, r, entry_t, best_t, best_u, best_v, best_prim, found, st);
            if (sp == 0) break;
            ni = stack[--sp];
            n0 = ldg4(B.nodes + 2 * (size_t)ni);
            n1 = ldg4(B.nodes + 2 * (size_t)ni + 1);
        } else {
            // children are adjacent: 4 float4 = 64 contiguous, 64-byte aligned bytes (left index is even)
            st.add(3);
            const float4* c = B.nodes + 2 * (size_t)lf;
            float4 l0 = ldg4(c + 0), l1 = ldg4(c + 1), r0 = ldg4(c + 2), r1 = ldg4(c + 3);
            float ld, rd;
            bool lh = slab_test(l0, l1, r, best_t, ld);
            bool rh = slab_test(r0, r1, r, best_t, rd);
            float lkey = lh ? ld : FLT_MAX;
            float rkey = rh ? rd : FLT_MAX;
            bool left_first = lkey < rkey;         // ties and double-miss: right first (bvh.rs:271-275)
            bool near_hit = left_first ? lh : rh;
            bool far_hit = left_first ? rh : lh;
            if (near_hit) {
                if (far_hit) stack[sp++] = left_first ? lf + 1 : lf;
                if (left_first) { ni = lf; n0 = l0; n1 = l1; } else { ni = lf + 1; n0 = r0; n1 = r1; }
                continue;
            }
            if (sp == 0) break;
            ni = stack[--sp];
            n0 = ldg4(B.nodes + 2 * (size_t)ni);
            n1 = ldg4(B.nodes + 2 * (size_t)ni + 1);
        }
    }
}

// scene/tlas.rs:123-177 + scene_object.rs:78-89.  World ray (ox..dz, recip), initial t = tmax.
//
// Chain skipping (ACCEL, `skip` != nullptr; built per pixel block by build_tlas_skip).  Under a candidate mask a TLAS
// subtree without candidate instances is never entered, so most interior nodes have ONE relevant child -- and because
// every .tri instance box reaches +999 (the sentinel), the reference's boxes almost never cull: the walk degenerates into
// long chains (the clustering of sixteen_armadillos is 12-15 levels deep).  skip[x] is the node at the end of the chain
// that starts at x: the first descendant that is a leaf or has two relevant children (x itself if it is one).  The
// reference, having arrived at x, tests the boxes of the chain x -> c1 -> ... -> s one after the other with an unchanged
// `closest` (nothing else is visited in between, and culled subtrees have no side effects).  Those boxes are nested
// (a parent is the exact min/max union of its children, tlas.rs:233-234) and Aabb::intersect is monotonic in the box
// for finite reciprocals (rounding of (b - o) * rd is monotonic in b), so "every box of the chain is hit" <=> "the
// last one is hit": one slab test replaces the chain.  Rays with a non-finite reciprocal take the plain walk.
template <bool ACCEL, bool PRUNE = false>
__device__ __forceinline__ HitRec scene_intersect(const SceneDev& S, const RayM& w, float tmax, uint32_t cand, Stat& st,
                                                  const uint8_t* skip = nullptr, const float4* origins = nullptr) {
    // cand (ACCEL): bit i clear = instance i cannot be hit by this ray (tile-level screen rectangles); all ones = unknown
    st.add(0);
    HitRec best;
    best.t = FLT_MAX; best.u = 0.0f; best.v = 0.0f; best.id = 0xFFFFFFFFu;
    float closest = tmax;
    bool have = false;
    uint32_t stack[kTlasStack];
    int sp = 0;
    float4 n0 = ldg4(S.tlas + 0);
    float4 n1 = ldg4(S.tlas + 1);
    // ACCEL: conservative world-space boxes of the REAL geometry under each TLAS node ("tight" boxes).  The reference's
    // instance boxes are inflated to +999 by the sentinel triangle of every .tri asset, so most rays enter most
    // instances only to find nothing.  A subtree whose tight box the ray misses cannot produce a hit, and entering it
    // has no side effect in the reference (closest is only updated by hits), so skipping it leaves the rest of the
    // reference's ordered walk -- and therefore the result -- unchanged.
    float wd2 = 0.0f, wo2 = 0.0f;
    RayM wf = w;
    if (ACCEL) {
        wd2 = (w.dx * w.dx + w.dy * w.dy) + w.dz * w.dz;
        {
            float cx = w.ox - S.tight_center[0], cy = w.oy - S.tight_center[1], cz = w.oz - S.tight_center[2];
            wo2 = (cx * cx + cy * cy) + cz * cz;
        }
        ray_prepare_fma(wf);
        float4 t0 = ldg4(S.tlas_tight + 0), t1 = ldg4(S.tlas_tight + 1);
        float tt;
        if (wd2 <= t0.w && wo2 <= t1.w && !slab_test_sub(t0, t1, wf, closest, tt)) return best;   // nothing reachable at all
        if (cand != 0xFFFFFFFFu && (__ldg(S.tlas_mask + 0) & cand) == 0u) return best;
    }
    uint32_t cur = 0u;                     // index of the node held in n0 / n1 (chain skipping only)
    if (ACCEL && PRUNE) {
        if (skip && !(fabsf(w.rdx) < 3.0e38f && fabsf(w.rdy) < 3.0e38f && fabsf(w.rdz) < 3.0e38f)) skip = nullptr;
    }
    for (;;) {
        if (ACCEL && PRUNE) {
            if (skip) {
                uint32_t sn = skip[cur];
                if (sn != cur) {
                    // arrived at the head of a chain: only the box at its end decides (see above)
                    st.add(11);
                    n0 = ldg4(S.tlas + 2 * (size_t)sn);
                    n1 = ldg4(S.tlas + 2 * (size_t)sn + 1);
                    bool ok = true;
                    float tt;
                    if (__float_as_uint(n0.w) == 0u) {          // a leaf: its tight box first
                        float4 a0 = ldg4(S.tlas_tight + 2 * (size_t)sn), a1 = ldg4(S.tlas_tight + 2 * (size_t)sn + 1);
                        if (wd2 <= a0.w && wo2 <= a1.w) ok = slab_test_sub(a0, a1, wf, closest, tt);
                    }
                    if (ok) ok = slab_test(n0, n1, w, closest, tt);
                    if (!ok) {
                        if (sp == 0) break;
                        cur = stack[--sp];
                        n0 = ldg4(S.tlas + 2 * (size_t)cur);
                        n1 = ldg4(S.tlas + 2 * (size_t)cur + 1);
                        continue;
                    }
                    cur = sn;
                }
            }
        }
        uint32_t lr = __float_as_uint(n0.w);
        if (lr == 0u) {
            uint32_t inst = __float_as_uint(n1.w);
            const float4* m = S.inst_cols + 4 * (size_t)inst;
            float4 c0 = ldg4(m + 0), c1 = ldg4(m + 1), c2 = ldg4(m + 2), c3 = ldg4(m + 3);
            RayM r;
            // transform.rs:219-234 over cglinalg Matrix4x4 * Vector4: ((c0*x + c1*y) + c2*z) + c3*w
            // (round-to-nearest intrinsics: never contracted, so the model-space ray is the reference's in BOTH builds)
            if (origins) {                    // primary rays: the shared origin was transformed once per instance on the host
                const float4 oo = origins[inst];
                r.ox = oo.x; r.oy = oo.y; r.oz = oo.z;
            } else {
                r.ox = BVHT_MV4(c0.x, c1.x, c2.x, c3.x, w.ox, w.oy, w.oz, 1.0f);
                r.oy = BVHT_MV4(c0.y, c1.y, c2.y, c3.y, w.ox, w.oy, w.oz, 1.0f);
                r.oz = BVHT_MV4(c0.z, c1.z, c2.z, c3.z, w.ox, w.oy, w.oz, 1.0f);
            }
            r.dx = BVHT_MV4(c0.x, c1.x, c2.x, c3.x, w.dx, w.dy, w.dz, 0.0f);
            r.dy = BVHT_MV4(c0.y, c1.y, c2.y, c3.y, w.dx, w.dy, w.dz, 0.0f);
            r.dz = BVHT_MV4(c0.z, c1.z, c2.z, c3.z, w.dx, w.dy, w.dz, 0.0f);
            r.rdx = __fdiv_rn(1.0f, r.dx); r.rdy = __fdiv_rn(1.0f, r.dy); r.rdz = __fdiv_rn(1.0f, r.dz);   // Ray::new, ray.rs:23-31
            // descriptor copied into registers once per instance entry (the loops below must not re-read it from memory)
            const BlasDesc* Bp = S.blas + __ldg(S.inst_blas + inst);
            BlasDesc B;
            B.nodes = Bp->nodes; B.tri = Bp->tri;
            if (ACCEL) {
                B.sub_nodes = Bp->sub_nodes; B.stri = Bp->stri; B.leaf_sub_root = Bp->leaf_sub_root;
                B.accel_d_max = Bp->accel_d_max; B.accel_o_max = Bp->accel_o_max;
            }
            bool use_accel = false;
            if (ACCEL) {
                // the sub boxes were inflated for model-space rays with |d| <= d_max and |o| <= o_max; anything
                // else takes the brute-force leaves (still exact)
                float dn2 = (r.dx * r.dx + r.dy * r.dy) + r.dz * r.dz;
                float on2 = (r.ox * r.ox + r.oy * r.oy) + r.oz * r.oz;
                use_accel = (dn2 <= B.accel_d_max * B.accel_d_max) && (on2 <= B.accel_o_max * B.accel_o_max);
                if (!use_accel) st.add(8);
                ray_prepare_fma(r);
            }
            st.add(2);
            float bt, bu, bv; uint32_t bp; bool found;
            blas_intersect<ACCEL>(B, r, closest, use_accel, bt, bu, bv, bp, found, st);
            if (found && bt < closest) {                                     // tlas.rs:131
                closest = bt;
                best.t = bt; best.u = bu; best.v = bv;
                // InstancePrimitiveIndex::from_primitive: instance bits are 0 in the reference (bvh.rs:296)
                best.id = (bp & 0x000FFFFFu) | ((S.flags & 0x4u) ? ((inst & 0xFFFu) << 20) : 0u);
                have = true;
            }
            if (sp == 0) break;
            uint32_t ni = stack[--sp];
            cur = ni;
            n0 = ldg4(S.tlas + 2 * (size_t)ni);
            n1 = ldg4(S.tlas + 2 * (size_t)ni + 1);
        } else {
            st.add(1);
            uint32_t li = (lr & 0xFFFF0000u) >> 16;    // left_blas()  = upper half (tlas.rs:25-27)
            uint32_t ri = lr & 0x0000FFFFu;            // right_blas() = lower half (tlas.rs:30-32)
            bool lt = true, rt = true;
            if (ACCEL) {
                // candidate masks first (no memory traffic beyond one word per child), then the tight boxes (one FMA per
                // plane): a child whose real geometry the ray cannot reach closer than `closest` is dropped without
                // evaluating the reference's box test for it
                bool l_tight = true, r_tight = true;
                if (cand != 0xFFFFFFFFu) {
                    uint32_t ml = __ldg(S.tlas_mask + li), mr = __ldg(S.tlas_mask + ri);
                    lt = (ml & cand) != 0u;
                    rt = (mr & cand) != 0u;
                    // with candidate masks active the interior tight boxes add little: test them at the leaves only
                    l_tight = __popc(ml) == 1;
                    r_tight = __popc(mr) == 1;
                }
                float tt;
                if (lt && l_tight) {
                    float4 a0 = ldg4(S.tlas_tight + 2 * (size_t)li), a1 = ldg4(S.tlas_tight + 2 * (size_t)li + 1);
                    if (wd2 <= a0.w && wo2 <= a1.w) lt = slab_test_sub(a0, a1, wf, closest, tt);
                }
                if (rt && r_tight) {
                    float4 b0 = ldg4(S.tlas_tight + 2 * (size_t)ri), b1 = ldg4(S.tlas_tight + 2 * (size_t)ri + 1);
                    if (wd2 <= b0.w && wo2 <= b1.w) rt = slab_test_sub(b0, b1, wf, closest, tt);
                }
            }
            float4 l0 = n0, l1 = n1, r0 = n0, r1 = n1;
            float ld = 0.0f, rd = 0.0f;
            bool lh = false, rh = false;
            if (lt) { l0 = ldg4(S.tlas + 2 * (size_t)li); l1 = ldg4(S.tlas + 2 * (size_t)li + 1); lh = slab_test(l0, l1, w, closest, ld); }
            if (rt) { r0 = ldg4(S.tlas + 2 * (size_t)ri); r1 = ldg4(S.tlas + 2 * (size_t)ri + 1); rh = slab_test(r0, r1, w, closest, rd); }
            float lkey = lh ? ld : FLT_MAX;
            float rkey = rh ? rd : FLT_MAX;
            // the order between two children only matters when both are entered, and then both keys are the reference's
            bool left_first = lkey < rkey;
            bool near_hit = left_first ? lh : rh;
            bool far_hit = left_first ? rh : lh;
            if (near_hit) {
                if (far_hit) stack[sp++] = left_first ? ri : li;
                if (left_first) { n0 = l0; n1 = l1; cur = li; } else { n0 = r0; n1 = r1; cur = ri; }
                continue;
            }
            if (ACCEL) {
                if (far_hit) {             // near child dropped (or missed): the far child is next in the reference's order
                    if (left_first) { n0 = r0; n1 = r1; cur = ri; } else { n0 = l0; n1 = l1; cur = li; }
                    continue;
                }
            }
            if (sp == 0) break;
            uint32_t ni = stack[--sp];
            cur = ni;
            n0 = ldg4(S.tlas + 2 * (size_t)ni);
            n1 = ldg4(S.tlas + 2 * (size_t)ni + 1);
        }
    }
    if (!((closest < FLT_MAX) && have)) { best.t = FLT_MAX; best.u = 0.0f; best.v = 0.0f; best.id = 0xFFFFFFFFu; }
    else st.add(9);
    return best;
}

// camera.rs:994-1010 with u, v from renderer.rs:358-361
__device__ __forceinline__ RayM primary_ray(const CameraDev& C, uint32_t px, uint32_t py, uint32_t width, uint32_t height) {
    float u = __fdiv_rn((float)px, (float)width);
    float v = __fdiv_rn((float)py, (float)height);
    float p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float origin = 0.0f;
        p[k] = __fadd_rn(__fadd_rn(__fadd_rn(origin, C.tl[k]), __fmul_rn(__fsub_rn(C.tr[k], C.tl[k]), u)),
                         __fmul_rn(__fsub_rn(C.bl[k], C.tl[k]), v));
        p[k] = __fsub_rn(p[k], origin);
    }
    // normalize = v / |v| (pinned, test_tri_mesh.rs:57-59)
    float m = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(p[0], p[0]), __fmul_rn(p[1], p[1])), __fmul_rn(p[2], p[2])));
    float ex = __fdiv_rn(p[0], m), ey = __fdiv_rn(p[1], m), ez = __fdiv_rn(p[2], m);
    RayM w;
    const float* M = C.vinv;
    w.ox = BVHT_MV4(M[0], M[4], M[8], M[12], 0.0f, 0.0f, 0.0f, 1.0f);
    w.oy = BVHT_MV4(M[1], M[5], M[9], M[13], 0.0f, 0.0f, 0.0f, 1.0f);
    w.oz = BVHT_MV4(M[2], M[6], M[10], M[14], 0.0f, 0.0f, 0.0f, 1.0f);
    w.dx = BVHT_MV4(M[0], M[4], M[8], M[12], ex, ey, ez, 0.0f);
    w.dy = BVHT_MV4(M[1], M[5], M[9], M[13], ex, ey, ez, 0.0f);
    w.dz = BVHT_MV4(M[2], M[6], M[10], M[14], ex, ey, ez, 0.0f);
    w.rdx = __fdiv_rn(1.0f, w.dx); w.rdy = __fdiv_rn(1.0f, w.dy); w.rdz = __fdiv_rn(1.0f, w.dz);
    return w;
}

// Rust's saturating `f32 as usize` (negative and NaN -> 0, >= 2^64 -> usize::MAX) followed by `% n` (material.rs:47-48).
__device__ __forceinline__ uint32_t texel_index(float x, uint32_t n) {
    if (!(x >= 1.0f)) return 0u;
    if (x < 4294967296.0f) return __float2uint_rz(x) % n;
    unsigned long long v = x >= 18446744073709551616.0f ? ~0ull : __float2ull_rz(x);
    return (uint32_t)(v % (unsigned long long)n);
}

// Fused accumulator + pixel shader (renderer.rs:116-245): what the examples turn a hit into.
//   kind 1: DepthAccumulator (:184-194) + DepthMappingShader (:207-222)   two/sixteen_armadillos
//   kind 2: IntersectionAccumulator (:145-153) + IntersectionShader (:166-174)   big_ben_clock
//   kind 3: UvMappingAccumulator (:233-245) + RadianceToRgbShader (:124-132)
//   kind 4: NormalMappingAccumulator (:256-286) + RadianceToRgbShader   cube, trippy_teapots
//   kind 5: TextureMaterialAccumulator (:289-334) + RadianceToRgbShader   quad
// Integer tricks are kept: `as i32` / `as u8` are saturating casts (NaN -> 0), u32 arithmetic wraps.
__device__ __forceinline__ uint32_t shade_pixel(const PrimaryParams& P, const HitRec& h) {
    const bool hit = h.id != 0xFFFFFFFFu;
    if (P.shade_kind == 1u) {
        float nearest_t = hit ? h.t : FLT_MAX;
        if (nearest_t < FLT_MAX) {
            int xi = __float2int_rz((nearest_t - P.shade_offset) * P.shade_scale);   // `as i32`
            uint32_t color = 255u - (uint32_t)xi;
            uint32_t c = color * 0x010101u;
            uint32_t r = (c & 0x00FF0000u) >> 16, g = (c & 0x0000FF00u) >> 8, b = c & 0x000000FFu;
            return r | (g << 8) | (b << 16) | 0xFF000000u;
        }
        return 0xFF000000u;                                                          // Rgba::new(0, 0, 0, 255)
    }
    if (P.shade_kind == 2u) return hit ? P.hit_rgba : P.miss_rgba;
    if (P.shade_kind == 3u) {
        float rx = hit ? h.u : 0.0f, ry = hit ? h.v : 0.0f, rz = hit ? 1.0f - (h.u + h.v) : 0.0f;
        uint32_t r = min(255u, __float2uint_rz(255.0f * rx));
        uint32_t g = min(255u, __float2uint_rz(255.0f * ry));
        uint32_t b = min(255u, __float2uint_rz(255.0f * rz));
        return r | (g << 8) | (b << 16) | 0xFF000000u;
    }
    if (P.shade_kind == 4u) {
        // NormalMappingAccumulator (renderer.rs:256-286): normals[prim] of object 0's model (instance index is always 0),
        // n = n0 * (1 - u - v) + n1 * u + n2 * v, object 0's transform_vector, normalize, (n + 1) * 0.5
        float rx = 0.0f, ry = 0.0f, rz = 0.0f;
        uint32_t prim = h.id & 0x000FFFFFu;
        if (hit && prim < P.shade_n_prims) {
            const float4* np = P.shade_normals + 3 * (size_t)prim;
            float4 a = __ldg(np + 0), b = __ldg(np + 1), c = __ldg(np + 2);
            float w0 = __fsub_rn(__fsub_rn(1.0f, h.u), h.v);
            float mx = __fadd_rn(__fadd_rn(__fmul_rn(a.x, w0), __fmul_rn(b.x, h.u)), __fmul_rn(c.x, h.v));
            float my = __fadd_rn(__fadd_rn(__fmul_rn(a.y, w0), __fmul_rn(b.y, h.u)), __fmul_rn(c.y, h.v));
            float mz = __fadd_rn(__fadd_rn(__fmul_rn(a.z, w0), __fmul_rn(b.z, h.u)), __fmul_rn(c.z, h.v));
            const float* M = P.shade_m;     // c0 = M[0..2], c1 = M[3..5], c2 = M[6..8], c3 = M[9..11]
            float wx = BVHT_MV4(M[0], M[3], M[6], M[9], mx, my, mz, 0.0f);
            float wy = BVHT_MV4(M[1], M[4], M[7], M[10], mx, my, mz, 0.0f);
            float wz = BVHT_MV4(M[2], M[5], M[8], M[11], mx, my, mz, 0.0f);
            float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)), __fmul_rn(wz, wz)));
            rx = __fmul_rn(__fadd_rn(__fdiv_rn(wx, len), 1.0f), 0.5f);
            ry = __fmul_rn(__fadd_rn(__fdiv_rn(wy, len), 1.0f), 0.5f);
            rz = __fmul_rn(__fadd_rn(__fdiv_rn(wz, len), 1.0f), 0.5f);
        }
        uint32_t r = min(255u, __float2uint_rz(__fmul_rn(255.0f, rx)));
        uint32_t g = min(255u, __float2uint_rz(__fmul_rn(255.0f, ry)));
        uint32_t b = min(255u, __float2uint_rz(__fmul_rn(255.0f, rz)));
        return r | (g << 8) | (b << 16) | 0xFF000000u;
    }
    if (P.shade_kind == 5u) {
        // TextureMaterialAccumulator (renderer.rs:297-333): tex_coords[prim] of object 0's model, uv = t0 * (1 - u - v) +
        // t1 * u + t2 * v, nearest texel of object 0's texture (material.rs:44-52), texel * (1 / 256).
        float rx = 0.0f, ry = 0.0f, rz = 0.0f;
        uint32_t prim = h.id & 0x000FFFFFu;
        if (hit && prim < P.shade_n_prims) {
            const float2* tp = P.shade_tex + 3 * (size_t)prim;
            float2 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
            float w0 = __fsub_rn(__fsub_rn(1.0f, h.u), h.v);
            float tu = __fadd_rn(__fadd_rn(__fmul_rn(a.x, w0), __fmul_rn(b.x, h.u)), __fmul_rn(c.x, h.v));
            float tv = __fadd_rn(__fadd_rn(__fmul_rn(a.y, w0), __fmul_rn(b.y, h.u)), __fmul_rn(c.y, h.v));
            uint32_t iu = texel_index(__fmul_rn(tu, (float)P.tex_w), P.tex_w);
            uint32_t iv = texel_index(__fmul_rn(tv, (float)P.tex_h), P.tex_h);
            const uint8_t* px = P.shade_texels + ((size_t)iv * P.tex_w + iu) * 3;
            const float s = 1.0f / 256.0f;
            rx = __fmul_rn((float)__ldg(px + 0), s); ry = __fmul_rn((float)__ldg(px + 1), s); rz = __fmul_rn((float)__ldg(px + 2), s);
        }
        uint32_t r = min(255u, __float2uint_rz(__fmul_rn(255.0f, rx)));
        uint32_t g = min(255u, __float2uint_rz(__fmul_rn(255.0f, ry)));
        uint32_t b = min(255u, __float2uint_rz(__fmul_rn(255.0f, rz)));
        return r | (g << 8) | (b << 16) | 0xFF000000u;
    }
    return 0u;
}

// Per pixel block (one warp): skip[x] for every TLAS node under the block's candidate mask (scene_intersect, "Chain
// skipping").  Lane j owns node j (and j + 32); `rounds` = ceil(log2(tree depth)) pointer-jumping rounds -- with shuffles
// when the tree has at most 32 nodes (<= 16 instances), else over the byte table in shared memory.
__device__ __forceinline__ uint32_t tlas_skip_init(const SceneDev& S, uint32_t id, uint32_t cand) {
    uint32_t lr = __float_as_uint(__ldg(reinterpret_cast<const float*>(S.tlas + 2 * (size_t)id) + 3));
    uint32_t sn = id;
    if (lr != 0u) {
        uint32_t li = lr >> 16, ri = lr & 0xFFFFu;
        bool ml = (__ldg(S.tlas_mask + li) & cand) != 0u, mr = (__ldg(S.tlas_mask + ri) & cand) != 0u;
        if (ml != mr) sn = ml ? li : ri;
    }
    return sn;
}

__device__ __forceinline__ void build_tlas_skip(const SceneDev& S, uint32_t n_nodes, uint32_t rounds, uint32_t cand, uint8_t* skip,
                                                unsigned lane) {
    if (n_nodes <= 32u) {
        uint32_t sn = lane < n_nodes ? tlas_skip_init(S, lane, cand) : lane;
        for (uint32_t r = 0; r < rounds; ++r) sn = __shfl_sync(0xFFFFFFFFu, sn, (int)sn);
        __syncwarp();                               // the previous block's walks are over
        skip[lane] = (uint8_t)sn;
        __syncwarp();
        return;
    }
    __syncwarp();
    for (uint32_t id = lane; id < n_nodes; id += 32u) skip[id] = (uint8_t)tlas_skip_init(S, id, cand);
    __syncwarp();
    for (uint32_t r = 0; r < rounds; ++r) {
        uint32_t a = lane < n_nodes ? skip[skip[lane]] : 0u;
        uint32_t b = lane + 32u < n_nodes ? skip[skip[lane + 32u]] : 0u;
        __syncwarp();
        if (lane < n_nodes) skip[lane] = (uint8_t)a;
        if (lane + 32u < n_nodes) skip[lane + 32u] = (uint8_t)b;
        __syncwarp();
    }
}

// K1: persistent, tile-pulling primary closest-hit kernel.
// 32-pixel slices pulled per work-counter atomic.  Measured on B200 (tools/sweep_grab.sh): 2/4/8 cut the all-miss frame
// from 0.197 to 0.148 ms (the single counter serialises in L2) but cost real frames 2 % / 6 % / 19 % through load imbalance
// (neighbouring slices are similarly expensive), so the finest granularity stays the default.
#ifndef BVHT_GRAB
#define BVHT_GRAB 1
#endif
// PRUNE: chain-skipping TLAS walk (a separate instantiation, chosen by the host when the scene qualifies: its extra
// registers cost single-instance scenes 2-3 % otherwise).
template <bool ACCEL, bool PRUNE = false>
__global__ void __launch_bounds__(128, BVHT_MIN_BLOCKS)
trace_primary_kernel(const __grid_constant__ PrimaryParams P) {
    const unsigned lane = threadIdx.x & 31u;
    Stat st;
    __shared__ uint8_t s_skip[PRUNE ? 4 : 1][64];
    // Bands (device_types.cuh): s_band_end[j] = END of pull position j's run of blocks (inclusive prefix sum of the bands' block
    // counts in pull order; shared memory, not registers: the trace below needs every one of its 64).  With a work list (K0 ran
    // first) only the listed blocks are pulled; the others already hold their miss records.
    __shared__ unsigned s_band_end[32];
    __shared__ unsigned s_slot[4];
    if (threadIdx.x < 32u) {
        unsigned cnt = 0, band = 0;
        if (lane < P.n_bands) {
            band = P.band_order[lane];
            if (P.work_list) cnt = __ldcg(P.band_count + band);
            else {
                const unsigned lo = band * P.band_items;
                cnt = lo < P.n_items ? min(P.band_items, P.n_items - lo) : 0u;
            }
        }
        unsigned end = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned v = __shfl_up_sync(0xFFFFFFFFu, end, off);
            if (lane >= (unsigned)off) end += v;
        }
        s_band_end[lane] = end;
        // a band without any block to trace is complete as soon as K0 is (this kernel started after it)
        if (P.band_done && blockIdx.x == 0 && lane < P.n_bands && cnt == 0u)
            *reinterpret_cast<volatile unsigned*>(P.band_flag + band) = P.band_seq;
    }
    __syncthreads();
    const unsigned n_work = s_band_end[31];
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(P.work_counter, (unsigned)BVHT_GRAB);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n_work) break;
#pragma unroll 1
      for (unsigned g = 0; g < (unsigned)BVHT_GRAB; ++g) {
        if (base + g >= n_work) break;
        // pull position -> band -> block (row-major inside a band: strided permutations measured 2-15 % slower)
        const unsigned pos = base + g;
        unsigned item;
        {
            const unsigned slot = (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, pos >= s_band_end[lane]));
            const unsigned slot_begin = slot ? s_band_end[slot - 1] : 0u;
            const unsigned in_band = (unsigned)P.band_order[slot] * P.band_items + (pos - slot_begin);
            item = P.work_list ? __ldcg(P.work_list + in_band) : in_band;
            if (lane == 0) s_slot[threadIdx.x >> 5] = slot;
        }
#ifdef BVHT_SLICE_LOG
        unsigned long long log_t0 = 0;                 // profiling build only (tools/slice_timeline.py): start / end of every block
        if (P.stats) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(log_t0));
#endif
        uint32_t tile_idx = item / P.items_per_tile;
        uint32_t sub = item - tile_idx * P.items_per_tile;
        uint32_t tx = P.tx0 + tile_idx % P.ntx;
        uint32_t ty = P.ty0 + (tile_idx / P.ntx) * P.row_stride;
        uint32_t p = sub * 32u + lane;                 // pixel within the tile, row-major (renderer.rs:356-357)
        uint32_t iu = p % P.tile, iv = p / P.tile;
        uint32_t px = tx * P.tile + iu, py = ty * P.tile + iv;
        bool active = (iv < P.tile) && px >= P.x0 && px < P.x1 && py >= P.y0 && py < P.y1;
        if (lane == 0) st.add(13);
        uint32_t cand = 0xFFFFFFFFu;
        if (ACCEL) {
            if (P.n_rect) {
                // which instances can this warp's pixel block see at all?  lane i tests instance i's screen rectangle
                int bx0 = (int)(tx * P.tile), bx1 = bx0 + (int)P.tile - 1;
                int by0 = (int)(ty * P.tile + (sub * 32u) / P.tile), by1 = (int)(ty * P.tile + (sub * 32u + 31u) / P.tile);
                bool ov = false;
                if (lane < P.n_rect) {
                    int4 rc = P.inst_rect[lane];
                    ov = !(rc.z < bx0 || rc.x > bx1 || rc.w < by0 || rc.y > by1);
                }
                cand = __ballot_sync(0xFFFFFFFFu, ov);
                // per-triangle coverage (cover_kernels.cu): instances none of whose triangles can be seen from this block
                if (P.cover) cand &= __ldg(P.cover + ((size_t)ty * P.cover_ntx + tx) * 2u + sub) | __ldg(P.cover_full);
            }
        }
        const uint8_t* skip = nullptr;
        if (ACCEL && PRUNE) {
            if (P.n_tlas_nodes != 0u && cand != 0u && cand != 0xFFFFFFFFu) {       // warp-uniform
                build_tlas_skip(P.scene, P.n_tlas_nodes, P.skip_rounds, cand, s_skip[threadIdx.x >> 5], lane);
                skip = s_skip[threadIdx.x >> 5];
                if (lane == 0) st.add(14);
            }
        }
        if (active) {
            HitRec h;
            if (ACCEL && cand == 0u) {
                h.t = FLT_MAX; h.u = 0.0f; h.v = 0.0f; h.id = 0xFFFFFFFFu;      // no instance can be seen from this block
                st.add(12);
            } else {
                RayM w = primary_ray(P.cam, px, py, P.width, P.height);
                h = scene_intersect<ACCEL, PRUNE>(P.scene, w, FLT_MAX, cand, st, skip, P.n_origin ? P.inst_origin : nullptr);
            }
            if (P.out) {
                uint4 o;
                o.x = __float_as_uint(h.t); o.y = __float_as_uint(h.u); o.z = __float_as_uint(h.v); o.w = h.id;
                store_hit_k1(P.out + (size_t)py * P.width + px, o);
            }
            if (P.out_rgba) store_px_k1(P.out_rgba + (size_t)py * P.width + px, shade_pixel(P, h));
        }
        if (P.band_done) {
            // A block counts as finished only when its pixels are in memory.  A warp-wide __threadfence() here cost 0.09 ms of a
            // 0.80 ms frame (profiles/r02_e2e_timeline.txt); instead the warp synchronises (the lanes' stores happen-before lane 0's
            // next operation) and lane 0 counts the block with a RELEASE read-modify-write at GPU scope, which is cumulative over
            // what it has observed.  Whoever counts a band's last block fences (acquire side of the chain, system scope for the
            // copy engine) and raises the band's flag.
            __syncwarp();
            if (lane == 0) {
                const unsigned slot = s_slot[threadIdx.x >> 5];
                const unsigned band_total = s_band_end[slot] - (slot ? s_band_end[slot - 1] : 0u);
                const unsigned band_id = P.band_order[slot];
                unsigned before;
                asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(before) : "l"(P.band_done + band_id) : "memory");
                if (before + 1u == band_total) {
                    __threadfence_system();
                    *reinterpret_cast<volatile unsigned*>(P.band_flag + band_id) = P.band_seq;
                }
            }
        }
#ifdef BVHT_SLICE_LOG
        __syncwarp();
        if (P.stats && lane == 0) {
            unsigned long long log_t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(log_t1));
            unsigned smid;
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            P.stats[2 * (size_t)item] = log_t0;
            P.stats[2 * (size_t)item + 1] = (log_t1 - log_t0) | ((unsigned long long)smid << 48) | ((unsigned long long)(threadIdx.x >> 5) << 44);
        }
#endif
      }
    }
#ifdef BVHT_STATS
    for (int i = 0; i < kStatSlots; ++i) {
        unsigned long long v = st.c[i];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
        if (lane == 0 && v) atomicAdd(P.stats + i, v);
    }
#endif
}

// K0: classify the 32-pixel blocks of a launch and finish the empty ones.  One LANE per block: the block's pixel rectangle
// against every instance's screen rectangle (the same test K1 makes with one lane per instance).  Blocks that see an
// instance are appended to the work list (one atomicAdd per warp = per 32 blocks); for the others the warp writes the 32
// miss records (and shaded pixels) right away.  K1 alone would spend one work-counter atomic on every block, and that single
// address serialises in L2 at ~1.3 G atomics/s: an empty 4K frame costs 0.2 ms, an 8K frame 0.8 ms, of which big_ben_clock
// (87 % empty blocks) paid most.  Streaming kernel: 16 B (+ 4 B) written per empty pixel.
__global__ void __launch_bounds__(128)
classify_fill_kernel(const __grid_constant__ PrimaryParams P) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned item = blockIdx.x * 128u + threadIdx.x;          // this lane's block
    const bool valid = item < P.n_items;
    uint32_t tile_idx = 0, sub = 0;
    bool seen = false;
    if (valid) {
        tile_idx = item / P.items_per_tile;
        sub = item - tile_idx * P.items_per_tile;
        uint32_t tx = P.tx0 + tile_idx % P.ntx;
        uint32_t ty = P.ty0 + (tile_idx / P.ntx) * P.row_stride;
        int bx0 = (int)(tx * P.tile), bx1 = bx0 + (int)P.tile - 1;
        int by0 = (int)(ty * P.tile + (sub * 32u) / P.tile), by1 = (int)(ty * P.tile + (sub * 32u + 31u) / P.tile);
        uint32_t mask = 0u;
        for (uint32_t i = 0; i < P.n_rect; ++i) {
            int4 rc = P.inst_rect[i];
            if (!(rc.z < bx0 || rc.x > bx1 || rc.w < by0 || rc.y > by1)) mask |= 1u << i;
        }
        if (P.cover) mask &= __ldg(P.cover + ((size_t)ty * P.cover_ntx + tx) * 2u + sub) | __ldg(P.cover_full);
        seen = mask != 0u;
    }
    const unsigned listed = __ballot_sync(0xFFFFFFFFu, valid && seen);
    unsigned empty = __ballot_sync(0xFFFFFFFFu, valid && !seen);
    if (valid && seen) {
        // append to the block's BAND's segment of the list (device_types.cuh): the lanes of one band share one atomic
        const unsigned band = item / P.band_items;
        const unsigned peers = __match_any_sync(listed, band);
        const int leader = __ffs((int)peers) - 1;
        unsigned at = 0;
        if ((int)lane == leader) at = atomicAdd(P.band_count + band, (unsigned)__popc(peers));
        at = __shfl_sync(peers, at, leader);
        P.work_list[(size_t)band * P.band_items + at + __popc(peers & ((1u << lane) - 1u))] = item;
    }
    HitRec miss;
    miss.t = FLT_MAX; miss.u = 0.0f; miss.v = 0.0f; miss.id = 0xFFFFFFFFu;
    const uint32_t miss_px = P.out_rgba ? shade_pixel(P, miss) : 0u;
    while (empty) {
        const int src = __ffs((int)empty) - 1;
        empty &= empty - 1u;
        const uint32_t e_tile = __shfl_sync(0xFFFFFFFFu, tile_idx, src), e_sub = __shfl_sync(0xFFFFFFFFu, sub, src);
        uint32_t tx = P.tx0 + e_tile % P.ntx;
        uint32_t ty = P.ty0 + (e_tile / P.ntx) * P.row_stride;
        uint32_t p = e_sub * 32u + lane;
        uint32_t iu = p % P.tile, iv = p / P.tile;
        uint32_t px = tx * P.tile + iu, py = ty * P.tile + iv;
        if ((iv < P.tile) && px >= P.x0 && px < P.x1 && py >= P.y0 && py < P.y1) {
            if (P.out) store_hit_k0(P.out + (size_t)py * P.width + px, make_uint4(__float_as_uint(miss.t), 0u, 0u, miss.id));
            if (P.out_rgba) store_px_k0(P.out_rgba + (size_t)py * P.width + px, miss_px);
        }
    }
}

// K1b: Scene::intersect(&Ray) for an arbitrary ray buffer; warps pull 32 rays at a time.
template <bool ACCEL>
__global__ void __launch_bounds__(128, BVHT_MIN_BLOCKS)
trace_rays_kernel(const __grid_constant__ RaysParams P) {
    const unsigned lane = threadIdx.x & 31u;
    Stat st;
    const uint64_t n_items = (P.n + 31u) / 32u;
    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(P.work_counter, 1u);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= n_items) break;
        uint64_t i = (uint64_t)item * 32u + lane;
        if (i < P.n) {
            const float* rp = P.rays + i * 7;
            RayM w;
            w.ox = __ldg(rp + 0); w.oy = __ldg(rp + 1); w.oz = __ldg(rp + 2);
            w.dx = __ldg(rp + 3); w.dy = __ldg(rp + 4); w.dz = __ldg(rp + 5);
            float t = __ldg(rp + 6);
            w.rdx = __fdiv_rn(1.0f, w.dx); w.rdy = __fdiv_rn(1.0f, w.dy); w.rdz = __fdiv_rn(1.0f, w.dz);
            HitRec h = scene_intersect<ACCEL>(P.scene, w, t, 0xFFFFFFFFu, st);
            uint4 o;
            o.x = __float_as_uint(h.t); o.y = __float_as_uint(h.u); o.z = __float_as_uint(h.v); o.w = h.id;
            P.out[i] = o;
        }
    }
}

} // namespace BVHT_MODE_NS
} // namespace bvht
