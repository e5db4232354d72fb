// leaf_accel.hpp -- host-side build of the conservative sub-BVHs placed INSIDE the reference's leaves.
//
// Not a reference component.  The reference's BVHs have giant leaves (armadillo: 48 leaves, mean 625 / max
// 5,157 triangles) because its SAH sweep boxes start at the origin (bvh.rs:365-366), and
// Bvh::intersect_subtree brute-forces every triangle of every visited leaf (bvh.rs:250-258).  Results must stay
// identical, so the reference tree, its child order and its leaf boundaries are kept exactly; only the inner
// loop of one leaf is replaced by an equivalent search:
//
//   reference leaf result = lexicographic min (t, primitive index) over the triangles that
//   Triangle::intersect accepts (against the ENTRY ray) with t < closest-at-leaf-entry.
//
// A binary BVH over the leaf's triangles finds the same minimum provided it never skips an accepting
// triangle.  Triangle::intersect in f32 accepts rays that miss the exact triangle by a small residual
//   |o + t d - (v0 + u e1 + v e2)|  <~  c * eps * |d| |e1| |e2| (|o - v0| + |e|) / |det|,   |det| >= 1e-4,
// so every sub box is inflated by that bound evaluated at the limits D_MAX, O_MAX below (plus the slab
// test's own rounding), and the trace kernel only uses the sub-BVH for rays inside those limits
// (|d| <= D_MAX, |o| <= O_MAX in model space); other rays take the brute-force leaf.  tests/ compare
// accel against brute force bit-for-bit on every config.
#pragma once
#include <cstdint>
#include <vector>

namespace bvht {

struct LeafAccelHost {
    // RAW (un-inflated) sub nodes, 16 floats each: c0.lo.xyz, kappa0 | c0.hi.xyz, ref0 | c1.lo.xyz, kappa1 | c1.hi.xyz, ref1
    // kappa = max |e1| |e2| over the child's triangles.  The device inflates them for the current ray limits
    // (bake_sub_nodes_kernel, bottom-up from the triangles, per-triangle inflation), see accel_deltas().
    std::vector<float>    sub_raw;
    std::vector<uint32_t> sub_parent;     // per sub node: (parent << 1) | child slot, 0xFFFFFFFF for the root of a leaf's tree
    std::vector<uint32_t> order;          // sub position -> reference primitive index
    std::vector<uint32_t> leaf_sub_root;  // per reference node: sub root node index, 0xFFFFFFFF = brute force
    uint32_t max_depth = 0;
    double radius = 1.0;                  // max vertex norm over non-degenerate triangles
    double max_edge = 0.0;
    // raw box of the NON-degenerate triangles of the whole model (the 999-sentinel is degenerate: never accepted)
    float  model_lo[3] = { 0, 0, 0 }, model_hi[3] = { 0, 0, 0 };
    double model_kappa = 0.0;
    bool   model_valid = false;
};

struct LeafAccelConfig {
    uint32_t min_leaf_tris = 12;   // reference leaves smaller than this stay brute force
    uint32_t max_sub_leaf = 1;     // triangles per sub leaf (<= 8: 3-bit count field); 1 measured fastest on B200
    float    d_max = 2.0f;         // default |d| limit in model space (instance scale >= 0.5)
    float    o_max_radii = 16.0f;  // default |o| limit as a multiple of the model radius
    float    c_mt = 48.0f;         // constant of the Moeller-Trumbore residual bound: 34.5 by first-order error analysis
                                   //   (below), x1.4 safety; tools/equivalence_sweep.py finds no mismatch even at c = 0
};

// Inflation of a box whose triangles have edge product <= kappa, valid for model-space rays with |d| <= d_max and
// |o| <= o_max:  delta = scale * kappa + abs.
//   scale * kappa : residual |o + t d - (v0 + u e1 + v e2)| of Triangle::intersect in f32 at its |det| >= 1e-4 cut-off,
//                   c_mt * eps * d_max * (o_max + radius + max_edge) * kappa / 1e-4.
//                   Derivation (first order in eps = 2^-24, s = o - v0, D = |d||e1||e2|): with the exact identity
//                   det*s + T*d - U*e1 - V*e2 = 0 the residual is (ddet*s + dT*d - dU*e1 - dV*e2) / det, and the rounding
//                   errors of the reference's cross/dot sequences are |ddet| <= 9 eps D, |dU|, |dV|, |dT| <= 8.5 eps |s| times
//                   the other two lengths, so |residual| <= (9 + 3 * 8.5) eps |s| D / |det| = 34.5 eps |s| D / |det|.
//   abs           : rounding of s = o - v0 and of the slab test itself, 16 * eps * (o_max + radius + max_edge)
inline void accel_deltas(const LeafAccelConfig& cfg, double d_max, double o_max, double radius, double max_edge,
                         double& scale, double& abs_) {
    const double eps = 5.9604644775390625e-08;      // 2^-24
    double s_max = o_max + radius + max_edge;
    scale = cfg.c_mt * eps * d_max * s_max / 1e-4;
    abs_ = 16.0 * eps * s_max;
}

// Whole-model statistics the inflation depends on (recomputed on the host after every vertex update: one O(n) pass).
struct ModelStats {
    double radius = 1.0, max_edge = 0.0, model_kappa = 0.0;
    double mean_edge = 0.0, mean_kappa = 0.0;     // over the non-degenerate triangles: "typical" size for the usefulness cap
    float  model_lo[3] = { 0, 0, 0 }, model_hi[3] = { 0, 0, 0 };
    bool   model_valid = false;
};
ModelStats compute_model_stats(const float* tris, uint32_t n_tris, std::vector<float>* kappa_out = nullptr);

// tris: n_tris x 9 floats in reference order; nodes: reference nodes (bvht_bvh_node layout: min[3], max[3], count, left_first)
bool build_leaf_accel(const float* tris, uint32_t n_tris, const void* nodes, uint32_t nodes_used,
                      const LeafAccelConfig& cfg, LeafAccelHost& out);

} // namespace bvht
