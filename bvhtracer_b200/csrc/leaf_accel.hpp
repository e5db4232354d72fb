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
    std::vector<float>    sub_nodes;      // 16 floats per sub node: c0.lo.xyz, ref0 | c0.hi.xyz, ref1 | c1.lo.xyz, 0 | c1.hi.xyz, 0
    std::vector<uint32_t> order;          // sub position -> reference primitive index
    std::vector<uint32_t> leaf_sub_root;  // per reference node: sub root node index, 0xFFFFFFFF = brute force
    float d_max = 0.0f;                   // limits under which the inflation is valid
    float o_max = 0.0f;
    uint32_t max_depth = 0;
    // box of the NON-degenerate triangles of the whole model, inflated like the sub boxes: under the same limits no
    // triangle can be accepted by a ray that misses it (the 999-sentinel is degenerate and never accepted)
    float tight_lo[3] = { 0, 0, 0 }, tight_hi[3] = { 0, 0, 0 };
    bool  tight_valid = false;
};

struct LeafAccelConfig {
    uint32_t min_leaf_tris = 12;   // reference leaves smaller than this stay brute force
    uint32_t max_sub_leaf = 4;     // triangles per sub leaf (<= 8: 3-bit count field)
    float    d_max = 2.0f;         // |d| limit in model space (instance scale >= 0.5)
    float    o_max_radii = 16.0f;  // |o| limit as a multiple of the model radius
    float    c_mt = 80.0f;         // safety constant of the Moeller-Trumbore residual bound (first-order estimate ~40)
};

// tris: n_tris x 9 floats in reference order; nodes: reference nodes (bvht_bvh_node layout: min[3], max[3], count, left_first)
bool build_leaf_accel(const float* tris, uint32_t n_tris, const void* nodes, uint32_t nodes_used,
                      const LeafAccelConfig& cfg, LeafAccelHost& out);

} // namespace bvht
