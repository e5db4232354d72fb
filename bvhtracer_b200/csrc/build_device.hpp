// build_device.hpp -- device-side BVH construction (build_kernels.cu): interface towards bvht_api.cu.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

namespace bvht {

// node under construction (creation-order numbering); min/max are order-preserving u32 encodings of the floats
struct BuildNode {
    uint32_t vmin[3], vmax[3];     // vertex bounds (update_node_bounds, bvh.rs:317-330)
    uint32_t cmin[3], cmax[3];     // centroid bounds per axis (find_best_split_plane, bvh.rs:339-342)
    uint32_t first, count;         // primitive range
    uint32_t left;                 // creation-order index of the left child (right = left + 1); 0 = leaf
    uint32_t depth;
};
static_assert(sizeof(BuildNode) == 64, "BuildNode is 64 bytes");

// finished node in the reference's numbering and layout (bvht_bvh_node / BvhNode, bvh.rs:88-134)
struct BuildNodeHost {
    float aabb_min[3] = { 0, 0, 0 }, aabb_max[3] = { 0, 0, 0 };
    uint32_t prim_count = 0, left_first = 0;
};
static_assert(sizeof(BuildNodeHost) == 32, "BuildNodeHost must match bvht_bvh_node");

enum { kCtrNodes = 0, kCtrNext = 1, kCtrChunks = 2, kCtrDepth = 3, kCtrCount = 4 };

struct DeviceBuildStats { uint32_t levels = 0, launches = 0, temp_nodes = 0, max_depth = 0; };

// grow-only device scratch, reused across builds of one context
class BuildWorkspace {
public:
    BuildWorkspace();
    ~BuildWorkspace();
    BuildWorkspace(const BuildWorkspace&) = delete;
    BuildWorkspace& operator=(const BuildWorkspace&) = delete;
    void release();
    struct Impl;
    Impl* impl_;
};

// tris: n_tris x 9 floats on the device, reordered IN PLACE exactly like BvhBuilder::build_for reorders the mesh;
// perm: n_tris u32 on the device, perm[i] = original index of the triangle that ends at position i;
// out_nodes: the reference's node pool (first nodes_used entries: node 1 is the alignment dummy).
cudaError_t device_build_reference_bvh(BuildWorkspace& ws, float* tris, uint32_t* perm, uint32_t n_tris, cudaStream_t s,
                                       std::vector<BuildNodeHost>& out_nodes, DeviceBuildStats* stats);

// ---- leaf accelerator (leaf_accel.hpp) on the device
struct SubRoot { uint32_t first_prim, count, base; };     // one accelerated reference leaf: its primitives, its slice of the index array
struct DeviceSubResult { uint32_t n_nodes = 0, max_depth = 0, levels = 0, launches = 0; };

// tris: the model's triangles on the device in the reference's (BVH) order; roots_host: the accelerated leaves (root r becomes
// sub node r); n_sub = sum of the counts.  Outputs (device): order[n_sub] (sub position -> primitive index), sub_raw (4 float4 per sub
// node, only the child references are set: run refit_sub_nodes next), sub_parent.  Capacity of sub_raw / sub_parent: n_sub nodes.
cudaError_t device_build_leaf_accel(BuildWorkspace& ws, const float* tris, const SubRoot* roots_host, uint32_t n_roots, uint32_t n_sub,
                                    float4* sub_raw, uint32_t* sub_parent, uint32_t* order, uint32_t max_sub_leaf, uint32_t sah_depth_limit,
                                    cudaStream_t s, DeviceSubResult* res);

} // namespace bvht
