// launchers.hpp -- host-callable entry points of the separately compiled kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "device_types.cuh"

namespace bvht {

#define BVHT_DECLARE_MODE(sfx)                                                                                   \
    cudaError_t launch_primary_##sfx(const PrimaryParams& p, bool accel, int grid, int block, cudaStream_t s,   \
                                     cudaEvent_t k1_begin = nullptr, cudaEvent_t k1_end = nullptr);              \
    cudaError_t launch_rays_##sfx(const RaysParams& p, bool accel, int grid, int block, cudaStream_t s);         \
    int blocks_per_sm_primary_##sfx(bool accel, bool prune, int block);                                                      \
    int blocks_per_sm_rays_##sfx(bool accel, int block);
BVHT_DECLARE_MODE(strict)
BVHT_DECLARE_MODE(fast)
BVHT_DECLARE_MODE(stats)
#undef BVHT_DECLARE_MODE

// upload_kernels.cu
cudaError_t launch_repack_triangles(const float* tris_aos, uint32_t n_tris, float4* out, cudaStream_t s);
cudaError_t launch_repack_sub_triangles(const float* tris_aos, const uint32_t* sub_order, uint32_t n, float4* out, cudaStream_t s);

// whole-model statistics / tight box of a model on the device (upload_kernels.cu); maxima and boxes as order-preserving integers
struct ModelStatsDev {
    unsigned long long radius2, max_edge2, kappa2_max;      // enc_f64 of the maxima of the SQUARES (0 = no triangle)
    double sum_edge, sum_kappa;
    unsigned long long n_good;
    unsigned int lo[3], hi[3];                               // enc_f32 of the vertex box of the non-degenerate triangles
};
cudaError_t launch_model_stats(const float* tris_aos, uint32_t n_tris, ModelStatsDev* out, cudaStream_t s);
cudaError_t launch_tight_box(const float* tris_aos, uint32_t n_tris, double scale, double abs_, unsigned long long* out6, cudaStream_t s);

// up to 8 word copies (from page-locked host memory) / word fills in one launch
struct SmallOp { void* dst; const void* src; uint32_t words; uint32_t fill; };
struct SmallOps { SmallOp op[8]; uint32_t n; };
cudaError_t launch_small_ops(const SmallOps& ops, cudaStream_t s);

cudaError_t launch_read_bw(const void* buf, size_t bytes, uint32_t iters, int grid, unsigned long long* sink, cudaStream_t s);
cudaError_t launch_ray_bounds(const float* rays, unsigned long long n, unsigned int* out2, cudaStream_t s);
cudaError_t launch_refit_sub_nodes(float4* raw, const uint32_t* parent, unsigned int* counters, const float* tris_aos,
                                   const uint32_t* order, uint32_t n_nodes, cudaStream_t s);
cudaError_t launch_bake_sub_nodes(const float4* raw, const uint32_t* parent, unsigned int* counters, const float* tris_aos,
                                  const uint32_t* order, uint32_t n_nodes, float scale, float abs_, float4* lohi_scratch, float4* out,
                                  cudaStream_t s);

// cover_kernels.cu
cudaError_t launch_raster_cover(const CoverParams& p, cudaStream_t s);

// scene_kernels.cu
size_t      scene_rebuild_smem_bytes(uint32_t n_inst);
cudaError_t launch_scene_rebuild(const SceneRebuildParams& p, cudaStream_t s);

// refit_kernels.cu
struct RefitPlan {
    float4*         nodes;          // device node pool (2 float4 per node)
    const float*    tris_aos;       // n_tris x 9 floats (reference layout)
    const uint32_t* chunk_leaf;     // per chunk: leaf node index
    const uint32_t* chunk_first;    // per chunk: first triangle
    const uint32_t* chunk_count;    // per chunk: triangle count
    const uint32_t* leaf_chunks;    // per node: number of chunks (leaves only)
    const uint32_t* parent;         // per node: parent index (root: 0xFFFFFFFF)
    float*          scratch;        // per node: 6 floats of running min/max
    unsigned int*   counters;       // per node: chunk / child arrival counters
    uint32_t        n_chunks;
    uint32_t        nodes_used;
};
cudaError_t launch_refit(const RefitPlan& plan, cudaStream_t s);

} // namespace bvht
