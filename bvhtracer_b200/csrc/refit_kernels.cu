// refit_kernels.cu -- K2: bottom-up BLAS refit with atomic arrival counters.
//
// Replaces Bvh::refit (model/bvh.rs:469-493) and update_node_bounds (bvh.rs:317-330): the reference sweeps
// node indices in reverse (children always have larger indices than their parent, bvh.rs:601-612), recomputing
// each leaf's box from its triangles' vertices and each interior box as the union of its two children.
// min/max are exact and order-independent, so any evaluation order gives the reference's boxes bit for bit
// (up to the sign of a zero, which no comparison can observe).
//
// The reference's leaves are huge (up to 9,404 triangles), so the leaf stage is a streaming
// reduction: every leaf is cut into chunks of kChunkTris triangles; one CTA reduces one chunk (coalesced float
// loads, warp shuffles), merges into the leaf's running box with float atomics, and bumps the leaf's arrival
// counter.  The CTA that completes a leaf publishes the box and climbs: at every parent the first arriving
// child stops, the second merges both children and continues (no second launch, no grid sync).
// HBM-bound: 36 B read per triangle; node traffic is negligible.
#include <cfloat>
#include "launchers.hpp"

namespace bvht {
namespace {

constexpr int kRefitBlock = 256;

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    v += 0.0f;                                    // -0.0 -> +0.0 (int ordering trick below needs a canonical zero)
    if (v >= 0.0f) atomicMin((int*)addr, __float_as_int(v));
    else           atomicMax((unsigned int*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    v += 0.0f;
    if (v >= 0.0f) atomicMax((int*)addr, __float_as_int(v));
    else           atomicMin((unsigned int*)addr, __float_as_uint(v));
}

__global__ void refit_init_kernel(float* scratch, unsigned int* counters, uint32_t nodes_used) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nodes_used) return;
    scratch[i * 6 + 0] = FLT_MAX;  scratch[i * 6 + 1] = FLT_MAX;  scratch[i * 6 + 2] = FLT_MAX;
    scratch[i * 6 + 3] = -FLT_MAX; scratch[i * 6 + 4] = -FLT_MAX; scratch[i * 6 + 5] = -FLT_MAX;
    counters[i] = 0u;
}

__device__ __forceinline__ void store_node_box(float4* nodes, uint32_t ni, const float mn[3], const float mx[3]) {
    float* lo = reinterpret_cast<float*>(nodes + 2 * (size_t)ni);
    float* hi = reinterpret_cast<float*>(nodes + 2 * (size_t)ni + 1);
    // .w of each float4 holds left_first / prim_count: untouched
    __stcg(lo + 0, mn[0]); __stcg(lo + 1, mn[1]); __stcg(lo + 2, mn[2]);
    __stcg(hi + 0, mx[0]); __stcg(hi + 1, mx[1]); __stcg(hi + 2, mx[2]);
}

__global__ void __launch_bounds__(kRefitBlock)
refit_kernel(const RefitPlan P) {
    __shared__ float red[6][kRefitBlock / 32];
    const uint32_t chunk = blockIdx.x;
    const uint32_t leaf = __ldg(P.chunk_leaf + chunk);
    const uint32_t first = __ldg(P.chunk_first + chunk);
    const uint32_t count = __ldg(P.chunk_count + chunk);

    // flat float index j over the chunk's 9*count floats; axis = j % 3 (x,y,z repeat with period 3)
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
    float mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    const float* src = P.tris_aos + (size_t)first * 9;
    const uint32_t n_floats = count * 9u;
    for (uint32_t j = threadIdx.x; j < n_floats; j += kRefitBlock) {
        float x = __ldg(src + j);
        uint32_t axis = j % 3u;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (axis == (uint32_t)k) { mn[k] = fminf(mn[k], x); mx[k] = fmaxf(mx[k], x); }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xFFFFFFFFu, mn[k], off));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xFFFFFFFFu, mx[k], off));
        }
    }
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { red[k][warp] = mn[k]; red[3 + k][warp] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int w = 1; w < kRefitBlock / 32; ++w) { mn[k] = fminf(mn[k], red[k][w]); mx[k] = fmaxf(mx[k], red[3 + k][w]); }
    }
    float* sc = P.scratch + (size_t)leaf * 6;
#pragma unroll
    for (int k = 0; k < 3; ++k) { atomic_min_float(sc + k, mn[k]); atomic_max_float(sc + 3 + k, mx[k]); }
    __threadfence();
    unsigned int arrived = atomicAdd(P.counters + leaf, 1u);
    if (arrived + 1u != __ldg(P.leaf_chunks + leaf)) return;

    // this CTA completed the leaf: publish its box, then climb
    __threadfence();
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] = __ldcg(sc + k); mx[k] = __ldcg(sc + 3 + k); }
    store_node_box(P.nodes, leaf, mn, mx);
    uint32_t node = leaf;
    for (;;) {
        uint32_t parent = __ldg(P.parent + node);
        if (parent == 0xFFFFFFFFu) break;
        __threadfence();
        unsigned int prev = atomicAdd(P.counters + parent, 1u);
        if (prev == 0u) break;                     // the sibling subtree is still being refitted
        __threadfence();
        // children of `parent` are adjacent: left = left_first, right = left_first + 1
        uint32_t lf = __float_as_uint(__ldcg(reinterpret_cast<const float*>(P.nodes + 2 * (size_t)parent) + 3));
        const float* l_lo = reinterpret_cast<const float*>(P.nodes + 2 * (size_t)lf);
        const float* l_hi = reinterpret_cast<const float*>(P.nodes + 2 * (size_t)lf + 1);
        const float* r_lo = reinterpret_cast<const float*>(P.nodes + 2 * (size_t)(lf + 1));
        const float* r_hi = reinterpret_cast<const float*>(P.nodes + 2 * (size_t)(lf + 1) + 1);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mn[k] = fminf(__ldcg(l_lo + k), __ldcg(r_lo + k));     // bvh.rs:489-490
            mx[k] = fmaxf(__ldcg(l_hi + k), __ldcg(r_hi + k));
        }
        store_node_box(P.nodes, parent, mn, mx);
        node = parent;
    }
}

} // namespace

cudaError_t launch_refit(const RefitPlan& plan, cudaStream_t s) {
    if (plan.nodes_used == 0) return cudaSuccess;
    int ib = 128;
    refit_init_kernel<<<(plan.nodes_used + ib - 1) / ib, ib, 0, s>>>(plan.scratch, plan.counters, plan.nodes_used);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (plan.n_chunks == 0) return cudaSuccess;
    refit_kernel<<<plan.n_chunks, kRefitBlock, 0, s>>>(plan);
    return cudaGetLastError();
}

} // namespace bvht
