// upload_kernels.cu -- K3: repack the reference's triangle buffer (Triangle<f32> = 3 x Vector3, 36 B,
// geometry/triangle.rs:9-16 viewed through Mesh::primitives, mesh.rs:126-134) into 3 aligned float4 per triangle
//   {v0, prim} | {e1 = v1 - v0, 0} | {e2 = v2 - v0, 0}
// The two edge subtractions are the first two operations of Triangle::intersect (triangle.rs:43-44); doing
// them once on upload is bit-identical (one IEEE subtraction either way; no FMA can form here).
// HBM-bound streaming kernel: 36 B read + 48 B written per triangle, loads coalesced through shared memory.
#include "launchers.hpp"

namespace bvht {
namespace {

constexpr int kRepackBlock = 256;

__global__ void __launch_bounds__(kRepackBlock)
repack_triangles_kernel(const float* __restrict__ tris, uint32_t n_tris, float4* __restrict__ out) {
    __shared__ float stage[kRepackBlock * 9];
    const uint32_t first = blockIdx.x * kRepackBlock;
    const uint32_t count = min((uint32_t)kRepackBlock, n_tris - first);
    const float* src = tris + (size_t)first * 9;
    for (uint32_t j = threadIdx.x; j < count * 9; j += kRepackBlock) stage[j] = __ldg(src + j);   // coalesced
    __syncthreads();
    if (threadIdx.x < count) {
        const float* t = stage + threadIdx.x * 9;      // stride 9 words: conflict-free (gcd(9, 32) = 1)
        uint32_t i = first + threadIdx.x;
        float ax = t[0], ay = t[1], az = t[2];
        out[3 * (size_t)i + 0] = make_float4(ax, ay, az, __uint_as_float(i));
        out[3 * (size_t)i + 1] = make_float4(t[3] - ax, t[4] - ay, t[5] - az, 0.0f);
        out[3 * (size_t)i + 2] = make_float4(t[6] - ax, t[7] - ay, t[8] - az, 0.0f);
    }
}

// Same repack, gathered through the sub-BVH order (leaf_accel.hpp); v0.w carries the REFERENCE primitive index.
__global__ void __launch_bounds__(kRepackBlock)
repack_sub_triangles_kernel(const float* __restrict__ tris, const uint32_t* __restrict__ order, uint32_t n,
                            float4* __restrict__ out) {
    uint32_t i = blockIdx.x * kRepackBlock + threadIdx.x;
    if (i >= n) return;
    uint32_t p = __ldg(order + i);
    const float* t = tris + (size_t)p * 9;
    float ax = __ldg(t + 0), ay = __ldg(t + 1), az = __ldg(t + 2);
    out[3 * (size_t)i + 0] = make_float4(ax, ay, az, __uint_as_float(p));
    out[3 * (size_t)i + 1] = make_float4(__ldg(t + 3) - ax, __ldg(t + 4) - ay, __ldg(t + 5) - az, 0.0f);
    out[3 * (size_t)i + 2] = make_float4(__ldg(t + 6) - ax, __ldg(t + 7) - ay, __ldg(t + 8) - az, 0.0f);
}

// Bake the sub-BVH for the current ray limits: out = raw boxes grown by delta = scale * kappa + abs (leaf_accel.hpp
// accel_deltas), every operation rounded AWAY from the box (directed-rounding intrinsics), plus a relative 1e-6.
// 64 B read + 64 B written per sub node; runs at upload and whenever the limits change (bvht_api.cu ensure_bake).
__global__ void __launch_bounds__(kRepackBlock)
inflate_sub_nodes_kernel(const float4* __restrict__ raw, float4* __restrict__ out, uint32_t n_nodes, float scale, float abs_) {
    uint32_t i = blockIdx.x * kRepackBlock + threadIdx.x;
    if (i >= n_nodes) return;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        float4 lo = __ldg(raw + 4 * (size_t)i + 2 * c);        // lo.xyz, kappa
        float4 hi = __ldg(raw + 4 * (size_t)i + 2 * c + 1);    // hi.xyz, ref
        float d = __fmaf_ru(scale, lo.w, abs_);
        float4 olo, ohi;
        olo.x = __fsub_rd(__fsub_rd(lo.x, d), __fmul_ru(fabsf(lo.x), 1e-6f));
        olo.y = __fsub_rd(__fsub_rd(lo.y, d), __fmul_ru(fabsf(lo.y), 1e-6f));
        olo.z = __fsub_rd(__fsub_rd(lo.z, d), __fmul_ru(fabsf(lo.z), 1e-6f));
        ohi.x = __fadd_ru(__fadd_ru(hi.x, d), __fmul_ru(fabsf(hi.x), 1e-6f));
        ohi.y = __fadd_ru(__fadd_ru(hi.y, d), __fmul_ru(fabsf(hi.y), 1e-6f));
        ohi.z = __fadd_ru(__fadd_ru(hi.z, d), __fmul_ru(fabsf(hi.z), 1e-6f));
        // trace layout: n[0] = c0.lo | ref0, n[1] = c0.hi | ref1, n[2] = c1.lo | 0, n[3] = c1.hi | 0
        olo.w = 0.0f; ohi.w = 0.0f;
        out[4 * (size_t)i + 2 * c] = olo;
        out[4 * (size_t)i + 2 * c + 1] = ohi;
    }
    // child references live in the .w of the first two float4s
    float r0 = __ldg(raw + 4 * (size_t)i + 1).w, r1 = __ldg(raw + 4 * (size_t)i + 3).w;
    reinterpret_cast<float*>(out + 4 * (size_t)i)[3] = r0;
    reinterpret_cast<float*>(out + 4 * (size_t)i + 1)[3] = r1;
}

} // namespace

cudaError_t launch_inflate_sub_nodes(const float4* raw, float4* out, uint32_t n_nodes, float scale, float abs_, cudaStream_t s) {
    if (n_nodes == 0) return cudaSuccess;
    int grid = (int)((n_nodes + kRepackBlock - 1) / kRepackBlock);
    inflate_sub_nodes_kernel<<<grid, kRepackBlock, 0, s>>>(raw, out, n_nodes, scale, abs_);
    return cudaGetLastError();
}

cudaError_t launch_repack_triangles(const float* tris_aos, uint32_t n_tris, float4* out, cudaStream_t s) {
    if (n_tris == 0) return cudaSuccess;
    int grid = (int)((n_tris + kRepackBlock - 1) / kRepackBlock);
    repack_triangles_kernel<<<grid, kRepackBlock, 0, s>>>(tris_aos, n_tris, out);
    return cudaGetLastError();
}

cudaError_t launch_repack_sub_triangles(const float* tris_aos, const uint32_t* sub_order, uint32_t n, float4* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    int grid = (int)((n + kRepackBlock - 1) / kRepackBlock);
    repack_sub_triangles_kernel<<<grid, kRepackBlock, 0, s>>>(tris_aos, sub_order, n, out);
    return cudaGetLastError();
}

} // namespace bvht
