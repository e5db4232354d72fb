// cover_kernels.cu -- per-frame conservative coverage of the 8x4-pixel blocks by every instance's triangles.
//
// Not a reference component.  The trace kernel decides per 32-pixel block which instances its rays can see at all; until now
// from ONE screen rectangle per instance (the projected box of the whole model).  For an armadillo two units from the camera
// that rectangle is a third of the frame although the silhouette is not: in sixteen_armadillos frame 30 every block carries
// some instance as a candidate, and 45 % of the summed block time is spent in blocks whose rays hit nothing (trippy_teapots:
// 66 %, measured with tools/slice_timeline.py).  Here every TRIANGLE marks the blocks its own conservative box projects onto:
//   box = the triangle's model-space AABB grown by ITS delta = scale * |e1||e2| + abs of the current bake (leaf_accel.hpp: a
//   ray the reference's Triangle::intersect accepts passes through that box), its 8 corners through the instance's transform
//   and the view matrix, their screen bounding rectangle +- 2 pixels, all blocks it touches: bit i of cover[block].
// A block without bit i holds no ray that any triangle of instance i can accept, so dropping instance i from the block's
// candidate mask changes nothing (entering an instance that yields no hit has no side effect, scene_object.rs:78-89).
// Boxes that reach behind the eye plane, non-finite projections and list overflows set the instance's bit in `full`
// (visible everywhere).  Rectangles of more than 64 blocks go to a list (room for every triangle) that a second kernel fills,
// one CTA per rectangle: close to the camera, or at 8K, most triangles are that large.
#include <cfloat>
#include "device_types.cuh"
#include "launchers.hpp"

namespace bvht {
namespace {


__device__ __forceinline__ void mark(uint32_t* cover, uint32_t ntx, int bx, int by, uint32_t bit) {
    uint32_t* p = cover + ((size_t)(by >> 1) * ntx + (uint32_t)bx) * 2u + (uint32_t)(by & 1);
    if (!(*reinterpret_cast<volatile uint32_t*>(p) & bit)) atomicOr(p, bit);
}

__global__ void __launch_bounds__(256)
raster_cover_kernel(const __grid_constant__ CoverParams P) {
    const uint32_t g = blockIdx.x * 256u + threadIdx.x;
    if (g == 0u && P.full_init) atomicOr(P.full, P.full_init);
    if (g >= P.tri_offset[P.n_inst]) return;
    uint32_t i = 0;
    while (i + 1 < P.n_inst && g >= P.tri_offset[i + 1]) ++i;
    const uint32_t j = g - P.tri_offset[i];
    const uint32_t bit = 1u << i;
    const float4* tp = P.inst_tri[i] + 3 * (size_t)j;
    const float4 v0 = __ldg(tp + 0), e1 = __ldg(tp + 1), e2 = __ldg(tp + 2);
    const bool z1 = e1.x == 0.0f && e1.y == 0.0f && e1.z == 0.0f, z2 = e2.x == 0.0f && e2.y == 0.0f && e2.z == 0.0f;
    if (z1 || z2) return;                                   // a zero edge: area == 0 exactly, never accepted
    // |e1||e2| rounded up; NaN / inf vertices end in `full` through the projection below
    const float l1 = __fsqrt_ru(__fmaf_ru(e1.x, e1.x, __fmaf_ru(e1.y, e1.y, __fmul_ru(e1.z, e1.z))));
    const float l2 = __fsqrt_ru(__fmaf_ru(e2.x, e2.x, __fmaf_ru(e2.y, e2.y, __fmul_ru(e2.z, e2.z))));
    const float delta = __fmaf_ru(P.inst_scale[i], __fmul_ru(__fmul_ru(l1, l2), 1.00001f), P.inst_abs[i]);
    float lo[3], hi[3];
    {
        const float a[3] = { v0.x, v0.y, v0.z }, b1[3] = { e1.x, e1.y, e1.z }, b2[3] = { e2.x, e2.y, e2.z };
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // v1 = v0 + e1 only up to one rounding of the stored edge: 4e-6 relative covers it (and the bake's own 1e-6)
            const float p1 = a[k] + b1[k], p2 = a[k] + b2[k];
            const float mn = fminf(a[k], fminf(p1, p2)), mx = fmaxf(a[k], fmaxf(p1, p2));
            const float m = fmaxf(fabsf(mn), fabsf(mx));
            lo[k] = __fsub_rd(__fsub_rd(mn, delta), __fmul_ru(m, 4e-6f));
            hi[k] = __fadd_ru(__fadd_ru(mx, delta), __fmul_ru(m, 4e-6f));
        }
    }
    const float* M = P.mv[i];                               // eye = M (3 x 4, row-major) * (p, 1)
    float umin = FLT_MAX, umax = -FLT_MAX, vmin = FLT_MAX, vmax = -FLT_MAX;
    bool behind = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float x = (c & 1) ? hi[0] : lo[0], y = (c & 2) ? hi[1] : lo[1], z = (c & 4) ? hi[2] : lo[2];
        const float ex = M[0] * x + M[1] * y + M[2] * z + M[3];
        const float ey = M[4] * x + M[5] * y + M[6] * z + M[7];
        const float ez = M[8] * x + M[9] * y + M[10] * z + M[11];
        if (!(ez < -P.z_eps)) behind = true;
        const float s = P.near_ / -ez;
        const float u = (ex * s - P.tlx) * P.inv_ex, v = (ey * s - P.tly) * P.inv_ey;
        umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
    }
    const float W = (float)P.width, H = (float)P.height;
    const float fx0 = floorf(umin * W) - 2.0f, fx1 = ceilf(umax * W) + 2.0f, fy0 = floorf(vmin * H) - 2.0f, fy1 = ceilf(vmax * H) + 2.0f;
    if (behind || !(fabsf(fx0) < 1e9f) || !(fabsf(fx1) < 1e9f) || !(fabsf(fy0) < 1e9f) || !(fabsf(fy1) < 1e9f)) {
        atomicOr(P.full, bit);
        return;
    }
    if (fx1 < 0.0f || fy1 < 0.0f || fx0 > W - 1.0f || fy0 > H - 1.0f) return;          // entirely off screen
    const int px0 = (int)fmaxf(fx0, 0.0f), px1 = (int)fminf(fx1, W - 1.0f), py0 = (int)fmaxf(fy0, 0.0f), py1 = (int)fminf(fy1, H - 1.0f);
    const int bx0 = px0 >> 3, bx1 = px1 >> 3, by0 = py0 >> 2, by1 = py1 >> 2;
    const uint32_t count = (uint32_t)(bx1 - bx0 + 1) * (uint32_t)(by1 - by0 + 1);
    if (count <= 64u) {
        // The kernel is latency bound (ncu: 29 % issue slots, long-scoreboard stalls): a dependent read-then-atomic per block
        // serialises a triangle's 4-6 marks.  So the words of up to eight blocks are READ first (independent L2 loads in flight
        // together), and only the blocks whose bit is still clear get the atomic.
        const int w = bx1 - bx0 + 1;
        for (uint32_t c0 = 0; c0 < count; c0 += 8u) {
            uint32_t* ptr[8];
            uint32_t seen[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t c = c0 + (uint32_t)k;
                ptr[k] = nullptr; seen[k] = bit;
                if (c < count) {
                    const int by = by0 + (int)(c / (uint32_t)w), bx = bx0 + (int)(c % (uint32_t)w);
                    if (P.shard_count > 1u && ((uint32_t)(by >> 1) % P.shard_count) != P.shard_index) continue;     // another GPU's tile row
                    ptr[k] = P.cover + ((size_t)(by >> 1) * P.ntx + (uint32_t)bx) * 2u + (uint32_t)(by & 1);
                    seen[k] = __ldcg(ptr[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (ptr[k] && !(seen[k] & bit)) atomicOr(ptr[k], bit);
        }
    } else {
        const uint32_t at = atomicAdd(P.big_count, 1u);
        if (at < P.big_cap) P.big_list[at] = make_int4((int)i, (bx0 << 16) | bx1, (by0 << 16) | by1, 0);
        else atomicOr(P.full, bit);
    }
}

__global__ void __launch_bounds__(256)
raster_big_kernel(const __grid_constant__ CoverParams P) {
    const uint32_t n = min(*P.big_count, P.big_cap);
    for (uint32_t e = blockIdx.x; e < n; e += gridDim.x) {
        const int4 r = P.big_list[e];
        const uint32_t bit = 1u << r.x;
        const int bx0 = r.y >> 16, bx1 = r.y & 0xFFFF, by0 = r.z >> 16, by1 = r.z & 0xFFFF;
        const int w = bx1 - bx0 + 1, total = w * (by1 - by0 + 1);
        for (int k = (int)threadIdx.x; k < total; k += 256) {
            const int by = by0 + k / w;
            if (P.shard_count > 1u && ((uint32_t)(by >> 1) % P.shard_count) != P.shard_index) continue;
            mark(P.cover, P.ntx, bx0 + k % w, by, bit);
        }
    }
}

} // namespace

cudaError_t launch_raster_cover(const CoverParams& p, cudaStream_t s) {
    const uint32_t total = p.tri_offset[p.n_inst];
    if (total == 0) return cudaSuccess;
    raster_cover_kernel<<<(total + 255u) / 256u, 256, 0, s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    raster_big_kernel<<<296, 256, 0, s>>>(p);
    return cudaGetLastError();
}

} // namespace bvht
