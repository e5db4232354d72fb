// scene_kernels.cu -- K6: `SceneObject::set_transform` for every object + `Tlas::rebuild`, on the device.
//
// Reference: scene/scene_object.rs:60-75 (world AABB = Aabb::new_empty grown by the 8 transformed corners of the model's
// bounds; inverse transform cached, transform_component.rs:17-27), scene/tlas.rs:179-202 (find_best_match) and
// :204-250 (rebuild: one leaf per object, then agglomerative clustering).  Compiled with --fmad=false and IEEE division,
// so every f32 operation is the single operation the Rust code performs; min/max are fminf/fmaxf like the oracle.
//
// One CTA.  The clustering is a chain of dependent arg-min searches (about 3n of them); each search is spread over the
// threads of the CTA and reduced with shuffles, the chain itself runs redundantly in every thread (a, b, count and
// nodes_used are uniform).  The active list lives in shared memory as SoA arrays indexed by LIST POSITION (box + node
// index), and is edited with exactly the reference's assignments in the reference's order -- including the case where
// `a` is the last list position and therefore falls outside the shortened list (tlas.rs:238-243): the stale slot stays
// readable, as it does in the reference's Vec.
#include <cfloat>
#include <climits>
#include "device_types.cuh"
#include "launchers.hpp"

namespace bvht {

namespace {

struct Cand { float s; int b; };

__device__ __forceinline__ Cand better(Cand x, Cand y) {
    // the serial loop keeps the FIRST strictly smaller area: lexicographic minimum of (area, b)
    return (y.s < x.s || (y.s == x.s && y.b < x.b)) ? y : x;
}

__device__ __forceinline__ Cand warp_best(Cand c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Cand y;
        y.s = __shfl_xor_sync(0xFFFFFFFFu, c.s, o);
        y.b = __shfl_xor_sync(0xFFFFFFFFu, c.b, o);
        c = better(c, y);
    }
    return c;
}

struct ListView {
    float* lo[3];
    float* hi[3];
    int*   idx;
};

// tlas.rs:179-202.  Returns -1 when no candidate has an area < f32::MAX.
__device__ int find_best_match(const ListView& L, int count, int a, Cand (*red)[32], unsigned& round) {
    const float alx = L.lo[0][a], aly = L.lo[1][a], alz = L.lo[2][a];
    const float ahx = L.hi[0][a], ahy = L.hi[1][a], ahz = L.hi[2][a];
    Cand best; best.s = FLT_MAX; best.b = INT_MAX;
    for (int b = (int)threadIdx.x; b < count; b += (int)blockDim.x) {
        if (b == a) continue;
        float ex = fmaxf(ahx, L.hi[0][b]) - fminf(alx, L.lo[0][b]);
        float ey = fmaxf(ahy, L.hi[1][b]) - fminf(aly, L.lo[1][b]);
        float ez = fmaxf(ahz, L.hi[2][b]) - fminf(alz, L.lo[2][b]);
        float area = (ex * ey + ey * ez) + ez * ex;
        if (area < best.s) { best.s = area; best.b = b; }
    }
    best = warp_best(best);
    if (blockDim.x > 32) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
        Cand* r = red[round & 1u];                       // double-buffered: a warp is at most one search ahead
        if (lane == 0) r[warp] = best;
        __syncthreads();
        Cand c; c.s = FLT_MAX; c.b = INT_MAX;
        if (lane < n_warps) c = r[lane];
        best = warp_best(c);
        round += 1;
    }
    return best.b == INT_MAX ? -1 : best.b;
}

// cglinalg Matrix4x4::inverse as restated by the oracle (adjugate over determinant, 2x2 minors); false when det == 0
__device__ bool mat4_inverse(const float* m, float* out) {
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
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    // out is column-major: out[c * 4 + r] = b(r, c)
    out[0 * 4 + 0] = ((A(1,1) * c5 - A(1,2) * c4) + A(1,3) * c3) * inv;
    out[1 * 4 + 0] = ((-A(0,1) * c5 + A(0,2) * c4) - A(0,3) * c3) * inv;
    out[2 * 4 + 0] = ((A(3,1) * s5 - A(3,2) * s4) + A(3,3) * s3) * inv;
    out[3 * 4 + 0] = ((-A(2,1) * s5 + A(2,2) * s4) - A(2,3) * s3) * inv;
    out[0 * 4 + 1] = ((-A(1,0) * c5 + A(1,2) * c2) - A(1,3) * c1) * inv;
    out[1 * 4 + 1] = ((A(0,0) * c5 - A(0,2) * c2) + A(0,3) * c1) * inv;
    out[2 * 4 + 1] = ((-A(3,0) * s5 + A(3,2) * s2) - A(3,3) * s1) * inv;
    out[3 * 4 + 1] = ((A(2,0) * s5 - A(2,2) * s2) + A(2,3) * s1) * inv;
    out[0 * 4 + 2] = ((A(1,0) * c4 - A(1,1) * c2) + A(1,3) * c0) * inv;
    out[1 * 4 + 2] = ((-A(0,0) * c4 + A(0,1) * c2) - A(0,3) * c0) * inv;
    out[2 * 4 + 2] = ((A(3,0) * s4 - A(3,1) * s2) + A(3,3) * s0) * inv;
    out[3 * 4 + 2] = ((-A(2,0) * s4 + A(2,1) * s2) - A(2,3) * s0) * inv;
    out[0 * 4 + 3] = ((-A(1,0) * c3 + A(1,1) * c1) - A(1,2) * c0) * inv;
    out[1 * 4 + 3] = ((A(0,0) * c3 - A(0,1) * c1) + A(0,2) * c0) * inv;
    out[2 * 4 + 3] = ((-A(3,0) * s3 + A(3,1) * s1) - A(3,2) * s0) * inv;
    out[3 * 4 + 3] = ((A(2,0) * s3 - A(2,1) * s1) + A(2,2) * s0) * inv;
#undef A
    return true;
}

__device__ __forceinline__ void store_tlas_node(float4* tlas, int ni, const float lo[3], const float hi[3], uint32_t left_right,
                                                uint32_t blas) {
    tlas[2 * ni + 0] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(left_right));
    tlas[2 * ni + 1] = make_float4(hi[0], hi[1], hi[2], __uint_as_float(blas));
}

__global__ void scene_rebuild_kernel(SceneRebuildParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Cand red[2][32];
    const int n = (int)p.n_inst;
    ListView L;
    {
        float* f = reinterpret_cast<float*>(smem_raw);
        for (int k = 0; k < 3; ++k) { L.lo[k] = f + (size_t)k * n; L.hi[k] = f + (size_t)(3 + k) * n; }
        L.idx = reinterpret_cast<int*>(f + (size_t)6 * n);
    }

    // ---- SceneObject::set_transform for every object (scene_object.rs:60-75) + the TLAS leaves (tlas.rs:205-218)
    for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) {
        float m[16], inv[16];
        const float4* src = reinterpret_cast<const float4*>(p.transforms) + (size_t)i * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) { float4 v = src[c]; m[c * 4 + 0] = v.x; m[c * 4 + 1] = v.y; m[c * 4 + 2] = v.z; m[c * 4 + 3] = v.w; }
        if (!mat4_inverse(m, inv)) { atomicOr(p.status, 1u); for (int k = 0; k < 16; ++k) inv[k] = 0.0f; }
#pragma unroll
        for (int c = 0; c < 4; ++c) p.inst_cols[(size_t)i * 4 + c] = make_float4(inv[c * 4 + 0], inv[c * 4 + 1], inv[c * 4 + 2], inv[c * 4 + 3]);
        const uint32_t blas_id = p.blas_ids[i];
        p.inst_blas[i] = blas_id;
        const float4 n0 = p.blas[blas_id].nodes[0], n1 = p.blas[blas_id].nodes[1];      // Model::bounds = root box (bvh.rs:499)
        const float omin[3] = { n0.x, n0.y, n0.z }, omax[3] = { n1.x, n1.y, n1.z };
        float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
        for (int c = 0; c < 8; ++c) {
            const float px = (c & 1) ? omax[0] : omin[0], py = (c & 2) ? omax[1] : omin[1], pz = (c & 4) ? omax[2] : omin[2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                // Mat4 x Vec4 with w = 1: ((c0*x + c1*y) + c2*z) + c3*1   (transform.rs:219-223)
                float q = ((m[0 + r] * px + m[4 + r] * py) + m[8 + r] * pz) + m[12 + r] * 1.0f;
                lo[r] = fminf(lo[r], q); hi[r] = fmaxf(hi[r], q);
            }
        }
        if (p.inst_bounds) {
            p.inst_bounds[(size_t)i * 6 + 0] = lo[0]; p.inst_bounds[(size_t)i * 6 + 1] = lo[1]; p.inst_bounds[(size_t)i * 6 + 2] = lo[2];
            p.inst_bounds[(size_t)i * 6 + 3] = hi[0]; p.inst_bounds[(size_t)i * 6 + 4] = hi[1]; p.inst_bounds[(size_t)i * 6 + 5] = hi[2];
        }
        store_tlas_node(p.tlas, i + 1, lo, hi, 0u, (uint32_t)i);
        for (int k = 0; k < 3; ++k) { L.lo[k][i] = lo[k]; L.hi[k][i] = hi[k]; }
        L.idx[i] = i + 1;
    }
    __syncthreads();

    // ---- agglomerative clustering (tlas.rs:220-247); every thread runs the chain, the searches are shared
    unsigned round = 0;
    int count = n, used = n + 1;
    int a = 0;
    int b = find_best_match(L, count, a, red, round);
    bool broken = false;
    while (count > 1) {
        if (b < 0) { broken = true; break; }             // the reference indexes list[-1 as usize] here and panics
        const int c = find_best_match(L, count, b, red, round);
        if (a == c) {
            __syncthreads();                             // every thread is done reading the list
            if (threadIdx.x == 0) {
                const int ia = L.idx[a], ib = L.idx[b];
                float lo[3], hi[3];
                for (int k = 0; k < 3; ++k) { lo[k] = fminf(L.lo[k][a], L.lo[k][b]); hi[k] = fmaxf(L.hi[k][a], L.hi[k][b]); }
                store_tlas_node(p.tlas, used, lo, hi, (uint32_t)ia + ((uint32_t)ib << 16), 0u);   // LeftRightIndex::new(ia, ib), blas = 0 (Default)
                for (int k = 0; k < 3; ++k) { L.lo[k][a] = lo[k]; L.hi[k][a] = hi[k]; }
                L.idx[a] = used;
                for (int k = 0; k < 3; ++k) { L.lo[k][b] = L.lo[k][count - 1]; L.hi[k][b] = L.hi[k][count - 1]; }
                L.idx[b] = L.idx[count - 1];
            }
            __syncthreads();
            used += 1;
            count -= 1;
            b = find_best_match(L, count, a, red, round);
        } else {
            a = b;
            b = c;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (broken) atomicOr(p.status, 2u);
        // self.nodes[0] = self.nodes[node_indices[a]]  (tlas.rs:248); __syncthreads made every node store visible
        const int ni = L.idx[a];
        const float4 q0 = p.tlas[2 * ni + 0], q1 = p.tlas[2 * ni + 1];
        p.tlas[0] = q0; p.tlas[1] = q1;
        p.status[1] = (unsigned)used;
    }
    if (p.pack) {
        // everything the host wants back, contiguous, so that ONE device->host copy fetches it:
        // [tlas: 2n nodes x 2 float4 | inverse transforms: n x 4 float4 | world bounds: n x 6 floats | status: 4 words]
        __syncthreads();
        float4* out = reinterpret_cast<float4*>(p.pack);
        const int n_tl = 4 * n, n_ic = 4 * n;
        for (int i = (int)threadIdx.x; i < n_tl; i += (int)blockDim.x) out[i] = p.tlas[i];
        for (int i = (int)threadIdx.x; i < n_ic; i += (int)blockDim.x) out[n_tl + i] = p.inst_cols[i];
        float* fb = reinterpret_cast<float*>(out + n_tl + n_ic);
        if (p.inst_bounds) for (int i = (int)threadIdx.x; i < 6 * n; i += (int)blockDim.x) fb[i] = p.inst_bounds[i];
        unsigned* st = reinterpret_cast<unsigned*>(fb + 6 * n);
        if (threadIdx.x < 4) st[threadIdx.x] = p.status[threadIdx.x];
    }
}

} // namespace

size_t scene_rebuild_smem_bytes(uint32_t n_inst) { return (size_t)n_inst * 28; }

cudaError_t launch_scene_rebuild(const SceneRebuildParams& p, cudaStream_t s) {
    const size_t smem = scene_rebuild_smem_bytes(p.n_inst);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(scene_rebuild_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int block = p.n_inst <= 64 ? 32 : (p.n_inst <= 1024 ? 256 : 1024);
    scene_rebuild_kernel<<<1, block, smem, s>>>(p);
    return cudaGetLastError();
}

} // namespace bvht
