// build_kernels.cu -- K3: the reference's binned-SAH BVH build, on the device, bit-compatible with the host build.
//
// Replaces BvhBuilder::build_for / subdivide / find_best_split_plane (model/bvh.rs:333-467, 505-541) for callers that
// want the tree built (or rebuilt, the alternative to Bvh::refit the reference benchmarks in bench_bvh_refit_rebuild.rs)
// where the triangles already live.  Everything the reference's result depends on is reproduced:
//
//   * update_node_bounds (bvh.rs:317-330): min/max over the vertices -- exact, order-independent.
//   * find_best_split_plane (bvh.rs:333-394): centroid bounds per axis with bounds_max starting at 1e-30, eight bins whose
//     boxes START AS THE ORIGIN POINT BOX (Aabb::default(), bvh.rs:349,365-366) -- the quirk that makes the reference's
//     trees shallow -- bin index by saturating cast, the two sweeps, cost = n_l * area_l + n_r * area_r in f32, first
//     strictly smaller plane wins (axis 0,1,2; plane 0..6).  Bin boxes and counts are min/max/integer reductions, so
//     any evaluation order gives the same bits; this TU is compiled with --fmad=false.
//   * the in-place partition (bvh.rs:419-430) defines the ORDER of the triangles (and hit ids are positions in that
//     order).  `while i <= j { if left(a[i]) { i++ } else { swap(a[i], a[j]); j-- } }` has a closed form: with
//     n_l = #left, q = first + n_l, "front" = positions < q (plus q itself when it holds a right element), H_k = k-th
//     right element of the front (ascending), G_k = k-th left element of the back (descending):
//         front left  -> stays          back right -> position - 1
//         G_k         -> pos(H_k)       H_k        -> pos(G_{k-1}) - 1     (H_0 -> last)
//     (checked against the sequential loop on random inputs in tests/test_build_permutation.py), so one prefix sum per
//     node gives every triangle its final position.  The permutation is applied even when the split is then abandoned
//     because one side is empty (bvh.rs:432-435), like the reference.
//   * node indices: children pairs are allocated in DFS pre-order of the successful splits (bvh.rs:437-466 recurses left
//     then right); the level-synchronous build numbers nodes in creation order and the host renumbers the (small) tree.
//
// Level-synchronous: every level runs the same ten small kernels over all nodes that may still split; a node's range is
// cut into 512-triangle chunks, one CTA per chunk, reductions merge through order-preserving u32 encodings of the floats
// with integer atomics.  HBM traffic per level: 36 B/triangle read four times + moved once; trees of the reference's
// shape (armadillo: 96 nodes, ~48 levels) are launch-latency bound, not bandwidth bound.
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <vector>

#include "build_device.hpp"

namespace bvht {
namespace {

constexpr int kBB = 128;                       // threads per CTA
constexpr int kPerThread = 4;
constexpr int kChunk = kBB * kPerThread;       // triangles per CTA job
constexpr int kBins = 8;                       // bvh.rs:17 BINS
constexpr int kBinWords = 3 * kBins * 7;       // per node: axis x bin x {min.xyz, max.xyz, count}
constexpr uint32_t kNoNode = 0xFFFFFFFFu;

struct Split {                                 // per ACTIVE node of the current level
    int      axis; float pos;                  // reference policy: left <=> centroid[axis] < pos
    uint32_t split;                            // 1: partition this node
    uint32_t n_left, q, q_is_right, left_before_q;
    uint32_t chunk0, n_chunks;
    float    lo, scale; int bin;               // leaf-accelerator policy: left <=> bin(centroid[axis]) <= bin ...
    uint32_t first, half;                      // ... or, axis == 3, the first `half` positions of the range (balanced fallback)
};
struct Chunk { uint32_t slot, node, begin, end; };

// order-preserving float <-> u32 (after canonicalising -0.0): min/max become integer atomics
__device__ __forceinline__ uint32_t enc(float f) {
    f += 0.0f;
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

__device__ __forceinline__ float centroid_axis(const float* t, int a) {      // triangle.rs:33-39
    const float one_third = 1.0f / 3.0f;
    return ((t[a] + t[3 + a]) + t[6 + a]) * one_third;
}
__device__ __forceinline__ float area(const float mn[3], const float mx[3]) {   // aabb.rs:48-57
    float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    return (ex * ey + ey * ez) + ez * ex;
}
__device__ __forceinline__ void load_tri(const float* tris, uint32_t p, float t[9]) {
    const float* s = tris + (size_t)p * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) t[k] = __ldg(s + k);
}

__device__ __forceinline__ void init_node(BuildNode& n, uint32_t first, uint32_t count, uint32_t depth) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        n.vmin[k] = enc(FLT_MAX); n.vmax[k] = enc(-FLT_MAX);          // Aabb::new_empty (bvh.rs:318)
        n.cmin[k] = enc(1e30f);   n.cmax[k] = enc(1e-30f);            // bvh.rs:339-340 (1e-30 is the reference's)
    }
    n.first = first; n.count = count; n.left = 0u; n.depth = depth;
}

__global__ void init_root_kernel(BuildNode* nodes, uint32_t* active, uint32_t* counters, uint32_t* perm, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        init_node(nodes[0], 0u, n, 0u);
        active[0] = 0u;
        counters[kCtrNodes] = 1u; counters[kCtrNext] = 0u; counters[kCtrChunks] = 0u; counters[kCtrDepth] = 0u;
    }
    for (; i < n; i += gridDim.x * blockDim.x) perm[i] = i;
}

// ---- plan: chunk list of the level
template <typename N, int BIN_WORDS, bool ORIGIN_BINS>
__global__ void plan_count_kernel(const N* nodes, const uint32_t* active, uint32_t n_active, Split* split, uint32_t* bins,
                                  uint32_t* counters) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot == 0) counters[kCtrNext] = 0u;
    if (slot >= n_active) return;
    const N& n = nodes[active[slot]];
    Split s; memset(&s, 0, sizeof s);
    s.axis = -1; s.n_chunks = (n.count + kChunk - 1) / kChunk;
    split[slot] = s;
    uint32_t* b = bins + (size_t)slot * BIN_WORDS;
    // reference bins: Aabb::default(), the point box at the origin; accelerator bins: empty boxes
    const uint32_t lo0 = ORIGIN_BINS ? enc(0.0f) : enc(FLT_MAX), hi0 = ORIGIN_BINS ? enc(0.0f) : enc(-FLT_MAX);
    for (int i = 0; i < BIN_WORDS; ++i) { const int w = i % 7; b[i] = w == 6 ? 0u : (w < 3 ? lo0 : hi0); }
}

__global__ void plan_scan_kernel(Split* split, uint32_t n_active, uint32_t* counters) {     // one CTA
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_active; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n_active ? split[i].n_chunks : 0u, x = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = lane < (int)(blockDim.x >> 5) ? warp_sum[lane] : 0u, z = w;
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, z, o); if (lane >= o) z += y; }
            warp_sum[lane] = z - w;                                      // exclusive
        }
        __syncthreads();
        uint32_t excl = carry + warp_sum[warp] + x - v;
        if (i < n_active) split[i].chunk0 = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) counters[kCtrChunks] = carry;
}

template <typename N>
__global__ void plan_fill_kernel(const N* nodes, const uint32_t* active, uint32_t n_active, const Split* split, Chunk* chunks) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_active) return;
    const N& n = nodes[active[slot]];
    const Split& s = split[slot];
    for (uint32_t j = 0; j < s.n_chunks; ++j) {
        Chunk c; c.slot = slot; c.node = active[slot];
        c.begin = n.first + j * kChunk; c.end = min(n.first + n.count, c.begin + kChunk);
        chunks[s.chunk0 + j] = c;
    }
}

// ---- bounds: update_node_bounds + the centroid bounds of find_best_split_plane, one pass
__global__ void __launch_bounds__(kBB) bounds_kernel(const float* __restrict__ tris, BuildNode* nodes, const Chunk* chunks,
                                                    const uint32_t* counters) {
    if (blockIdx.x >= counters[kCtrChunks]) return;
    const Chunk c = chunks[blockIdx.x];
    float v[12];
#pragma unroll
    for (int k = 0; k < 3; ++k) { v[k] = FLT_MAX; v[3 + k] = -FLT_MAX; v[6 + k] = 1e30f; v[9 + k] = 1e-30f; }
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        uint32_t p = c.begin + threadIdx.x * kPerThread + e;
        if (p < c.end) {
            float t[9]; load_tri(tris, p, t);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                v[k] = fminf(fminf(v[k], t[k]), fminf(t[3 + k], t[6 + k]));
                v[3 + k] = fmaxf(fmaxf(v[3 + k], t[k]), fmaxf(t[3 + k], t[6 + k]));
                float cen = centroid_axis(t, k);
                v[6 + k] = fminf(v[6 + k], cen); v[9 + k] = fmaxf(v[9 + k], cen);
            }
        }
    }
    __shared__ float red[kBB / 32][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const bool is_min = (k < 3) || (k >= 6 && k < 9);
        for (int o = 16; o > 0; o >>= 1) {
            float y = __shfl_xor_sync(0xFFFFFFFFu, v[k], o);
            v[k] = is_min ? fminf(v[k], y) : fmaxf(v[k], y);
        }
        if (lane == 0) red[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        const int k = threadIdx.x;
        const bool is_min = (k < 3) || (k >= 6 && k < 9);
        float x = red[0][k];
        for (int w = 1; w < kBB / 32; ++w) x = is_min ? fminf(x, red[w][k]) : fmaxf(x, red[w][k]);
        if (x == x) {                                                    // f32::min/max ignore NaN (bvh.rs:322-327)
            BuildNode& n = nodes[c.node];
            uint32_t* dst = k < 3 ? &n.vmin[k] : k < 6 ? &n.vmax[k - 3] : k < 9 ? &n.cmin[k - 6] : &n.cmax[k - 9];
            if (is_min) atomicMin(dst, enc(x)); else atomicMax(dst, enc(x));
        }
    }
}

// ---- bins (bvh.rs:343-358)
__global__ void __launch_bounds__(kBB) bins_kernel(const float* __restrict__ tris, const BuildNode* nodes, const Chunk* chunks,
                                                  uint32_t* bins, const uint32_t* counters) {
    if (blockIdx.x >= counters[kCtrChunks]) return;
    const Chunk c = chunks[blockIdx.x];
    __shared__ uint32_t sb[kBinWords];
    const uint32_t zero = enc(0.0f);
    for (int i = threadIdx.x; i < kBinWords; i += kBB) sb[i] = (i % 7 == 6) ? 0u : zero;
    float bmin[3], scale[3]; bool on[3];
    {
        const BuildNode& n = nodes[c.node];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            bmin[a] = dec(n.cmin[a]);
            float bmax = dec(n.cmax[a]);
            on[a] = !(bmin[a] == bmax);                                   // bvh.rs:344-346
            scale[a] = (float)kBins / (bmax - bmin[a]);                    // bvh.rs:350
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        uint32_t p = c.begin + threadIdx.x * kPerThread + e;
        if (p >= c.end) continue;
        float t[9]; load_tri(tris, p, t);
        uint32_t lo[3], hi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = enc(fminf(t[k], fminf(t[3 + k], t[6 + k])));
            hi[k] = enc(fmaxf(t[k], fmaxf(t[3 + k], t[6 + k])));
        }
        const bool tri_nan[3] = { !(t[0] == t[0] && t[3] == t[3] && t[6] == t[6]), !(t[1] == t[1] && t[4] == t[4] && t[7] == t[7]),
                                  !(t[2] == t[2] && t[5] == t[5] && t[8] == t[8]) };
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (!on[a]) continue;
            float f = (centroid_axis(t, a) - bmin[a]) * scale[a];
            int idx = f >= (float)kBins ? kBins - 1 : (f > 0.0f ? (int)f : 0);   // `as usize` saturates, then min(BINS - 1, .)
            uint32_t* b = sb + (a * kBins + idx) * 7;
            atomicAdd(b + 6, 1u);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (tri_nan[k]) continue;       // NaN vertices: min/max ignore them; finite inputs never get here
                atomicMin(b + k, lo[k]); atomicMax(b + 3 + k, hi[k]);
            }
        }
    }
    __syncthreads();
    uint32_t* g = bins + (size_t)c.slot * kBinWords;
    for (int i = threadIdx.x; i < kBinWords; i += kBB) {
        const int w = i % 7;
        if (sb[i - w + 6] == 0u) continue;                                // empty bin in this chunk: nothing to merge
        if (w == 6) atomicAdd(g + i, sb[i]);
        else if (w < 3) atomicMin(g + i, sb[i]);
        else atomicMax(g + i, sb[i]);
    }
}

// ---- plane selection (bvh.rs:360-393) and the split decision (bvh.rs:399-408)
__global__ void select_kernel(const BuildNode* nodes, const uint32_t* active, uint32_t n_active, const uint32_t* bins, Split* split) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_active) return;
    const BuildNode& n = nodes[active[slot]];
    int best_axis = -1; float best_pos = 0.0f, best_cost = FLT_MAX;
    for (int a = 0; a < 3; ++a) {
        const float bmin = dec(n.cmin[a]), bmax = dec(n.cmax[a]);
        if (bmin == bmax) continue;
        const uint32_t* b = bins + (size_t)slot * kBinWords + a * kBins * 7;
        float la[kBins - 1], ra[kBins - 1]; uint32_t lc[kBins - 1], rc[kBins - 1];
        float lmn[3] = { 0, 0, 0 }, lmx[3] = { 0, 0, 0 }, rmn[3] = { 0, 0, 0 }, rmx[3] = { 0, 0, 0 };
        uint32_t ls = 0, rs = 0;
        for (int i = 0; i < kBins - 1; ++i) {
            const uint32_t* bl = b + i * 7;
            const uint32_t* br = b + (kBins - 1 - i) * 7;
            ls += bl[6]; lc[i] = ls;
            if (dec(bl[0]) != FLT_MAX) {                                  // grow_aabb's emptiness test (aabb.rs:41-46)
                for (int k = 0; k < 3; ++k) {
                    lmn[k] = fminf(fminf(lmn[k], dec(bl[k])), dec(bl[3 + k]));
                    lmx[k] = fmaxf(fmaxf(lmx[k], dec(bl[k])), dec(bl[3 + k]));
                }
            }
            la[i] = area(lmn, lmx);
            rs += br[6]; rc[kBins - 2 - i] = rs;
            if (dec(br[0]) != FLT_MAX) {
                for (int k = 0; k < 3; ++k) {
                    rmn[k] = fminf(fminf(rmn[k], dec(br[k])), dec(br[3 + k]));
                    rmx[k] = fmaxf(fmaxf(rmx[k], dec(br[k])), dec(br[3 + k]));
                }
            }
            ra[kBins - 2 - i] = area(rmn, rmx);
        }
        const float scale = (bmax - bmin) / (float)kBins;
        for (int i = 0; i < kBins - 1; ++i) {
            float cost = (float)lc[i] * la[i] + (float)rc[i] * ra[i];
            if (cost < best_cost) { best_axis = a; best_pos = bmin + scale * (float)(i + 1); best_cost = cost; }
        }
    }
    float mn[3], mx[3];
    for (int k = 0; k < 3; ++k) { mn[k] = dec(n.vmin[k]); mx[k] = dec(n.vmax[k]); }
    const float no_split = (float)n.count * area(mn, mx);
    Split& s = split[slot];
    s.axis = best_axis; s.pos = best_pos;
    s.split = (best_axis >= 0 && !(best_cost >= no_split)) ? 1u : 0u;
}

// What differs between the two builders in the partition passes: the predicate and what is moved.
struct RefPolicy {                       // the reference's build: triangles themselves are permuted (bvh.rs:419-430)
    const float* tris; const uint32_t* perm; float* tris_tmp; uint32_t* perm_tmp; float* tris_rw; uint32_t* perm_rw;
    __device__ __forceinline__ bool left(uint32_t p, const Split& s) const {
        const float* t = tris + (size_t)p * 9;
        return ((__ldg(t + s.axis) + __ldg(t + 3 + s.axis)) + __ldg(t + 6 + s.axis)) * (1.0f / 3.0f) < s.pos;    // bvh.rs:423
    }
    __device__ __forceinline__ void move(uint32_t p, uint32_t dest) const {
        float t[9]; load_tri(tris, p, t);
        float* o = tris_tmp + (size_t)dest * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) o[k] = t[k];
        perm_tmp[dest] = perm[p];
    }
    __device__ __forceinline__ void copy_back(uint32_t begin, uint32_t end) const {
        const size_t b = (size_t)begin * 9, e = (size_t)end * 9;
        for (size_t i = b + threadIdx.x; i < e; i += kBB) tris_rw[i] = tris_tmp[i];
        for (uint32_t p = begin + threadIdx.x; p < end; p += kBB) perm_rw[p] = perm_tmp[p];
    }
};
constexpr int kSubBins = 16;
struct SubPolicy {                       // the leaf accelerator's build: an index array is permuted, triangles stay
    const float* tris; const uint32_t* order; uint32_t* order_tmp; uint32_t* order_rw;
    __device__ __forceinline__ bool left(uint32_t p, const Split& s) const {
        if (s.axis == 3) return (p - s.first) < s.half;
        const float* t = tris + (size_t)__ldg(order + p) * 9;
        float c = ((__ldg(t + s.axis) + __ldg(t + 3 + s.axis)) + __ldg(t + 6 + s.axis)) * (1.0f / 3.0f);
        int bi = (int)((c - s.lo) * s.scale);
        bi = max(0, min(kSubBins - 1, bi));
        return bi <= s.bin;
    }
    __device__ __forceinline__ void move(uint32_t p, uint32_t dest) const { order_tmp[dest] = order[p]; }
    __device__ __forceinline__ void copy_back(uint32_t begin, uint32_t end) const {
        for (uint32_t p = begin + threadIdx.x; p < end; p += kBB) order_rw[p] = order_tmp[p];
    }
};

// ---- partition, pass 1: left elements per chunk
template <typename P>
__global__ void __launch_bounds__(kBB) count_kernel(const P pol, const Chunk* chunks, const Split* split,
                                                   uint32_t* chunk_left, const uint32_t* counters) {
    if (blockIdx.x >= counters[kCtrChunks]) return;
    const Chunk c = chunks[blockIdx.x];
    const Split& s = split[c.slot];
    if (!s.split) return;
    int local = 0;
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        uint32_t p = c.begin + threadIdx.x * kPerThread + e;
        if (p < c.end && pol.left(p, s)) ++local;
    }
    __shared__ uint32_t acc;
    if (threadIdx.x == 0) acc = 0u;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc, (uint32_t)local);
    __syncthreads();
    if (threadIdx.x == 0) chunk_left[blockIdx.x] = acc;
}

// ---- partition, pass 2: per node, prefix over its chunks; n_left, q and what sits at q   (one warp per node)
template <typename P, typename N>
__global__ void scan_kernel(const P pol, const N* nodes, const uint32_t* active, uint32_t n_active,
                            Split* split, const uint32_t* chunk_left, uint32_t* chunk_left_before) {
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (slot >= n_active) return;
    Split& s = split[slot];
    if (!s.split) return;
    const N& n = nodes[active[slot]];
    uint32_t carry = 0u;
    for (uint32_t base = 0; base < s.n_chunks; base += 32) {
        uint32_t j = base + lane;
        uint32_t v = j < s.n_chunks ? chunk_left[s.chunk0 + j] : 0u, x = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
        if (j < s.n_chunks) chunk_left_before[s.chunk0 + j] = carry + x - v;
        carry += __shfl_sync(0xFFFFFFFFu, x, 31);
    }
    __syncwarp();
    const uint32_t n_left = carry, q = n.first + n_left, last = n.first + n.count - 1u;
    uint32_t q_is_right = 0u, left_before_q = n_left;
    if (q <= last) {
        q_is_right = pol.left(q, s) ? 0u : 1u;
        const uint32_t jq = (q - n.first) / kChunk, cb = n.first + jq * kChunk;
        uint32_t cnt = 0u;
        for (uint32_t p = cb + lane; p < q; p += 32) cnt += pol.left(p, s) ? 1u : 0u;
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        left_before_q = chunk_left_before[s.chunk0 + jq] + cnt;         // written by this warp above
    }
    if (lane == 0) { s.n_left = n_left; s.q = q; s.q_is_right = q_is_right; s.left_before_q = left_before_q; }
}

// block-wide exclusive prefix of per-thread counts
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v) {
    __shared__ uint32_t ws[kBB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    uint32_t before = 0u;
    for (int w = 0; w < warp; ++w) before += ws[w];
    __syncthreads();
    return before + x - v;
}

// ---- partition, pass 3: positions of the front's right elements (H) and the back's left elements (G)
template <typename P, typename N>
__global__ void __launch_bounds__(kBB) tables_kernel(const P pol, const N* nodes, const Chunk* chunks,
                                                    const Split* split, const uint32_t* chunk_left_before, uint32_t* table_h,
                                                    uint32_t* table_g, const uint32_t* counters) {
    if (blockIdx.x >= counters[kCtrChunks]) return;
    const Chunk c = chunks[blockIdx.x];
    const Split s = split[c.slot];
    if (!s.split) return;
    const uint32_t first = nodes[c.node].first;
    bool left[kPerThread]; uint32_t local = 0u;
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        uint32_t p = c.begin + threadIdx.x * kPerThread + e;
        left[e] = p < c.end && pol.left(p, s);
        local += left[e] ? 1u : 0u;
    }
    uint32_t before = chunk_left_before[blockIdx.x] + block_exclusive(local);
    const uint32_t n_back_left = s.n_left - s.left_before_q;
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        uint32_t p = c.begin + threadIdx.x * kPerThread + e;
        if (p < c.end) {
            const bool front = p < s.q || (p == s.q && s.q_is_right);
            if (front && !left[e]) table_h[first + ((p - first) - before)] = p;
            if (!front && left[e]) table_g[first + (n_back_left - 1u - (before - s.left_before_q))] = p;
        }
        before += left[e] ? 1u : 0u;
    }
}

// ---- partition, pass 4: move every triangle of a splitting node to its final position (into the scratch copy)
template <typename P, typename N>
__global__ void __launch_bounds__(kBB) scatter_kernel(const P pol, const N* nodes, const Chunk* chunks, const Split* split,
                                                     const uint32_t* chunk_left_before, const uint32_t* table_h, const uint32_t* table_g,
                                                     const uint32_t* counters) {
    if (blockIdx.x >= counters[kCtrChunks]) return;
    const Chunk c = chunks[blockIdx.x];
    const Split s = split[c.slot];
    if (!s.split) return;
    const uint32_t first = nodes[c.node].first, last = first + nodes[c.node].count - 1u;
    bool left[kPerThread]; uint32_t local = 0u;
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        uint32_t p = c.begin + threadIdx.x * kPerThread + e;
        left[e] = p < c.end && pol.left(p, s);
        local += left[e] ? 1u : 0u;
    }
    uint32_t before = chunk_left_before[blockIdx.x] + block_exclusive(local);
    const uint32_t n_back_left = s.n_left - s.left_before_q;
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        uint32_t p = c.begin + threadIdx.x * kPerThread + e;
        if (p < c.end) {
            const bool front = p < s.q || (p == s.q && s.q_is_right);
            uint32_t dest;
            if (front) {
                if (left[e]) dest = p;
                else { uint32_t k = (p - first) - before; dest = k == 0u ? last : table_g[first + k - 1u] - 1u; }
            } else {
                if (left[e]) dest = table_h[first + (n_back_left - 1u - (before - s.left_before_q))];
                else dest = p - 1u;
            }
            pol.move(p, dest);
        }
        before += left[e] ? 1u : 0u;
    }
}

template <typename P>
__global__ void __launch_bounds__(kBB) copy_back_kernel(const P pol, const Chunk* chunks, const Split* split, const uint32_t* counters) {
    if (blockIdx.x >= counters[kCtrChunks]) return;
    const Chunk c = chunks[blockIdx.x];
    if (!split[c.slot].split) return;
    pol.copy_back(c.begin, c.end);
}

// ---- children (bvh.rs:432-461): allocate the pair, queue both for the next level
__global__ void children_kernel(BuildNode* nodes, const uint32_t* active, uint32_t n_active, const Split* split, uint32_t* next_active,
                                uint32_t* counters) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_active) return;
    const Split& s = split[slot];
    if (!s.split) return;
    BuildNode& n = nodes[active[slot]];
    if (s.n_left == 0u || s.n_left == n.count) return;                    // bvh.rs:432-435: stays a leaf (already permuted)
    const uint32_t left = atomicAdd(&counters[kCtrNodes], 2u);
    init_node(nodes[left], n.first, s.n_left, n.depth + 1u);
    init_node(nodes[left + 1u], n.first + s.n_left, n.count - s.n_left, n.depth + 1u);
    n.left = left;
    const uint32_t at = atomicAdd(&counters[kCtrNext], 2u);
    next_active[at] = left; next_active[at + 1u] = left + 1u;
    atomicMax(&counters[kCtrDepth], n.depth + 1u);
}


// ======================================================================================================================
// Leaf accelerator (leaf_accel.hpp): the conservative sub-BVHs inside the reference's leaves, built on the device.
// Same level-synchronous machinery, different policy: an INDEX array is partitioned (the reference's triangle order must
// not change), 16 bins with empty boxes, no cost-based termination (one triangle per sub leaf), SAH splits down to
// `sah_depth_limit`, then balanced half splits of the current arrangement (bounds the depth for the traversal stack).
// Only the topology is produced here (child references, parents, the index array); boxes and kappa are then filled
// bottom-up by refit_sub_nodes_kernel (upload_kernels.cu), the kernel that also runs after every vertex update.
struct SubNode {
    uint32_t cmin[3], cmax[3];     // encoded centroid bounds
    uint32_t first, count;         // range in the index array
    uint32_t depth, parent_ref;    // parent_ref = (parent << 1) | slot, 0xFFFFFFFF for the root of a reference leaf
    uint32_t n_left;
    uint32_t child[2];             // inner children (node ids) or kNoNode for sub leaves
    uint32_t pad[3];
};
static_assert(sizeof(SubNode) == 64, "SubNode is 64 bytes");
constexpr int kSubBinWords = 3 * kSubBins * 7;

__device__ __forceinline__ void init_sub_node(SubNode& n, uint32_t first, uint32_t count, uint32_t depth, uint32_t parent_ref) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { n.cmin[k] = enc(FLT_MAX); n.cmax[k] = enc(-FLT_MAX); }
    n.first = first; n.count = count; n.depth = depth; n.parent_ref = parent_ref; n.n_left = 0u;
    n.child[0] = kNoNode; n.child[1] = kNoNode;
}

__global__ void sub_init_kernel(const SubRoot* roots, uint32_t n_roots, SubNode* nodes, uint32_t* active, uint32_t* order, uint32_t* counters) {
    const uint32_t r = blockIdx.x;
    const SubRoot root = roots[r];
    if (threadIdx.x == 0) {
        init_sub_node(nodes[r], root.base, root.count, 0u, kNoNode);
        active[r] = r;
        if (r == 0) { counters[kCtrNodes] = n_roots; counters[kCtrNext] = 0u; counters[kCtrChunks] = 0u; counters[kCtrDepth] = 0u; }
    }
    for (uint32_t k = threadIdx.x; k < root.count; k += blockDim.x) order[root.base + k] = root.first_prim + k;
}

__global__ void __launch_bounds__(kBB) sub_bounds_kernel(const float* __restrict__ tris, const uint32_t* __restrict__ order, SubNode* nodes,
                                                        const Chunk* chunks, const uint32_t* counters) {
    if (blockIdx.x >= counters[kCtrChunks]) return;
    const Chunk c = chunks[blockIdx.x];
    float v[6] = { FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX };
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        uint32_t p = c.begin + threadIdx.x * kPerThread + e;
        if (p < c.end) {
            float t[9]; load_tri(tris, __ldg(order + p), t);
#pragma unroll
            for (int k = 0; k < 3; ++k) { float cen = centroid_axis(t, k); v[k] = fminf(v[k], cen); v[3 + k] = fmaxf(v[3 + k], cen); }
        }
    }
    __shared__ float red[kBB / 32][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            float y = __shfl_xor_sync(0xFFFFFFFFu, v[k], o);
            v[k] = k < 3 ? fminf(v[k], y) : fmaxf(v[k], y);
        }
        if (lane == 0) red[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int k = threadIdx.x;
        float x = red[0][k];
        for (int w = 1; w < kBB / 32; ++w) x = k < 3 ? fminf(x, red[w][k]) : fmaxf(x, red[w][k]);
        if (x == x) { SubNode& n = nodes[c.node]; if (k < 3) atomicMin(&n.cmin[k], enc(x)); else atomicMax(&n.cmax[k - 3], enc(x)); }
    }
}

__global__ void __launch_bounds__(kBB) sub_bins_kernel(const float* __restrict__ tris, const uint32_t* __restrict__ order, const SubNode* nodes,
                                                      const Chunk* chunks, uint32_t* bins, const uint32_t* counters) {
    if (blockIdx.x >= counters[kCtrChunks]) return;
    const Chunk c = chunks[blockIdx.x];
    __shared__ uint32_t sb[kSubBinWords];
    const uint32_t lo0 = enc(FLT_MAX), hi0 = enc(-FLT_MAX);
    for (int i = threadIdx.x; i < kSubBinWords; i += kBB) { const int w = i % 7; sb[i] = w == 6 ? 0u : (w < 3 ? lo0 : hi0); }
    float bmin[3], scale[3]; bool on[3];
    {
        const SubNode& n = nodes[c.node];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            bmin[a] = dec(n.cmin[a]);
            float bmax = dec(n.cmax[a]);
            on[a] = bmax > bmin[a];
            scale[a] = (float)kSubBins / (bmax - bmin[a]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kPerThread; ++e) {
        uint32_t p = c.begin + threadIdx.x * kPerThread + e;
        if (p >= c.end) continue;
        float t[9]; load_tri(tris, __ldg(order + p), t);
        uint32_t lo[3], hi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = enc(fminf(t[k], fminf(t[3 + k], t[6 + k])));
            hi[k] = enc(fmaxf(t[k], fmaxf(t[3 + k], t[6 + k])));
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (!on[a]) continue;
            int bi = (int)((centroid_axis(t, a) - bmin[a]) * scale[a]);       // must match SubPolicy::left
            bi = max(0, min(kSubBins - 1, bi));
            uint32_t* b = sb + (a * kSubBins + bi) * 7;
            atomicAdd(b + 6, 1u);
#pragma unroll
            for (int k = 0; k < 3; ++k) { atomicMin(b + k, lo[k]); atomicMax(b + 3 + k, hi[k]); }
        }
    }
    __syncthreads();
    uint32_t* g = bins + (size_t)c.slot * kSubBinWords;
    for (int i = threadIdx.x; i < kSubBinWords; i += kBB) {
        const int w = i % 7;
        if (sb[i - w + 6] == 0u) continue;
        if (w == 6) atomicAdd(g + i, sb[i]);
        else if (w < 3) atomicMin(g + i, sb[i]);
        else atomicMax(g + i, sb[i]);
    }
}

__global__ void sub_select_kernel(const SubNode* nodes, const uint32_t* active, uint32_t n_active, const uint32_t* bins, Split* split,
                                  uint32_t sah_depth_limit) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_active) return;
    const SubNode& n = nodes[active[slot]];
    Split& s = split[slot];
    int best_axis = -1, best_bin = -1; float best_cost = FLT_MAX, best_lo = 0.0f, best_scale = 0.0f;
    if (n.depth < sah_depth_limit) {
        for (int a = 0; a < 3; ++a) {
            const float lo = dec(n.cmin[a]), hi = dec(n.cmax[a]);
            if (!(hi > lo)) continue;
            const uint32_t* b = bins + (size_t)slot * kSubBinWords + a * kSubBins * 7;
            // right-to-left sweep first (areas + counts of the right side of every plane), then left-to-right with the cost
            float ra[kSubBins - 1]; uint32_t rc[kSubBins - 1];
            float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
            uint32_t cnt = 0;
            for (int i = kSubBins - 1; i >= 1; --i) {
                const uint32_t* bb = b + i * 7;
                if (bb[6]) { for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], dec(bb[k])); mx[k] = fmaxf(mx[k], dec(bb[3 + k])); } }
                cnt += bb[6];
                rc[i - 1] = cnt; ra[i - 1] = cnt ? area(mn, mx) : 0.0f;
            }
            for (int k = 0; k < 3; ++k) { mn[k] = FLT_MAX; mx[k] = -FLT_MAX; }
            cnt = 0;
            for (int i = 0; i < kSubBins - 1; ++i) {
                const uint32_t* bb = b + i * 7;
                if (bb[6]) { for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], dec(bb[k])); mx[k] = fmaxf(mx[k], dec(bb[3 + k])); } }
                cnt += bb[6];
                if (cnt == 0u || rc[i] == 0u) continue;
                float cost = (float)cnt * area(mn, mx) + (float)rc[i] * ra[i];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = i; best_lo = lo; best_scale = (float)kSubBins / (hi - lo); }
            }
        }
    }
    s.split = 1u;
    if (best_axis >= 0) { s.axis = best_axis; s.lo = best_lo; s.scale = best_scale; s.bin = best_bin; }
    else { s.axis = 3; s.first = n.first; s.half = n.count / 2u; }      // identical centroids or depth limit: balanced half split
}

__global__ void sub_children_kernel(SubNode* nodes, const uint32_t* active, uint32_t n_active, const Split* split, uint32_t* next_active,
                                    uint32_t* counters, uint32_t max_sub_leaf) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_active) return;
    const uint32_t id = active[slot];
    SubNode& n = nodes[id];
    uint32_t n_left = split[slot].n_left;
    if (n_left == 0u || n_left >= n.count) n_left = n.count / 2u;          // cannot happen (the plane had both sides populated)
    n.n_left = n_left;
    const uint32_t first[2] = { n.first, n.first + n_left }, cnt[2] = { n_left, n.count - n_left };
    for (int c = 0; c < 2; ++c) {
        if (cnt[c] <= max_sub_leaf) continue;
        const uint32_t child = atomicAdd(&counters[kCtrNodes], 1u);
        init_sub_node(nodes[child], first[c], cnt[c], n.depth + 1u, (id << 1) | (uint32_t)c);
        n.child[c] = child;
        next_active[atomicAdd(&counters[kCtrNext], 1u)] = child;
    }
    atomicMax(&counters[kCtrDepth], n.depth + 1u);
}

__global__ void sub_emit_kernel(const SubNode* nodes, uint32_t n_nodes, float4* raw, uint32_t* parent) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_nodes) return;
    const SubNode& n = nodes[id];
    const uint32_t first[2] = { n.first, n.first + n.n_left }, cnt[2] = { n.n_left, n.count - n.n_left };
    for (int c = 0; c < 2; ++c) {
        uint32_t ref = n.child[c] != kNoNode ? n.child[c] : (0x80000000u | ((cnt[c] - 1u) << 28) | first[c]);
        raw[4 * (size_t)id + 2 * c] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        raw[4 * (size_t)id + 2 * c + 1] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(ref));
    }
    parent[id] = n.parent_ref;
}

template <typename T> cudaError_t grow(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t n = need + need / 4 + 16;
    cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
}

} // namespace

struct BuildWorkspace::Impl {
    BuildNode* nodes = nullptr;   size_t nodes_cap = 0;
    uint32_t* active[2] = { nullptr, nullptr }; size_t active_cap[2] = { 0, 0 };
    Split* split = nullptr;       size_t split_cap = 0;
    uint32_t* bins = nullptr;     size_t bins_cap = 0;
    Chunk* chunks = nullptr;      size_t chunks_cap = 0;
    uint32_t* chunk_left = nullptr; size_t chunk_left_cap = 0;
    uint32_t* chunk_before = nullptr; size_t chunk_before_cap = 0;
    uint32_t* table_h = nullptr;  size_t table_h_cap = 0;
    uint32_t* table_g = nullptr;  size_t table_g_cap = 0;
    float* tris_tmp = nullptr;    size_t tris_tmp_cap = 0;
    uint32_t* perm_tmp = nullptr; size_t perm_tmp_cap = 0;
    uint32_t* counters = nullptr; size_t counters_cap = 0;
    SubNode* sub_nodes = nullptr; size_t sub_nodes_cap = 0;
    SubRoot* roots = nullptr;     size_t roots_cap = 0;
};

BuildWorkspace::BuildWorkspace() : impl_(new Impl()) {}
BuildWorkspace::~BuildWorkspace() { release(); delete impl_; }
void BuildWorkspace::release() {
    Impl& w = *impl_;
    void* ptrs[] = { w.nodes, w.active[0], w.active[1], w.split, w.bins, w.chunks, w.chunk_left, w.chunk_before, w.table_h, w.table_g,
                     w.tris_tmp, w.perm_tmp, w.counters, w.sub_nodes, w.roots };
    for (void* p : ptrs) if (p) cudaFree(p);
    *impl_ = Impl();
}

#define BVHT_BUILD_CHECK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return e__; } while (0)

cudaError_t device_build_reference_bvh(BuildWorkspace& ws, float* tris, uint32_t* perm, uint32_t n_tris, cudaStream_t s,
                                       std::vector<BuildNodeHost>& out_nodes, DeviceBuildStats* stats) {
    BuildWorkspace::Impl& w = *ws.impl_;
    out_nodes.clear();
    if (n_tris == 0) return cudaSuccess;
    const size_t n = n_tris;
    BVHT_BUILD_CHECK(grow(w.nodes, w.nodes_cap, 2 * n + 2));
    BVHT_BUILD_CHECK(grow(w.active[0], w.active_cap[0], n + 2));
    BVHT_BUILD_CHECK(grow(w.active[1], w.active_cap[1], n + 2));
    BVHT_BUILD_CHECK(grow(w.table_h, w.table_h_cap, n));
    BVHT_BUILD_CHECK(grow(w.table_g, w.table_g_cap, n));
    BVHT_BUILD_CHECK(grow(w.tris_tmp, w.tris_tmp_cap, n * 9));
    BVHT_BUILD_CHECK(grow(w.perm_tmp, w.perm_tmp_cap, n));
    BVHT_BUILD_CHECK(grow(w.counters, w.counters_cap, (size_t)kCtrCount));

    init_root_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 1024), 256, 0, s>>>(w.nodes, w.active[0], w.counters, perm, n_tris);
    uint32_t n_active = 1, levels = 0, launches = 1;
    int cur = 0;
    while (n_active > 0) {
        const size_t ub = n / kChunk + n_active + 1;                      // upper bound of this level's chunk count
        BVHT_BUILD_CHECK(grow(w.split, w.split_cap, (size_t)n_active));
        BVHT_BUILD_CHECK(grow(w.bins, w.bins_cap, (size_t)n_active * kBinWords));
        BVHT_BUILD_CHECK(grow(w.chunks, w.chunks_cap, ub));
        BVHT_BUILD_CHECK(grow(w.chunk_left, w.chunk_left_cap, ub));
        BVHT_BUILD_CHECK(grow(w.chunk_before, w.chunk_before_cap, ub));
        const unsigned gs = (n_active + 127) / 128, gc = (unsigned)ub;
        plan_count_kernel<BuildNode, kBinWords, true><<<gs, 128, 0, s>>>(w.nodes, w.active[cur], n_active, w.split, w.bins, w.counters);
        plan_scan_kernel<<<1, 1024, 0, s>>>(w.split, n_active, w.counters);
        plan_fill_kernel<BuildNode><<<gs, 128, 0, s>>>(w.nodes, w.active[cur], n_active, w.split, w.chunks);
        bounds_kernel<<<gc, kBB, 0, s>>>(tris, w.nodes, w.chunks, w.counters);
        bins_kernel<<<gc, kBB, 0, s>>>(tris, w.nodes, w.chunks, w.bins, w.counters);
        select_kernel<<<gs, 128, 0, s>>>(w.nodes, w.active[cur], n_active, w.bins, w.split);
        const RefPolicy pol = { tris, perm, w.tris_tmp, w.perm_tmp, tris, perm };
        count_kernel<<<gc, kBB, 0, s>>>(pol, w.chunks, w.split, w.chunk_left, w.counters);
        scan_kernel<<<(n_active + 3) / 4, 128, 0, s>>>(pol, w.nodes, w.active[cur], n_active, w.split, w.chunk_left, w.chunk_before);
        tables_kernel<<<gc, kBB, 0, s>>>(pol, w.nodes, w.chunks, w.split, w.chunk_before, w.table_h, w.table_g, w.counters);
        scatter_kernel<<<gc, kBB, 0, s>>>(pol, w.nodes, w.chunks, w.split, w.chunk_before, w.table_h, w.table_g, w.counters);
        copy_back_kernel<<<gc, kBB, 0, s>>>(pol, w.chunks, w.split, w.counters);
        children_kernel<<<gs, 128, 0, s>>>(w.nodes, w.active[cur], n_active, w.split, w.active[cur ^ 1], w.counters);
        launches += 12;
        BVHT_BUILD_CHECK(cudaGetLastError());
        uint32_t next = 0;
        BVHT_BUILD_CHECK(cudaMemcpyAsync(&next, w.counters + kCtrNext, 4, cudaMemcpyDeviceToHost, s));
        BVHT_BUILD_CHECK(cudaStreamSynchronize(s));
        n_active = next; cur ^= 1; ++levels;
    }
    uint32_t ctr[kCtrCount];
    BVHT_BUILD_CHECK(cudaMemcpyAsync(ctr, w.counters, sizeof ctr, cudaMemcpyDeviceToHost, s));
    BVHT_BUILD_CHECK(cudaStreamSynchronize(s));
    std::vector<BuildNode> raw(ctr[kCtrNodes]);
    BVHT_BUILD_CHECK(cudaMemcpyAsync(raw.data(), w.nodes, raw.size() * sizeof(BuildNode), cudaMemcpyDeviceToHost, s));
    BVHT_BUILD_CHECK(cudaStreamSynchronize(s));

    // renumber: children pairs in DFS pre-order of the splits, node 1 is the alignment dummy (bvh.rs:514-517)
    auto dec_h = [](uint32_t u) { uint32_t b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u; float f; memcpy(&f, &b, 4); return f; };
    out_nodes.assign(2, BuildNodeHost());
    std::vector<std::pair<uint32_t, uint32_t>> stack;                     // (temp index, final index)
    stack.emplace_back(0u, 0u);
    while (!stack.empty()) {
        auto [ti, fi] = stack.back(); stack.pop_back();
        const BuildNode& r = raw[ti];
        BuildNodeHost h;
        for (int k = 0; k < 3; ++k) { h.aabb_min[k] = dec_h(r.vmin[k]); h.aabb_max[k] = dec_h(r.vmax[k]); }
        if (r.left != 0u) {
            const uint32_t l = (uint32_t)out_nodes.size();
            out_nodes.resize(out_nodes.size() + 2);
            h.prim_count = 0u; h.left_first = l;
            stack.emplace_back(r.left + 1u, l + 1u);                      // right is visited after the whole left subtree
            stack.emplace_back(r.left, l);
        } else {
            h.prim_count = r.count; h.left_first = r.first;
        }
        out_nodes[fi] = h;
    }
    if (stats) { stats->levels = levels; stats->launches = launches; stats->temp_nodes = ctr[kCtrNodes]; stats->max_depth = ctr[kCtrDepth]; }
    return cudaSuccess;
}

cudaError_t device_build_leaf_accel(BuildWorkspace& ws, const float* tris, const SubRoot* roots_host, uint32_t n_roots, uint32_t n_sub,
                                    float4* sub_raw, uint32_t* sub_parent, uint32_t* order, uint32_t max_sub_leaf, uint32_t sah_depth_limit,
                                    cudaStream_t s, DeviceSubResult* res) {
    BuildWorkspace::Impl& w = *ws.impl_;
    if (res) *res = DeviceSubResult();
    if (n_roots == 0 || n_sub == 0) return cudaSuccess;
    const size_t n = n_sub;
    BVHT_BUILD_CHECK(grow(w.sub_nodes, w.sub_nodes_cap, n + 1));
    BVHT_BUILD_CHECK(grow(w.roots, w.roots_cap, (size_t)n_roots));
    BVHT_BUILD_CHECK(grow(w.active[0], w.active_cap[0], n + 2));
    BVHT_BUILD_CHECK(grow(w.active[1], w.active_cap[1], n + 2));
    BVHT_BUILD_CHECK(grow(w.table_h, w.table_h_cap, n));
    BVHT_BUILD_CHECK(grow(w.table_g, w.table_g_cap, n));
    BVHT_BUILD_CHECK(grow(w.perm_tmp, w.perm_tmp_cap, n));
    BVHT_BUILD_CHECK(grow(w.counters, w.counters_cap, (size_t)kCtrCount));
    BVHT_BUILD_CHECK(cudaMemcpyAsync(w.roots, roots_host, (size_t)n_roots * sizeof(SubRoot), cudaMemcpyHostToDevice, s));
    sub_init_kernel<<<n_roots, 128, 0, s>>>(w.roots, n_roots, w.sub_nodes, w.active[0], order, w.counters);
    uint32_t n_active = n_roots, levels = 0, launches = 1;
    int cur = 0;
    const SubPolicy pol = { tris, order, w.perm_tmp, order };
    while (n_active > 0) {
        const size_t ub = n / kChunk + n_active + 1;
        BVHT_BUILD_CHECK(grow(w.split, w.split_cap, (size_t)n_active));
        BVHT_BUILD_CHECK(grow(w.bins, w.bins_cap, (size_t)n_active * kSubBinWords));
        BVHT_BUILD_CHECK(grow(w.chunks, w.chunks_cap, ub));
        BVHT_BUILD_CHECK(grow(w.chunk_left, w.chunk_left_cap, ub));
        BVHT_BUILD_CHECK(grow(w.chunk_before, w.chunk_before_cap, ub));
        const unsigned gs = (n_active + 127) / 128, gc = (unsigned)ub;
        plan_count_kernel<SubNode, kSubBinWords, false><<<gs, 128, 0, s>>>(w.sub_nodes, w.active[cur], n_active, w.split, w.bins, w.counters);
        plan_scan_kernel<<<1, 1024, 0, s>>>(w.split, n_active, w.counters);
        plan_fill_kernel<SubNode><<<gs, 128, 0, s>>>(w.sub_nodes, w.active[cur], n_active, w.split, w.chunks);
        sub_bounds_kernel<<<gc, kBB, 0, s>>>(tris, order, w.sub_nodes, w.chunks, w.counters);
        sub_bins_kernel<<<gc, kBB, 0, s>>>(tris, order, w.sub_nodes, w.chunks, w.bins, w.counters);
        sub_select_kernel<<<gs, 128, 0, s>>>(w.sub_nodes, w.active[cur], n_active, w.bins, w.split, sah_depth_limit);
        count_kernel<<<gc, kBB, 0, s>>>(pol, w.chunks, w.split, w.chunk_left, w.counters);
        scan_kernel<<<(n_active + 3) / 4, 128, 0, s>>>(pol, w.sub_nodes, w.active[cur], n_active, w.split, w.chunk_left, w.chunk_before);
        tables_kernel<<<gc, kBB, 0, s>>>(pol, w.sub_nodes, w.chunks, w.split, w.chunk_before, w.table_h, w.table_g, w.counters);
        scatter_kernel<<<gc, kBB, 0, s>>>(pol, w.sub_nodes, w.chunks, w.split, w.chunk_before, w.table_h, w.table_g, w.counters);
        copy_back_kernel<<<gc, kBB, 0, s>>>(pol, w.chunks, w.split, w.counters);
        sub_children_kernel<<<gs, 128, 0, s>>>(w.sub_nodes, w.active[cur], n_active, w.split, w.active[cur ^ 1], w.counters, max_sub_leaf);
        launches += 12;
        BVHT_BUILD_CHECK(cudaGetLastError());
        uint32_t next = 0;
        BVHT_BUILD_CHECK(cudaMemcpyAsync(&next, w.counters + kCtrNext, 4, cudaMemcpyDeviceToHost, s));
        BVHT_BUILD_CHECK(cudaStreamSynchronize(s));
        n_active = next; cur ^= 1; ++levels;
    }
    uint32_t ctr[kCtrCount];
    BVHT_BUILD_CHECK(cudaMemcpyAsync(ctr, w.counters, sizeof ctr, cudaMemcpyDeviceToHost, s));
    BVHT_BUILD_CHECK(cudaStreamSynchronize(s));
    sub_emit_kernel<<<(ctr[kCtrNodes] + 127) / 128, 128, 0, s>>>(w.sub_nodes, ctr[kCtrNodes], sub_raw, sub_parent);
    BVHT_BUILD_CHECK(cudaGetLastError());
    if (res) { res->n_nodes = ctr[kCtrNodes]; res->max_depth = ctr[kCtrDepth]; res->levels = levels; res->launches = launches + 1; }
    return cudaSuccess;
}

} // namespace bvht
