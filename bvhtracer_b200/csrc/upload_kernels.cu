// upload_kernels.cu -- K3: repack the reference's triangle buffer (Triangle<f32> = 3 x Vector3, 36 B,
// geometry/triangle.rs:9-16 viewed through Mesh::primitives, mesh.rs:126-134) into 3 aligned float4 per triangle
//   {v0, prim} | {e1 = v1 - v0, 0} | {e2 = v2 - v0, 0}
// The two edge subtractions are the first two operations of Triangle::intersect (triangle.rs:43-44); doing
// them once on upload is bit-identical (one IEEE subtraction either way; no FMA can form here).
// HBM-bound streaming kernel: 36 B read + 48 B written per triangle, loads coalesced through shared memory.
#include <algorithm>
#include <cstring>
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

// One baked child box in the form the trace kernel reads (device_types.cuh BVHT_SUB_CH): {lo, hi} or {centre, half extent}.
// [c - h, c + h] contains [lo, hi]; h also absorbs the two extra roundings of slab_test_ch (eps * |t_centre| and
// eps * |h f|, together < 2 eps (|o| + |c| + h) in space units: abs_ / 4 = 4 eps s_max covers it).
__device__ __forceinline__ void store_baked_child(float4* __restrict__ out, size_t node, int slot, float4 olo, float4 ohi, float abs_,
                                                  float w0, float w1) {
#if BVHT_SUB_CH
    float4 ce, ha;
    ce.x = __fmul_rn(0.5f, __fadd_rn(olo.x, ohi.x));
    ce.y = __fmul_rn(0.5f, __fadd_rn(olo.y, ohi.y));
    ce.z = __fmul_rn(0.5f, __fadd_rn(olo.z, ohi.z));
    ha.x = fmaxf(__fsub_ru(ohi.x, ce.x), __fsub_ru(ce.x, olo.x));
    ha.y = fmaxf(__fsub_ru(ohi.y, ce.y), __fsub_ru(ce.y, olo.y));
    ha.z = fmaxf(__fsub_ru(ohi.z, ce.z), __fsub_ru(ce.z, olo.z));
    const float extra = __fmul_ru(abs_, 0.25f);
    if (olo.x <= ohi.x && olo.y <= ohi.y && olo.z <= ohi.z) {
        ha.x = __fadd_ru(__fmul_ru(ha.x, 1.000001f), extra);
        ha.y = __fadd_ru(__fmul_ru(ha.y, 1.000001f), extra);
        ha.z = __fadd_ru(__fmul_ru(ha.z, 1.000001f), extra);
        // boxes blown up to infinity (absurd ray limits): an always-hit axis instead of inf - inf = NaN
        if (!(fabsf(ce.x) <= 1e37f) || !(ha.x <= 1e37f)) { ce.x = 0.0f; ha.x = 3.0e38f; }
        if (!(fabsf(ce.y) <= 1e37f) || !(ha.y <= 1e37f)) { ce.y = 0.0f; ha.y = 3.0e38f; }
        if (!(fabsf(ce.z) <= 1e37f) || !(ha.z <= 1e37f)) { ce.z = 0.0f; ha.z = 3.0e38f; }
    } else {
        // empty box (only degenerate triangles below): a negative half extent can never be hit
        ce.x = ce.y = ce.z = 0.0f; ha.x = ha.y = ha.z = -1.0f;
    }
    olo = ce; ohi = ha;
#endif
    olo.w = w0; ohi.w = w1;
    out[4 * node + 2 * slot] = olo;
    out[4 * node + 2 * slot + 1] = ohi;
}

// Bake the sub-BVH for the current ray limits, bottom-up.  Every TRIANGLE's box is grown by its own
// delta = scale * |e1||e2| + abs (leaf_accel.hpp accel_deltas; every operation rounded AWAY from the box, plus a relative
// 1e-6), and an inner box is the exact min/max union of its children's baked boxes -- far tighter than growing the union
// by the largest delta below it when triangle sizes are uneven (big_ben_clock: 1.21 -> measured in DESIGN.md).  Triangles
// with a zero edge are never accepted by Triangle::intersect (area == 0 exactly) and contribute no box: the 999-sentinel
// of every .tri asset no longer stretches the boxes above it.  Same arrival-counter climb as refit_sub_nodes_kernel;
// `lohi` is scratch in {lo, hi} form (the climb must not lose exactness to the centre/half conversion).
// Runs at upload, after every vertex update and whenever the limits change (bvht_api.cu ensure_bake).
__global__ void __launch_bounds__(kRepackBlock)
bake_sub_nodes_kernel(const float4* __restrict__ raw, const uint32_t* __restrict__ parent, unsigned int* __restrict__ counters,
                      const float* __restrict__ tris, const uint32_t* __restrict__ order, uint32_t n_nodes, float scale, float abs_,
                      float4* __restrict__ lohi, float4* __restrict__ out) {
    uint32_t node = blockIdx.x * kRepackBlock + threadIdx.x;
    if (node >= n_nodes) return;
    const float ref0 = __ldg(reinterpret_cast<const float*>(raw + 4 * (size_t)node + 1) + 3);
    const float ref1 = __ldg(reinterpret_cast<const float*>(raw + 4 * (size_t)node + 3) + 3);
    unsigned int arrivals = 0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t ref = __float_as_uint(c ? ref1 : ref0);
        if (!(ref & 0x80000000u)) continue;
        uint32_t first = ref & 0x0FFFFFFFu, cnt = ((ref >> 28) & 7u) + 1u;
        float4 lo = make_float4(3.402823466e38f, 3.402823466e38f, 3.402823466e38f, 0.0f);
        float4 hi = make_float4(-3.402823466e38f, -3.402823466e38f, -3.402823466e38f, 0.0f);
        for (uint32_t k = 0; k < cnt; ++k) {
            const float* t = tris + (size_t)__ldg(order + first + k) * 9;
            float v[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) v[j] = __ldg(t + j);
            double e1x = (double)v[3] - v[0], e1y = (double)v[4] - v[1], e1z = (double)v[5] - v[2];
            double e2x = (double)v[6] - v[0], e2y = (double)v[7] - v[1], e2z = (double)v[8] - v[2];
            double kp = sqrt(e1x * e1x + e1y * e1y + e1z * e1z) * sqrt(e2x * e2x + e2y * e2y + e2z * e2z);
            if (!(kp > 0.0)) continue;                       // a zero edge: never accepted
            float d = __fmaf_ru(scale, __double2float_ru(kp * (1.0 + 1e-12)), abs_);
            float tl[3], th[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                tl[a] = fminf(fminf(v[a], v[3 + a]), v[6 + a]);
                th[a] = fmaxf(fmaxf(v[a], v[3 + a]), v[6 + a]);
                tl[a] = __fsub_rd(__fsub_rd(tl[a], d), __fmul_ru(fabsf(tl[a]), 1e-6f));
                th[a] = __fadd_ru(__fadd_ru(th[a], d), __fmul_ru(fabsf(th[a]), 1e-6f));
            }
            lo.x = fminf(lo.x, tl[0]); lo.y = fminf(lo.y, tl[1]); lo.z = fminf(lo.z, tl[2]);
            hi.x = fmaxf(hi.x, th[0]); hi.y = fmaxf(hi.y, th[1]); hi.z = fmaxf(hi.z, th[2]);
        }
        __stcg(lohi + 4 * (size_t)node + 2 * c, lo);
        __stcg(lohi + 4 * (size_t)node + 2 * c + 1, hi);
        store_baked_child(out, node, c, lo, hi, abs_, c ? 0.0f : ref0, c ? 0.0f : ref1);
        arrivals++;
    }
    if (arrivals == 0) return;                       // both children are inner nodes: whoever completes them climbs through here
    __threadfence();
    unsigned int old = atomicAdd(counters + node, arrivals);
    if (old + arrivals != 2u) return;
    for (;;) {                                       // this thread completed `node`: climb
        __threadfence();
        uint32_t pp = __ldg(parent + node);
        if (pp == 0xFFFFFFFFu) break;                // root of a sub tree: its children ARE what the trace kernel tests
        float4 a = __ldcg(lohi + 4 * (size_t)node + 0), b = __ldcg(lohi + 4 * (size_t)node + 1);
        float4 c = __ldcg(lohi + 4 * (size_t)node + 2), d = __ldcg(lohi + 4 * (size_t)node + 3);
        uint32_t p = pp >> 1, slot = pp & 1u;
        float4 plo = make_float4(fminf(a.x, c.x), fminf(a.y, c.y), fminf(a.z, c.z), 0.0f);
        float4 phi = make_float4(fmaxf(b.x, d.x), fmaxf(b.y, d.y), fmaxf(b.z, d.z), 0.0f);
        __stcg(lohi + 4 * (size_t)p + 2 * slot, plo);
        __stcg(lohi + 4 * (size_t)p + 2 * slot + 1, phi);
        const float pr0 = __ldg(reinterpret_cast<const float*>(raw + 4 * (size_t)p + 1) + 3);
        const float pr1 = __ldg(reinterpret_cast<const float*>(raw + 4 * (size_t)p + 3) + 3);
        store_baked_child(out, p, (int)slot, plo, phi, abs_, slot ? 0.0f : pr0, slot ? 0.0f : pr1);
        __threadfence();
        unsigned int o = atomicAdd(counters + p, 1u);
        if (o + 1u != 2u) break;
        node = p;
    }
}

// Refit of the sub-BVHs after a vertex update (topology kept, raw boxes and kappa recomputed bottom-up): one thread per
// sub node computes its LEAF children from the new vertices, then nodes complete when both children have arrived
// (atomic arrival counters) and the completing thread climbs, writing the node's box into its slot of the parent.
// Any topology is valid for the conservative search; only its quality degrades with deformation.  HBM/L2-bound:
// 36 B per triangle + 64 B per node.
__device__ __forceinline__ void leaf_child_box(const float* __restrict__ tris, const uint32_t* __restrict__ order, uint32_t ref,
                                               float lo[3], float hi[3], float& kappa) {
    uint32_t first = ref & 0x0FFFFFFFu, cnt = ((ref >> 28) & 7u) + 1u;
    lo[0] = lo[1] = lo[2] = 3.402823466e38f; hi[0] = hi[1] = hi[2] = -3.402823466e38f;
    double kmax = 0.0;
    for (uint32_t k = 0; k < cnt; ++k) {
        const float* t = tris + (size_t)__ldg(order + first + k) * 9;
        float v[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] = __ldg(t + j);
#pragma unroll
        for (int j = 0; j < 9; ++j) { lo[j % 3] = fminf(lo[j % 3], v[j]); hi[j % 3] = fmaxf(hi[j % 3], v[j]); }
        double e1x = (double)v[3] - v[0], e1y = (double)v[4] - v[1], e1z = (double)v[5] - v[2];
        double e2x = (double)v[6] - v[0], e2y = (double)v[7] - v[1], e2z = (double)v[8] - v[2];
        double kp = sqrt(e1x * e1x + e1y * e1y + e1z * e1z) * sqrt(e2x * e2x + e2y * e2y + e2z * e2z);
        kmax = fmax(kmax, kp);
    }
    kappa = __double2float_ru(kmax * (1.0 + 1e-12));
}

__global__ void __launch_bounds__(kRepackBlock)
refit_sub_nodes_kernel(float4* __restrict__ raw, const uint32_t* __restrict__ parent, unsigned int* __restrict__ counters,
                       const float* __restrict__ tris, const uint32_t* __restrict__ order, uint32_t n_nodes) {
    uint32_t node = blockIdx.x * kRepackBlock + threadIdx.x;
    if (node >= n_nodes) return;
    unsigned int arrivals = 0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t ref = __float_as_uint(__ldcg(reinterpret_cast<const float*>(raw + 4 * (size_t)node + 2 * c + 1) + 3));
        if (ref & 0x80000000u) {
            float lo[3], hi[3], kp;
            leaf_child_box(tris, order, ref, lo, hi, kp);
            __stcg(raw + 4 * (size_t)node + 2 * c, make_float4(lo[0], lo[1], lo[2], kp));
            __stcg(raw + 4 * (size_t)node + 2 * c + 1, make_float4(hi[0], hi[1], hi[2], __uint_as_float(ref)));
            arrivals++;
        }
    }
    if (arrivals == 0) return;                       // both children are inner nodes: whoever completes them climbs through here
    __threadfence();
    unsigned int old = atomicAdd(counters + node, arrivals);
    if (old + arrivals != 2u) return;
    // this thread completed `node`: climb
    for (;;) {
        __threadfence();
        uint32_t pp = __ldg(parent + node);
        if (pp == 0xFFFFFFFFu) break;                // root of a leaf's sub tree: its children ARE what the trace kernel tests
        float4 a = __ldcg(raw + 4 * (size_t)node + 0), b = __ldcg(raw + 4 * (size_t)node + 1);
        float4 c = __ldcg(raw + 4 * (size_t)node + 2), d = __ldcg(raw + 4 * (size_t)node + 3);
        uint32_t p = pp >> 1, slot = pp & 1u;
        float4 plo = make_float4(fminf(a.x, c.x), fminf(a.y, c.y), fminf(a.z, c.z), fmaxf(a.w, c.w));
        float keep = __ldcg(reinterpret_cast<const float*>(raw + 4 * (size_t)p + 2 * slot + 1) + 3);   // the child reference (= node)
        float4 phi = make_float4(fmaxf(b.x, d.x), fmaxf(b.y, d.y), fmaxf(b.z, d.z), keep);
        __stcg(raw + 4 * (size_t)p + 2 * slot, plo);
        __stcg(raw + 4 * (size_t)p + 2 * slot + 1, phi);
        __threadfence();
        unsigned int o = atomicAdd(counters + p, 1u);
        if (o + 1u != 2u) break;
        node = p;
    }
}

// max |o|^2 and max |d|^2 over a ray buffer (7 floats per ray): lets the leaf accelerator be baked for an arbitrary
// batch of rays before it is traced.  Non-negative floats order like their bit patterns, so atomicMax on uint works.
__global__ void __launch_bounds__(kRepackBlock)
ray_bounds_kernel(const float* __restrict__ rays, unsigned long long n, unsigned int* __restrict__ out) {
    float mo = 0.0f, md = 0.0f;
    for (unsigned long long i = blockIdx.x * (unsigned long long)kRepackBlock + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * kRepackBlock) {
        const float* r = rays + i * 7;
        float ox = __ldg(r), oy = __ldg(r + 1), oz = __ldg(r + 2), dx = __ldg(r + 3), dy = __ldg(r + 4), dz = __ldg(r + 5);
        float o2 = ox * ox + oy * oy + oz * oz, d2 = dx * dx + dy * dy + dz * dz;
        if (!(o2 <= 3.0e38f)) o2 = 3.0e38f;          // NaN / inf: make the limits unreachable -> brute-force leaves
        if (!(d2 <= 3.0e38f)) d2 = 3.0e38f;
        mo = fmaxf(mo, o2); md = fmaxf(md, d2);
    }
    for (int off = 16; off > 0; off >>= 1) {
        mo = fmaxf(mo, __shfl_xor_sync(0xFFFFFFFFu, mo, off));
        md = fmaxf(md, __shfl_xor_sync(0xFFFFFFFFu, md, off));
    }
    if ((threadIdx.x & 31) == 0) { atomicMax(out + 0, __float_as_uint(mo)); atomicMax(out + 1, __float_as_uint(md)); }
}

// Read-bandwidth probe: every CTA streams the whole buffer `iters` times with 16-byte loads.  With a buffer that fits in L2
// this measures the L2 -> SM delivery rate the roofline of an L2-resident traversal should be compared with (SURVEY.md 8d ii);
// with a buffer far larger than L2 it measures HBM.
__global__ void __launch_bounds__(256)
read_bw_kernel(const uint4* __restrict__ buf, size_t n_vec, uint32_t iters, unsigned long long* sink) {
    unsigned int acc = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        // rotate the start per CTA and iteration so that CTAs do not march in lock step over the same lines
        size_t start = ((size_t)blockIdx.x * 977 + (size_t)it * 131) * 256 % n_vec;
        for (size_t i = threadIdx.x; i < n_vec; i += 256 * (size_t)gridDim.x) {
            size_t j = start + i + (size_t)blockIdx.x * 256;
            j = j % n_vec;
            uint4 v = __ldcg(buf + j);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x9E3779B9u) atomicAdd(sink, 1ull);      // keeps the loads alive
}


// ---- whole-model statistics and tight box of a deforming model, on the device (bvht_blas_update_vertices runs per frame for
// big_ben_clock: the two host passes over the triangles cost 0.46 ms of a 1.9 ms frame; here they are two streaming kernels).
// Arithmetic is the host functions' (leaf_accel.cpp compute_model_stats, bvht_api.cu bake_accel), in double: the maxima and the
// boxes are order-independent, only the two means (a heuristic's inputs) depend on the summation order.
__device__ __forceinline__ unsigned long long enc_f64(double d) {        // order-preserving u64 of a double
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned int enc_f32(float f) {
    unsigned int b = __float_as_uint(f);
    return (b >> 31) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(kRepackBlock)
model_stats_kernel(const float* __restrict__ tris, uint32_t n_tris, ModelStatsDev* __restrict__ out) {
    double radius2 = 0.0, edge2 = 0.0, kappa2max = 0.0, sum_edge = 0.0, sum_kappa = 0.0;
    unsigned long long good = 0;
    float lo[3] = { 3.402823466e38f, 3.402823466e38f, 3.402823466e38f }, hi[3] = { -3.402823466e38f, -3.402823466e38f, -3.402823466e38f };
    for (uint32_t i = blockIdx.x * kRepackBlock + threadIdx.x; i < n_tris; i += gridDim.x * kRepackBlock) {
        const float* t = tris + (size_t)i * 9;
        float v[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] = __ldg(t + j);
        double l1 = 0.0, l2 = 0.0, l3 = 0.0, n0 = 0.0, n1 = 0.0, n2 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double a = v[k], b = v[3 + k], c = v[6 + k];
            const double e1 = b - a, e2 = c - a, e3 = c - b;
            l1 += e1 * e1; l2 += e2 * e2; l3 += e3 * e3;
            n0 += a * a; n1 += b * b; n2 += c * c;
        }
        const double kappa2 = l1 * l2;
        if (!(kappa2 > 0.0)) continue;          // degenerate (e.g. the 999 sentinel): never accepted
        radius2 = fmax(radius2, fmax(n0, fmax(n1, n2)));
        edge2 = fmax(edge2, fmax(l1, fmax(l2, l3)));
        kappa2max = fmax(kappa2max, kappa2);
        sum_edge += sqrt(l1); sum_kappa += sqrt(kappa2); ++good;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], fminf(v[k], fminf(v[3 + k], v[6 + k])));
            hi[k] = fmaxf(hi[k], fmaxf(v[k], fmaxf(v[3 + k], v[6 + k])));
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        radius2 = fmax(radius2, __shfl_xor_sync(0xFFFFFFFFu, radius2, off));
        edge2 = fmax(edge2, __shfl_xor_sync(0xFFFFFFFFu, edge2, off));
        kappa2max = fmax(kappa2max, __shfl_xor_sync(0xFFFFFFFFu, kappa2max, off));
        sum_edge += __shfl_xor_sync(0xFFFFFFFFu, sum_edge, off);
        sum_kappa += __shfl_xor_sync(0xFFFFFFFFu, sum_kappa, off);
        good += __shfl_xor_sync(0xFFFFFFFFu, good, off);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], off));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], off));
        }
    }
    if ((threadIdx.x & 31) == 0 && good) {
        atomicMax(&out->radius2, enc_f64(radius2));
        atomicMax(&out->max_edge2, enc_f64(edge2));
        atomicMax(&out->kappa2_max, enc_f64(kappa2max));
        atomicAdd(&out->sum_edge, sum_edge);
        atomicAdd(&out->sum_kappa, sum_kappa);
        atomicAdd(&out->n_good, good);
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&out->lo[k], enc_f32(lo[k])); atomicMax(&out->hi[k], enc_f32(hi[k])); }
    }
}

// union over the non-degenerate triangles of (triangle box grown by ITS delta = scale * |e1||e2| (1 + 1e-9) + abs and 1e-6
// relative), in double like the host loop it replaces; the host rounds the six results outwards to float
__global__ void __launch_bounds__(kRepackBlock)
tight_box_kernel(const float* __restrict__ tris, uint32_t n_tris, double scale, double abs_, unsigned long long* __restrict__ out6) {
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for (uint32_t i = blockIdx.x * kRepackBlock + threadIdx.x; i < n_tris; i += gridDim.x * kRepackBlock) {
        const float* t = tris + (size_t)i * 9;
        float v[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] = __ldg(t + j);
        double l1 = 0.0, l2 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double e1 = (double)v[3 + k] - v[k], e2 = (double)v[6 + k] - v[k];
            l1 += e1 * e1; l2 += e2 * e2;
        }
        const double kp = sqrt(l1) * sqrt(l2);
        if (!(kp > 0.0)) continue;
        const double delta = scale * kp * (1.0 + 1e-9) + abs_;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double l = fmin((double)v[k], fmin((double)v[3 + k], (double)v[6 + k]));
            const double h = fmax((double)v[k], fmax((double)v[3 + k], (double)v[6 + k]));
            lo[k] = fmin(lo[k], l - delta - fabs(l) * 1e-6);
            hi[k] = fmax(hi[k], h + delta + fabs(h) * 1e-6);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], off));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], off));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(out6 + k, enc_f64(lo[k])); atomicMax(out6 + 3 + k, enc_f64(hi[k])); }
    }
}

// Small per-frame uploads and fills done by the SMs: `src` is page-locked host memory read over PCIe by the kernel itself (or
// null: store `fill`).  A cudaMemcpyAsync of the same bytes goes through a copy engine, where it queues behind whatever large
// device->host copy another stream has in progress -- with frames in flight that is the previous frame's band copy, ~30 us per
// upload (tools/ce_interference_probe.py) and the next frame's kernels wait behind it.
__global__ void __launch_bounds__(256) small_ops_kernel(SmallOps ops) {
    for (uint32_t k = 0; k < ops.n; ++k) {
        const SmallOp op = ops.op[k];
        uint32_t* dst = (uint32_t*)op.dst;
        const uint32_t* src = (const uint32_t*)op.src;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < op.words; i += gridDim.x * blockDim.x)
            dst[i] = src ? src[i] : op.fill;
    }
}

} // namespace

cudaError_t launch_small_ops(const SmallOps& ops, cudaStream_t s) {
    if (ops.n == 0) return cudaSuccess;
    uint32_t most = 0;
    for (uint32_t k = 0; k < ops.n; ++k) most = std::max(most, ops.op[k].words);
    int grid = (int)std::min<uint32_t>(std::max<uint32_t>((most + 1023u) / 1024u, 1u), 148u);
    small_ops_kernel<<<grid, 256, 0, s>>>(ops);
    return cudaGetLastError();
}

cudaError_t launch_read_bw(const void* buf, size_t bytes, uint32_t iters, int grid, unsigned long long* sink, cudaStream_t s) {
    read_bw_kernel<<<grid, 256, 0, s>>>((const uint4*)buf, bytes / 16, iters, sink);
    return cudaGetLastError();
}

cudaError_t launch_ray_bounds(const float* rays, unsigned long long n, unsigned int* out2, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out2, 0, 8, s);
    if (e != cudaSuccess || n == 0) return e;
    int grid = (int)std::min<unsigned long long>((n + kRepackBlock - 1) / kRepackBlock, 148ull * 8);
    ray_bounds_kernel<<<grid, kRepackBlock, 0, s>>>(rays, n, out2);
    return cudaGetLastError();
}

cudaError_t launch_model_stats(const float* tris_aos, uint32_t n_tris, ModelStatsDev* out, cudaStream_t s) {
    ModelStatsDev init;
    memset(&init, 0, sizeof init);
    for (int k = 0; k < 3; ++k) { init.lo[k] = 0xFFFFFFFFu; init.hi[k] = 0u; }
    cudaError_t e = cudaMemcpyAsync(out, &init, sizeof init, cudaMemcpyHostToDevice, s);        // pageable, 80 bytes: staged before return
    if (e != cudaSuccess || n_tris == 0) return e;
    int grid = (int)std::min<uint32_t>((n_tris + kRepackBlock - 1) / kRepackBlock, 148u * 4);
    model_stats_kernel<<<grid, kRepackBlock, 0, s>>>(tris_aos, n_tris, out);
    return cudaGetLastError();
}

cudaError_t launch_tight_box(const float* tris_aos, uint32_t n_tris, double scale, double abs_, unsigned long long* out6, cudaStream_t s) {
    unsigned long long init[6] = { ~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull };
    cudaError_t e = cudaMemcpyAsync(out6, init, sizeof init, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess || n_tris == 0) return e;
    int grid = (int)std::min<uint32_t>((n_tris + kRepackBlock - 1) / kRepackBlock, 148u * 4);
    tight_box_kernel<<<grid, kRepackBlock, 0, s>>>(tris_aos, n_tris, scale, abs_, out6);
    return cudaGetLastError();
}

cudaError_t launch_refit_sub_nodes(float4* raw, const uint32_t* parent, unsigned int* counters, const float* tris_aos,
                                   const uint32_t* order, uint32_t n_nodes, cudaStream_t s) {
    if (n_nodes == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(counters, 0, (size_t)n_nodes * 4, s);
    if (e != cudaSuccess) return e;
    int grid = (int)((n_nodes + kRepackBlock - 1) / kRepackBlock);
    refit_sub_nodes_kernel<<<grid, kRepackBlock, 0, s>>>(raw, parent, counters, tris_aos, order, n_nodes);
    return cudaGetLastError();
}

cudaError_t launch_bake_sub_nodes(const float4* raw, const uint32_t* parent, unsigned int* counters, const float* tris_aos,
                                  const uint32_t* order, uint32_t n_nodes, float scale, float abs_, float4* lohi_scratch, float4* out,
                                  cudaStream_t s) {
    if (n_nodes == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(counters, 0, (size_t)n_nodes * 4, s);
    if (e != cudaSuccess) return e;
    int grid = (int)((n_nodes + kRepackBlock - 1) / kRepackBlock);
    bake_sub_nodes_kernel<<<grid, kRepackBlock, 0, s>>>(raw, parent, counters, tris_aos, order, n_nodes, scale, abs_, lohi_scratch, out);
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
