// leaf_accel.cpp -- see leaf_accel.hpp.  Host-only C++ (runs once per upload / vertex update).
#include "leaf_accel.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace bvht {
namespace {

struct RefNode { float mn[3]; float mx[3]; uint32_t count; uint32_t left_first; };

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; ++k) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; } }
    void grow(const float* p) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    void grow(const Box& b) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    double area() const {
        double ex = (double)hi[0] - lo[0], ey = (double)hi[1] - lo[1], ez = (double)hi[2] - lo[2];
        if (ex < 0 || ey < 0 || ez < 0) return 0.0;
        return ex * ey + ey * ez + ez * ex;
    }
};

struct TriInfo { Box box; float centroid[3]; double kappa; };

struct Builder {
    const float* tris;
    const LeafAccelConfig& cfg;
    LeafAccelHost& out;
    std::vector<TriInfo> info;       // per reference primitive
    uint32_t depth_limit = 28;      // SAH splits up to here, balanced median splits below: depth <= 28 + log2(n)

    Builder(const float* t, const LeafAccelConfig& c, LeafAccelHost& o) : tris(t), cfg(c), out(o) {}

    static uint32_t leaf_ref(uint32_t first, uint32_t count) { return 0x80000000u | ((count - 1u) << 28) | first; }

    void range_box(const std::vector<uint32_t>& idx, uint32_t b, uint32_t e, Box& box, double& kappa) const {
        box.reset(); kappa = 0.0;
        for (uint32_t i = b; i < e; ++i) { box.grow(info[idx[i]].box); kappa = std::max(kappa, info[idx[i]].kappa); }
    }

    // Partition idx[b,e) into two non-empty halves; binned SAH (16 bins) with a median fallback.
    uint32_t split(std::vector<uint32_t>& idx, uint32_t b, uint32_t e, uint32_t depth) {
        const int NB = 16;
        Box cb; cb.reset();
        for (uint32_t i = b; i < e; ++i) cb.grow(info[idx[i]].centroid);
        int best_axis = -1; double best_cost = DBL_MAX; int best_bin = -1;
        if (depth < depth_limit) {
            for (int axis = 0; axis < 3; ++axis) {
                double lo = cb.lo[axis], hi = cb.hi[axis];
                if (!(hi > lo)) continue;
                Box bins[NB]; uint32_t cnt[NB];
                for (int i = 0; i < NB; ++i) { bins[i].reset(); cnt[i] = 0; }
                double scale = NB / (hi - lo);
                for (uint32_t i = b; i < e; ++i) {
                    int bi = (int)(((double)info[idx[i]].centroid[axis] - lo) * scale);
                    bi = std::max(0, std::min(NB - 1, bi));
                    bins[bi].grow(info[idx[i]].box); cnt[bi]++;
                }
                double la[NB - 1], ra[NB - 1]; uint32_t lc[NB - 1], rc[NB - 1];
                Box lb, rb; lb.reset(); rb.reset(); uint32_t ls = 0, rs = 0;
                for (int i = 0; i < NB - 1; ++i) {
                    ls += cnt[i]; lc[i] = ls; if (cnt[i]) lb.grow(bins[i]); la[i] = lb.area();
                    rs += cnt[NB - 1 - i]; rc[NB - 2 - i] = rs; if (cnt[NB - 1 - i]) rb.grow(bins[NB - 1 - i]); ra[NB - 2 - i] = rb.area();
                }
                for (int i = 0; i < NB - 1; ++i) {
                    if (lc[i] == 0 || rc[i] == 0) continue;
                    double cost = lc[i] * la[i] + rc[i] * ra[i];
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = i; }
                }
            }
        }
        if (best_axis >= 0) {
            double lo = cb.lo[best_axis], hi = cb.hi[best_axis];
            double scale = NB / (hi - lo);
            auto mid_it = std::partition(idx.begin() + b, idx.begin() + e, [&](uint32_t p) {
                int bi = (int)(((double)info[p].centroid[best_axis] - lo) * scale);
                bi = std::max(0, std::min(NB - 1, bi));
                return bi <= best_bin;
            });
            uint32_t mid = (uint32_t)(mid_it - idx.begin());
            if (mid > b && mid < e) return mid;
        }
        // median split along the widest centroid axis (always balanced: bounds the depth)
        int axis = 0; double ext = -1.0;
        for (int k = 0; k < 3; ++k) { double ex = (double)cb.hi[k] - cb.lo[k]; if (ex > ext) { ext = ex; axis = k; } }
        uint32_t mid = b + (e - b) / 2;
        std::nth_element(idx.begin() + b, idx.begin() + mid, idx.begin() + e, [&](uint32_t p, uint32_t q) {
            if (info[p].centroid[axis] != info[q].centroid[axis]) return info[p].centroid[axis] < info[q].centroid[axis];
            return p < q;
        });
        return mid;
    }

    // Builds the subtree over idx[b,e) (count > max_sub_leaf); returns its inner-node index.
    // `base` = sub position of idx[0] in the BLAS-wide order array.
    uint32_t build(std::vector<uint32_t>& idx, uint32_t b, uint32_t e, uint32_t base, uint32_t depth) {
        out.max_depth = std::max(out.max_depth, depth + 1);
        uint32_t node = (uint32_t)(out.sub_raw.size() / 16);
        out.sub_raw.resize(out.sub_raw.size() + 16, 0.0f);
        out.sub_parent.push_back(0xFFFFFFFFu);
        uint32_t mid = split(idx, b, e, depth);
        uint32_t refs[2];
        uint32_t rb[2] = { b, mid }, re[2] = { mid, e };
        for (int c = 0; c < 2; ++c) {
            uint32_t cnt = re[c] - rb[c];
            if (cnt <= cfg.max_sub_leaf) {
                std::sort(idx.begin() + rb[c], idx.begin() + re[c]);      // ascending reference index inside a sub leaf
                refs[c] = leaf_ref(base + rb[c], cnt);
            } else {
                refs[c] = build(idx, rb[c], re[c], base, depth + 1);
                out.sub_parent[refs[c]] = (node << 1) | (uint32_t)c;
            }
        }
        float rec[16];
        for (int c = 0; c < 2; ++c) {
            Box box; double kappa;
            range_box(idx, rb[c], re[c], box, kappa);
            float kf = (float)kappa; if ((double)kf < kappa) kf = std::nextafterf(kf, FLT_MAX);
            for (int k = 0; k < 3; ++k) { rec[8 * c + k] = box.lo[k]; rec[8 * c + 4 + k] = box.hi[k]; }
            rec[8 * c + 3] = kf;
            std::memcpy(&rec[8 * c + 7], &refs[c], 4);
        }
        std::memcpy(&out.sub_raw[(size_t)node * 16], rec, sizeof rec);
        return node;
    }
};

} // namespace

// One pass, on the frame path of deforming models (bvht_blas_update_vertices): the conservative quantities (radius, longest
// edge, largest |e1||e2|) are maxima, taken over SQUARES with one square root at the end (sqrt is monotonic: same value as the
// maximum of the roots); only the two means, which steer a heuristic, need a root per triangle, in single precision.
// kappa_out (optional): |e1||e2| of every triangle, rounded up, 0 for degenerate ones -- what bake_accel needs per triangle.
ModelStats compute_model_stats(const float* tris, uint32_t n_tris, std::vector<float>* kappa_out) {
    ModelStats m;
    double radius2 = 0.0, max_edge2 = 0.0, kmax2 = 0.0, sum_edge = 0.0, sum_kappa = 0.0;
    uint64_t n_good = 0;
    Box mb; mb.reset(); bool any = false;
    if (kappa_out) kappa_out->assign(n_tris, 0.0f);
    for (uint32_t i = 0; i < n_tris; ++i) {
        const float* t = tris + (size_t)i * 9;
        double l1 = 0.0, l2 = 0.0, l3 = 0.0, n0 = 0.0, n1 = 0.0, n2 = 0.0;
        for (int k = 0; k < 3; ++k) {
            const double a = t[k], b_ = t[3 + k], c = t[6 + k];
            const double e1 = b_ - a, e2 = c - a, e3 = c - b_;
            l1 += e1 * e1; l2 += e2 * e2; l3 += e3 * e3;
            n0 += a * a; n1 += b_ * b_; n2 += c * c;
        }
        const double kappa2 = l1 * l2;
        if (!(kappa2 > 0.0)) continue;          // degenerate (e.g. the 999 sentinel): area == 0 exactly, never accepted
        radius2 = std::max(radius2, std::max(n0, std::max(n1, n2)));
        max_edge2 = std::max(max_edge2, std::max(l1, std::max(l2, l3)));
        kmax2 = std::max(kmax2, kappa2);
        const float kf = std::sqrt((float)kappa2);
        sum_edge += (double)std::sqrt((float)l1); sum_kappa += (double)kf; ++n_good;
        if (kappa_out) (*kappa_out)[i] = kf * 1.000001f + 1e-37f;           // >= the exact product (float sqrt is within 1 ulp)
        mb.grow(t); mb.grow(t + 3); mb.grow(t + 6); any = true;
    }
    if (n_good) { m.mean_edge = sum_edge / (double)n_good; m.mean_kappa = sum_kappa / (double)n_good; }
    const double radius = std::sqrt(radius2);
    m.radius = radius > 0.0 ? radius : 1.0;
    m.max_edge = std::sqrt(max_edge2); m.model_kappa = std::sqrt(kmax2); m.model_valid = any;
    if (any) for (int k = 0; k < 3; ++k) { m.model_lo[k] = mb.lo[k]; m.model_hi[k] = mb.hi[k]; }
    return m;
}

bool build_leaf_accel(const float* tris, uint32_t n_tris, const void* nodes_v, uint32_t nodes_used,
                      const LeafAccelConfig& cfg, LeafAccelHost& out) {
    const RefNode* nodes = (const RefNode*)nodes_v;
    out = LeafAccelHost();
    out.leaf_sub_root.assign(nodes_used, 0xFFFFFFFFu);
    if (cfg.max_sub_leaf == 0 || cfg.max_sub_leaf > 8) return false;

    Builder bld(tris, cfg, out);
    bld.info.resize(n_tris);
    double radius = 0.0, max_edge = 0.0;
    for (uint32_t i = 0; i < n_tris; ++i) {
        const float* t = tris + (size_t)i * 9;
        TriInfo& ti = bld.info[i];
        ti.box.reset(); ti.box.grow(t); ti.box.grow(t + 3); ti.box.grow(t + 6);
        double e1[3], e2[3], e3[3];
        for (int k = 0; k < 3; ++k) {
            ti.centroid[k] = (float)(((double)t[k] + t[3 + k] + t[6 + k]) / 3.0);
            e1[k] = (double)t[3 + k] - t[k]; e2[k] = (double)t[6 + k] - t[k]; e3[k] = (double)t[6 + k] - t[3 + k];
        }
        double l1 = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
        double l2 = std::sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
        double l3 = std::sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
        ti.kappa = l1 * l2;
        if (ti.kappa > 0.0) {      // degenerate triangles (e.g. the 999 sentinel) have area == 0 exactly: never accepted
            for (int v = 0; v < 3; ++v) {
                double n = std::sqrt((double)t[3 * v] * t[3 * v] + (double)t[3 * v + 1] * t[3 * v + 1] + (double)t[3 * v + 2] * t[3 * v + 2]);
                radius = std::max(radius, n);
            }
            max_edge = std::max(max_edge, std::max(l1, std::max(l2, l3)));
        }
    }
    if (!(radius > 0.0)) radius = 1.0;
    out.radius = radius;
    out.max_edge = max_edge;
    {   // whole-model raw box over the non-degenerate triangles
        Box mb; mb.reset(); double kmax = 0.0; bool any = false;
        for (uint32_t i = 0; i < n_tris; ++i) {
            if (!(bld.info[i].kappa > 0.0)) continue;
            mb.grow(bld.info[i].box); kmax = std::max(kmax, bld.info[i].kappa); any = true;
        }
        if (any) {
            for (int k = 0; k < 3; ++k) { out.model_lo[k] = mb.lo[k]; out.model_hi[k] = mb.hi[k]; }
            out.model_kappa = kmax; out.model_valid = true;
        }
    }

    std::vector<uint32_t> idx;
    for (uint32_t ni = 0; ni < nodes_used; ++ni) {
        if (ni == 1) continue;
        const RefNode& n = nodes[ni];
        if (n.count == 0 || n.count < cfg.min_leaf_tris || n.count <= cfg.max_sub_leaf) continue;
        if ((uint64_t)n.left_first + n.count > n_tris) return false;
        idx.resize(n.count);
        for (uint32_t k = 0; k < n.count; ++k) idx[k] = n.left_first + k;
        uint32_t base = (uint32_t)out.order.size();
        if (base + n.count >= 0x0FFFFFFFu) return false;
        uint32_t root = bld.build(idx, 0, n.count, base, 0);
        out.leaf_sub_root[ni] = root;
        out.order.insert(out.order.end(), idx.begin(), idx.end());
    }
    return true;
}

} // namespace bvht
