// rt_bvh.cpp -- host BVH builder of libb200rt: the reference's median-split tree, emitted
// directly in the flattened 32-byte node layout the traversal kernels read.
#include "rt_bvh.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <map>

namespace b200rt {

void sphere_boxes(const float* cr, int64_t n, PrimBoxes& out) {
    out.lo.resize(3 * n); out.hi.resize(3 * n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) {       // Sphere::update_bbox, cpp_raytracer/raytracer_core.h:187-190
            out.lo[3 * i + c] = cr[4 * i + c] - cr[4 * i + 3];
            out.hi[3 * i + c] = cr[4 * i + c] + cr[4 * i + 3];
        }
}

void triangle_boxes(const float* v, int64_t n, PrimBoxes& out) {
    out.lo.resize(3 * n); out.hi.resize(3 * n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) {
            float a = v[9 * i + c], b = v[9 * i + 3 + c], d = v[9 * i + 6 + c];
            out.lo[3 * i + c] = std::fmin(a, std::fmin(b, d));
            out.hi[3 * i + c] = std::fmax(a, std::fmax(b, d));
        }
}

namespace {

struct Builder {
    const PrimBoxes& bx;
    std::vector<float> ctr;             // box centres, n x 3
    std::vector<int32_t>& idx;
    std::vector<rt_bvh_node>& nodes;
    std::map<int64_t, int64_t> memo;    // span -> node records used by the subtree's descendants
    int64_t leaf = 4;                   // a range of at most this many primitives is a leaf (4: the reference's)

    // Number of node records (sibling pairs x 2) below a node spanning `span` primitives.  The
    // tree shape depends on span only (always split at span/2), which is what lets subtrees be
    // built in parallel at fixed, deterministic positions.
    int64_t records_below(int64_t span) {
        if (span <= leaf) return 0;
        auto it = memo.find(span);
        if (it != memo.end()) return it->second;
        int64_t l = span / 2;
        int64_t r = 2 + records_below(l) + records_below(span - l);
        memo[span] = r;
        return r;
    }
    int64_t records_below_ro(int64_t span) const {
        if (span <= leaf) return 0;
        return memo.at(span);
    }

    void build(int64_t at, int64_t start, int64_t end, int64_t pair_base) {
        rt_bvh_node& nd = nodes[at];
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int64_t k = start; k < end; ++k) {
            const float* l = &bx.lo[3 * (int64_t)idx[k]];
            const float* h = &bx.hi[3 * (int64_t)idx[k]];
            for (int c = 0; c < 3; ++c) { if (l[c] < lo[c]) lo[c] = l[c]; if (h[c] > hi[c]) hi[c] = h[c]; }
        }
        for (int c = 0; c < 3; ++c) { nd.bmin[c] = lo[c]; nd.bmax[c] = hi[c]; }
        int64_t span = end - start;
        if (span <= leaf) {
            std::sort(idx.begin() + start, idx.begin() + end);
            nd.a = (int32_t)start; nd.b = (int32_t)span;
            return;
        }
        float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
        int axis = 0;
        if (ey > ex) axis = 1;
        if (ez > ey && ez > ex) axis = 2;
        int64_t mid = start + span / 2;
        const float* key = ctr.data();
        // The reference sorts the whole range; only the two halves' membership matters for the
        // tree, so a selection with the same strict total order (centre, then number) suffices.
        std::nth_element(idx.begin() + start, idx.begin() + mid, idx.begin() + end, [key, axis](int32_t a, int32_t b) {
            float ka = key[3 * (int64_t)a + axis], kb = key[3 * (int64_t)b + axis];
            return ka < kb || (ka == kb && a < b);
        });
        nd.a = (int32_t)pair_base; nd.b = 0;
        int64_t left_base = pair_base + 2;
        int64_t right_base = left_base + records_below_ro(mid - start);
        if (span > 8192) {
#pragma omp task firstprivate(pair_base, start, mid, left_base)
            build(pair_base, start, mid, left_base);
#pragma omp task firstprivate(pair_base, mid, end, right_base)
            build(pair_base + 1, mid, end, right_base);
#pragma omp taskwait
        } else {
            build(pair_base, start, mid, left_base);
            build(pair_base + 1, mid, end, right_base);
        }
    }
};

}  // namespace

void build_median_split(const PrimBoxes& boxes, int64_t n, std::vector<rt_bvh_node>& nodes,
                        std::vector<int32_t>& prim_index, int leaf_size) {
    nodes.clear(); prim_index.clear();
    if (n == 0) return;
    prim_index.resize(n);
    Builder b{boxes, {}, prim_index, nodes, {}};
    b.leaf = leaf_size < 1 ? 1 : (leaf_size > 4 ? 4 : leaf_size);
    b.ctr.resize(3 * n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        prim_index[i] = (int32_t)i;
        for (int c = 0; c < 3; ++c) b.ctr[3 * i + c] = (boxes.lo[3 * i + c] + boxes.hi[3 * i + c]) * 0.5f;
    }
    int64_t total = 2 + b.records_below(n);
    nodes.assign(total, rt_bvh_node{{0, 0, 0}, 0, {0, 0, 0}, 0});
#pragma omp parallel
#pragma omp single
    b.build(0, 0, n, 2);
    float scale = 0.0f;
    for (int c = 0; c < 3; ++c) {
        scale = std::fmax(scale, std::fabs(nodes[0].bmin[c]));
        scale = std::fmax(scale, std::fabs(nodes[0].bmax[c]));
    }
    float pad = scale * 0x1p-16f;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < total; ++k) {
        if (k == 1) continue;
        for (int c = 0; c < 3; ++c) { nodes[k].bmin[c] -= pad; nodes[k].bmax[c] += pad; }
    }
}

// ------------------------------------------------------------------------------- binned SAH (builder 2)
// Not the reference's tree: the same node layout and closest hits (they do not depend on the tree), fewer nodes entered and
// fewer primitives tested per ray.  Top-down; a range is cut where the surface-area cost area(L) * |L| + area(R) * |R| over
// 16 centroid bins per axis is smallest; ranges of <= leaf_size primitives are cut further only while that beats testing them
// all in one leaf (cost of a traversal step = one primitive test).  Subtrees are built as independent tasks into their own
// record vectors and concatenated left before right, so the layout is deterministic for any thread count.
namespace {

struct SahBuilder {
    const PrimBoxes& bx;
    std::vector<float> ctr;
    std::vector<int32_t>& idx;
    int64_t leaf = 4;
    float trav_cost = 1.0f;             // cost of one traversal step in primitive tests (when to cut a small range further)
    static constexpr int kBins = 16;

    static float half_area(const float* lo, const float* hi) {
        const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
        return ex * ey + ey * ez + ez * ex;
    }
    struct Box {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        void add(const float* l, const float* h) { for (int c = 0; c < 3; ++c) { if (l[c] < lo[c]) lo[c] = l[c]; if (h[c] > hi[c]) hi[c] = h[c]; } }
        void add(const Box& o) { add(o.lo, o.hi); }
    };

    // node = this subtree's root record (child index relative to `below`), below = its descendants' records (pairs)
    void build(int64_t start, int64_t end, rt_bvh_node& node, std::vector<rt_bvh_node>& below) {
        Box box, cb;
        for (int64_t k = start; k < end; ++k) {
            const int64_t p = idx[k];
            box.add(&bx.lo[3 * p], &bx.hi[3 * p]);
            cb.add(&ctr[3 * p], &ctr[3 * p]);
        }
        for (int c = 0; c < 3; ++c) { node.bmin[c] = box.lo[c]; node.bmax[c] = box.hi[c]; }
        const int64_t span = end - start;
        auto make_leaf = [&]() {
            std::sort(idx.begin() + start, idx.begin() + end);
            node.a = (int32_t)start; node.b = (int32_t)span;
        };
        if (span == 1) { make_leaf(); return; }
        // best binned split
        float best = INFINITY;
        int best_axis = -1, best_bin = 0;
        for (int axis = 0; axis < 3; ++axis) {
            const float c0 = cb.lo[axis], c1 = cb.hi[axis];
            if (!(c1 > c0)) continue;
            const float scale = (float)kBins / (c1 - c0);
            Box bins[kBins];
            int64_t cnt[kBins] = {0};
            for (int64_t k = start; k < end; ++k) {
                const int64_t p = idx[k];
                int b = (int)((ctr[3 * p + axis] - c0) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                bins[b].add(&bx.lo[3 * p], &bx.hi[3 * p]);
                cnt[b] += 1;
            }
            float right_cost[kBins];
            Box acc;
            int64_t n = 0;
            for (int b = kBins - 1; b >= 1; --b) {
                if (cnt[b]) acc.add(bins[b]);
                n += cnt[b];
                right_cost[b] = n ? half_area(acc.lo, acc.hi) * (float)n : INFINITY;
            }
            Box accl;
            n = 0;
            for (int b = 1; b < kBins; ++b) {                      // split before bin b
                if (cnt[b - 1]) accl.add(bins[b - 1]);
                n += cnt[b - 1];
                if (n == 0 || n == span) continue;
                const float cost = half_area(accl.lo, accl.hi) * (float)n + right_cost[b];
                if (cost < best) { best = cost; best_axis = axis; best_bin = b; }
            }
        }
        const float node_area = half_area(box.lo, box.hi);
        if (span <= leaf && (best_axis < 0 || !(best + trav_cost * node_area < node_area * (float)span))) { make_leaf(); return; }
        int64_t mid;
        if (best_axis < 0) {                                       // all centroids equal: halve by number
            std::sort(idx.begin() + start, idx.begin() + end);
            mid = start + span / 2;
        } else {
            const float c0 = cb.lo[best_axis], scale = (float)kBins / (cb.hi[best_axis] - c0);
            const float* key = ctr.data();
            const int axis = best_axis, bin = best_bin;
            auto it = std::partition(idx.begin() + start, idx.begin() + end, [key, axis, c0, scale, bin](int32_t p) {
                int b = (int)((key[3 * (int64_t)p + axis] - c0) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                return b < bin;
            });
            mid = it - idx.begin();
        }
        rt_bvh_node l{{0, 0, 0}, 0, {0, 0, 0}, 0}, r{{0, 0, 0}, 0, {0, 0, 0}, 0};
        std::vector<rt_bvh_node> lb, rb;
        if (span > 16384) {
#pragma omp task shared(l, lb) firstprivate(start, mid)
            build(start, mid, l, lb);
#pragma omp task shared(r, rb) firstprivate(mid, end)
            build(mid, end, r, rb);
#pragma omp taskwait
        } else {
            build(start, mid, l, lb);
            build(mid, end, r, rb);
        }
        // below = [l, r] + lb + rb; child indices are relative to the start of `below`
        const int32_t off_l = 2, off_r = 2 + (int32_t)lb.size();
        below.reserve(2 + lb.size() + rb.size());
        if (l.b == 0) l.a += off_l;
        if (r.b == 0) r.a += off_r;
        below.push_back(l); below.push_back(r);
        for (auto& nd : lb) { if (nd.b == 0) nd.a += off_l; below.push_back(nd); }
        for (auto& nd : rb) { if (nd.b == 0) nd.a += off_r; below.push_back(nd); }
        node.a = 0; node.b = 0;
    }
};

}  // namespace

void build_sah(const PrimBoxes& boxes, int64_t n, std::vector<rt_bvh_node>& nodes, std::vector<int32_t>& prim_index, int leaf_size, float trav_cost) {
    nodes.clear(); prim_index.clear();
    if (n == 0) return;
    prim_index.resize(n);
    SahBuilder b{boxes, {}, prim_index};
    b.leaf = leaf_size < 1 ? 1 : (leaf_size > 4 ? 4 : leaf_size);
    b.trav_cost = trav_cost;
    b.ctr.resize(3 * n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        prim_index[i] = (int32_t)i;
        for (int c = 0; c < 3; ++c) b.ctr[3 * i + c] = (boxes.lo[3 * i + c] + boxes.hi[3 * i + c]) * 0.5f;
    }
    rt_bvh_node root{{0, 0, 0}, 0, {0, 0, 0}, 0};
    std::vector<rt_bvh_node> below;
#pragma omp parallel
#pragma omp single
    b.build(0, n, root, below);
    nodes.assign(2 + below.size(), rt_bvh_node{{0, 0, 0}, 0, {0, 0, 0}, 0});
    if (root.b == 0) root.a += 2;                                  // records 0 (root), 1 (pad), then `below`
    nodes[0] = root;
    const int64_t total = (int64_t)nodes.size();
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)below.size(); ++k) {
        rt_bvh_node nd = below[k];
        if (nd.b == 0) nd.a += 2;
        nodes[2 + k] = nd;
    }
    float scale = 0.0f;
    for (int c = 0; c < 3; ++c) {
        scale = std::fmax(scale, std::fabs(nodes[0].bmin[c]));
        scale = std::fmax(scale, std::fabs(nodes[0].bmax[c]));
    }
    const float pad = scale * 0x1p-16f;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < total; ++k) {
        if (k == 1) continue;
        for (int c = 0; c < 3; ++c) { nodes[k].bmin[c] -= pad; nodes[k].bmax[c] += pad; }
    }
}

int validate_bvh(const rt_bvh_node* nodes, int64_t n_nodes, int64_t n_prims, const char** msg) {
    static const char* ok = "";
    *msg = ok;
    if (n_nodes == 0) return 0;
    if (n_nodes < 2) { *msg = "bvh: need at least root + pad record"; return -1; }
    std::vector<std::pair<int64_t, int>> stack;
    stack.emplace_back(0, 1);
    int max_depth = 0;
    int64_t visited = 0, covered = 0;
    while (!stack.empty()) {
        auto [k, depth] = stack.back();
        stack.pop_back();
        if (++visited > n_nodes) { *msg = "bvh: cycle or shared node"; return -1; }
        if (depth > max_depth) max_depth = depth;
        const rt_bvh_node& nd = nodes[k];
        if (nd.b < 0 || nd.b > 7) { *msg = "bvh: leaf count out of range (1..7)"; return -1; }
        if (nd.b > 0) {
            if (nd.a < 0 || (int64_t)nd.a + nd.b > n_prims) { *msg = "bvh: leaf slot range out of bounds"; return -1; }
            covered += nd.b;
        } else {
            if (nd.a < 2 || (nd.a & 1) || (int64_t)nd.a + 1 >= n_nodes) { *msg = "bvh: child pair index invalid"; return -1; }
            stack.emplace_back(nd.a, depth + 1);
            stack.emplace_back(nd.a + 1, depth + 1);
        }
    }
    if (covered != n_prims) { *msg = "bvh: leaves do not cover every primitive exactly once"; return -1; }
    return max_depth;
}

}  // namespace b200rt
