// bvh.cpp -- the reference's BVH construction, then flattening into the device layout.
//
// Construction follows BVH::construct_tree (reference Code/acceleration.cpp:20-64): node box =
// union of the shapes' boxes; <= 4 shapes -> leaf; otherwise sort the range by box centre along
// the node box's longest axis (AABB::get_longest_axis, shapes.cpp:46-53) with std::sort and split
// at the median index. The reference recomputes every shape's box inside the comparator; we
// cache boxes and centres once, which yields the same comparator outcomes and therefore (same
// libstdc++ std::sort on the same initial sequence) the same permutation and the same tree.
//
// Why the exact tree matters: the reference's traversal (acceleration.cpp:67-117) tests a shape
// iff the ray passes the box test of its leaf (ancestor boxes contain the leaf box and IEEE
// rounding is monotonic, so they pass too). Which <=4 shapes share a leaf box therefore decides
// which shapes are tested at all, and hit IDs are only bit-exact with identical leaves.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstring>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <stdexcept>
#include <thread>

#include "scene.hpp"

namespace rtb {
namespace {

unsigned build_threads() {
    static const unsigned n = [] {
        const char* e = std::getenv("RT_B200_HOST_THREADS");
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        return e ? (unsigned)std::max(1, std::atoi(e)) : std::min(hw, 64u);
    }();
    return n;
}

// fn(begin, end) over [0, n) in contiguous chunks, one per thread.
template <class Fn>
void parallel_chunks(size_t n, size_t min_chunk, Fn fn) {
    const unsigned threads = (unsigned)std::max<size_t>(1, std::min<size_t>(build_threads(), n / std::max<size_t>(1, min_chunk)));
    if (threads <= 1) { fn((size_t)0, n); return; }
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < threads; ++t) pool.emplace_back([=] { fn(n * t / threads, n * (t + 1) / threads); });
    fn((size_t)0, n / threads);
    for (std::thread& th : pool) th.join();
}

struct BuildCtx {
    std::vector<Box> box;          // the shapes' boxes, compact (24 B each: the builder touches them in sorted order)
    std::vector<int>& order;
    std::vector<float> centre[3];  // (lo+hi)/2 per axis, acceleration.cpp:49-50
};

// union of the boxes of order[start, end): min / max are exact, so any grouping gives the reference's union
Box range_box(const BuildCtx& c, int start, int end) {
    Box b;
    for (int i = 0; i < 3; ++i) { b.lo[i] = FLT_MAX; b.hi[i] = -FLT_MAX; }
    auto scan = [&](size_t a0, size_t a1, Box& out) {
        for (size_t k = (size_t)start + a0; k < (size_t)start + a1; ++k) {
            const Box& pb = c.box[(size_t)c.order[k]];
            for (int i = 0; i < 3; ++i) { out.lo[i] = std::min(out.lo[i], pb.lo[i]); out.hi[i] = std::max(out.hi[i], pb.hi[i]); }
        }
    };
    if (end - start >= (1 << 17)) {
        std::mutex mu;
        parallel_chunks((size_t)(end - start), 1 << 15, [&](size_t a0, size_t a1) {
            Box part;
            for (int i = 0; i < 3; ++i) { part.lo[i] = FLT_MAX; part.hi[i] = -FLT_MAX; }
            scan(a0, a1, part);
            std::lock_guard<std::mutex> lock(mu);
            for (int i = 0; i < 3; ++i) { b.lo[i] = std::min(b.lo[i], part.lo[i]); b.hi[i] = std::max(b.hi[i], part.hi[i]); }
        });
    } else {
        scan(0, (size_t)(end - start), b);
    }
    return b;
}

int longest_axis(const Box& b) {  // shapes.cpp:46-53
    const float x = b.hi[0] - b.lo[0], y = b.hi[1] - b.lo[1], z = b.hi[2] - b.lo[2];
    if (x > y && x > z) return 0;
    if (y > z) return 1;
    return 2;
}

// Pass 1: only permutes `order` (disjoint sub-ranges are independent, so sub-trees can run on
// other threads without changing the result).
// The reference sorts shape POINTERS with a comparator that looks the centres up; we sort (centre, index)
// pairs with the same comparison on the centre. std::sort's moves are a function of the comparator's
// outcomes and of the positions only -- never of the element type -- so the indices end up in the
// permutation the reference's sort produces (equal centres included), while the sort itself streams
// through memory instead of chasing an index per comparison.
struct Keyed { float key; int idx; };

void sort_ranges(BuildCtx& c, int start, int end, int par_depth) {
    if (end - start <= 4) return;
    const Box b = range_box(c, start, end);
    const int axis = longest_axis(b);
    const std::vector<float>& ctr = c.centre[axis];
    if (end - start >= 64) {
        std::vector<Keyed> tmp((size_t)(end - start));
        for (int k = start; k < end; ++k) tmp[(size_t)(k - start)] = {ctr[c.order[k]], c.order[k]};
        std::sort(tmp.begin(), tmp.end(), [](const Keyed& a, const Keyed& b2) { return a.key < b2.key; });
        for (int k = start; k < end; ++k) c.order[k] = tmp[(size_t)(k - start)].idx;
    } else {
        std::sort(c.order.begin() + start, c.order.begin() + end, [&ctr](int a, int b2) { return ctr[a] < ctr[b2]; });
    }
    const int mid = (start + end) / 2;
    if (par_depth > 0 && end - start > (1 << 13)) {
        auto fut = std::async(std::launch::async, [&c, start, mid, par_depth] { sort_ranges(c, start, mid, par_depth - 1); });
        sort_ranges(c, mid, end, par_depth - 1);
        fut.get();
    } else {
        sort_ranges(c, start, mid, 0);
        sort_ranges(c, mid, end, 0);
    }
}

// Pass 2: pre-order node array; boxes bottom-up (min/max are exact, so the union of the two
// children equals the reference's union over the whole range). The split is always at floor(m / 2), so the
// number of nodes below a range depends on its size only: sub-trees know their place in the array and
// are written on different threads.
struct NodeCount {
    std::vector<std::pair<int, int>> memo;  // (range size, nodes): two sizes per level
    int operator()(int m) {
        if (m <= 4) return 1;
        for (const auto& e : memo) if (e.first == m) return e.second;
        const int v = 1 + (*this)(m / 2) + (*this)(m - m / 2);
        memo.emplace_back(m, v);
        return v;
    }
};

void emit_nodes(HostScene& s, const std::vector<Box>& box, NodeCount& count, int start, int end, int me, int par_depth) {
    TreeNode& n = s.tree[(size_t)me];
    n.first = start;
    n.count = end - start;
    if (end - start <= 4) {
        n.left = n.right = -1;
        for (int i = 0; i < 3; ++i) { n.box.lo[i] = FLT_MAX; n.box.hi[i] = -FLT_MAX; }
        for (int k = start; k < end; ++k) {
            const Box& pb = box[(size_t)s.order[(size_t)k]];
            for (int i = 0; i < 3; ++i) { n.box.lo[i] = std::min(n.box.lo[i], pb.lo[i]); n.box.hi[i] = std::max(n.box.hi[i], pb.hi[i]); }
        }
        return;
    }
    const int mid = (start + end) / 2;
    const int l = me + 1, r = me + 1 + count(mid - start);
    if (par_depth > 0 && end - start > (1 << 14)) {
        NodeCount other = count;  // the memo is not shared between threads
        auto fut = std::async(std::launch::async, [&s, &box, other, start, mid, l, par_depth]() mutable { emit_nodes(s, box, other, start, mid, l, par_depth - 1); });
        emit_nodes(s, box, count, mid, end, r, par_depth - 1);
        fut.get();
    } else {
        emit_nodes(s, box, count, start, mid, l, 0);
        emit_nodes(s, box, count, mid, end, r, 0);
    }
    n.left = l;
    n.right = r;
    for (int i = 0; i < 3; ++i) {
        n.box.lo[i] = std::min(s.tree[(size_t)l].box.lo[i], s.tree[(size_t)r].box.lo[i]);
        n.box.hi[i] = std::max(s.tree[(size_t)l].box.hi[i], s.tree[(size_t)r].box.hi[i]);
    }
}

inline float bits_f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

}  // namespace

void build_bvh(HostScene& s) {
    const auto t0 = std::chrono::steady_clock::now();
    const int n = (int)s.prims.size();
    if (n >= (1 << 28)) throw std::runtime_error("too many shapes for the 28-bit leaf encoding");
    s.order.resize(n);
    for (int i = 0; i < n; ++i) s.order[i] = i;
    s.tree.clear();
    s.n_leaves = 0;
    if (n == 0) return;
    BuildCtx c{{}, s.order, {}};
    c.box.resize((size_t)n);
    for (int a = 0; a < 3; ++a) c.centre[a].resize((size_t)n);
    parallel_chunks((size_t)n, 1 << 15, [&](size_t i0, size_t i1) {
        for (size_t i = i0; i < i1; ++i) {
            c.box[i] = s.prims[i].box;
            for (int a = 0; a < 3; ++a) c.centre[a][i] = (s.prims[i].box.lo[a] + s.prims[i].box.hi[a]) / 2.0f;
        }
    });
    const auto t1 = std::chrono::steady_clock::now();
    sort_ranges(c, 0, n, 6);  // up to 64 sub-trees in flight
    const auto t2 = std::chrono::steady_clock::now();
    NodeCount count;
    s.tree.assign((size_t)count(n), TreeNode{});
    emit_nodes(s, c.box, count, 0, n, 0, 6);
    for (const TreeNode& t : s.tree) if (t.left < 0) s.n_leaves++;
    const auto t3 = std::chrono::steady_clock::now();
    s.build_seconds = std::chrono::duration<double>(t3 - t0).count();
    if (std::getenv("RT_B200_DEBUG"))
        std::fprintf(stderr, "[rt_b200] reference BVH: centres %.3f s, sorts %.3f s, nodes %.3f s\n", std::chrono::duration<double>(t1 - t0).count(),
                     std::chrono::duration<double>(t2 - t1).count(), std::chrono::duration<double>(t3 - t2).count());
}

namespace {
// ---------------------------------------------------------------------------------------------
// Culling boxes. A primitive's culling box is only used to SKIP its intersection test, so it must
// contain every ray the reference's test could report as a hit, including hits that exist only
// through rounding. Bounds (u = 2^-24, D = distance ray origin -> shape):
//   * every shape: the reported point lies within ~16u (|o| + |c| + D) of both the ray and the
//     shape (transform and slab rounding; o = ray origin, c = shape position). Ray origins are the
//     camera / lens or points on shapes, so |o|, |c| and D are bounded by the scene: the pad is
//     8e-6 G (= 134u G) with G = the largest |x|+|y|+|z| over all boxes and the camera;
//   * sphere: the discriminant b^2 - 4ac (shapes.cpp:220-225) cancels catastrophically for distant
//     origins: a ray is accepted up to 7.5u |o_obj|^2 (object units) outside the unit sphere,
//     i.e. up to Q D^2 in world units with Q = 7.5u s_max / s_min^2 -- a per-primitive
//     coefficient the traversal multiplies by its own distance bound (stored doubled, and with a
//     1.5x margin, because the traversal's bound is (|tn|+|tf|)^2 <= 2 D'^2 + 2 diag^2);
//   * plane: the inside tests tolerate cross(e, P - v) . n >= -1e-6 (shapes.cpp:31-38), i.e. a
//     point up to 1e-6 (|e_a|+|e_b|) / (2 area) beyond a vertex; degenerate triangles (repeated
//     corner) accept a whole line, and corner 3 may lie off the plane of corners 0..2 -- such
//     quads get an unbounded culling box (the test is simply never skipped).
// `pad` is the static outward push in world units; returns false when the box must be unbounded.
// ---------------------------------------------------------------------------------------------
bool cull_pad(const HostPrim& p, double scene_g, double& pad, float& q) {
    const double u = 5.9604644775390625e-8;
    q = 0.0f;
    double ext = 0.0, mag = 0.0;
    for (int a = 0; a < 3; ++a) {
        ext = std::max(ext, (double)p.box.hi[a] - (double)p.box.lo[a]);
        mag = std::max(mag, std::max(std::fabs((double)p.box.lo[a]), std::fabs((double)p.box.hi[a])));
    }
    if (!(ext < 1e30) || !(mag < 1e30)) return false;
    pad = 1e-5 * (ext + mag) + 1e-6 + 8e-6 * scene_g;
    if (p.type == RT_SPHERE) {
        double smin = 1e300, smax = 0.0;
        for (int j = 0; j < 3; ++j) {  // |scale_j| = length of column j of the rotation*scale block
            const double n = std::sqrt((double)p.o2w[0][j] * p.o2w[0][j] + (double)p.o2w[1][j] * p.o2w[1][j] +
                                       (double)p.o2w[2][j] * p.o2w[2][j]);
            smin = std::min(smin, n);
            smax = std::max(smax, n);
        }
        if (!(smin > 1e-12) || !(smax < 1e12)) return false;
        const double Q = 7.5 * u * smax / (smin * smin);
        const double qq = 2.0 * 1.5 * Q;
        if (!(qq < 1e10)) return false;
        q = (float)qq;
        if (!((double)q >= qq)) q = std::nextafter(q, FLT_MAX);
        const double v2 = (double)p.velocity[0] * p.velocity[0] + (double)p.velocity[1] * p.velocity[1] +
                          (double)p.velocity[2] * p.velocity[2];
        pad += 2.0 * 1.5 * Q * (12.0 * smax * smax + 2.0 * v2);  // the diag^2 share of the bound
    } else if (p.type == RT_PLANE) {
        if (!p.normal_valid) { pad = 0.0; return true; }  // never hit (shapes.cpp:449)
        const double n[3] = {p.normal[0], p.normal[1], p.normal[2]};
        auto sub = [](const float* a, const float* b, double* o) { for (int i = 0; i < 3; ++i) o[i] = (double)a[i] - (double)b[i]; };
        auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
        auto len = [&](const double* a) { return std::sqrt(dot(a, a)); };
        double h3[3];
        sub(p.corners[3], p.corners[0], h3);
        const double off = std::fabs(dot(h3, n));  // corner 3 off the plane of corners 0..2
        const int tri[2][3] = {{1, 3, 2}, {0, 1, 2}};  // isPointInQuad, shapes.cpp:491-492
        double worst = 0.0;
        for (const auto& t : tri) {
            double e1[3], e2[3], e3[3];
            sub(p.corners[t[1]], p.corners[t[0]], e1);
            sub(p.corners[t[2]], p.corners[t[1]], e2);
            sub(p.corners[t[0]], p.corners[t[2]], e3);
            const double cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            const double area2 = std::fabs(dot(cr, n));  // twice the projected area
            const double l1 = len(e1), l2 = len(e2), l3 = len(e3);
            if (!(area2 > 1e-9 * (l1 + l2 + l3) * (l1 + l2 + l3)) || !(std::min(l1, std::min(l2, l3)) > 1e-7)) return false;
            worst = std::max(worst, 2e-6 * (l1 + l2 + l3) / area2);
        }
        pad += off * 1.001 + worst;
        if (!(pad < 1e-2 * ext + 1e-3)) return false;  // tolerance comparable to the quad itself: do not cull
    }
    return true;
}

DWide empty_wide() {
    DWide w;
    for (float& x : w.f) x = 0.0f;
    return w;
}

void set_child_box(DWide& w, int c, const Box& b) {
    w.f[0 + c] = b.lo[0];  w.f[4 + c] = b.hi[0];
    w.f[8 + c] = b.lo[1];  w.f[12 + c] = b.hi[1];
    w.f[16 + c] = b.lo[2]; w.f[20 + c] = b.hi[2];
}

// ---------------------------------------------------------------------------------------------
// The device tree: a 4-wide BVH over the PRIMITIVES, built for traversal cost.
//
// What the reference prescribes is only this: a shape is tested iff the EXACT box of
// its reference leaf passes AABB::intersect, and the answer is (min t, first shape in shape_list order among
// equal t). That condition is carried per primitive -- the "gate": HostScene::dleafbox holds the leaf box of
// every primitive, and the traversal evaluates the reference's test on it before it runs the primitive's
// intersection routine. Everything else is ours: the boxes of this tree are CULLING boxes (cull_pad: they
// contain every ray for which the primitive's routine could report a hit) and their unions, used only to skip
// work, so the tree may group primitives in any way:
//   * binary tree by binned surface-area heuristic (16 bins per axis on the culling-box centroids, cost =
//     SA_left * n_left + SA_right * n_right) down to single primitives,
//   * collapsed to <= 4 children per node by repeatedly opening the inner child with the largest area; a node's
//     children are inner nodes (slots 0..ni-1, consecutive node indices first + slot) and primitives (the
//     following slots, sorted positions in f[27..30]),
//   * nodes numbered breadth-first.
// Compared with a tree whose leaves are the reference's leaves (round-2 first version, tag r2-sah-leaves) a ray
// visits about a third fewer nodes: the level of "leaf nodes" (4 culling boxes behind every leaf box) is gone.
// ---------------------------------------------------------------------------------------------
struct UpNode {
    Box box;                    // union of the ANCESTOR boxes (see flatten_scene) of the primitives below
    int left = -1, right = -1;  // UpNode indices; -1 for a leaf
    int prim = -1;              // sorted position of the primitive (leaves only)
};

inline double half_area(const Box& b) {
    const double x = (double)b.hi[0] - b.lo[0], y = (double)b.hi[1] - b.lo[1], z = (double)b.hi[2] - b.lo[2];
    return x * y + y * z + z * x;
}
inline void box_reset(Box& b) { for (int i = 0; i < 3; ++i) { b.lo[i] = FLT_MAX; b.hi[i] = -FLT_MAX; } }
inline void box_merge(Box& b, const Box& o) {
    for (int i = 0; i < 3; ++i) { b.lo[i] = std::min(b.lo[i], o.lo[i]); b.hi[i] = std::max(b.hi[i], o.hi[i]); }
}

struct SahBuilder {
    struct Item { Box box; int prim; };  // a primitive's ancestor box travels with it: every pass streams through memory
    double min_frac;                // a split must leave at least this share of the primitives on each side
    std::vector<Item> items;        // (ancestor box, sorted position), permuted in place
    std::vector<UpNode> nodes;      // pre-sized to 2n - 1: sub-trees built on other threads claim slots with `next`
    std::atomic<int> next{0};

    int alloc() { return next.fetch_add(1, std::memory_order_relaxed); }
    void set(int id, const UpNode& n) { nodes[(size_t)id] = n; }

    static constexpr int NB = 16;
    struct Bins {
        Box node, cb;            // box of the items, bounds of their centroids
        Box bins[3][NB];
        int cnt[3][NB];
    };
    static int bin_of(const Box& b, int a, float lo, float scale) {
        return std::min(NB - 1, std::max(0, (int)((0.5f * (b.lo[a] + b.hi[a]) - lo) * scale)));
    }
    // Big ranges (the top of the tree) are scanned by all threads; the outcome does not depend on how the range is
    // cut (min / max and counts), and the partition below keeps the items' relative order.
    static constexpr int PAR_RANGE = 1 << 17;

    // builds the sub-tree over items[lo, hi) and returns its node id
    int build(int lo, int hi, int par_depth) {
        const int me = alloc();
        UpNode n;
        const bool wide_scan = hi - lo >= PAR_RANGE;
        Bins acc0;
        box_reset(acc0.node);
        box_reset(acc0.cb);
        {   // pass 1: node box and centroid bounds
            std::mutex mu;
            auto scan = [&](size_t a, size_t b) {
                Box nb, cb;
                box_reset(nb);
                box_reset(cb);
                for (size_t k = (size_t)lo + a; k < (size_t)lo + b; ++k) {
                    const Box& bx = items[k].box;
                    box_merge(nb, bx);
                    for (int ax = 0; ax < 3; ++ax) {
                        const float c = 0.5f * (bx.lo[ax] + bx.hi[ax]);
                        cb.lo[ax] = std::min(cb.lo[ax], c);
                        cb.hi[ax] = std::max(cb.hi[ax], c);
                    }
                }
                std::lock_guard<std::mutex> lock(mu);
                box_merge(acc0.node, nb);
                box_merge(acc0.cb, cb);
            };
            if (wide_scan) parallel_chunks((size_t)(hi - lo), 1 << 15, scan);
            else scan(0, (size_t)(hi - lo));
        }
        n.box = acc0.node;
        if (hi - lo == 1) {
            n.prim = items[lo].prim;
            set(me, n);
            return me;
        }
        const Box cb = acc0.cb;
        float scale3[3];
        bool axis_ok[3];
        for (int a = 0; a < 3; ++a) {
            const float ext = cb.hi[a] - cb.lo[a];
            axis_ok[a] = ext > 0.0f && ext < 1e30f;
            scale3[a] = axis_ok[a] ? (float)NB / ext : 0.0f;
        }
        // pass 2: the items into 16 bins per axis (on the stack: the recursion is ~40 deep, 2.6 KB per level)
        Bins total_store;
        Bins* const total = &total_store;
        for (int a = 0; a < 3; ++a) for (int b = 0; b < NB; ++b) { box_reset(total->bins[a][b]); total->cnt[a][b] = 0; }
        {
            std::mutex mu;
            auto scan = [&](size_t a0, size_t b0) {
                if (!wide_scan) {  // the only scanner: straight into the totals
                    for (size_t k = (size_t)lo + a0; k < (size_t)lo + b0; ++k) {
                        const Box& bx = items[k].box;
                        for (int a = 0; a < 3; ++a) {
                            if (!axis_ok[a]) continue;
                            const int bi = bin_of(bx, a, cb.lo[a], scale3[a]);
                            box_merge(total->bins[a][bi], bx);
                            total->cnt[a][bi]++;
                        }
                    }
                    return;
                }
                std::unique_ptr<Bins> loc(new Bins());
                for (int a = 0; a < 3; ++a) for (int b = 0; b < NB; ++b) { box_reset(loc->bins[a][b]); loc->cnt[a][b] = 0; }
                for (size_t k = (size_t)lo + a0; k < (size_t)lo + b0; ++k) {
                    const Box& bx = items[k].box;
                    for (int a = 0; a < 3; ++a) {
                        if (!axis_ok[a]) continue;
                        const int bi = bin_of(bx, a, cb.lo[a], scale3[a]);
                        box_merge(loc->bins[a][bi], bx);
                        loc->cnt[a][bi]++;
                    }
                }
                std::lock_guard<std::mutex> lock(mu);
                for (int a = 0; a < 3; ++a) for (int b = 0; b < NB; ++b) {
                    if (loc->cnt[a][b]) box_merge(total->bins[a][b], loc->bins[a][b]);
                    total->cnt[a][b] += loc->cnt[a][b];
                }
            };
            if (wide_scan) parallel_chunks((size_t)(hi - lo), 1 << 15, scan);
            else scan(0, (size_t)(hi - lo));
        }
        const int min_side = std::max(1, (int)(min_frac * (hi - lo)));
        int best_axis = -1, best_bin = -1;
        double best_cost = 1e300;
        for (int a = 0; a < 3; ++a) {
            if (!axis_ok[a]) continue;
            const Box* bins = total->bins[a];
            const int* cnt = total->cnt[a];
            double right_area[NB];
            int right_cnt[NB];
            Box acc;
            box_reset(acc);
            int c = 0;
            for (int b = NB - 1; b >= 1; --b) {
                if (cnt[b]) box_merge(acc, bins[b]);
                c += cnt[b];
                right_area[b] = c ? half_area(acc) : 0.0;
                right_cnt[b] = c;
            }
            box_reset(acc);
            c = 0;
            for (int b = 0; b + 1 < NB; ++b) {
                if (cnt[b]) box_merge(acc, bins[b]);
                c += cnt[b];
                if (c < min_side || right_cnt[b + 1] < min_side) continue;
                const double cost = half_area(acc) * c + right_area[b + 1] * right_cnt[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
            }
        }
        int mid = -1;
        if (best_axis >= 0) {
            const int a = best_axis;
            const float scale = scale3[a];
            auto left = [&](const Item& t) { return bin_of(t.box, a, cb.lo[a], scale) <= best_bin; };
            if (wide_scan) {
                // parallel stable partition through a scratch copy: per-chunk counts, prefix, scatter
                const size_t cnt_all = (size_t)(hi - lo);
                const size_t chunks = std::min<size_t>(64, std::max<size_t>(1, cnt_all >> 15));
                std::vector<size_t> nl(chunks + 1, 0);
                std::vector<Item> tmp(items.begin() + lo, items.begin() + hi);
                parallel_chunks(chunks, 1, [&](size_t c0, size_t c1) {
                    for (size_t c = c0; c < c1; ++c) {
                        size_t k = 0;
                        for (size_t i = cnt_all * c / chunks; i < cnt_all * (c + 1) / chunks; ++i) k += left(tmp[i]) ? 1 : 0;
                        nl[c + 1] = k;
                    }
                });
                for (size_t c = 0; c < chunks; ++c) nl[c + 1] += nl[c];
                const size_t n_left = nl[chunks];
                parallel_chunks(chunks, 1, [&](size_t c0, size_t c1) {
                    for (size_t c = c0; c < c1; ++c) {
                        const size_t i0 = cnt_all * c / chunks, i1 = cnt_all * (c + 1) / chunks;
                        size_t l = (size_t)lo + nl[c], r = (size_t)lo + n_left + (i0 - nl[c]);
                        for (size_t i = i0; i < i1; ++i) {
                            if (left(tmp[i])) items[l++] = tmp[i];
                            else items[r++] = tmp[i];
                        }
                    }
                });
                mid = lo + (int)n_left;
            } else {
                auto it = std::partition(items.begin() + lo, items.begin() + hi, left);
                mid = (int)(it - items.begin());
            }
        }
        if (mid <= lo || mid >= hi) {
            // no admissible SAH split (all centroids equal, or the balance bound excludes every bin boundary):
            // median of the centroids along the longest axis
            int a = 0;
            for (int i = 1; i < 3; ++i) if (cb.hi[i] - cb.lo[i] > cb.hi[a] - cb.lo[a]) a = i;
            mid = (lo + hi) / 2;
            std::nth_element(items.begin() + lo, items.begin() + mid, items.begin() + hi, [&](const Item& x, const Item& y) {
                const float cx = x.box.lo[a] + x.box.hi[a], cy = y.box.lo[a] + y.box.hi[a];
                return cx < cy || (cx == cy && x.prim < y.prim);
            });
        }
        if (par_depth > 0 && hi - lo > (1 << 14)) {
            auto fut = std::async(std::launch::async, [this, lo, mid, par_depth] { return build(lo, mid, par_depth - 1); });
            n.right = build(mid, hi, par_depth - 1);
            n.left = fut.get();
        } else {
            n.left = build(lo, mid, 0);
            n.right = build(mid, hi, 0);
        }
        set(me, n);
        return me;
    }
};

// Collapses the binary tree to <= 4 children per node and writes the wide nodes breadth-first: a serial pass
// decides the structure (which sub-trees become the children of which node, node numbers), a parallel pass fills
// the 128-byte records.
void emit_wide_tree(HostScene& s, const std::vector<UpNode>& up, int root, const std::vector<Box>& cull, const std::vector<float>& cq) {
    struct Plan { int up, level, sp, first; int slots[4]; int8_t n, ni; };
    std::vector<float> area(up.size());
    parallel_chunks(up.size(), 1 << 16, [&](size_t a, size_t b) { for (size_t i = a; i < b; ++i) area[i] = (float)half_area(up[i].box); });
    // Slot order = visit order of the any-hit PACKETS (every other loop sorts the children by entry distance).
    // Shadow rays all end at the light, so "farthest from the light first" is near-first for them at no run-time
    // cost: measured on configs[2] (one area light) 378 -> 363 ms for the shadow kernel, against 439 / 496 ms for
    // lowest-first / nearest-to-the-light-first and 434 / 445 ms for largest / smallest box first
    // (profiles/ab_r2l*.jsonl). With several lights one static order cannot suit them all (measured neutral for two
    // opposite lights): the construction order is kept. (Per-light ranks in the node, chosen per packet at run time,
    // were measured too: the sort they need in the packet loop costs 8 %, twice what the better order gives:
    // profiles/ab_r2m_ab.jsonl.) RT_B200_CHILD_ORDER = default | area | small | low |
    // light_near | light_far overrides.
    const int child_order = [&] {
        const char* e = std::getenv("RT_B200_CHILD_ORDER");
        const std::string v = e ? e : "";
        if (v.empty()) return s.lights.size() == 1 ? 4 : 0;
        return v == "area" ? 1 : (v == "small" ? -1 : (v == "low" ? 2 : (v == "light_near" ? 3 : (v == "light_far" ? 4 : 0))));
    }();
    float light0[3] = {0, 0, 0};
    if (!s.lights.empty()) for (int a = 0; a < 3; ++a) light0[a] = s.lights[0].location[a];
    // breadth-first, one level at a time: the nodes of a level pick their children in parallel, a prefix sum over the
    // level numbers the children (node index == position in `plan`)
    std::vector<Plan> plan;
    plan.reserve(up.size() / 2 + 2);
    plan.push_back({root, 1, 0, 0, {0, 0, 0, 0}, 0, 0});
    s.stack_need = 1;
    s.wide_depth = 0;
    size_t level_begin = 0;
    while (level_begin < plan.size()) {
        const size_t level_end = plan.size();
        s.wide_depth = std::max(s.wide_depth, plan[level_begin].level);
        parallel_chunks(level_end - level_begin, 1 << 12, [&](size_t a, size_t b) {
            for (size_t qi = level_begin + a; qi < level_begin + b; ++qi) {
                Plan& it = plan[qi];
                const UpNode& u = up[(size_t)it.up];
                int* slots = it.slots;
                int n = 0;
                if (u.left < 0) slots[n++] = it.up;  // a scene of one primitive: the root holds it
                else {
                    slots[n++] = u.left;
                    slots[n++] = u.right;
                }
                while (n < 4) {  // open the inner child with the largest surface area
                    int pick = -1;
                    float best = -1.0f;
                    for (int k = 0; k < n; ++k)
                        if (up[(size_t)slots[k]].left >= 0 && area[(size_t)slots[k]] > best) { best = area[(size_t)slots[k]]; pick = k; }
                    if (pick < 0) break;
                    const int c = slots[pick];
                    slots[pick] = up[(size_t)c].left;
                    slots[n++] = up[(size_t)c].right;
                }
                // inner children first (consecutive node indices), primitives behind them
                std::stable_partition(slots, slots + n, [&](int c) { return up[(size_t)c].left >= 0; });
                int ni = 0;
                while (ni < n && up[(size_t)slots[ni]].left >= 0) ++ni;
                // slot order is the visit order of the any-hit packets (the other loops sort by entry distance):
                // RT_B200_CHILD_ORDER=area puts the largest boxes first, =small the smallest
                if (child_order != 0) {
                    auto key = [&](int x) -> float {
                        const Box& b = up[(size_t)x].box;
                        if (child_order == 1) return -area[(size_t)x];
                        if (child_order == -1) return area[(size_t)x];
                        if (child_order == 2) return b.lo[2] + b.hi[2];  // lowest first
                        float d2 = 0.0f;                                   // distance of the box centre from the first light
                        for (int a = 0; a < 3; ++a) { const float c = 0.5f * (b.lo[a] + b.hi[a]) - light0[a]; d2 += c * c; }
                        return child_order == 3 ? d2 : -d2;               // 3: nearest to the light first, 4: farthest first
                    };
                    std::stable_sort(slots, slots + ni, [&](int x, int y) { return key(x) < key(y); });
                    std::stable_sort(slots + ni, slots + n, [&](int x, int y) { return key(x) < key(y); });
                }
                it.n = (int8_t)n;
                it.ni = (int8_t)ni;
            }
        });
        size_t next = level_end;
        for (size_t qi = level_begin; qi < level_end; ++qi) {
            Plan& it = plan[qi];
            it.first = (int)next;
            next += (size_t)it.ni;
            // the traversal pushes up to ni - 1 sibling nodes before it descends: the stack a ray can need below here
            s.stack_need = std::max(s.stack_need, it.sp + std::max(0, (int)it.ni - 1) + 1);
        }
        if (next >= (1u << 30)) throw std::runtime_error("too many BVH nodes");
        plan.resize(next);
        parallel_chunks(level_end - level_begin, 1 << 12, [&](size_t a, size_t b) {
            for (size_t qi = level_begin + a; qi < level_begin + b; ++qi) {
                const Plan it = plan[qi];
                const int pushed = std::max(0, (int)it.ni - 1);
                for (int k = 0; k < it.ni; ++k) plan[(size_t)it.first + (size_t)k] = {it.slots[k], it.level + 1, it.sp + pushed, 0, {0, 0, 0, 0}, 0, 0};
            }
        });
        level_begin = level_end;
    }
    const size_t n_nodes = plan.size();
    s.dwide.assign(n_nodes, empty_wide());
    parallel_chunks(plan.size(), 1 << 14, [&](size_t a, size_t b) {
        for (size_t qi = a; qi < b; ++qi) {
            const Plan& it = plan[qi];
            DWide& w = s.dwide[qi];  // breadth-first: the qi-th planned node is node qi
            uint32_t meta = 0;
            float qmax = 0.0f;
            for (int k = 0; k < it.n; ++k) {
                const UpNode& c = up[(size_t)it.slots[k]];
                meta |= 1u << k;
                if (c.left < 0) {
                    // a primitive child carries its own (tight) culling box; Q of the node = the largest of its spheres
                    const HostPrim& p = s.prims[(size_t)s.order[(size_t)c.prim]];
                    set_child_box(w, k, cull[(size_t)c.prim]);
                    meta |= (16u << k) | ((uint32_t)p.type << (16 + 2 * k));
                    w.f[27 + k] = bits_f((uint32_t)c.prim);
                    qmax = std::max(qmax, cq[(size_t)c.prim]);
                } else {
                    set_child_box(w, k, c.box);
                }
            }
            w.f[24] = bits_f((uint32_t)it.first);
            w.f[25] = bits_f(meta);
            w.f[26] = qmax;
        }
    });
}

}  // namespace

void flatten_scene(HostScene& s) {
    const bool timing = std::getenv("RT_B200_DEBUG") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        const auto now = std::chrono::steady_clock::now();
        if (timing) std::fprintf(stderr, "[rt_b200] flatten: %s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    const int n = (int)s.prims.size();
    // primitives in sorted order
    s.dprims.assign((size_t)n, DPrim{});
    parallel_chunks((size_t)n, 16384, [&](size_t k_lo, size_t k_hi) {
    for (size_t k = k_lo; k < k_hi; ++k) {
        const HostPrim& p = s.prims[s.order[k]];
        DPrim& d = s.dprims[k];
        const uint32_t tag = (uint32_t)p.type | ((uint32_t)p.material << 2);
        d.q[0] = {p.velocity[0], p.velocity[1], p.velocity[2], bits_f(tag)};
        if (p.type == RT_PLANE) {
            d.q[1] = {p.corners[0][0], p.corners[0][1], p.corners[0][2], p.corners[3][0]};
            d.q[2] = {p.corners[1][0], p.corners[1][1], p.corners[1][2], p.corners[3][1]};
            d.q[3] = {p.corners[2][0], p.corners[2][1], p.corners[2][2], p.corners[3][2]};
            d.q[4] = {p.normal[0], p.normal[1], p.normal[2], p.normal_valid ? 1.0f : 0.0f};
        } else {
            for (int r = 0; r < 3; ++r) {
                d.q[1 + r] = {p.w2o[r][0], p.w2o[r][1], p.w2o[r][2], p.w2o[r][3]};
                d.q[4 + r] = {p.o2w[r][0], p.o2w[r][1], p.o2w[r][2], p.o2w[r][3]};
            }
        }
        d.q[7] = {bits_f((uint32_t)s.order[k]), 0, 0, 0};
    }
    });

    s.dleafbox.assign((size_t)2 * n, F4{0, 0, 0, 0});
    parallel_chunks(s.tree.size(), 16384, [&](size_t t_lo, size_t t_hi) {
    for (size_t ti = t_lo; ti < t_hi; ++ti) {
        const TreeNode& t = s.tree[ti];
        if (t.left >= 0) continue;
        for (int k = 0; k < t.count; ++k) {
            s.dleafbox[2 * (size_t)(t.first + k)] = {t.box.lo[0], t.box.lo[1], t.box.lo[2], 0.0f};
            s.dleafbox[2 * (size_t)(t.first + k) + 1] = {t.box.hi[0], t.box.hi[1], t.box.hi[2], 0.0f};
        }
    }
    });
    lap("primitive records + gate boxes");
    s.dwide.clear();
    s.wide_depth = 0;
    s.stack_need = 1;
    if (!s.tree.empty()) {
        // bound of |x|+|y|+|z| over every ray origin: the scene box and the camera (+ lens radius)
        double scene_g = std::fabs((double)s.cam.location[0]) + std::fabs((double)s.cam.location[1]) +
                         std::fabs((double)s.cam.location[2]) + 2.0 * std::fabs((double)s.cam.aperture);
        {
            double g = 0.0;
            for (int a = 0; a < 3; ++a)
                g += std::max(std::fabs((double)s.tree[0].box.lo[a]), std::fabs((double)s.tree[0].box.hi[a]));
            if (g < 1e30) scene_g = std::max(scene_g, g);
        }
        // Per primitive (sorted order):
        //   cull[k], cq[k] : its culling box + sphere coefficient -- what the node that holds it tests;
        //   abox[k]        : what its ANCESTORS must contain. A shape is only ever tested when its reference leaf's
        //                    box passes the reference's test (the gate), so every ray that matters passes that box:
        //                    the culling box, or -- where the culling box is unbounded (degenerate quads) or needs
        //                    the distance-dependent sphere term Q, which must not leak into every level of the
        //                    tree -- the (padded) box of the reference leaf.
        std::vector<Box> cull((size_t)n), abox((size_t)n);
        std::vector<float> cq((size_t)n, 0.0f);
        parallel_chunks((size_t)n, 8192, [&](size_t k_lo, size_t k_hi) {
            for (size_t k = k_lo; k < k_hi; ++k) {
                const HostPrim& p = s.prims[(size_t)s.order[k]];
                double pad = 0.0;
                float q = 0.0f;
                Box cb, lb;
                const F4 lo = s.dleafbox[2 * k], hi = s.dleafbox[2 * k + 1];
                const float l3[3] = {lo.x, lo.y, lo.z}, h3[3] = {hi.x, hi.y, hi.z};
                for (int a = 0; a < 3; ++a) {
                    const double m = 1e-6 * (std::fabs((double)l3[a]) + std::fabs((double)h3[a])) + 1e-6;
                    lb.lo[a] = std::nextafter((float)((double)l3[a] - m), -FLT_MAX);
                    lb.hi[a] = std::nextafter((float)((double)h3[a] + m), FLT_MAX);
                }
                if (cull_pad(p, scene_g, pad, q)) {
                    for (int a = 0; a < 3; ++a) {
                        cb.lo[a] = std::nextafter((float)((double)p.box.lo[a] - pad), -FLT_MAX);
                        cb.hi[a] = std::nextafter((float)((double)p.box.hi[a] + pad), FLT_MAX);
                    }
                    abox[k] = q > 0.0f ? lb : cb;
                } else {
                    cb = lb;
                    q = 0.0f;
                    abox[k] = lb;
                }
                cull[k] = cb;
                cq[k] = q;
            }
        });
        lap("culling boxes");
        // the traversal stacks live in shared memory (36 entries x 6 blocks is what an SM holds): a tree that could
        // need more is rebuilt with a balance bound
        for (double min_frac : {0.0, 0.2, 0.35, 0.5}) {
            SahBuilder b{min_frac, {}, {}};
            b.items.resize((size_t)n);
            for (int k = 0; k < n; ++k) b.items[(size_t)k] = {abox[(size_t)k], k};
            b.nodes.resize(2 * (size_t)n);
            static const int sah_par = [] { const char* e = std::getenv("RT_B200_SAH_PAR"); return e ? std::atoi(e) : 8; }();  // sub-trees built on other threads down to this depth
            const int root = b.build(0, n, sah_par);
            lap("SAH build");
            emit_wide_tree(s, b.nodes, root, cull, cq);
            lap("collapse + emit");
            if (std::getenv("RT_B200_DEBUG")) std::fprintf(stderr, "[rt_b200] device tree: min_frac %.2f -> %zu nodes, depth %d, stack need %d\n", min_frac, s.dwide.size(), s.wide_depth, s.stack_need);
            if (s.stack_need <= 36) break;
        }
    }

    s.dmaterials.assign(s.materials.size(), DMaterial{});
    for (size_t i = 0; i < s.materials.size(); ++i) {
        const rt_material_desc& m = s.materials[i];
        DMaterial& d = s.dmaterials[i];
        d.m[0] = {m.diffuse_color[0], m.diffuse_color[1], m.diffuse_color[2], m.k_ambient};
        d.m[1] = {m.specular_color[0], m.specular_color[1], m.specular_color[2], m.k_diffuse};
        d.m[2] = {m.k_specular, m.shininess, m.roughness, m.reflectivity};
        d.m[3] = {m.transparency, m.refractive_index, bits_f((uint32_t)m.texture), 0};
    }
    s.dlights.assign(s.lights.size(), DLight{});
    for (size_t i = 0; i < s.lights.size(); ++i) {
        const rt_light_desc& l = s.lights[i];
        s.dlights[i].l[0] = {l.location[0], l.location[1], l.location[2], l.intensity};
        s.dlights[i].l[1] = {l.color[0], l.color[1], l.color[2], l.radius};
    }
    s.dtextures.clear();
    s.texels.clear();
    for (const Texture& t : s.textures) {
        DTexture d{t.width, t.height, (uint32_t)s.texels.size(), 0};
        s.dtextures.push_back(d);
        s.texels.insert(s.texels.end(), t.rgb.begin(), t.rgb.end());
    }
}

}  // namespace rtb
