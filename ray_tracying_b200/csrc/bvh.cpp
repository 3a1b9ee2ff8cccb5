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
#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstring>
#include <future>
#include <stdexcept>

#include "scene.hpp"

namespace rtb {
namespace {

struct BuildCtx {
    const std::vector<HostPrim>& prims;
    std::vector<int>& order;
    std::vector<float> centre[3];  // (lo+hi)/2 per axis, acceleration.cpp:49-50
};

Box range_box(const BuildCtx& c, int start, int end) {
    Box b;
    for (int i = 0; i < 3; ++i) { b.lo[i] = FLT_MAX; b.hi[i] = -FLT_MAX; }
    for (int k = start; k < end; ++k) {
        const Box& pb = c.prims[c.order[k]].box;
        for (int i = 0; i < 3; ++i) { b.lo[i] = std::min(b.lo[i], pb.lo[i]); b.hi[i] = std::max(b.hi[i], pb.hi[i]); }
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
void sort_ranges(BuildCtx& c, int start, int end, int par_depth) {
    if (end - start <= 4) return;
    const Box b = range_box(c, start, end);
    const int axis = longest_axis(b);
    const std::vector<float>& ctr = c.centre[axis];
    std::sort(c.order.begin() + start, c.order.begin() + end, [&ctr](int a, int b2) { return ctr[a] < ctr[b2]; });
    const int mid = (start + end) / 2;
    if (par_depth > 0 && end - start > (1 << 15)) {
        auto fut = std::async(std::launch::async, [&c, start, mid, par_depth] { sort_ranges(c, start, mid, par_depth - 1); });
        sort_ranges(c, mid, end, par_depth - 1);
        fut.get();
    } else {
        sort_ranges(c, start, mid, 0);
        sort_ranges(c, mid, end, 0);
    }
}

// Pass 2: pre-order node array; boxes bottom-up (min/max are exact, so the union of the two
// children equals the reference's union over the whole range).
int emit_nodes(HostScene& s, int start, int end) {
    const int me = (int)s.tree.size();
    s.tree.emplace_back();
    if (end - start <= 4) {
        TreeNode& n = s.tree[me];
        n.first = start;
        n.count = end - start;
        for (int i = 0; i < 3; ++i) { n.box.lo[i] = FLT_MAX; n.box.hi[i] = -FLT_MAX; }
        for (int k = start; k < end; ++k) {
            const Box& pb = s.prims[s.order[k]].box;
            for (int i = 0; i < 3; ++i) { n.box.lo[i] = std::min(n.box.lo[i], pb.lo[i]); n.box.hi[i] = std::max(n.box.hi[i], pb.hi[i]); }
        }
        s.n_leaves++;
        return me;
    }
    const int mid = (start + end) / 2;
    const int l = emit_nodes(s, start, mid);
    const int r = emit_nodes(s, mid, end);
    TreeNode& n = s.tree[me];
    n.left = l;
    n.right = r;
    n.first = start;
    n.count = end - start;
    for (int i = 0; i < 3; ++i) {
        n.box.lo[i] = std::min(s.tree[l].box.lo[i], s.tree[r].box.lo[i]);
        n.box.hi[i] = std::max(s.tree[l].box.hi[i], s.tree[r].box.hi[i]);
    }
    return me;
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
    BuildCtx c{s.prims, s.order, {}};
    for (int a = 0; a < 3; ++a) {
        c.centre[a].resize(n);
        for (int i = 0; i < n; ++i) c.centre[a][i] = (s.prims[i].box.lo[a] + s.prims[i].box.hi[a]) / 2.0f;
    }
    sort_ranges(c, 0, n, 3);
    s.tree.reserve((size_t)n);
    emit_nodes(s, 0, n);
    s.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void flatten_scene(HostScene& s) {
    const int n = (int)s.prims.size();
    // primitives in sorted order
    s.dprims.assign((size_t)n, DPrim{});
    for (int k = 0; k < n; ++k) {
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

    s.dnodes.clear();
    s.dleaves.clear();
    s.dnodes.clear();
    s.root_ref = 0;
    for (int i = 0; i < 3; ++i) { s.root_box.lo[i] = FLT_MAX; s.root_box.hi[i] = -FLT_MAX; }
    if (!s.tree.empty()) {
        s.root_box = s.tree[0].box;
        std::vector<int> dev_index(s.tree.size(), -1);
        int n_internal = 0, n_leaf = 0;
        for (size_t i = 0; i < s.tree.size(); ++i)
            dev_index[i] = (s.tree[i].left >= 0) ? n_internal++ : n_leaf++;
        auto ref_of = [&](int ti) -> int32_t {
            const TreeNode& t = s.tree[ti];
            return t.left >= 0 ? dev_index[ti] : leaf_ref(dev_index[ti]);
        };
        // leaf records (pre-order, so neighbouring leaves are neighbours in memory)
        s.dleaves.assign((size_t)n_leaf, DLeaf{});
        for (size_t i = 0; i < s.tree.size(); ++i) {
            const TreeNode& t = s.tree[i];
            if (t.left >= 0) continue;
            DLeaf& L = s.dleaves[dev_index[i]];
            L.l[0] = {t.box.lo[0], t.box.lo[1], t.box.lo[2], bits_f((uint32_t)t.first)};
            // bits 0-2 count; bits 4+4T..7+4T = which of the (up to 4) primitives have type T
            uint32_t meta = (uint32_t)t.count;
            for (int k = 0; k < t.count; ++k) meta |= (1u << k) << (4 + 4 * s.prims[s.order[t.first + k]].type);
            L.l[1] = {t.box.hi[0], t.box.hi[1], t.box.hi[2], bits_f(meta)};
            float pb[24];
            for (int k = 0; k < 4; ++k) {
                for (int a = 0; a < 3; ++a) { pb[6 * k + a] = FLT_MAX; pb[6 * k + 3 + a] = -FLT_MAX; }
                if (k >= t.count) continue;
                const Box& b = s.prims[s.order[t.first + k]].box;
                for (int a = 0; a < 3; ++a) {
                    // Culling box: the primitive's own box pushed outward by far more than the
                    // rounding noise of the intersection routines (~1e-7 relative), so a ray that
                    // misses it cannot be reported as a hit by the reference's primitive test.
                    const double ext = (double)b.hi[a] - (double)b.lo[a];
                    const double mag = std::max(std::fabs((double)b.lo[a]), std::fabs((double)b.hi[a]));
                    const double pad = 1e-5 * (ext + mag) + 1e-6;
                    pb[6 * k + a] = std::nextafter((float)((double)b.lo[a] - pad), -FLT_MAX);
                    pb[6 * k + 3 + a] = std::nextafter((float)((double)b.hi[a] + pad), FLT_MAX);
                }
            }
            for (int q = 0; q < 6; ++q) L.l[2 + q] = {pb[4 * q], pb[4 * q + 1], pb[4 * q + 2], pb[4 * q + 3]};
        }
        s.dnodes.assign((size_t)n_internal, DNode{});
        for (size_t i = 0; i < s.tree.size(); ++i) {
            const TreeNode& t = s.tree[i];
            if (t.left < 0) continue;
            const Box& L = s.tree[t.left].box;
            const Box& R = s.tree[t.right].box;
            DNode& d = s.dnodes[dev_index[i]];
            d.a = {L.lo[0], L.lo[1], L.lo[2], L.hi[0]};
            d.b = {L.hi[1], L.hi[2], R.lo[0], R.lo[1]};
            d.c = {R.lo[2], R.hi[0], R.hi[1], R.hi[2]};
            d.d = {bits_f((uint32_t)ref_of(t.left)), bits_f((uint32_t)ref_of(t.right)), 0, 0};
        }
        s.root_ref = ref_of(0);
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
