// rt_oracle.cpp -- CPU ORACLE (test infrastructure only; never on the product path).
//
// A scalar restatement of the reference renderer's algorithm, written from the reference's
// behaviour (file:line cited at every function), used as the checker for the CUDA path:
//   * same float operations in the same order as the reference's x86-64 build (this file is
//     compiled with -ffp-contract=off and no -march, so no FMA),
//   * the LITERAL traversal semantics of the reference: visit every node whose box the ray
//     passes, collect all leaf hits, return the first minimum -- no ordering, no pruning,
//   * recursion for Trace() exactly like the reference.
// The one deliberate difference: random numbers. The reference pulls them from one serial
// std::mt19937 (raytracer.cpp:425-427), which no parallel renderer can reproduce; the oracle
// uses the same counter-based Philox4x32-10 keying as the product (key = pixel, seed; counter =
// sample, purpose|tree-node, light|shadow-sample, attempt) and draws from the reference's
// distributions. Deterministic scenes (no area lights, roughness, aperture, motion) use no
// random numbers at all and are compared bit for bit with the real reference
// (oracle/_ref/ref_driver); stochastic ones are compared statistically with it.
//
// Parity pinning: tests/test_oracle_golden.py checks this file against the unmodified
// reference compiled from /root/reference (hit IDs, hit distances, BVH dump, 8-bit images) and
// against committed golden vectors generated from it (tests/golden/).
#include "rt_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct V { float x, y, z; };
inline V sub(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V add(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V mul(V a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// raytracer.cpp:75-79, camera.cpp:60-68
inline V unit(V v) {
    float m = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    if (m == 0.0f) return {0, 0, 0};
    return {v.x / m, v.y / m, v.z / m};
}
inline float comp(V v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

struct M4 { float m[4][4]; };
M4 matmul(const M4& A, const M4& B) {  // shapes.cpp:141-149
    M4 R;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += A.m[i][k] * B.m[k][j];
            R.m[i][j] = s;
        }
    return R;
}
M4 rows(float a, float b, float c, float d, float e, float f, float g, float h, float i, float j, float k, float l) {
    M4 R = {{{a, b, c, d}, {e, f, g, h}, {i, j, k, l}, {0, 0, 0, 1}}};
    return R;
}
V apply_point(const M4& M, V p) {  // shapes.cpp:151-158
    float w = M.m[3][0] * p.x + M.m[3][1] * p.y + M.m[3][2] * p.z + M.m[3][3];
    V r = {M.m[0][0] * p.x + M.m[0][1] * p.y + M.m[0][2] * p.z + M.m[0][3],
           M.m[1][0] * p.x + M.m[1][1] * p.y + M.m[1][2] * p.z + M.m[1][3],
           M.m[2][0] * p.x + M.m[2][1] * p.y + M.m[2][2] * p.z + M.m[2][3]};
    if (fabs(w - 1.0f) > 1e-6f && w != 0) { r.x /= w; r.y /= w; r.z /= w; }
    return r;
}
V apply_vector(const M4& M, V v) {  // shapes.cpp:160-165
    return {M.m[0][0] * v.x + M.m[0][1] * v.y + M.m[0][2] * v.z, M.m[1][0] * v.x + M.m[1][1] * v.y + M.m[1][2] * v.z,
            M.m[2][0] * v.x + M.m[2][1] * v.y + M.m[2][2] * v.z};
}

struct Bounds {
    float lo[3], hi[3];
    Bounds() { for (int i = 0; i < 3; i++) { lo[i] = FLT_MAX; hi[i] = -FLT_MAX; } }  // shapes.hpp:35-40
    void grow(V p) {
        const float c[3] = {p.x, p.y, p.z};
        for (int i = 0; i < 3; i++) { lo[i] = std::min(lo[i], c[i]); hi[i] = std::max(hi[i], c[i]); }
    }
    void grow(const Bounds& o) {
        for (int i = 0; i < 3; i++) { lo[i] = std::min(lo[i], o.lo[i]); hi[i] = std::max(hi[i], o.hi[i]); }
    }
};

struct Beam { V o, d; float time; };

// AABB::intersect, shapes.cpp:55-72 (note the double-precision 1e-6 there)
bool passes_box(const Bounds& b, const Beam& r) {
    float tn = -FLT_MAX, tf = FLT_MAX;
    for (int i = 0; i < 3; i++) {
        const float d = comp(r.d, i), o = comp(r.o, i);
        if (fabs(d) < 1e-6) {
            if (o < b.lo[i] || o > b.hi[i]) return false;
        } else {
            float t1 = (b.lo[i] - o) / d, t2 = (b.hi[i] - o) / d;
            if (t1 > t2) std::swap(t1, t2);
            tn = std::max(tn, t1);
            tf = std::min(tf, t2);
            if (tn > tf || tf < 0) return false;
        }
    }
    return true;
}

struct Surface {  // what the reference calls Hit
    V point, normal;
    float t;
    int shape;  // load-order index, -1 = none
    float u, v;
};

struct Solid {
    int kind, material;
    V velocity;
    M4 to_object, to_world;
    V corner[4];
    Bounds box;

    // Shapes::buildTransformationMatrices, shapes.cpp:92-139
    void set_transform(const float t[3], const float r[3], const float s[3]) {
        M4 S = rows(s[0], 0, 0, 0, 0, s[1], 0, 0, 0, 0, s[2], 0);
        float cx = cos(r[0]), sx = sin(r[0]);  // double overloads, narrowed (as in the reference)
        float cy = cos(r[1]), sy = sin(r[1]);
        float cz = cos(r[2]), sz = sin(r[2]);
        M4 R = rows(cy * cz, sx * sy * cz - cx * sz, cx * sy * cz + sx * sz, 0, cy * sz, sx * sy * sz + cx * cz,
                    cx * sy * sz - sx * cz, 0, -sy, sx * cy, cx * cy, 0);
        M4 T = rows(1, 0, 0, t[0], 0, 1, 0, t[1], 0, 0, 1, t[2]);
        to_world = matmul(T, matmul(R, S));
        M4 iS = rows(1.0f / s[0], 0, 0, 0, 0, 1.0f / s[1], 0, 0, 0, 0, 1.0f / s[2], 0);
        M4 iR = rows(R.m[0][0], R.m[1][0], R.m[2][0], 0, R.m[0][1], R.m[1][1], R.m[2][1], 0, R.m[0][2], R.m[1][2],
                     R.m[2][2], 0);
        M4 iT = rows(1, 0, 0, -t[0], 0, 1, 0, -t[1], 0, 0, 1, -t[2]);
        to_object = matmul(matmul(iS, iR), iT);
    }

    // Shapes::transformNormal, shapes.cpp:167-187 (always uses world_to_object transposed)
    V world_normal(V n) const {
        V r = {to_object.m[0][0] * n.x + to_object.m[1][0] * n.y + to_object.m[2][0] * n.z,
               to_object.m[0][1] * n.x + to_object.m[1][1] * n.y + to_object.m[2][1] * n.z,
               to_object.m[0][2] * n.x + to_object.m[1][2] * n.y + to_object.m[2][2] * n.z};
        float len = sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
        if (len > 1e-6f) { r.x /= len; r.y /= len; r.z /= len; }
        return r;
    }

    void compute_box() {
        box = Bounds();
        if (kind == 0) {  // Sphere::get_bounding_box, shapes.cpp:264-287
            for (int k = 0; k < 8; k++) {
                V c = {(k & 1) ? 1.0f : -1.0f, (k & 2) ? 1.0f : -1.0f, (k & 4) ? 1.0f : -1.0f};
                V w = apply_point(to_world, c);
                box.grow(w);
                box.grow(V{w.x + velocity.x, w.y + velocity.y, w.z + velocity.z});
            }
        } else if (kind == 1) {  // Cube, shapes.cpp:425-433
            for (int k = 0; k < 8; k++)
                box.grow(apply_point(to_world, V{(k & 1) ? 0.5f : -0.5f, (k & 2) ? 0.5f : -0.5f, (k & 4) ? 0.5f : -0.5f}));
        } else if (kind == 2) {  // Rectangle, shapes.cpp:335-343
            for (int k = 0; k < 4; k++) box.grow(apply_point(to_world, V{(k & 1) ? 0.5f : -0.5f, (k & 2) ? 0.5f : -0.5f, 0.0f}));
        } else {  // Plane, shapes.cpp:496-503
            for (int k = 0; k < 4; k++) box.grow(corner[k]);
            for (int i = 0; i < 3; i++) { box.lo[i] -= 1e-4f; box.hi[i] += 1e-4f; }
        }
    }
};

// shapes.cpp:24-40
bool inside_triangle(V P, V A, V B, V C, V n) {
    if (dot(cross(sub(B, A), sub(P, A)), n) < -1e-6f) return false;
    if (dot(cross(sub(C, B), sub(P, B)), n) < -1e-6f) return false;
    if (dot(cross(sub(A, C), sub(P, C)), n) < -1e-6f) return false;
    return true;
}

bool hit_solid(const Solid& s, int index, const Beam& ray, Surface& out) {
    if (s.kind == 3) {  // Plane::intersect, shapes.cpp:444-483
        V n = cross(sub(s.corner[1], s.corner[0]), sub(s.corner[2], s.corner[0]));
        float len = sqrt(dot(n, n));
        if (len < 1e-6f) return false;
        n = {n.x / len, n.y / len, n.z / len};
        float denom = dot(n, ray.d);
        if (fabs(denom) < 1e-6f) return false;
        float t = dot(sub(s.corner[0], ray.o), n) / denom;
        if (t < 0) return false;
        V P = {ray.o.x + t * ray.d.x, ray.o.y + t * ray.d.y, ray.o.z + t * ray.d.z};
        // isPointInQuad, shapes.cpp:485-494
        if (!inside_triangle(P, s.corner[1], s.corner[3], s.corner[2], n) &&
            !inside_triangle(P, s.corner[0], s.corner[1], s.corner[2], n))
            return false;
        V eu = sub(s.corner[1], s.corner[0]), ev = sub(s.corner[3], s.corner[0]), hv = sub(P, s.corner[0]);
        float u = dot(hv, eu) / dot(eu, eu), v = dot(hv, ev) / dot(ev, ev);
        out.u = std::max(0.0f, std::min(1.0f, u));
        out.v = std::max(0.0f, std::min(1.0f, v));
        out.point = P; out.normal = n; out.t = t; out.shape = index;
        return true;
    }

    Beam local;
    V origin = ray.o;
    if (s.kind == 0) {  // motion blur, shapes.cpp:203-209
        origin.x -= s.velocity.x * ray.time;
        origin.y -= s.velocity.y * ray.time;
        origin.z -= s.velocity.z * ray.time;
    }
    local.o = apply_point(s.to_object, origin);
    local.d = apply_vector(s.to_object, ray.d);
    V pl, nl;
    float u = 0, v = 0;

    if (s.kind == 0) {  // Sphere::intersect, shapes.cpp:200-262
        float a = dot(local.d, local.d);
        float b = 2.0f * dot(local.o, local.d);
        float c = dot(local.o, local.o) - 1.0f;
        float disc = b * b - 4 * a * c;
        if (disc < 0) return false;
        float sq = sqrt(disc);
        float t1 = (-b - sq) / (2.0f * a), t2 = (-b + sq) / (2.0f * a);
        float tl = (t1 > 0.001f) ? t1 : ((t2 > 0.001f) ? t2 : -1.0f);
        if (tl < 0) return false;
        pl = {local.o.x + tl * local.d.x, local.o.y + tl * local.d.y, local.o.z + tl * local.d.z};
        nl = pl;
        const float PI = 3.1415926535f;
        u = 0.5f + atan2(nl.z, nl.x) / (2.0f * PI);
        v = 0.5f - asin(nl.y) / PI;
    } else if (s.kind == 2) {  // Rectangle::intersect, shapes.cpp:299-333
        if (fabs(local.d.z) < 1e-6f) return false;
        float tl = -local.o.z / local.d.z;
        if (tl < 0.001f) return false;
        float hx = local.o.x + tl * local.d.x, hy = local.o.y + tl * local.d.y;
        if (hx < -0.5f || hx > 0.5f || hy < -0.5f || hy > 0.5f) return false;
        pl = {hx, hy, 0.0f};
        nl = {0.0f, 0.0f, 1.0f};
        u = hx + 0.5f;
        v = hy + 0.5f;
    } else {  // Cube::intersect, shapes.cpp:355-423
        float tn = -FLT_MAX, tf = FLT_MAX;
        int axis = -1, sign = 0;
        for (int i = 0; i < 3; i++) {
            float d = comp(local.d, i), o = comp(local.o, i);
            if (fabs(d) < 1e-6f) {
                if (o < -0.5f || o > 0.5f) return false;
            } else {
                float t1 = (-0.5f - o) / d, t2 = (0.5f - o) / d;
                float te = std::min(t1, t2), tx = std::max(t1, t2);
                if (te > tn) { tn = te; axis = i; sign = (t1 < t2) ? -1 : 1; }
                if (tx < tf) tf = tx;
                if (tn > tf || tf < 0) return false;
            }
        }
        float tl = (tn > 0) ? tn : tf;
        if (tl < 0) return false;
        pl = {local.o.x + tl * local.d.x, local.o.y + tl * local.d.y, local.o.z + tl * local.d.z};
        nl = {0, 0, 0};
        if (axis == 0) nl.x = (float)sign; else if (axis == 1) nl.y = (float)sign; else if (axis == 2) nl.z = (float)sign;
        float uc = pl.x + 0.5f, vc = pl.y + 0.5f, wc = pl.z + 0.5f;
        if (axis == 0) { u = (sign > 0) ? wc : (1.0f - wc); v = vc; }
        else if (axis == 1) { u = uc; v = (sign > 0) ? wc : (1.0f - wc); }
        else { u = (sign > 0) ? uc : (1.0f - uc); v = vc; }
    }
    V P = apply_point(s.to_world, pl);
    if (s.kind == 0) { P.x += s.velocity.x * ray.time; P.y += s.velocity.y * ray.time; P.z += s.velocity.z * ray.time; }
    out.point = P;
    out.normal = s.world_normal(nl);
    V dv = sub(P, ray.o);
    out.t = sqrt(dot(dv, dv));
    out.shape = index;
    out.u = u; out.v = v;
    return true;
}

struct TreeNode { Bounds box; int left = -1, right = -1, first = 0, count = 0; };

// ---- Philox4x32-10 (own copy; the product has its own in csrc/philox.cuh) ---------------------
void philox(const uint32_t c_in[4], uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t c[4] = {c_in[0], c_in[1], c_in[2], c_in[3]};
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n[4] = {(uint32_t)(p1 >> 32) ^ c[1] ^ k0, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c[3] ^ k1, (uint32_t)p0};
        std::memcpy(c, n, sizeof(c));
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    std::memcpy(out, c, sizeof(c));
}

struct Draws {
    uint32_t pixel, seed_lo, seed_hi, sample;
    void block(uint32_t purpose, uint32_t node, uint32_t sub, uint32_t attempt, uint32_t out[4]) const {
        uint32_t c[4] = {sample, (purpose << 28) | node, sub, attempt};
        philox(c, pixel ^ (seed_hi * 0x9E3779B1u), seed_lo, out);
    }
};
inline float unit_float(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
inline double unit_double(uint32_t x) { return (double)x * (1.0 / 4294967296.0); }

// raytracer.cpp:152-171
V ball_sample(const Draws& g, uint32_t purpose, uint32_t node, uint32_t sub) {
    V p = {0, 0, 0};
    for (uint32_t attempt = 0;; attempt++) {
        uint32_t r[4];
        g.block(purpose, node, sub, attempt, r);
        p = {2.0f * unit_float(r[0]) - 1.0f, 2.0f * unit_float(r[1]) - 1.0f, 2.0f * unit_float(r[2]) - 1.0f};
        if (dot(p, p) < 1.0f || attempt >= 63u) break;
    }
    return p;
}

struct Rgb { float r, g, b; };

}  // namespace

struct orc_scene {
    orc_camera cam;
    V cx, cy, cz;
    std::vector<orc_light> lights;
    std::vector<orc_material> materials;
    struct Tex { int w, h; std::vector<uint8_t> rgb; };
    std::vector<Tex> textures;
    std::vector<Solid> solids;      // load order
    std::vector<int> order;         // shape_list after construction
    std::vector<TreeNode> nodes;    // pre-order
};

namespace {

// BVH::construct_tree, acceleration.cpp:20-64
int build(orc_scene& sc, const std::vector<float> centre[3], int start, int end) {
    int me = (int)sc.nodes.size();
    sc.nodes.emplace_back();
    Bounds b;
    for (int i = start; i < end; i++) b.grow(sc.solids[sc.order[i]].box);
    sc.nodes[me].box = b;
    sc.nodes[me].first = start;
    sc.nodes[me].count = end - start;
    if (end - start <= 4) return me;
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    int axis = (dx > dy && dx > dz) ? 0 : ((dy > dz) ? 1 : 2);  // shapes.cpp:46-53
    const std::vector<float>& c = centre[axis];
    std::sort(sc.order.begin() + start, sc.order.begin() + end, [&c](int a, int b2) { return c[a] < c[b2]; });
    int mid = (start + end) / 2;
    int l = build(sc, centre, start, mid);
    int r = build(sc, centre, mid, end);
    sc.nodes[me].left = l;
    sc.nodes[me].right = r;
    return me;
}

struct Tracer {
    const orc_scene& sc;
    const orc_params& pr;
    uint64_t n_primary = 0, n_shadow = 0, n_secondary = 0;
    std::vector<Surface> found;

    // BVH::intersect_helper, acceleration.cpp:67-100
    void collect(const Beam& ray, int node) {
        const TreeNode& n = sc.nodes[node];
        if (!passes_box(n.box, ray)) return;
        if (n.left >= 0) {
            collect(ray, n.left);
            collect(ray, n.right);
        } else {
            for (int k = 0; k < n.count; k++) {
                Surface h;
                int id = sc.order[n.first + k];
                if (hit_solid(sc.solids[id], id, ray, h)) found.push_back(h);
            }
        }
    }

    // BVH::get_intersection, acceleration.cpp:103-150
    Surface nearest(const Beam& ray) {
        Surface best;
        best.t = FLT_MAX;
        best.shape = -1;
        if (sc.solids.empty()) return best;
        if (pr.use_bvh) {
            found.clear();
            collect(ray, 0);
            for (const Surface& h : found)
                if (h.t < best.t) best = h;  // std::min_element: first minimum
        } else {
            for (int id : sc.order) {
                Surface h;
                if (hit_solid(sc.solids[id], id, ray, h) && h.t < best.t) best = h;
            }
        }
        return best;
    }

    // Material::getDiffuseColor, material.hpp:99-134
    Rgb albedo(const orc_material& m, float u, float v) const {
        if (m.texture < 0) return {m.diffuse[0], m.diffuse[1], m.diffuse[2]};
        const orc_scene::Tex& t = sc.textures[m.texture];
        float fx = u * (t.w - 1), fy = (1.0f - v) * (t.h - 1);
        int r = 0, g = 0, b = 0;
        if (fx == fx && fy == fy) {
            int x = static_cast<int>(fx), y = static_cast<int>(fy);
            if (x >= 0 && x < t.w && y >= 0 && y < t.h) {
                const uint8_t* p = &t.rgb[((size_t)y * t.w + x) * 3];
                r = p[0]; g = p[1]; b = p[2];
            }
        }
        return {(r / 255.0f) * m.diffuse[0], (g / 255.0f) * m.diffuse[1], (b / 255.0f) * m.diffuse[2]};
    }

    // shade(), raytracer.cpp:180-274
    Rgb local_colour(const Surface& h, const Beam& view, const Draws& g, uint32_t node) {
        const orc_material& m = sc.materials[sc.solids[h.shape].material];
        Rgb base = albedo(m, h.u, h.v);
        Rgb out = {base.r * m.k_ambient, base.g * m.k_ambient, base.b * m.k_ambient};
        V view_dir = unit(sub(view.o, h.point));
        for (size_t li = 0; li < sc.lights.size(); li++) {
            const orc_light& L = sc.lights[li];
            V lpos = {L.location[0], L.location[1], L.location[2]};
            float visible = 0.0f;
            int n = (L.radius > 0.0f) ? pr.light_samples : 1;
            for (int s = 0; s < n; s++) {
                V target = lpos;
                if (L.radius > 0.0f) target = add(target, mul(ball_sample(g, 2u, node, ((uint32_t)li << 16) | (uint32_t)s), L.radius));
                V lv = sub(target, h.point);
                float ldist = std::sqrt(dot(lv, lv));
                Beam sh;
                sh.o = add(h.point, mul(h.normal, 1e-4f));
                sh.d = unit(lv);
                sh.time = 0.0f;
                n_shadow++;
                Surface blocker = nearest(sh);
                if (blocker.shape < 0 || blocker.t > ldist) visible += 1.0f;
            }
            visible /= (float)n;
            if (visible <= 0.0f) continue;
            V lc = sub(lpos, h.point);
            float d2 = dot(lc, lc);
            float d = std::sqrt(d2);
            V ldir = unit(lc);
            float ndl = std::max(0.0f, dot(h.normal, ldir));
            V half = unit(add(ldir, view_dir));
            float ndh = std::max(0.0f, dot(h.normal, half));
            float spec = std::pow(ndh, m.shininess);
            float att = 10.0f * L.intensity / (25.0f + 10.0f * d + 150.0f * d2);
            float cr = L.color[0] * ((base.r * ndl) * m.k_diffuse + (m.specular[0] * spec) * m.k_specular) * att;
            float cg = L.color[1] * ((base.g * ndl) * m.k_diffuse + (m.specular[1] * spec) * m.k_specular) * att;
            float cb = L.color[2] * ((base.b * ndl) * m.k_diffuse + (m.specular[2] * spec) * m.k_specular) * att;
            out = {out.r + cr * visible, out.g + cg * visible, out.b + cb * visible};
        }
        return out;
    }

    // Trace(), raytracer.cpp:280-351. `node` numbers the ray tree (root 1, reflection 2k, refraction 2k+1).
    Rgb trace(const Beam& ray, int depth, const Draws& g, uint32_t node, int* first_shape, float* first_t) {
        if (depth > pr.max_depth) return {0, 0, 0};
        if (depth == 0) n_primary++; else n_secondary++;
        Surface h = nearest(ray);
        if (first_shape) { *first_shape = h.shape; if (first_t) *first_t = h.t; }
        if (h.shape < 0) return {0.1f, 0.1f, 0.1f};
        Rgb local = local_colour(h, ray, g, node);
        const orc_material& m = sc.materials[sc.solids[h.shape].material];
        Rgb refl = {0, 0, 0}, refr = {0, 0, 0};
        if (m.reflectivity > 0.0f) {  // createReflectionRay raytracer.cpp:101-115 + glossy :312-327
            float idn = dot(ray.d, h.normal);
            V dir = sub(ray.d, mul(h.normal, 2.0f * idn));
            V origin = add(h.point, mul(h.normal, 1e-4f));
            if (m.roughness > 0.0f) {
                V fuzz = ball_sample(g, 3u, node, 0u);
                dir = unit(add(dir, mul(fuzz, m.roughness)));
                if (dot(dir, h.normal) < 0.0f) dir = {0, 0, 0};
            }
            if (dot(dir, dir) > 0.001f) {
                Beam nr = {origin, dir, 0.0f};
                refl = trace(nr, depth + 1, g, node * 2u, nullptr, nullptr);
            }
        }
        if (m.transparency > 0.0f) {  // createRefractionRay raytracer.cpp:118-150
            V N = h.normal;
            float n_in = 1.0f, n_out = m.refractive_index;
            float cos_i = dot(ray.d, N);
            if (cos_i > 0) { std::swap(n_in, n_out); N = mul(N, -1.0f); }
            float eta = n_in / n_out;
            float ca = std::abs(cos_i);
            float disc = 1.0f - eta * eta * (1.0f - ca * ca);
            if (!(disc < 0)) {
                float ct = std::sqrt(disc);
                V dir = unit(add(mul(ray.d, eta), mul(N, (eta * ca - ct))));
                V origin = add(h.point, mul(N, -1e-4f));
                if (dot(dir, dir) > 1e-6f) {
                    Beam nr = {origin, dir, 0.0f};
                    refr = trace(nr, depth + 1, g, node * 2u + 1u, nullptr, nullptr);
                }
            }
        }
        float keep = std::max(0.0f, 1.0f - m.reflectivity - m.transparency);
        return {keep * local.r + m.reflectivity * refl.r + m.transparency * refr.r,
                keep * local.g + m.reflectivity * refl.g + m.transparency * refr.g,
                keep * local.b + m.reflectivity * refl.b + m.transparency * refr.b};
    }

    // compute_pixel_color raytracer.cpp:18-70 + Camera::pixelToRay_thin_lens camera.cpp:97-178
    Beam camera_ray(int x, int y, int s, const Draws& g) const {
        uint32_t r[4];
        g.block(0u, 0u, 0u, 0u, r);
        float fx, fy;
        if (pr.samples_sqrt <= 1) { fx = x + 0.5f; fy = y + 0.5f; }
        else {
            int i = s % pr.samples_sqrt, j = s / pr.samples_sqrt;
            double sx = (i + unit_double(r[0])) / pr.samples_sqrt, sy = (j + unit_double(r[1])) / pr.samples_sqrt;
            fx = (float)(x + sx); fy = (float)(y + sy);
        }
        const orc_camera& c = sc.cam;
        float nx = 1 - (fx / (float)c.res_x) * 2, ny = 1 - (fy / (float)c.res_y) * 2;
        float nxr = nx * ((float)c.sensor_width / 2.0f), nyr = ny * ((float)c.sensor_height / 2.0f);
        V dir = {sc.cx.x * nxr + sc.cy.x * nyr + sc.cz.x * c.focal_length, sc.cx.y * nxr + sc.cy.y * nyr + sc.cz.y * c.focal_length,
                 sc.cx.z * nxr + sc.cy.z * nyr + sc.cz.z * c.focal_length};
        dir = unit(dir);
        V loc = {c.location[0], c.location[1], c.location[2]};
        Beam out = {loc, dir, 0.0f};
        if (c.aperture > 0.0f) {
            V focus = {loc.x + dir.x * c.focus_dist, loc.y + dir.y * c.focus_dist, loc.z + dir.z * c.focus_dist};
            float rx = 0, ry = 0;
            for (uint32_t attempt = 0;; attempt++) {  // random_in_unit_disk camera.cpp:89-95
                uint32_t l[4];
                g.block(1u, 0u, 0u, attempt, l);
                rx = unit_float(l[0]) * 2.0f - 1.0f;
                ry = unit_float(l[1]) * 2.0f - 1.0f;
                if (rx * rx + ry * ry < 1.0f || attempt >= 63u) break;
            }
            float lens = c.aperture / 2.0f;
            rx *= lens; ry *= lens;
            V off = {sc.cx.x * rx + sc.cy.x * ry, sc.cx.y * rx + sc.cy.y * ry, sc.cx.z * rx + sc.cy.z * ry};
            out.o = {loc.x + off.x, loc.y + off.y, loc.z + off.z};
            out.d = unit(V{focus.x - out.o.x, focus.y - out.o.y, focus.z - out.o.z});
        }
        out.time = pr.fixed_time >= 0.0f ? pr.fixed_time : unit_float(r[2]);
        return out;
    }
};

}  // namespace

extern "C" {

int orc_scene_create(const orc_scene_desc* d, orc_scene** out) {
    if (!d || !out) return -1;
    orc_scene* sc = new orc_scene();
    sc->cam = d->camera;
    V gaze = {d->camera.gaze[0], d->camera.gaze[1], d->camera.gaze[2]}, up = {d->camera.up[0], d->camera.up[1], d->camera.up[2]};
    sc->cz = unit(gaze);                 // camera.cpp:109-115
    sc->cx = unit(cross(up, sc->cz));
    sc->cy = unit(cross(sc->cz, sc->cx));
    sc->lights.assign(d->lights, d->lights + d->n_lights);
    sc->materials.assign(d->materials, d->materials + d->n_materials);
    for (int i = 0; i < d->n_textures; i++) {
        orc_scene::Tex t;
        t.w = d->textures[i].width; t.h = d->textures[i].height;
        t.rgb.assign(d->textures[i].rgb, d->textures[i].rgb + (size_t)t.w * t.h * 3);
        sc->textures.push_back(std::move(t));
    }
    sc->solids.resize(d->n_shapes);
    for (int i = 0; i < d->n_shapes; i++) {
        const orc_shape& in = d->shapes[i];
        Solid& s = sc->solids[i];
        s.kind = in.type;
        s.material = in.material;
        s.velocity = {0, 0, 0};
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) s.to_object.m[r][c] = s.to_world.m[r][c] = (r == c) ? 1.0f : 0.0f;
        for (int k = 0; k < 4; k++) s.corner[k] = {in.corners[3 * k], in.corners[3 * k + 1], in.corners[3 * k + 2]};
        if (s.kind != 3) s.set_transform(in.translation, in.rotation, in.scale);
        if (s.kind == 0) s.velocity = {in.velocity[0], in.velocity[1], in.velocity[2]};
        s.compute_box();
    }
    int n = d->n_shapes;
    sc->order.resize(n);
    for (int i = 0; i < n; i++) sc->order[i] = i;
    if (n > 0) {
        std::vector<float> centre[3];
        for (int a = 0; a < 3; a++) {
            centre[a].resize(n);
            for (int i = 0; i < n; i++) centre[a][i] = (sc->solids[i].box.lo[a] + sc->solids[i].box.hi[a]) / 2.0f;
        }
        sc->nodes.reserve(n);
        build(*sc, centre, 0, n);
    }
    *out = sc;
    return 0;
}

void orc_scene_destroy(orc_scene* s) { delete s; }

int orc_scene_shape_order(const orc_scene* s, int32_t* out, int32_t n) {
    if (!s || !out || n != (int32_t)s->order.size()) return -1;
    for (int i = 0; i < n; i++) out[i] = s->order[i];
    return 0;
}

int orc_scene_dump_bvh(const orc_scene* s, orc_node_dump* out, int32_t max_nodes) {
    if (!s) return -1;
    int n = 0;
    for (const TreeNode& t : s->nodes) {
        if (n >= max_nodes) break;
        orc_node_dump& d = out[n++];
        std::memset(&d, 0, sizeof(d));
        d.is_leaf = t.left < 0;
        for (int i = 0; i < 3; i++) { d.lo[i] = t.box.lo[i]; d.hi[i] = t.box.hi[i]; }
        if (d.is_leaf) { d.count = t.count; for (int k = 0; k < t.count && k < 4; k++) d.prims[k] = s->order[t.first + k]; }
    }
    return n;
}

int orc_render(const orc_scene* s, const orc_params* p, uint8_t* rgb8, int32_t* hit_ids, float* hit_t, float* linear,
               uint64_t* rays) {
    if (!s || !p) return -1;
    const int W = s->cam.res_x, H = s->cam.res_y;
    if (W <= 0 || H <= 0) return -4;
    const int row0 = std::max(0, p->row0), row1 = (p->row1 <= 0 || p->row1 > H) ? H : p->row1;
    const int col0 = std::max(0, p->col0), col1 = (p->col1 <= 0 || p->col1 > W) ? W : p->col1;
    const int wcols = std::max(0, col1 - col0);
    const long long n_window = (long long)std::max(0, row1 - row0) * wcols;
    const int spp = p->samples_sqrt <= 1 ? 1 : p->samples_sqrt * p->samples_sqrt;
    const int nthreads = std::max(1, p->threads);
    std::vector<uint64_t> counts((size_t)nthreads * 3, 0);
    auto work = [&](int tid) {
        Tracer tr{*s, *p};
        // pixels of the window are dealt to the threads in runs of 8 (a band of a few rows still uses every core)
        for (long long run = tid; run * 8 < n_window; run += nthreads) {
            for (long long wi = run * 8; wi < std::min(n_window, run * 8 + 8); wi++) {
                const int y = row0 + (int)(wi / wcols), x = col0 + (int)(wi % wcols);
                Draws g = {(uint32_t)(y * W + x), (uint32_t)(p->seed & 0xffffffffu), (uint32_t)(p->seed >> 32), 0u};
                Rgb sum = {0, 0, 0};
                int shape0 = -1;
                float t0 = FLT_MAX;
                for (int k = 0; k < spp; k++) {
                    g.sample = (uint32_t)k;
                    Beam ray = tr.camera_ray(x, y, k, g);
                    int sh; float tt;
                    Rgb c = tr.trace(ray, 0, g, 1u, &sh, &tt);
                    if (k == 0) { shape0 = sh; t0 = tt; }
                    sum = {sum.r + c.r, sum.g + c.g, sum.b + c.b};
                }
                if (p->samples_sqrt > 1) { float n = (float)spp; sum = {sum.r / n, sum.g / n, sum.b / n}; }
                size_t idx = (size_t)y * W + x;
                if (hit_ids) hit_ids[idx] = shape0;
                if (hit_t) hit_t[idx] = t0;
                if (linear) { linear[3 * idx] = sum.r; linear[3 * idx + 1] = sum.g; linear[3 * idx + 2] = sum.b; }
                if (rgb8) {  // raytracer.cpp:446-457
                    float gamma = 1.1f;
                    float ch[3] = {std::pow(sum.r, 1.0f / gamma), std::pow(sum.g, 1.0f / gamma), std::pow(sum.b, 1.0f / gamma)};
                    for (int q = 0; q < 3; q++) {
                        int v = static_cast<int>(std::max(0.0f, std::min(1.0f, ch[q])) * 255.999);
                        rgb8[3 * idx + q] = (uint8_t)std::max(0, std::min(v, 255));
                    }
                }
            }
        }
        counts[3 * tid] = tr.n_primary; counts[3 * tid + 1] = tr.n_shadow; counts[3 * tid + 2] = tr.n_secondary;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    if (rays) {
        rays[0] = rays[1] = rays[2] = 0;
        for (int t = 0; t < nthreads; t++) for (int k = 0; k < 3; k++) rays[k] += counts[3 * t + k];
    }
    return 0;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox(ctr, key[0], key[1], out); }

}  // extern "C"
