// render.cu -- the hot path on the GPU (sm_100a): per-pixel ray generation, BVH traversal,
// ray-primitive intersection, Blinn-Phong shading with shadow / reflection / refraction rays.
//
// Replaces the reference frame loop raytracer.cpp:433-476 and everything it calls:
//   compute_pixel_color  raytracer.cpp:18-70      -> camera_ray() + render_kernel sample loop
//   Camera::pixelToRay_thin_lens camera.cpp:97-178 -> camera_ray()
//   Trace                raytracer.cpp:280-351    -> trace_tree() (explicit stack, same DFS order)
//   shade                raytracer.cpp:180-274    -> shade()
//   BVH::get_intersection acceleration.cpp:67-150 -> closest_hit() / occluded()
//   AABB::intersect      shapes.cpp:55-72         -> box_test()
//   Sphere/Cube/Rectangle/Plane::intersect shapes.cpp:200-262,299-333,355-423,444-483 -> intersect_prim()
//
// Arithmetic contract: this file is compiled with -fmad=false and the default IEEE division and
// square root, and every expression is written in the reference's operand order, so each float
// the reference computes on x86-64/SSE2 is reproduced bit for bit (transcendentals -- powf,
// atan2, asin -- excepted; they only touch shading and texture coordinates). That is what makes
// primary-ray hit IDs match exactly.
//
// Traversal semantics: the reference visits every node whose box the ray passes, collects ALL
// leaf hits and returns the one with the smallest t (first in leaf order on ties). Ancestor boxes
// contain leaf boxes and IEEE rounding is monotonic, so "shape is tested" == "its leaf's box
// passes the reference box test". We therefore may visit near-first and skip a sub-tree whose
// entry distance is beyond the best hit (with a safety margin far above rounding noise) and
// still return the identical (t, shape) pair.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "api_internal.hpp"
#include "philox.cuh"

namespace rtb {

// ---------------------------------------------------------------------------------------------
// Kernel parameters
// ---------------------------------------------------------------------------------------------
struct KParams {
    const float4* __restrict__ prims;   // 8 x float4 per primitive, sorted (leaf) order
    const float4* __restrict__ nodes;   // 4 x float4 per internal node
    const float4* __restrict__ mats;    // 4 x float4 per material
    const float4* __restrict__ lights;  // 2 x float4 per light
    const DTexture* __restrict__ textures;
    const uint8_t* __restrict__ texels;
    int n_prims, n_lights;
    int root_ref;
    float root_lo[3], root_hi[3];
    int use_bvh;
    int prune;  // 1: near-first with t-pruning (default), 0: visit everything like the reference
    // camera
    float cam_loc[3], xdir[3], ydir[3], zdir[3];
    float focal, half_sw, half_sh, aperture, focus_dist;
    int res_x, res_y;
    // sampling
    int samples_sqrt, spp, light_samples, max_depth;
    uint32_t seed_lo, seed_hi;
    float fixed_time;
    // work decomposition
    int tile_w, tile_h, tiles_x, n_tiles, rank, world, n_my_tiles;
    int sub_x, sub_per_tile;       // 8x4 sub-tiles per screen tile
    int chunk_samples, n_chunks;   // samples are split into chunks; one work item = (sub-tile, chunk)
    int n_items;
    // outputs
    float4* partial;               // [n_chunks][res_y * res_x] partial sums
    int* hit_ids;                  // may be null
    unsigned long long* counters;  // [8]
    unsigned int* work_counter;
    int collect_stats;
};

enum Counter { C_PRIMARY = 0, C_SHADOW = 1, C_SECONDARY = 2, C_NODES = 3, C_PRIMS = 4 };

struct Ray {
    float ox, oy, oz;
    float dx, dy, dz;
    float time;
};

struct Hit {
    float t;
    int prim;  // sorted index, -1 = miss
    float px, py, pz;
    float nx, ny, nz;
    float u, v;
};

#define RT_DEV __device__ __forceinline__

RT_DEV float dot3(float ax, float ay, float az, float bx, float by, float bz) { return ax * bx + ay * by + az * bz; }

// VecMath::normalize (raytracer.cpp:75-79) / Camera::normalize (camera.cpp:60-68)
RT_DEV void normalize3(float& x, float& y, float& z) {
    const float mag = sqrtf(x * x + y * y + z * z);
    if (mag == 0.0f) { x = 0.0f; y = 0.0f; z = 0.0f; return; }
    x = x / mag; y = y / mag; z = z / mag;
}

// ---------------------------------------------------------------------------------------------
// AABB::intersect (shapes.cpp:55-72). `fabs(d) < 1e-6` there is a DOUBLE comparison of a float
// against 1e-6; the largest float below 1e-6 is 1e-6f itself, hence `<=` here.
// Returns the entry distance in tnear (may be negative when the origin is inside).
// ---------------------------------------------------------------------------------------------
RT_DEV bool box_test(float lox, float loy, float loz, float hix, float hiy, float hiz, const Ray& r, float& tnear) {
    float tn = -FLT_MAX, tf = FLT_MAX;
    if (fabsf(r.dx) <= 1e-6f) {
        if (r.ox < lox || r.ox > hix) return false;
    } else {
        float t1 = (lox - r.ox) / r.dx, t2 = (hix - r.ox) / r.dx;
        tn = fmaxf(tn, fminf(t1, t2));
        tf = fminf(tf, fmaxf(t1, t2));
    }
    if (fabsf(r.dy) <= 1e-6f) {
        if (r.oy < loy || r.oy > hiy) return false;
    } else {
        float t1 = (loy - r.oy) / r.dy, t2 = (hiy - r.oy) / r.dy;
        tn = fmaxf(tn, fminf(t1, t2));
        tf = fminf(tf, fmaxf(t1, t2));
    }
    if (fabsf(r.dz) <= 1e-6f) {
        if (r.oz < loz || r.oz > hiz) return false;
    } else {
        float t1 = (loz - r.oz) / r.dz, t2 = (hiz - r.oz) / r.dz;
        tn = fmaxf(tn, fminf(t1, t2));
        tf = fminf(tf, fmaxf(t1, t2));
    }
    tnear = tn;
    return !(tn > tf || tf < 0.0f);
}

// Shapes::transformPoint with w == 1 (shapes.cpp:151-158) and transformVector (:160-165)
RT_DEV void xform_point(const float4 r0, const float4 r1, const float4 r2, float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = r0.x * x + r0.y * y + r0.z * z + r0.w;
    oy = r1.x * x + r1.y * y + r1.z * z + r1.w;
    oz = r2.x * x + r2.y * y + r2.z * z + r2.w;
}
RT_DEV void xform_vector(const float4 r0, const float4 r1, const float4 r2, float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = r0.x * x + r0.y * y + r0.z * z;
    oy = r1.x * x + r1.y * y + r1.z * z;
    oz = r2.x * x + r2.y * y + r2.z * z;
}
// Shapes::transformNormal (shapes.cpp:167-187): world_to_object transposed, then normalise.
RT_DEV void xform_normal(const float4 r0, const float4 r1, const float4 r2, float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = r0.x * x + r1.x * y + r2.x * z;
    oy = r0.y * x + r1.y * y + r2.y * z;
    oz = r0.z * x + r1.z * y + r2.z * z;
    const float len = sqrtf(ox * ox + oy * oy + oz * oz);
    if (len > 1e-6f) { ox /= len; oy /= len; oz /= len; }
}

// isPointInTriangle (shapes.cpp:24-40)
RT_DEV bool point_in_triangle(float px, float py, float pz, float ax, float ay, float az, float bx, float by, float bz,
                              float cx, float cy, float cz, float nx, float ny, float nz) {
    {
        const float ex = bx - ax, ey = by - ay, ez = bz - az;
        const float vx = px - ax, vy = py - ay, vz = pz - az;
        const float kx = ey * vz - ez * vy, ky = ez * vx - ex * vz, kz = ex * vy - ey * vx;
        if (dot3(kx, ky, kz, nx, ny, nz) < -1e-6f) return false;
    }
    {
        const float ex = cx - bx, ey = cy - by, ez = cz - bz;
        const float vx = px - bx, vy = py - by, vz = pz - bz;
        const float kx = ey * vz - ez * vy, ky = ez * vx - ex * vz, kz = ex * vy - ey * vx;
        if (dot3(kx, ky, kz, nx, ny, nz) < -1e-6f) return false;
    }
    {
        const float ex = ax - cx, ey = ay - cy, ez = az - cz;
        const float vx = px - cx, vy = py - cy, vz = pz - cz;
        const float kx = ey * vz - ez * vy, ky = ez * vx - ex * vz, kz = ex * vy - ey * vx;
        if (dot3(kx, ky, kz, nx, ny, nz) < -1e-6f) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Ray-primitive intersection. FULL = false: only t (what get_intersection compares);
// FULL = true: also point, normal and uv for shading. Both paths compute t with the same
// operations, so the winner's t recomputed in FULL mode is the same value.
// ---------------------------------------------------------------------------------------------
template <bool FULL>
RT_DEV bool intersect_prim(const float4* __restrict__ prims, int idx, const Ray& r, Hit& h) {
    const float4* q = prims + (size_t)idx * 8;
    const float4 q0 = __ldg(q + 0);
    const float4 q1 = __ldg(q + 1);
    const float4 q2 = __ldg(q + 2);
    const float4 q3 = __ldg(q + 3);
    const int type = (int)(__float_as_uint(q0.w) & 3u);

    if (type == RT_PLANE) {
        // Plane::intersect (shapes.cpp:444-483); q1..q3 = corners 0..2 (+ corner 3 in .w), q4 = normal
        const float4 q4 = __ldg(q + 4);
        if (q4.w == 0.0f) return false;  // |cross| < 1e-6
        const float nx = q4.x, ny = q4.y, nz = q4.z;
        const float denom = dot3(nx, ny, nz, r.dx, r.dy, r.dz);
        if (fabsf(denom) < 1e-6f) return false;
        const float t = dot3(q1.x - r.ox, q1.y - r.oy, q1.z - r.oz, nx, ny, nz) / denom;
        if (t < 0.0f) return false;
        const float px = r.ox + t * r.dx, py = r.oy + t * r.dy, pz = r.oz + t * r.dz;
        const float c3x = q1.w, c3y = q2.w, c3z = q3.w;
        // isPointInQuad (shapes.cpp:485-494): triangles (c1,c3,c2) then (c0,c1,c2)
        if (!point_in_triangle(px, py, pz, q2.x, q2.y, q2.z, c3x, c3y, c3z, q3.x, q3.y, q3.z, nx, ny, nz) &&
            !point_in_triangle(px, py, pz, q1.x, q1.y, q1.z, q2.x, q2.y, q2.z, q3.x, q3.y, q3.z, nx, ny, nz))
            return false;
        h.t = t;
        if (FULL) {
            const float ux = q2.x - q1.x, uy = q2.y - q1.y, uz = q2.z - q1.z;
            const float vx = c3x - q1.x, vy = c3y - q1.y, vz = c3z - q1.z;
            const float hx = px - q1.x, hy = py - q1.y, hz = pz - q1.z;
            const float u = dot3(hx, hy, hz, ux, uy, uz) / dot3(ux, uy, uz, ux, uy, uz);
            const float v = dot3(hx, hy, hz, vx, vy, vz) / dot3(vx, vy, vz, vx, vy, vz);
            h.u = fmaxf(0.0f, fminf(1.0f, u));
            h.v = fmaxf(0.0f, fminf(1.0f, v));
            h.px = px; h.py = py; h.pz = pz;
            h.nx = nx; h.ny = ny; h.nz = nz;
        }
        return true;
    }

    // Transformed shapes: ray to object space (q1..q3 = world_to_object rows).
    float mox = r.ox, moy = r.oy, moz = r.oz;
    if (type == RT_SPHERE) {  // motion blur: shift the origin back (shapes.cpp:203-209)
        mox = r.ox - q0.x * r.time;
        moy = r.oy - q0.y * r.time;
        moz = r.oz - q0.z * r.time;
    }
    float lox, loy, loz, ldx, ldy, ldz;
    xform_point(q1, q2, q3, mox, moy, moz, lox, loy, loz);
    xform_vector(q1, q2, q3, r.dx, r.dy, r.dz, ldx, ldy, ldz);

    float plx, ply, plz;       // local hit point
    float nlx, nly, nlz;       // local normal
    float u = 0.0f, v = 0.0f;

    if (type == RT_SPHERE) {
        // Sphere::intersect (shapes.cpp:200-262)
        const float a = dot3(ldx, ldy, ldz, ldx, ldy, ldz);
        const float b = 2.0f * dot3(lox, loy, loz, ldx, ldy, ldz);
        const float c = dot3(lox, loy, loz, lox, loy, loz) - 1.0f;
        const float disc = b * b - 4.0f * a * c;
        if (disc < 0.0f) return false;
        const float sq = sqrtf(disc);
        const float t1 = (-b - sq) / (2.0f * a);
        const float t2 = (-b + sq) / (2.0f * a);
        const float tl = (t1 > 0.001f) ? t1 : ((t2 > 0.001f) ? t2 : -1.0f);
        if (tl < 0.0f) return false;
        plx = lox + tl * ldx; ply = loy + tl * ldy; plz = loz + tl * ldz;
        nlx = plx; nly = ply; nlz = plz;
        if (FULL) {
            // the reference evaluates these in double (atan2/asin on floats, shapes.cpp:257-259)
            const float PI = 3.1415926535f;
            u = (float)((double)0.5f + atan2((double)nlz, (double)nlx) / (double)(2.0f * PI));
            v = (float)((double)0.5f - asin((double)nly) / (double)PI);
        }
    } else if (type == RT_RECTANGLE) {
        // Rectangle::intersect (shapes.cpp:299-333)
        if (fabsf(ldz) < 1e-6f) return false;
        const float tl = -loz / ldz;
        if (tl < 0.001f) return false;
        const float hx = lox + tl * ldx;
        const float hy = loy + tl * ldy;
        if (hx < -0.5f || hx > 0.5f || hy < -0.5f || hy > 0.5f) return false;
        plx = hx; ply = hy; plz = 0.0f;
        nlx = 0.0f; nly = 0.0f; nlz = 1.0f;
        u = hx + 0.5f; v = hy + 0.5f;
    } else {
        // Cube::intersect (shapes.cpp:355-423): slabs on [-0.5,0.5]^3, remembering the entry face
        float tn = -FLT_MAX, tf = FLT_MAX;
        int axis = -1, sign = 0;
        const float lo3[3] = {lox, loy, loz};
        const float ld3[3] = {ldx, ldy, ldz};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (fabsf(ld3[i]) < 1e-6f) {
                if (lo3[i] < -0.5f || lo3[i] > 0.5f) return false;
            } else {
                const float t1 = (-0.5f - lo3[i]) / ld3[i];
                const float t2 = (0.5f - lo3[i]) / ld3[i];
                const float te = fminf(t1, t2), tx = fmaxf(t1, t2);
                if (te > tn) { tn = te; axis = i; sign = (t1 < t2) ? -1 : 1; }
                if (tx < tf) tf = tx;
                if (tn > tf || tf < 0.0f) return false;
            }
        }
        const float tl = (tn > 0.0f) ? tn : tf;
        if (tl < 0.0f) return false;
        plx = lox + tl * ldx; ply = loy + tl * ldy; plz = loz + tl * ldz;
        nlx = (axis == 0) ? (float)sign : 0.0f;
        nly = (axis == 1) ? (float)sign : 0.0f;
        nlz = (axis == 2) ? (float)sign : 0.0f;
        if (FULL) {
            const float uc = plx + 0.5f, vc = ply + 0.5f, wc = plz + 0.5f;
            if (axis == 0) { u = (sign > 0) ? wc : (1.0f - wc); v = vc; }
            else if (axis == 1) { u = uc; v = (sign > 0) ? wc : (1.0f - wc); }
            else { u = (sign > 0) ? uc : (1.0f - uc); v = vc; }
        }
    }

    // back to world space (q4..q6 = object_to_world rows)
    const float4 q4 = __ldg(q + 4);
    const float4 q5 = __ldg(q + 5);
    const float4 q6 = __ldg(q + 6);
    float wx, wy, wz;
    xform_point(q4, q5, q6, plx, ply, plz, wx, wy, wz);
    if (type == RT_SPHERE) {
        wx += q0.x * r.time; wy += q0.y * r.time; wz += q0.z * r.time;
    }
    const float ex = wx - r.ox, ey = wy - r.oy, ez = wz - r.oz;
    h.t = sqrtf(dot3(ex, ey, ez, ex, ey, ez));
    if (FULL) {
        h.px = wx; h.py = wy; h.pz = wz;
        xform_normal(q1, q2, q3, nlx, nly, nlz, h.nx, h.ny, h.nz);
        h.u = u; h.v = v;
    }
    return true;
}

struct TraceStats {
    unsigned int nodes, prims;
};

// Safety margin for near-first pruning: a sub-tree is skipped only when the ray enters its box
// farther than best_t * (1 + 1e-4) + 1e-4. Hit distances are recomputed as |P - O| and can differ
// from the ray parameter by rounding (~1e-7 relative), four orders below this margin.
RT_DEV float prune_limit(float best_t) { return best_t * 1.0001f + 1e-4f; }

#define RT_STACK 48

// BVH::get_intersection (acceleration.cpp:142-150): closest hit, ties -> first in leaf order.
template <bool STATS>
RT_DEV void closest_hit(const KParams& p, const Ray& r, float& best_t, int& best_prim, TraceStats& st) {
    best_t = FLT_MAX;
    best_prim = -1;
    if (p.n_prims == 0) return;
    if (!p.use_bvh) {  // BVH::intersect_linear (acceleration.cpp:123-138)
        for (int i = 0; i < p.n_prims; ++i) {
            Hit h;
            if (STATS) st.prims++;
            if (intersect_prim<false>(p.prims, i, r, h) && h.t < best_t) { best_t = h.t; best_prim = i; }
        }
        return;
    }
    float tn;
    if (STATS) st.nodes++;
    if (!box_test(p.root_lo[0], p.root_lo[1], p.root_lo[2], p.root_hi[0], p.root_hi[1], p.root_hi[2], r, tn)) return;

    int stack[RT_STACK];
    int sp = 0;
    int cur = p.root_ref;
    while (true) {
        if (cur >= 0) {
            const float4* n = p.nodes + (size_t)cur * 4;
            const float4 a = __ldg(n + 0), b = __ldg(n + 1), c = __ldg(n + 2), d = __ldg(n + 3);
            float tl, tr;
            if (STATS) st.nodes += 2;
            bool hl = box_test(a.x, a.y, a.z, a.w, b.x, b.y, r, tl);
            bool hr = box_test(b.z, b.w, c.x, c.y, c.z, c.w, r, tr);
            if (p.prune) {
                const float lim = prune_limit(best_t);
                hl = hl && !(tl > lim);
                hr = hr && !(tr > lim);
            }
            const int li = __float_as_int(d.x), ri = __float_as_int(d.y);
            if (hl && hr) {
                const bool left_first = !p.prune || tl <= tr;
                stack[sp++] = left_first ? ri : li;
                cur = left_first ? li : ri;
                continue;
            } else if (hl) { cur = li; continue; }
            else if (hr) { cur = ri; continue; }
        } else {
            const int code = ~cur;
            const int first = code >> 3, count = code & 7;
            for (int k = 0; k < count; ++k) {
                Hit h;
                if (STATS) st.prims++;
                const int idx = first + k;
                if (intersect_prim<false>(p.prims, idx, r, h)) {
                    if (h.t < best_t || (h.t == best_t && idx < best_prim)) { best_t = h.t; best_prim = idx; }
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
}

// Shadow query of shade() (raytracer.cpp:230-235): occluded iff the closest hit has t <= max_t,
// i.e. iff ANY tested shape has t <= max_t.
template <bool STATS>
RT_DEV bool occluded(const KParams& p, const Ray& r, float max_t, TraceStats& st) {
    if (p.n_prims == 0) return false;
    if (!p.use_bvh) {
        for (int i = 0; i < p.n_prims; ++i) {
            Hit h;
            if (STATS) st.prims++;
            if (intersect_prim<false>(p.prims, i, r, h) && !(h.t > max_t)) return true;
        }
        return false;
    }
    float tn;
    if (STATS) st.nodes++;
    if (!box_test(p.root_lo[0], p.root_lo[1], p.root_lo[2], p.root_hi[0], p.root_hi[1], p.root_hi[2], r, tn)) return false;
    const float lim = prune_limit(max_t);
    int stack[RT_STACK];
    int sp = 0;
    int cur = p.root_ref;
    while (true) {
        if (cur >= 0) {
            const float4* n = p.nodes + (size_t)cur * 4;
            const float4 a = __ldg(n + 0), b = __ldg(n + 1), c = __ldg(n + 2), d = __ldg(n + 3);
            float tl, tr;
            if (STATS) st.nodes += 2;
            bool hl = box_test(a.x, a.y, a.z, a.w, b.x, b.y, r, tl);
            bool hr = box_test(b.z, b.w, c.x, c.y, c.z, c.w, r, tr);
            if (p.prune) { hl = hl && !(tl > lim); hr = hr && !(tr > lim); }
            const int li = __float_as_int(d.x), ri = __float_as_int(d.y);
            if (hl && hr) { stack[sp++] = ri; cur = li; continue; }
            else if (hl) { cur = li; continue; }
            else if (hr) { cur = ri; continue; }
        } else {
            const int code = ~cur;
            const int first = code >> 3, count = code & 7;
            for (int k = 0; k < count; ++k) {
                Hit h;
                if (STATS) st.prims++;
                if (intersect_prim<false>(p.prims, first + k, r, h) && !(h.t > max_t)) return true;
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// RNG-driven sampling (distributions of the reference)
// ---------------------------------------------------------------------------------------------
struct RngCtx {
    uint32_t pixel, seed_lo, seed_hi, sample;
};

// VecMath::random_in_unit_sphere (raytracer.cpp:152-171): rejection sampling in [-1,1]^3
RT_DEV void random_in_unit_sphere(const RngCtx& g, uint32_t purpose, uint32_t node, uint32_t sub, float& x, float& y, float& z) {
    for (uint32_t attempt = 0;; ++attempt) {
        const U4 u = rt_rng(g.pixel, g.seed_lo, g.seed_hi, g.sample, purpose, node, sub, attempt);
        x = 2.0f * u32_to_unit_float(u.x) - 1.0f;
        y = 2.0f * u32_to_unit_float(u.y) - 1.0f;
        z = 2.0f * u32_to_unit_float(u.z) - 1.0f;
        if (dot3(x, y, z, x, y, z) < 1.0f || attempt >= 63u) return;
    }
}

// compute_pixel_color + Camera::pixelToRay_thin_lens
RT_DEV Ray camera_ray(const KParams& p, int x, int y, int s, const RngCtx& g) {
    const U4 u = rt_rng(g.pixel, g.seed_lo, g.seed_hi, g.sample, RNG_CAMERA, 0u, 0u, 0u);
    float fx, fy;
    if (p.samples_sqrt <= 1) {
        fx = (float)x + 0.5f;  // raytracer.cpp:33
        fy = (float)y + 0.5f;
    } else {
        // stratified jitter in double, then narrowed to float by the tuple<float,float> (raytracer.cpp:50-58)
        const int i = s % p.samples_sqrt, j = s / p.samples_sqrt;
        const double sx = ((double)i + u32_to_unit_double(u.x)) / (double)p.samples_sqrt;
        const double sy = ((double)j + u32_to_unit_double(u.y)) / (double)p.samples_sqrt;
        fx = (float)((double)x + sx);
        fy = (float)((double)y + sy);
    }
    const float nx = 1.0f - (fx / (float)p.res_x) * 2.0f;
    const float ny = 1.0f - (fy / (float)p.res_y) * 2.0f;
    const float nxr = nx * p.half_sw;
    const float nyr = ny * p.half_sh;
    float dx = p.xdir[0] * nxr + p.ydir[0] * nyr + p.zdir[0] * p.focal;
    float dy = p.xdir[1] * nxr + p.ydir[1] * nyr + p.zdir[1] * p.focal;
    float dz = p.xdir[2] * nxr + p.ydir[2] * nyr + p.zdir[2] * p.focal;
    normalize3(dx, dy, dz);
    Ray r;
    r.ox = p.cam_loc[0]; r.oy = p.cam_loc[1]; r.oz = p.cam_loc[2];
    r.dx = dx; r.dy = dy; r.dz = dz;
    if (p.aperture > 0.0f) {  // thin lens (camera.cpp:141-177)
        const float fpx = p.cam_loc[0] + dx * p.focus_dist;
        const float fpy = p.cam_loc[1] + dy * p.focus_dist;
        const float fpz = p.cam_loc[2] + dz * p.focus_dist;
        float rx = 0.0f, ry = 0.0f;
        for (uint32_t attempt = 0;; ++attempt) {  // random_in_unit_disk (camera.cpp:89-95)
            const U4 l = rt_rng(g.pixel, g.seed_lo, g.seed_hi, g.sample, RNG_LENS, 0u, 0u, attempt);
            rx = u32_to_unit_float(l.x) * 2.0f - 1.0f;
            ry = u32_to_unit_float(l.y) * 2.0f - 1.0f;
            if (rx * rx + ry * ry < 1.0f || attempt >= 63u) break;
        }
        const float lens_radius = p.aperture / 2.0f;
        rx *= lens_radius;
        ry *= lens_radius;
        const float offx = p.xdir[0] * rx + p.ydir[0] * ry;
        const float offy = p.xdir[1] * rx + p.ydir[1] * ry;
        const float offz = p.xdir[2] * rx + p.ydir[2] * ry;
        r.ox = p.cam_loc[0] + offx; r.oy = p.cam_loc[1] + offy; r.oz = p.cam_loc[2] + offz;
        r.dx = fpx - r.ox; r.dy = fpy - r.oy; r.dz = fpz - r.oz;
        normalize3(r.dx, r.dy, r.dz);
    }
    r.time = (p.fixed_time >= 0.0f) ? p.fixed_time : u32_to_unit_float(u.z);  // raytracer.cpp:37,61
    return r;
}

// Material::getDiffuseColor (material.hpp:99-134)
RT_DEV void diffuse_color(const KParams& p, const float4 m0, int tex, float u, float v, float& r, float& g, float& b) {
    r = m0.x; g = m0.y; b = m0.z;
    if (tex < 0) return;
    const DTexture t = p.textures[tex];
    const float fx = u * (float)(t.width - 1);
    const float fy = (1.0f - v) * (float)(t.height - 1);
    int tr = 0, tg = 0, tb = 0;
    if (fx == fx && fy == fy) {  // NaN uv -> out of bounds -> black, like Image::getPixel
        const int x = (int)fx, y = (int)fy;
        if (x >= 0 && x < t.width && y >= 0 && y < t.height) {
            const uint8_t* px = p.texels + t.offset + ((size_t)y * t.width + x) * 3;
            tr = px[0]; tg = px[1]; tb = px[2];
        }
    }
    r = ((float)tr / 255.0f) * m0.x;
    g = ((float)tg / 255.0f) * m0.y;
    b = ((float)tb / 255.0f) * m0.z;
}

struct Counts {
    unsigned int primary, shadow, secondary;
};

// shade() (raytracer.cpp:180-274)
template <bool STATS>
RT_DEV void shade(const KParams& p, const Hit& h, const Ray& view, const float4 m0, const float4 m1, const float4 m2,
                  int tex, const RngCtx& g, uint32_t node, float& out_r, float& out_g, float& out_b, Counts& cnt,
                  TraceStats& st) {
    float br, bg, bb;
    diffuse_color(p, m0, tex, h.u, h.v, br, bg, bb);
    const float ka = m0.w, kd = m1.w, ks = m2.x, shininess = m2.y;
    float fr = br * ka, fg = bg * ka, fb = bb * ka;

    float vx = view.ox - h.px, vy = view.oy - h.py, vz = view.oz - h.pz;
    normalize3(vx, vy, vz);

    for (int li = 0; li < p.n_lights; ++li) {
        const float4 l0 = __ldg(p.lights + 2 * li);
        const float4 l1 = __ldg(p.lights + 2 * li + 1);
        const float radius = l1.w;
        const int shadow_samples = (radius > 0.0f) ? p.light_samples : 1;
        float visibility = 0.0f;
        for (int s = 0; s < shadow_samples; ++s) {
            float tx = l0.x, ty = l0.y, tz = l0.z;
            if (radius > 0.0f) {
                float rx, ry, rz;
                random_in_unit_sphere(g, RNG_LIGHT, node, ((uint32_t)li << 16) | (uint32_t)s, rx, ry, rz);
                tx = tx + rx * radius; ty = ty + ry * radius; tz = tz + rz * radius;
            }
            float lx = tx - h.px, ly = ty - h.py, lz = tz - h.pz;
            const float light_dist = sqrtf(dot3(lx, ly, lz, lx, ly, lz));
            normalize3(lx, ly, lz);
            Ray sr;
            sr.ox = h.px + h.nx * 1e-4f; sr.oy = h.py + h.ny * 1e-4f; sr.oz = h.pz + h.nz * 1e-4f;
            sr.dx = lx; sr.dy = ly; sr.dz = lz;
            sr.time = 0.0f;  // `Ray shadowRay;` keeps the default time (shapes.hpp:28)
            cnt.shadow++;
            if (!occluded<STATS>(p, sr, light_dist, st)) visibility += 1.0f;
        }
        visibility /= (float)shadow_samples;
        if (visibility <= 0.0f) continue;

        float cx = l0.x - h.px, cy = l0.y - h.py, cz = l0.z - h.pz;
        const float dist_sq = dot3(cx, cy, cz, cx, cy, cz);
        const float light_distance = sqrtf(dist_sq);
        normalize3(cx, cy, cz);
        const float ndl = fmaxf(0.0f, dot3(h.nx, h.ny, h.nz, cx, cy, cz));
        float hx = cx + vx, hy = cy + vy, hz = cz + vz;
        normalize3(hx, hy, hz);
        const float ndh = fmaxf(0.0f, dot3(h.nx, h.ny, h.nz, hx, hy, hz));
        const float spec = powf(ndh, shininess);
        const float att = 10.0f * l0.w / (25.0f + 10.0f * light_distance + 150.0f * dist_sq);
        // light_color * (diffuse * k_diffuse + specular * k_specular) * attenuation, then * visibility
        const float cr = l1.x * ((br * ndl) * kd + (m1.x * spec) * ks) * att;
        const float cg = l1.y * ((bg * ndl) * kd + (m1.y * spec) * ks) * att;
        const float cb = l1.z * ((bb * ndl) * kd + (m1.z * spec) * ks) * att;
        fr = fr + cr * visibility;
        fg = fg + cg * visibility;
        fb = fb + cb * visibility;
    }
    out_r = fr; out_g = fg; out_b = fb;
}

struct Pending {
    float ox, oy, oz, dx, dy, dz;
    float weight;
    uint32_t node;  // ray-tree node id: root 1, reflection child 2k, refraction child 2k+1
    int depth;
};

#define RT_MAX_DEPTH 16

// Trace() (raytracer.cpp:280-351) with an explicit stack. Children are pushed refraction first,
// reflection on top, so nodes are visited in the recursion's order. Colour is accumulated as
// sum over nodes of (product of reflectivity/transparency along the path) * local term.
template <bool STATS>
RT_DEV void trace_tree(const KParams& p, const Ray& primary, const RngCtx& g, float& out_r, float& out_g, float& out_b,
                       int& primary_prim, Counts& cnt, TraceStats& st) {
    Pending stack[RT_MAX_DEPTH + 2];
    int sp = 0;
    float acc_r = 0.0f, acc_g = 0.0f, acc_b = 0.0f;
    Ray r = primary;
    float weight = 1.0f;
    uint32_t node = 1u;
    int depth = 0;
    primary_prim = -1;
    while (true) {
        float t;
        int prim;
        closest_hit<STATS>(p, r, t, prim, st);
        if (depth == 0) { primary_prim = prim; cnt.primary++; } else cnt.secondary++;
        if (prim < 0) {
            acc_r += weight * 0.1f; acc_g += weight * 0.1f; acc_b += weight * 0.1f;  // background (raytracer.cpp:297)
        } else {
            Hit h;
            intersect_prim<true>(p.prims, prim, r, h);
            h.prim = prim;
            const uint32_t tag = __float_as_uint(__ldg(p.prims + (size_t)prim * 8).w);
            const int mat = (int)(tag >> 2);
            const float4 m0 = __ldg(p.mats + 4 * mat + 0);
            const float4 m1 = __ldg(p.mats + 4 * mat + 1);
            const float4 m2 = __ldg(p.mats + 4 * mat + 2);
            const float4 m3 = __ldg(p.mats + 4 * mat + 3);
            const float roughness = m2.z, reflectivity = m2.w, transparency = m3.x, ior = m3.y;
            const int tex = __float_as_int(m3.z);
            float lr, lg, lb;
            shade<STATS>(p, h, r, m0, m1, m2, tex, g, node, lr, lg, lb, cnt, st);
            const float local_w = fmaxf(0.0f, 1.0f - reflectivity - transparency);
            acc_r += weight * (local_w * lr);
            acc_g += weight * (local_w * lg);
            acc_b += weight * (local_w * lb);

            const bool deeper = depth + 1 <= p.max_depth;
            // refraction (raytracer.cpp:336-344, createRefractionRay :118-150) -- pushed first
            if (transparency > 0.0f && deeper) {
                float nx = h.nx, ny = h.ny, nz = h.nz;
                float n_in = 1.0f, n_out = ior;
                const float cos_i = dot3(r.dx, r.dy, r.dz, nx, ny, nz);
                if (cos_i > 0.0f) { const float tmp = n_in; n_in = n_out; n_out = tmp; nx = nx * -1.0f; ny = ny * -1.0f; nz = nz * -1.0f; }
                const float eta = n_in / n_out;
                const float cos_abs = fabsf(cos_i);
                const float disc = 1.0f - eta * eta * (1.0f - cos_abs * cos_abs);
                if (!(disc < 0.0f)) {
                    const float cos_t = sqrtf(disc);
                    const float k = eta * cos_abs - cos_t;
                    float tx = r.dx * eta + nx * k, ty = r.dy * eta + ny * k, tz = r.dz * eta + nz * k;
                    normalize3(tx, ty, tz);
                    if (dot3(tx, ty, tz, tx, ty, tz) > 1e-6f) {
                        Pending& q = stack[sp++];
                        q.ox = h.px + nx * -1e-4f; q.oy = h.py + ny * -1e-4f; q.oz = h.pz + nz * -1e-4f;
                        q.dx = tx; q.dy = ty; q.dz = tz;
                        q.weight = weight * transparency;
                        q.node = node * 2u + 1u;
                        q.depth = depth + 1;
                    }
                }
            }
            // reflection (raytracer.cpp:308-333, createReflectionRay :101-115)
            if (reflectivity > 0.0f && deeper) {
                const float idn = dot3(r.dx, r.dy, r.dz, h.nx, h.ny, h.nz);
                const float k = 2.0f * idn;
                float rx = r.dx - h.nx * k, ry = r.dy - h.ny * k, rz = r.dz - h.nz * k;
                if (roughness > 0.0f) {
                    float fx, fy, fz;
                    random_in_unit_sphere(g, RNG_GLOSSY, node, 0u, fx, fy, fz);
                    rx = rx + fx * roughness; ry = ry + fy * roughness; rz = rz + fz * roughness;
                    normalize3(rx, ry, rz);
                    if (dot3(rx, ry, rz, h.nx, h.ny, h.nz) < 0.0f) { rx = 0.0f; ry = 0.0f; rz = 0.0f; }
                }
                if (dot3(rx, ry, rz, rx, ry, rz) > 0.001f) {
                    Pending& q = stack[sp++];
                    q.ox = h.px + h.nx * 1e-4f; q.oy = h.py + h.ny * 1e-4f; q.oz = h.pz + h.nz * 1e-4f;
                    q.dx = rx; q.dy = ry; q.dz = rz;
                    q.weight = weight * reflectivity;
                    q.node = node * 2u;
                    q.depth = depth + 1;
                }
            }
        }
        if (sp == 0) break;
        const Pending& q = stack[--sp];
        r.ox = q.ox; r.oy = q.oy; r.oz = q.oz; r.dx = q.dx; r.dy = q.dy; r.dz = q.dz;
        r.time = 0.0f;  // secondary rays are built as {origin, direction}: default time (shapes.hpp:28)
        weight = q.weight; node = q.node; depth = q.depth;
    }
    out_r = acc_r; out_g = acc_g; out_b = acc_b;
}

// ---------------------------------------------------------------------------------------------
// Persistent render kernel: every warp repeatedly fetches a work item = (8x4 pixel sub-tile,
// sample chunk) with one atomic per warp; a lane owns one pixel of the sub-tile and loops over
// the chunk's samples. Partial sums go to partial[chunk][pixel]; finalize_kernel adds the chunks
// in a fixed order, so the image does not depend on scheduling.
// ---------------------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(128) render_kernel(const __grid_constant__ KParams p) {
    const int lane = threadIdx.x & 31;
    Counts cnt = {0u, 0u, 0u};
    TraceStats st = {0u, 0u};
    while (true) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= (unsigned int)p.n_items) break;
        // item = ((my_tile * n_chunks) + chunk) * sub_per_tile + sub
        const int sub = (int)(item % (unsigned int)p.sub_per_tile);
        const int rest = (int)(item / (unsigned int)p.sub_per_tile);
        const int chunk = rest % p.n_chunks;
        const int my_tile = rest / p.n_chunks;
        const int tile = my_tile * p.world + p.rank;
        const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
        const int x = tx * p.tile_w + (sub % p.sub_x) * 8 + (lane & 7);
        const int y = ty * p.tile_h + (sub / p.sub_x) * 4 + (lane >> 3);
        const bool in_tile = (sub % p.sub_x) * 8 + (lane & 7) < p.tile_w && (sub / p.sub_x) * 4 + (lane >> 3) < p.tile_h;
        if (!in_tile || x >= p.res_x || y >= p.res_y) continue;
        const int s0 = chunk * p.chunk_samples;
        const int s1 = min(p.spp, s0 + p.chunk_samples);
        float ar = 0.0f, ag = 0.0f, ab = 0.0f;
        RngCtx g;
        g.pixel = (uint32_t)(y * p.res_x + x);
        g.seed_lo = p.seed_lo;
        g.seed_hi = p.seed_hi;
        for (int s = s0; s < s1; ++s) {
            g.sample = (uint32_t)s;
            const Ray r = camera_ray(p, x, y, s, g);
            float cr, cg, cb;
            int prim;
            trace_tree<STATS>(p, r, g, cr, cg, cb, prim, cnt, st);
            ar = ar + cr; ag = ag + cg; ab = ab + cb;
            if (s == 0 && p.hit_ids) {
                int id = -1;
                if (prim >= 0) id = __float_as_int(__ldg(p.prims + (size_t)prim * 8 + 7).x);
                p.hit_ids[(size_t)y * p.res_x + x] = id;
            }
        }
        p.partial[(size_t)chunk * p.res_x * p.res_y + (size_t)y * p.res_x + x] = make_float4(ar, ag, ab, 0.0f);
    }
    // ray counters: one atomic per warp and counter
    unsigned int v[5] = {cnt.primary, cnt.shadow, cnt.secondary, st.nodes, st.prims};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        unsigned long long w = v[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) w += __shfl_xor_sync(0xffffffffu, w, off);
        if (lane == 0 && w) atomicAdd(p.counters + k, w);
    }
}

// Sum the chunks, average, gamma 1.1, clamp, * 255.999 (raytracer.cpp:69, 446-457).
__global__ void finalize_kernel(const __grid_constant__ KParams p, uint8_t* rgb8, float* linear) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = p.res_x * p.res_y;
    if (idx >= n) return;
    const int x = idx % p.res_x, y = idx / p.res_x;
    const int tile = (y / p.tile_h) * p.tiles_x + (x / p.tile_w);
    if (tile % p.world != p.rank) return;
    float r = 0.0f, g = 0.0f, b = 0.0f;
    for (int c = 0; c < p.n_chunks; ++c) {
        const float4 v = p.partial[(size_t)c * n + idx];
        r = r + v.x; g = g + v.y; b = b + v.z;
    }
    if (p.samples_sqrt > 1) {
        const float total = (float)p.spp;
        r = r / total; g = g / total; b = b / total;
    }
    if (linear) { linear[3 * (size_t)idx + 0] = r; linear[3 * (size_t)idx + 1] = g; linear[3 * (size_t)idx + 2] = b; }
    if (rgb8) {
        const float inv_gamma = 1.0f / 1.1f;
        const float ch[3] = {powf(r, inv_gamma), powf(g, inv_gamma), powf(b, inv_gamma)};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // std::max(0.0f, std::min(1.0f, c)): NaN -> 1
            float c = ch[k];
            c = (c < 1.0f) ? c : 1.0f;
            c = (0.0f < c) ? c : 0.0f;
            int v = (int)((double)c * 255.999);
            v = max(0, min(v, 255));
            rgb8[3 * (size_t)idx + k] = (uint8_t)v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host side: device scene, launches
// ---------------------------------------------------------------------------------------------
struct DeviceScene {
    int device = -1;
    float4* prims = nullptr;
    float4* nodes = nullptr;
    float4* mats = nullptr;
    float4* lights = nullptr;
    DTexture* textures = nullptr;
    uint8_t* texels = nullptr;
    uint64_t bytes = 0;
    // scratch
    float4* partial = nullptr;
    size_t partial_elems = 0;
    unsigned long long* counters = nullptr;  // 8 counters + work counter at [8]
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int blocks_per_sm[2] = {0, 0};
    int sm_count = 0;
    bool timed = false;               // ev[0..2] have been recorded at least once
    std::vector<void*> pinned;        // host ranges registered for fast H2D staging
};

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
            return RT_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

template <typename T>
static int upload_vec(const std::vector<T>& v, void** dst, uint64_t& bytes, DeviceScene* d) {
    *dst = nullptr;
    const size_t n = std::max<size_t>(v.size(), 1) * sizeof(T);
    CUDA_TRY(cudaMalloc(dst, n));
    if (!v.empty()) {
        // page-lock large host arrays so the copy runs at full PCIe rate (best effort)
        if (v.size() * sizeof(T) >= (1u << 20)) {
            if (cudaHostRegister((void*)v.data(), v.size() * sizeof(T), cudaHostRegisterDefault) == cudaSuccess)
                d->pinned.push_back((void*)v.data());
            else
                cudaGetLastError();
        }
        CUDA_TRY(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        bytes += v.size() * sizeof(T);
    }
    return RT_OK;
}

static void free_device(DeviceScene* d) {
    if (!d) return;
    cudaFree(d->prims); cudaFree(d->nodes); cudaFree(d->mats); cudaFree(d->lights);
    cudaFree(d->textures); cudaFree(d->texels); cudaFree(d->partial); cudaFree(d->counters);
    for (cudaEvent_t e : d->ev) if (e) cudaEventDestroy(e);
    for (void* p : d->pinned) cudaHostUnregister(p);
    delete d;
}

void device_release(HostScene& h) {
    free_device(h.dev);
    h.dev = nullptr;
}

static int upload_all(HostScene& h, uint64_t* bytes_out);

static int ensure_uploaded(HostScene& h, uint64_t* bytes_out) {
    if (h.dev) { if (bytes_out) *bytes_out = 0; return RT_OK; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device: the renderer has no CPU fallback");
        return RT_ERR_CUDA;
    }
    const int rc = upload_all(h, bytes_out);
    if (rc != RT_OK) device_release(h);  // never keep a half-initialised device scene
    return rc;
}

static int upload_all(HostScene& h, uint64_t* bytes_out) {
    DeviceScene* d = new DeviceScene();
    h.dev = d;
    int rc;
    CUDA_TRY(cudaGetDevice(&d->device));
    if ((rc = upload_vec(h.dprims, (void**)&d->prims, d->bytes, d)) != RT_OK) return rc;
    if ((rc = upload_vec(h.dnodes, (void**)&d->nodes, d->bytes, d)) != RT_OK) return rc;
    if ((rc = upload_vec(h.dmaterials, (void**)&d->mats, d->bytes, d)) != RT_OK) return rc;
    if ((rc = upload_vec(h.dlights, (void**)&d->lights, d->bytes, d)) != RT_OK) return rc;
    if ((rc = upload_vec(h.dtextures, (void**)&d->textures, d->bytes, d)) != RT_OK) return rc;
    if ((rc = upload_vec(h.texels, (void**)&d->texels, d->bytes, d)) != RT_OK) return rc;
    CUDA_TRY(cudaMalloc((void**)&d->counters, 16 * sizeof(unsigned long long)));
    for (auto& e : d->ev) CUDA_TRY(cudaEventCreate(&e));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, d->device));
    d->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->blocks_per_sm[0], render_kernel<false>, 128, 0));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->blocks_per_sm[1], render_kernel<true>, 128, 0));
    if (bytes_out) *bytes_out = d->bytes;
    return RT_OK;
}

static int fill_params(const HostScene& h, const rt_render_params& rp, KParams& k) {
    if (h.cam.res_x <= 0 || h.cam.res_y <= 0) { set_error("Camera resolution is 0. Check scene.json."); return RT_ERR_SCENE; }
    if (rp.world < 1 || rp.rank < 0 || rp.rank >= rp.world) { set_error("rank/world out of range"); return RT_ERR_INVALID; }
    if (rp.tile_w < 8 || rp.tile_h < 4 || rp.tile_w % 8 || rp.tile_h % 4) { set_error("tile_w must be a multiple of 8 and tile_h of 4"); return RT_ERR_INVALID; }
    if (rp.max_depth < 0 || rp.max_depth > RT_MAX_DEPTH) { set_error("max_depth must be in [0,16]"); return RT_ERR_INVALID; }
    if (rp.light_samples < 1) { set_error("light_samples must be >= 1"); return RT_ERR_INVALID; }
    std::memset(&k, 0, sizeof(k));
    k.n_prims = (int)h.dprims.size();
    k.n_lights = (int)h.dlights.size();
    k.root_ref = h.root_ref;
    for (int i = 0; i < 3; ++i) {
        k.root_lo[i] = h.root_box.lo[i]; k.root_hi[i] = h.root_box.hi[i];
        k.cam_loc[i] = h.cam.location[i]; k.xdir[i] = h.xdir[i]; k.ydir[i] = h.ydir[i]; k.zdir[i] = h.zdir[i];
    }
    k.use_bvh = rp.use_bvh ? 1 : 0;
    k.prune = (rp.reserved[0] & 1) ? 0 : 1;  // reserved[0] bit 0: disable pruning (test hook)
    k.focal = h.cam.focal_length;
    k.half_sw = (float)h.cam.sensor_width / 2.0f;   // camera.cpp:106-107
    k.half_sh = (float)h.cam.sensor_height / 2.0f;
    k.aperture = h.cam.aperture;
    k.focus_dist = h.cam.focus_dist;
    k.res_x = h.cam.res_x;
    k.res_y = h.cam.res_y;
    k.samples_sqrt = rp.samples_sqrt;
    k.spp = rp.samples_sqrt <= 1 ? 1 : rp.samples_sqrt * rp.samples_sqrt;
    k.light_samples = rp.light_samples;
    k.max_depth = rp.max_depth;
    k.seed_lo = (uint32_t)(rp.seed & 0xffffffffu);
    k.seed_hi = (uint32_t)(rp.seed >> 32);
    k.fixed_time = rp.fixed_time;
    k.tile_w = rp.tile_w;
    k.tile_h = rp.tile_h;
    k.tiles_x = (k.res_x + k.tile_w - 1) / k.tile_w;
    const int tiles_y = (k.res_y + k.tile_h - 1) / k.tile_h;
    k.n_tiles = k.tiles_x * tiles_y;
    k.rank = rp.rank;
    k.world = rp.world;
    k.n_my_tiles = (k.n_tiles - rp.rank + rp.world - 1) / rp.world;
    k.sub_x = k.tile_w / 8;
    k.sub_per_tile = k.sub_x * (k.tile_h / 4);
    k.collect_stats = rp.collect_stats;
    return RT_OK;
}

static int64_t shard_pixels(const KParams& k) {
    int64_t n = 0;
    const int tiles_y = (k.res_y + k.tile_h - 1) / k.tile_h;
    for (int t = k.rank; t < k.n_tiles; t += k.world) {
        const int tx = t % k.tiles_x, ty = t / k.tiles_x;
        (void)tiles_y;
        const int w = std::min(k.tile_w, k.res_x - tx * k.tile_w);
        const int hh = std::min(k.tile_h, k.res_y - ty * k.tile_h);
        n += (int64_t)w * hh;
    }
    return n;
}

static int render_impl(HostScene& h, const rt_render_params& rp, uint8_t* rgb8, int32_t* hit_ids, float* linear,
                       cudaStream_t stream, rt_render_stats* stats) {
    int rc = ensure_uploaded(h, nullptr);
    if (rc != RT_OK) return rc;
    DeviceScene* d = h.dev;
    KParams k;
    if ((rc = fill_params(h, rp, k)) != RT_OK) return rc;
    k.prims = d->prims; k.nodes = d->nodes; k.mats = d->mats; k.lights = d->lights;
    k.textures = d->textures; k.texels = d->texels;
    k.hit_ids = hit_ids;
    k.counters = d->counters;
    k.work_counter = reinterpret_cast<unsigned int*>(d->counters + 8);

    const int stats_idx = rp.collect_stats ? 1 : 0;
    const int resident_warps = d->sm_count * std::max(1, d->blocks_per_sm[stats_idx]) * 4;
    // enough (sub-tile, chunk) items for >= 8 waves of resident warps
    const int64_t subtiles = (int64_t)k.n_my_tiles * k.sub_per_tile;
    int n_chunks = 1;
    if (subtiles > 0 && k.spp > 1) {
        const int64_t want = 8LL * resident_warps;
        n_chunks = (int)std::min<int64_t>(k.spp, std::max<int64_t>(1, (want + subtiles - 1) / subtiles));
    }
    k.chunk_samples = (k.spp + n_chunks - 1) / n_chunks;
    k.n_chunks = (k.spp + k.chunk_samples - 1) / k.chunk_samples;
    const int64_t items = subtiles * k.n_chunks;
    if (items >= (1LL << 31)) { set_error("too many work items"); return RT_ERR_INVALID; }
    k.n_items = (int)items;

    const size_t need = (size_t)k.n_chunks * k.res_x * k.res_y;
    if (need > d->partial_elems) {
        cudaFree(d->partial);
        d->partial = nullptr;
        d->partial_elems = 0;
        CUDA_TRY(cudaMalloc((void**)&d->partial, need * sizeof(float4)));
        d->partial_elems = need;
    }
    k.partial = d->partial;

    CUDA_TRY(cudaEventRecord(d->ev[0], stream));
    CUDA_TRY(cudaMemsetAsync(d->counters, 0, 16 * sizeof(unsigned long long), stream));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)d->sm_count * std::max(1, d->blocks_per_sm[stats_idx]), (items + 3) / 4));
    CUDA_TRY(cudaEventRecord(d->ev[1], stream));
    if (rp.collect_stats) render_kernel<true><<<grid, 128, 0, stream>>>(k);
    else render_kernel<false><<<grid, 128, 0, stream>>>(k);
    CUDA_TRY(cudaGetLastError());
    const int n_pix = k.res_x * k.res_y;
    finalize_kernel<<<(n_pix + 255) / 256, 256, 0, stream>>>(k, rgb8, linear);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(d->ev[2], stream));
    d->timed = true;

    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        CUDA_TRY(cudaEventSynchronize(d->ev[2]));
        unsigned long long c[8];
        CUDA_TRY(cudaMemcpy(c, d->counters, sizeof(c), cudaMemcpyDeviceToHost));
        stats->primary_rays = c[C_PRIMARY];
        stats->shadow_rays = c[C_SHADOW];
        stats->secondary_rays = c[C_SECONDARY];
        stats->rays = c[C_PRIMARY] + c[C_SHADOW] + c[C_SECONDARY];
        stats->node_visits = c[C_NODES];
        stats->prim_tests = c[C_PRIMS];
        CUDA_TRY(cudaEventElapsedTime(&stats->kernel_ms, d->ev[1], d->ev[2]));
        CUDA_TRY(cudaEventElapsedTime(&stats->total_ms, d->ev[0], d->ev[2]));
        stats->launches = 2;
        stats->pixels = (int32_t)shard_pixels(k);
    }
    return RT_OK;
}

}  // namespace rtb

// ---------------------------------------------------------------------------------------------
// extern "C" device entry points
// ---------------------------------------------------------------------------------------------
extern "C" {

int rt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rt_scene_upload(rt_scene* scene, uint64_t* bytes) {
    if (!scene) { rtb::set_error("null scene"); return RT_ERR_INVALID; }
    return rtb::ensure_uploaded(*rtb::host_of(scene), bytes);
}

int rt_scene_evict(rt_scene* scene) {
    if (!scene) { rtb::set_error("null scene"); return RT_ERR_INVALID; }
    rtb::device_release(*rtb::host_of(scene));
    return RT_OK;
}

int rt_scene_last_timing(rt_scene* scene, float* kernel_ms, float* total_ms) {
    if (!scene) { rtb::set_error("null scene"); return RT_ERR_INVALID; }
    rtb::DeviceScene* d = rtb::host_of(scene)->dev;
    if (!d || !d->timed) { rtb::set_error("rt_scene_last_timing: no render has been recorded on this scene"); return RT_ERR_INVALID; }
    cudaError_t e = cudaEventSynchronize(d->ev[2]);
    float k = 0.0f, t = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&k, d->ev[1], d->ev[2]);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, d->ev[0], d->ev[2]);
    if (e != cudaSuccess) { rtb::set_error(std::string("rt_scene_last_timing: ") + cudaGetErrorString(e)); return RT_ERR_CUDA; }
    if (kernel_ms) *kernel_ms = k;
    if (total_ms) *total_ms = t;
    return RT_OK;
}

int rt_shard_pixels(const rt_scene* scene, const rt_render_params* p, int64_t* n_pixels) {
    if (!scene || !p || !n_pixels) { rtb::set_error("null argument"); return RT_ERR_INVALID; }
    rtb::KParams k;
    int rc = rtb::fill_params(*rtb::host_of(scene), *p, k);
    if (rc != RT_OK) return rc;
    *n_pixels = rtb::shard_pixels(k);
    return RT_OK;
}

int rt_render_device(rt_scene* scene, const rt_render_params* p, uint8_t* rgb8, int32_t* hit_ids, float* linear,
                     void* stream, rt_render_stats* stats) {
    if (!scene || !p) { rtb::set_error("null argument"); return RT_ERR_INVALID; }
    return rtb::render_impl(*rtb::host_of(scene), *p, rgb8, hit_ids, linear, (cudaStream_t)stream, stats);
}

int rt_render(rt_scene* scene, const rt_render_params* p, uint8_t* rgb8, int32_t* hit_ids, float* linear,
              rt_render_stats* stats) {
    if (!scene || !p) { rtb::set_error("null argument"); return RT_ERR_INVALID; }
    rtb::HostScene& h = *rtb::host_of(scene);
    if (h.cam.res_x <= 0 || h.cam.res_y <= 0) { rtb::set_error("Camera resolution is 0. Check scene.json."); return RT_ERR_SCENE; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        rtb::set_error("no CUDA device: the renderer has no CPU fallback");
        return RT_ERR_CUDA;
    }
    const size_t n = (size_t)h.cam.res_x * h.cam.res_y;
    uint8_t* d_rgb = nullptr;
    int32_t* d_ids = nullptr;
    float* d_lin = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = RT_OK;
    auto fail = [&](cudaError_t e, const char* what) {
        rtb::set_error(std::string(what) + ": " + cudaGetErrorString(e));
        rc = RT_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess || (e = cudaEventCreate(&e1)) != cudaSuccess) fail(e, "cudaEventCreate");
    if (rc == RT_OK && (e = cudaEventRecord(e0, 0)) != cudaSuccess) fail(e, "cudaEventRecord");
    if (rc == RT_OK && rgb8 && (e = cudaMalloc((void**)&d_rgb, n * 3)) != cudaSuccess) fail(e, "cudaMalloc rgb");
    if (rc == RT_OK && hit_ids && (e = cudaMalloc((void**)&d_ids, n * sizeof(int32_t))) != cudaSuccess) fail(e, "cudaMalloc ids");
    if (rc == RT_OK && linear && (e = cudaMalloc((void**)&d_lin, n * 3 * sizeof(float))) != cudaSuccess) fail(e, "cudaMalloc linear");
    if (rc == RT_OK && p->world > 1) {
        // other ranks' pixels stay zero / -1 in the host buffers
        if (d_rgb) cudaMemset(d_rgb, 0, n * 3);
        if (d_ids) cudaMemset(d_ids, 0xff, n * sizeof(int32_t));
        if (d_lin) cudaMemset(d_lin, 0, n * 3 * sizeof(float));
    }
    rt_render_stats local;
    if (rc == RT_OK) rc = rtb::render_impl(h, *p, d_rgb, d_ids, d_lin, 0, &local);
    if (rc == RT_OK && d_rgb && (e = cudaMemcpy(rgb8, d_rgb, n * 3, cudaMemcpyDeviceToHost)) != cudaSuccess) fail(e, "cudaMemcpy rgb");
    if (rc == RT_OK && d_ids && (e = cudaMemcpy(hit_ids, d_ids, n * sizeof(int32_t), cudaMemcpyDeviceToHost)) != cudaSuccess) fail(e, "cudaMemcpy ids");
    if (rc == RT_OK && d_lin && (e = cudaMemcpy(linear, d_lin, n * 3 * sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess) fail(e, "cudaMemcpy linear");
    if (rc == RT_OK) {
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        local.total_ms = ms;
        if (stats) *stats = local;
    }
    cudaFree(d_rgb); cudaFree(d_ids); cudaFree(d_lin);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

}  // extern "C"
