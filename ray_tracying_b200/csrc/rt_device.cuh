// rt_device.cuh -- device-side building blocks: the reference's arithmetic (box test, the four
// ray-primitive intersections, sampling) and the two BVH queries (closest hit, occlusion).
//
// Arithmetic contract: the including .cu is compiled with -fmad=false and the default IEEE
// division / square root, and every expression here is written in the reference's operand order,
// so each float the reference computes on x86-64/SSE2 is reproduced bit for bit (transcendentals
// excepted). Explicit __fmaf_rn is used ONLY in the conservative box test, which never decides a
// result on its own (see box_maybe()).
//
// Traversal semantics: the reference (acceleration.cpp:67-117) visits every node whose box the
// ray passes, collects ALL leaf hits and returns the first minimum. Ancestor boxes contain leaf
// boxes and IEEE rounding is monotonic, so a shape is tested iff the box of ITS LEAF passes
// AABB::intersect. Therefore:
//   * internal boxes only need a CONSERVATIVE test (never rejects what the reference accepts):
//     6 FMAs with a precomputed reciprocal instead of 6 IEEE divisions;
//   * leaf boxes get the conservative test first and, if it passes, the EXACT reference test;
//   * visiting near-first and skipping sub-trees that start beyond the best hit (plus a margin
//     four orders above rounding noise) cannot change the (t, shape) result.
#pragma once

#include <cuda_runtime.h>
#include <cfloat>
#include <stdint.h>

#include "../../include/rt_render.h"
#include "philox.cuh"

namespace rtb {

#define RT_DEV __device__ __forceinline__

struct Ray {
    float ox, oy, oz;
    float dx, dy, dz;
    float time;
};

struct Hit {
    float t;
    float px, py, pz;
    float nx, ny, nz;
    float u, v;
};

RT_DEV float dot3(float ax, float ay, float az, float bx, float by, float bz) { return ax * bx + ay * by + az * bz; }

// VecMath::normalize (raytracer.cpp:75-79) / Camera::normalize (camera.cpp:60-68)
RT_DEV void normalize3(float& x, float& y, float& z) {
    const float mag = sqrtf(x * x + y * y + z * z);
    if (mag == 0.0f) { x = 0.0f; y = 0.0f; z = 0.0f; return; }
    x = x / mag; y = y / mag; z = z / mag;
}

// ---------------------------------------------------------------------------------------------
// Exact AABB::intersect (shapes.cpp:55-72). `fabs(d) < 1e-6` there is a DOUBLE comparison of a
// float against 1e-6; the largest float below 1e-6 is 1e-6f itself, hence `<=`.
// ---------------------------------------------------------------------------------------------
RT_DEV bool box_exact(float lox, float loy, float loz, float hix, float hiy, float hiz, const Ray& r, float& tnear) {
    float tn = -FLT_MAX, tf = FLT_MAX;
    if (fabsf(r.dx) <= 1e-6f) {
        if (r.ox < lox || r.ox > hix) return false;
    } else {
        const float t1 = (lox - r.ox) / r.dx, t2 = (hix - r.ox) / r.dx;
        tn = fmaxf(tn, fminf(t1, t2));
        tf = fminf(tf, fmaxf(t1, t2));
    }
    if (fabsf(r.dy) <= 1e-6f) {
        if (r.oy < loy || r.oy > hiy) return false;
    } else {
        const float t1 = (loy - r.oy) / r.dy, t2 = (hiy - r.oy) / r.dy;
        tn = fmaxf(tn, fminf(t1, t2));
        tf = fminf(tf, fmaxf(t1, t2));
    }
    if (fabsf(r.dz) <= 1e-6f) {
        if (r.oz < loz || r.oz > hiz) return false;
    } else {
        const float t1 = (loz - r.oz) / r.dz, t2 = (hiz - r.oz) / r.dz;
        tn = fmaxf(tn, fminf(t1, t2));
        tf = fminf(tf, fmaxf(t1, t2));
    }
    tnear = tn;
    return !(tn > tf || tf < 0.0f);
}

// ---------------------------------------------------------------------------------------------
// Conservative slab test. t' = fma(plane, 1/d, -o/d). Against the reference's RN(RN(plane-o)/d):
//     |t' - t| <= 4 eps |t| + 1.01 eps |o/d|      (eps = 2^-24)
// so accepting when tn' <= tf' + slack and tf' >= -slack with
//     slack = 1e-6 (|tn'| + |tf'|) + 4e-7 max_i |o_i/d_i| + 1e-30
// never rejects a box the reference accepts. Rays with a component |d_i| <= 1e-6 (the reference's
// "parallel" rule) do not use this path at all (RayAux::slow).
// ---------------------------------------------------------------------------------------------
struct RayAux {
    float ix, iy, iz;     // 1/d
    float nx, ny, nz;     // -o/d
    float k;              // absolute slack
    bool slow;            // some |d_i| <= 1e-6: use box_exact everywhere
};

RT_DEV RayAux make_aux(const Ray& r) {
    RayAux a;
    a.slow = fabsf(r.dx) <= 1e-6f || fabsf(r.dy) <= 1e-6f || fabsf(r.dz) <= 1e-6f;
    a.ix = 1.0f / r.dx; a.iy = 1.0f / r.dy; a.iz = 1.0f / r.dz;
    a.nx = -(r.ox * a.ix); a.ny = -(r.oy * a.iy); a.nz = -(r.oz * a.iz);
    a.k = 4e-7f * fmaxf(fmaxf(fabsf(a.nx), fabsf(a.ny)), fabsf(a.nz)) + 1e-30f;
    return a;
}

RT_DEV bool box_maybe(float lox, float loy, float loz, float hix, float hiy, float hiz, const RayAux& a, float& tnear) {
    const float x1 = __fmaf_rn(lox, a.ix, a.nx), x2 = __fmaf_rn(hix, a.ix, a.nx);
    const float y1 = __fmaf_rn(loy, a.iy, a.ny), y2 = __fmaf_rn(hiy, a.iy, a.ny);
    const float z1 = __fmaf_rn(loz, a.iz, a.nz), z2 = __fmaf_rn(hiz, a.iz, a.nz);
    const float tn = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fminf(z1, z2));
    const float tf = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fmaxf(z1, z2));
    const float slack = __fmaf_rn(1e-6f, fabsf(tn) + fabsf(tf), a.k);
    tnear = tn - slack;
    return tn <= tf + slack && tf >= -slack;
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
// Primitive record: see scene.hpp (8 x float4, the first 4 are enough for a miss).
// ---------------------------------------------------------------------------------------------
// KNOWN_TYPE >= 0: the caller knows the primitive's type at compile time (the traversal groups a
// warp's pending tests by type), so only that type's code is generated.
template <bool FULL, int KNOWN_TYPE = -1>
RT_DEV bool intersect_prim(const float4* __restrict__ prims, int idx, const Ray& r, Hit& h) {
    const float4* q = prims + (size_t)idx * 8;
    const float4 q0 = __ldg(q + 0);
    const float4 q1 = __ldg(q + 1);
    const float4 q2 = __ldg(q + 2);
    const float4 q3 = __ldg(q + 3);
    const int type = KNOWN_TYPE >= 0 ? KNOWN_TYPE : (int)(__float_as_uint(q0.w) & 3u);

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

    float plx, ply, plz;  // local hit point
    float nlx, nly, nlz;  // local normal
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

// ---------------------------------------------------------------------------------------------
// BVH queries
// ---------------------------------------------------------------------------------------------
struct BvhView {
    const float4* __restrict__ prims;
    const float4* __restrict__ nodes;
    const float4* __restrict__ leaves;  // 8 x float4 per leaf (scene.hpp DLeaf)
    int n_prims;
    int root_ref;
    float root_lo[3], root_hi[3];
    int use_bvh;
    int prune;     // 0: visit everything, exact tests only (the reference's literal traversal)
};

struct TraceStats {
    unsigned int nodes, prims;
};

// A sub-tree is skipped only when the ray enters its box farther than best_t * (1 + 1e-4) + 1e-4.
// Hit distances are recomputed as |P - O| and can differ from the ray parameter by rounding
// (~1e-7 relative), four orders below this margin.
RT_DEV float prune_limit(float best_t) { return best_t * 1.0001f + 1e-4f; }

#define RT_STACK 40

// Conservative test with a three-way answer for leaf boxes: 0 = the reference test surely fails,
// 2 = it surely passes (the interval has more than twice the error bound to spare), 1 = too close
// to call -> evaluate box_exact.
RT_DEV int box_classify(float lox, float loy, float loz, float hix, float hiy, float hiz, const RayAux& a) {
    const float x1 = __fmaf_rn(lox, a.ix, a.nx), x2 = __fmaf_rn(hix, a.ix, a.nx);
    const float y1 = __fmaf_rn(loy, a.iy, a.ny), y2 = __fmaf_rn(hiy, a.iy, a.ny);
    const float z1 = __fmaf_rn(loz, a.iz, a.nz), z2 = __fmaf_rn(hiz, a.iz, a.nz);
    const float tn = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fminf(z1, z2));
    const float tf = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fmaxf(z1, z2));
    const float slack = __fmaf_rn(1e-6f, fabsf(tn) + fabsf(tf), a.k);
    if (!(tn <= tf + slack && tf >= -slack)) return 0;
    return (tn + slack <= tf - slack && tf - slack >= 0.0f) ? 2 : 1;
}

// BVH::get_intersection (acceleration.cpp:142-150): closest hit, ties -> first in leaf order.
// ANY = true: the shadow query of shade() (raytracer.cpp:230-235): true iff some tested shape has
// t <= max_t (== "closest hit exists and its t is not > light distance").
//
// "while-while" traversal: all lanes of a warp first descend through internal nodes (uniform
// code: one 64-byte node, two conservative box tests), and only when every lane has reached a
// leaf or finished do they process leaves together (exact leaf-box test, per-primitive culling
// boxes, primitive tests). This keeps the lanes of a warp in the same code far more often than
// interleaving the two per lane.
// Out-of-line copies keep the hot traversal loop small enough for the instruction cache (with
// everything inlined the trace kernel was 64 KB of SASS, twice the 32 KB L1.5 I-cache, and
// `no_instruction` was the top stall reason).
// (arguments by value: a reference parameter of a real call would pin the caller's ray state in
// local memory)
__device__ __noinline__ float box_exact_call_impl(float lox, float loy, float loz, float hix, float hiy, float hiz, float ox,
                                                  float oy, float oz, float dx, float dy, float dz) {
    Ray r;
    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz; r.time = 0.0f;
    float tn;
    return box_exact(lox, loy, loz, hix, hiy, hiz, r, tn) ? tn : __int_as_float(0x7fc00000);  // NaN = miss
}
RT_DEV bool box_exact_call(float lox, float loy, float loz, float hix, float hiy, float hiz, const Ray& r, float& tnear) {
    const float t = box_exact_call_impl(lox, loy, loz, hix, hiy, hiz, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz);
    tnear = t;
    return t == t;
}

// BVH::intersect_linear (acceleration.cpp:123-138): every shape, in shape_list order.
template <bool ANY>
__device__ __noinline__ int4 traverse_linear_impl(const float4* __restrict__ prims, int n_prims, float ox, float oy, float oz,
                                                  float dx, float dy, float dz, float time, float max_t) {
    Ray r;
    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz; r.time = time;
    float best_t = FLT_MAX;
    int best_prim = -1, tests = 0;
    for (int i = 0; i < n_prims; ++i) {
        Hit h;
        tests++;
        if (intersect_prim<false>(prims, i, r, h)) {
            if (ANY) { if (!(h.t > max_t)) return make_int4(1, i, __float_as_int(h.t), tests); }
            else if (h.t < best_t) { best_t = h.t; best_prim = i; }
        }
    }
    return make_int4(0, best_prim, __float_as_int(best_t), tests);
}
template <bool ANY>
RT_DEV bool traverse_linear(const BvhView& b, const Ray& r, float max_t, float& best_t, int& best_prim, unsigned int& n_tests) {
    const int4 v = traverse_linear_impl<ANY>(b.prims, b.n_prims, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, r.time, max_t);
    if (!ANY) { best_prim = v.y; best_t = __int_as_float(v.z); }
    n_tests += (unsigned int)v.w;
    return v.x != 0;
}

template <bool ANY, bool STATS>
RT_DEV bool traverse(const BvhView& b, const Ray& r, float max_t, float& best_t, int& best_prim, TraceStats& st) {
    best_t = FLT_MAX;
    best_prim = -1;
    if (b.n_prims == 0) return false;
    if (!b.use_bvh) return traverse_linear<ANY>(b, r, max_t, best_t, best_prim, st.prims);
    const RayAux a = make_aux(r);
    const bool exact_only = a.slow || !b.prune;
    // No separate root-box test: the root contains every leaf box, so a ray that fails it fails
    // every (exact) leaf test below as well.

    int stack[RT_STACK];
    int sp = 0;
    int cur = b.root_ref;
    float lim = ANY ? prune_limit(max_t) : FLT_MAX;
    bool done = false;
    while (true) {
        // ---- phase 1: internal nodes ----
        while (cur >= 0) {
            const float4* n = b.nodes + (size_t)cur * 4;
            const float4 na = __ldg(n + 0), nb = __ldg(n + 1), nc = __ldg(n + 2), nd = __ldg(n + 3);
            const int li = __float_as_int(nd.x), ri = __float_as_int(nd.y);
            float tl, tr;
            bool hl, hr;
            if (STATS) st.nodes += 2;
            if (exact_only) {
                hl = box_exact_call(na.x, na.y, na.z, na.w, nb.x, nb.y, r, tl);
                hr = box_exact_call(nb.z, nb.w, nc.x, nc.y, nc.z, nc.w, r, tr);
            } else {
                hl = box_maybe(na.x, na.y, na.z, na.w, nb.x, nb.y, a, tl);
                hr = box_maybe(nb.z, nb.w, nc.x, nc.y, nc.z, nc.w, a, tr);
            }
            if (b.prune) { hl = hl && !(tl > lim); hr = hr && !(tr > lim); }
            if (hl && hr) {
                const bool left_first = !b.prune || tl <= tr;
                stack[sp++] = left_first ? ri : li;
                cur = left_first ? li : ri;
            } else if (hl) {
                cur = li;
            } else if (hr) {
                cur = ri;
            } else {
                if (sp == 0) { done = true; break; }
                cur = stack[--sp];
            }
        }
        if (done) break;
        // ---- phase 2: one leaf ----
        {
            const float4* L = b.leaves + (size_t)(~cur) * 8;
            const float4 l0 = __ldg(L + 0), l1 = __ldg(L + 1);
            if (STATS) st.nodes++;
            int c = exact_only ? 1 : box_classify(l0.x, l0.y, l0.z, l1.x, l1.y, l1.z, a);
            if (c == 1) { float te; c = box_exact_call(l0.x, l0.y, l0.z, l1.x, l1.y, l1.z, r, te) ? 2 : 0; }
            if (c == 2) {
                const int first = __float_as_int(l0.w), count = __float_as_int(l1.w);
                unsigned int mask = (1u << count) - 1u;
                if (!exact_only) {
                    // skip primitives whose (inflated) own box the ray clearly misses or enters too far
                    const float4 v2 = __ldg(L + 2), v3 = __ldg(L + 3), v4 = __ldg(L + 4);
                    const float4 v5 = __ldg(L + 5), v6 = __ldg(L + 6), v7 = __ldg(L + 7);
                    float tp;
                    if (!(box_maybe(v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, a, tp) && !(tp > lim))) mask &= ~1u;
                    if (!(box_maybe(v3.z, v3.w, v4.x, v4.y, v4.z, v4.w, a, tp) && !(tp > lim))) mask &= ~2u;
                    if (!(box_maybe(v5.x, v5.y, v5.z, v5.w, v6.x, v6.y, a, tp) && !(tp > lim))) mask &= ~4u;
                    if (!(box_maybe(v6.z, v6.w, v7.x, v7.y, v7.z, v7.w, a, tp) && !(tp > lim))) mask &= ~8u;
                }
                while (mask) {  // ONE call site of the (large) primitive test
                    const int k = __ffs(mask) - 1;
                    mask &= mask - 1u;
                    Hit h;
                    if (STATS) st.prims++;
                    const int idx = first + k;
                    if (intersect_prim<false>(b.prims, idx, r, h)) {
                        if (ANY) { if (!(h.t > max_t)) return true; }
                        else if (h.t < best_t || (h.t == best_t && idx < best_prim)) { best_t = h.t; best_prim = idx; lim = prune_limit(best_t); }
                    }
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// Resumable traversal: the same algorithm as traverse(), cut into steps so that a persistent
// warp can hand a finished lane a new ray while its other lanes keep going.
//   trav_begin : set up the per-ray state
//   trav_step  : phase 1 (descend internal nodes until a leaf) + phase 2 (that leaf) + pop;
//                returns true when the ray is finished. For ANY queries best_prim >= 0 then
//                means "occluded".
// ---------------------------------------------------------------------------------------------
#define RT_CUR_IDLE ((int)0x80000000)

struct TravState {
    Ray r;
    RayAux a;
    float max_t;      // ANY: light distance
    float best_t;
    float lim;
    int best_prim;
    int cur;          // node ref being visited, RT_CUR_IDLE when the lane has no ray
    int sp;
    bool exact_only;
};

template <bool ANY>
RT_DEV bool trav_begin(const BvhView& b, TravState& s, const Ray& r, float max_t, TraceStats& st) {
    s.r = r;
    s.max_t = max_t;
    s.best_t = FLT_MAX;
    s.best_prim = -1;
    s.sp = 0;
    s.cur = RT_CUR_IDLE;
    if (b.n_prims == 0) return true;
    if (!b.use_bvh) {
        if (traverse_linear<ANY>(b, r, max_t, s.best_t, s.best_prim, st.prims)) s.best_prim = 0;
        return true;
    }
    s.a = make_aux(r);
    s.exact_only = s.a.slow || !b.prune;
    s.lim = ANY ? prune_limit(max_t) : FLT_MAX;
    s.cur = b.root_ref;
    return false;
}

// One primitive test of a known type inside the leaf phase (see trav_step).
template <bool ANY, bool STATS, int TYPE>
RT_DEV void leaf_tests_of_type(const BvhView& b, TravState& s, unsigned int& pending, unsigned int type_mask, int first,
                               bool& occluded, TraceStats& st) {
    unsigned int m = pending & type_mask;
    while (__any_sync(0xffffffffu, m != 0u)) {  // all 32 lanes are here: one type's routine at a time
        if (m != 0u) {
            const int k = __ffs(m) - 1;
            m &= m - 1u;
            Hit h;
            if (STATS) st.prims++;
            const int idx = first + k;
            if (intersect_prim<false, TYPE>(b.prims, idx, s.r, h)) {
                if (ANY) {
                    if (!(h.t > s.max_t)) { s.best_prim = idx; occluded = true; m = 0u; pending = 0u; }
                } else if (h.t < s.best_t || (h.t == s.best_t && idx < s.best_prim)) {
                    s.best_t = h.t; s.best_prim = idx; s.lim = prune_limit(s.best_t);
                }
            }
        }
    }
}

// MUST be called by all 32 lanes of a warp together (lanes without a ray have cur == RT_CUR_IDLE).
template <bool ANY, bool STATS>
RT_DEV bool trav_step(const BvhView& b, TravState& s, int* stack, TraceStats& st) {
    bool finished = false;
    // ---- phase 1: internal nodes, until this lane reaches a leaf or runs out of nodes ----
    while (s.cur >= 0) {
        const float4* n = b.nodes + (size_t)s.cur * 4;
        const float4 na = __ldg(n + 0), nb = __ldg(n + 1), nc = __ldg(n + 2), nd = __ldg(n + 3);
        const int li = __float_as_int(nd.x), ri = __float_as_int(nd.y);
        float tl, tr;
        bool hl, hr;
        if (STATS) st.nodes += 2;
        if (s.exact_only) {
            hl = box_exact_call(na.x, na.y, na.z, na.w, nb.x, nb.y, s.r, tl);
            hr = box_exact_call(nb.z, nb.w, nc.x, nc.y, nc.z, nc.w, s.r, tr);
        } else {
            hl = box_maybe(na.x, na.y, na.z, na.w, nb.x, nb.y, s.a, tl);
            hr = box_maybe(nb.z, nb.w, nc.x, nc.y, nc.z, nc.w, s.a, tr);
        }
        if (b.prune) { hl = hl && !(tl > s.lim); hr = hr && !(tr > s.lim); }
        if (hl && hr) {
            const bool left_first = !b.prune || tl <= tr;
            stack[s.sp++] = left_first ? ri : li;
            s.cur = left_first ? li : ri;
        } else if (hl) {
            s.cur = li;
        } else if (hr) {
            s.cur = ri;
        } else if (s.sp == 0) {
            s.cur = RT_CUR_IDLE;
            finished = true;
        } else {
            s.cur = stack[--s.sp];
        }
    }
    // ---- phase 2: one leaf per lane (lanes that are idle or just finished carry pending == 0) ----
    const bool has_leaf = s.cur != RT_CUR_IDLE;
    unsigned int pending = 0u, meta = 0u;
    int first = 0;
    if (has_leaf) {
        const float4* L = b.leaves + (size_t)(~s.cur) * 8;
        const float4 l0 = __ldg(L + 0), l1 = __ldg(L + 1);
        if (STATS) st.nodes++;
        int c = s.exact_only ? 1 : box_classify(l0.x, l0.y, l0.z, l1.x, l1.y, l1.z, s.a);
        if (c == 1) { float te; c = box_exact_call(l0.x, l0.y, l0.z, l1.x, l1.y, l1.z, s.r, te) ? 2 : 0; }
        if (c == 2) {
            first = __float_as_int(l0.w);
            meta = __float_as_uint(l1.w);
            pending = (1u << (meta & 7u)) - 1u;
            if (!s.exact_only) {
                // skip primitives whose (inflated) own box the ray clearly misses or enters too far
                const float4 v2 = __ldg(L + 2), v3 = __ldg(L + 3), v4 = __ldg(L + 4);
                const float4 v5 = __ldg(L + 5), v6 = __ldg(L + 6), v7 = __ldg(L + 7);
                float tp;
                if (!(box_maybe(v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, s.a, tp) && !(tp > s.lim))) pending &= ~1u;
                if (!(box_maybe(v3.z, v3.w, v4.x, v4.y, v4.z, v4.w, s.a, tp) && !(tp > s.lim))) pending &= ~2u;
                if (!(box_maybe(v5.x, v5.y, v5.z, v5.w, v6.x, v6.y, s.a, tp) && !(tp > s.lim))) pending &= ~4u;
                if (!(box_maybe(v6.z, v6.w, v7.x, v7.y, v7.z, v7.w, s.a, tp) && !(tp > s.lim))) pending &= ~8u;
            }
        }
    }
    // Primitive tests grouped by type across the warp: lanes sit in different leaves whose
    // primitives have different types; running sphere tests, then cube tests, ... keeps the lanes
    // in one routine at a time instead of serialising up to four routines per test.
    bool occluded = false;
    leaf_tests_of_type<ANY, STATS, RT_SPHERE>(b, s, pending, (meta >> 4) & 15u, first, occluded, st);
    leaf_tests_of_type<ANY, STATS, RT_CUBE>(b, s, pending, (meta >> 8) & 15u, first, occluded, st);
    leaf_tests_of_type<ANY, STATS, RT_RECTANGLE>(b, s, pending, (meta >> 12) & 15u, first, occluded, st);
    leaf_tests_of_type<ANY, STATS, RT_PLANE>(b, s, pending, (meta >> 16) & 15u, first, occluded, st);
    if (has_leaf) {
        if (occluded || s.sp == 0) { s.cur = RT_CUR_IDLE; finished = true; }
        else s.cur = stack[--s.sp];
    }
    return finished;
}

// Warp-level work distribution for persistent kernels: the warp owns a pool [pool_lo, pool_hi)
// of consecutive item indices (refilled with one atomic per RT_POOL items); every lane that needs
// an item gets one. Returns the item index or -1. `more` turns false when the global counter has
// passed n. All 32 lanes must call it together.
#define RT_POOL 128u
RT_DEV long long warp_take(unsigned int* counter, unsigned long long n, bool need, unsigned int& pool_lo, unsigned int& pool_hi,
                           bool& more) {
    const int lane = threadIdx.x & 31;
    long long item = -1;
    unsigned int mask = __ballot_sync(0xffffffffu, need);
    while (mask != 0u && (pool_lo < pool_hi || more)) {
        if (pool_lo >= pool_hi) {  // warp-uniform: get the next pool
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(counter, RT_POOL);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((unsigned long long)base >= n) { more = false; break; }
            pool_lo = base;
            pool_hi = (unsigned int)min((unsigned long long)base + RT_POOL, n);
        }
        const unsigned int avail = pool_hi - pool_lo;
        const unsigned int rank = (unsigned int)__popc(mask & ((1u << lane) - 1u));
        const bool mine = need && item < 0 && rank < avail;
        if (mine) item = (long long)pool_lo + rank;
        pool_lo += min(avail, (unsigned int)__popc(mask));
        mask = __ballot_sync(0xffffffffu, need && item < 0);
    }
    return item;
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

}  // namespace rtb
