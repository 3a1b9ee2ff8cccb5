// rt_device.cuh -- device-side building blocks: the reference's arithmetic (box test, the four
// ray-primitive intersections, sampling) and the two BVH queries (closest hit, occlusion).
//
// Arithmetic contract: the including .cu is compiled with -fmad=false and the default IEEE
// division / square root, and every expression here is written in the reference's operand order,
// so each float the reference computes on x86-64/SSE2 is reproduced bit for bit (transcendentals
// excepted). Explicit __fmaf_rn is used ONLY in the conservative box test, which never decides a
// result on its own (see wide_child_test()).
//
// Traversal semantics: the reference (acceleration.cpp:67-117) visits every node whose box the
// ray passes, collects ALL leaf hits and returns the first minimum. Ancestor boxes contain leaf
// boxes and IEEE rounding is monotonic, so a shape is tested iff the box of ITS LEAF passes
// AABB::intersect. Therefore:
//   * that condition is evaluated PER PRIMITIVE ("gate"): the exact box of the primitive's reference
//     leaf gets a conservative test first and, when the outcome is too close to call, the reference's
//     exact test (gate_passes());
//   * every box of the device tree is a CULLING box (a primitive's padded box or a union of such):
//     it only needs a CONSERVATIVE test -- never rejects a ray for which the primitive's routine
//     could report a hit -- 6 FMAs with a precomputed reciprocal instead of 6 IEEE divisions;
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

// Loads of data that is touched once per ray (primitive records, gate boxes, queue entries, shade records): they
// go around L1 (RT_STREAM_LOADS = 1) so that L1 keeps what the traversal re-reads -- the BVH nodes.
#ifndef RT_STREAM_LOADS
#define RT_STREAM_LOADS 0
#endif
RT_DEV float4 ld_once(const float4* p) {
#if RT_STREAM_LOADS
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
// ... the same for buffers another kernel wrote shortly before (no .nc)
RT_DEV float4 ld_once_rw(const float4* p) {
#if RT_STREAM_LOADS
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return *p;
#endif
}

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
// CLS: what the caller knows about the primitive's type (the traversal runs a warp's pending tests
// class by class, so only that class's code is generated at each call site):
//   PRIM_ANY any type; PRIM_XFORM sphere / cube / rectangle (they share the transform into object
//   space and back); PRIM_PLANE plane.
enum PrimClass { PRIM_ANY = 0, PRIM_XFORM = 1, PRIM_PLANE = 2 };
template <bool FULL, int CLS = PRIM_ANY>
RT_DEV bool intersect_prim(const float4* __restrict__ prims, int idx, const Ray& r, Hit& h) {
    const float4* q = prims + (size_t)idx * 8;
    const float4 q0 = ld_once(q + 0);
    const float4 q1 = ld_once(q + 1);
    const float4 q2 = ld_once(q + 2);
    const float4 q3 = ld_once(q + 3);
    const int type = CLS == PRIM_PLANE ? (int)RT_PLANE : (int)(__float_as_uint(q0.w) & 3u);

    if (CLS != PRIM_XFORM && type == RT_PLANE) {
        // Plane::intersect (shapes.cpp:444-483); q1..q3 = corners 0..2 (+ corner 3 in .w), q4 = normal
        const float4 q4 = ld_once(q + 4);
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
    const float4 q4 = ld_once(q + 4);
    const float4 q5 = ld_once(q + 5);
    const float4 q6 = ld_once(q + 6);
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
// BVH queries over the 4-wide tree (scene.hpp DWide)
// ---------------------------------------------------------------------------------------------
struct BvhView {
    const float4* __restrict__ prims;
    const float* __restrict__ wide;  // 32 floats per node
    const float4* __restrict__ leafbox;  // 2 x float4 per primitive: exact box of its reference leaf
    const float* __restrict__ ref_tree;  // literal mode only: the reference's binary tree, 10 words per node (RefNode)
    int n_prims;
    int use_bvh;
    int prune;        // 0: visit everything, exact tests only (the reference's literal traversal)
    int stack_depth;  // entries per thread of the shared-memory traversal stack (per-ray kernels)
    int packet_stack_depth;  // entries per warp (packet kernels)
    int n_staged;  // RT_STAGE_TOP builds: the first n_staged nodes (breadth-first = the top levels) are also in shared memory
};

struct TraceStats {
    unsigned int nodes, prims;
};

// A sub-tree is skipped only when the ray enters its box farther than best_t * (1 + 1e-4) + 1e-4.
// Hit distances are recomputed as |P - O| and can differ from the ray parameter by rounding
// (~1e-7 relative), four orders below this margin.
RT_DEV float prune_limit(float best_t) { return best_t * 1.0001f + 1e-4f; }


struct F8 { float v[8]; };
// One 256-bit read-only load (sm_100: LDG.E.256): a quarter of a wide node per instruction, i.e.
// a quarter of the L1 tag look-ups that 16-byte loads would need for the same line.
RT_DEV F8 ldg256(const float* p) {
    F8 r;
#ifdef RT_NO_LDG256
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
#endif
#ifdef RT_NODE_EVICT_LAST
    asm("ld.global.nc.L1::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
        : "l"(p));
    return r;
#endif
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
        : "l"(p));
    return r;
}

// RT_STAGE_TOP = K > 0 (A/B build): every block of the per-ray kernels copies the first K nodes of the tree (the
// top levels: nodes are numbered breadth-first) into shared memory, and trav_step reads those from there.
#ifndef RT_STAGE_TOP
#define RT_STAGE_TOP 0
#endif
RT_DEV F8 lds256(unsigned int addr) {
    F8 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "r"(addr));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "r"(addr + 16u));
    return r;
}

// Out-of-line copies keep the hot traversal loop small (instruction cache) and rare.
// (arguments by value: a reference parameter of a real call would pin the caller's ray state in
// local memory)
__device__ __noinline__ float box_exact_call_impl(float lox, float loy, float loz, float hix, float hiy, float hiz, float ox,
                                                  float oy, float oz, float dx, float dy, float dz) {
    Ray r;
    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz; r.time = 0.0f;
    float tn;
    return box_exact(lox, loy, loz, hix, hiy, hiz, r, tn) ? tn : __int_as_float(0x7fc00000);  // NaN = miss
}
RT_DEV bool box_exact_call(float lox, float loy, float loz, float hix, float hiy, float hiz, const Ray& r) {
    const float t = box_exact_call_impl(lox, loy, loz, hix, hiy, hiz, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz);
    return t == t;
}

// BVH::intersect_linear (acceleration.cpp:123-138): every shape, in shape_list order.
// Returns (occluded, best_prim, bits best_t, tests).
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

// BVH::intersect / intersect_helper (acceleration.cpp:67-117), literally: the reference's own binary tree
// (RefNode = TreeNode of scene.hpp: box lo[3] hi[3], left, right, first, count), the exact box test at every
// node, both children visited, every shape of every leaf reached tested, first minimum of t. This is the
// prune = 0 validation mode; it shares nothing with the production traversal but the intersection routines.
// Returns (occluded, best_prim, bits best_t, primitive tests); box tests in `boxes`.
template <bool ANY>
__device__ __noinline__ int4 traverse_reference_impl(const float4* __restrict__ prims, const float* __restrict__ ref_tree, float ox,
                                                     float oy, float oz, float dx, float dy, float dz, float time, float max_t,
                                                     unsigned int* boxes) {
    Ray r;
    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz; r.time = time;
    float best_t = FLT_MAX;
    int best_prim = -1, tests = 0;
    unsigned int nb = 0;
    int stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const float* nd = ref_tree + (size_t)stack[--sp] * 10;
        float t;
        nb++;
        if (!box_exact(__ldg(nd + 0), __ldg(nd + 1), __ldg(nd + 2), __ldg(nd + 3), __ldg(nd + 4), __ldg(nd + 5), r, t)) continue;
        const int left = __float_as_int(__ldg(nd + 6)), right = __float_as_int(__ldg(nd + 7));
        if (left < 0) {
            const int first = __float_as_int(__ldg(nd + 8)), count = __float_as_int(__ldg(nd + 9));
            for (int k = 0; k < count; ++k) {
                Hit h;
                tests++;
                const int idx = first + k;
                if (intersect_prim<false>(prims, idx, r, h)) {
                    if (ANY) { if (!(h.t > max_t)) { *boxes = nb; return make_int4(1, idx, __float_as_int(h.t), tests); } }
                    else if (h.t < best_t || (h.t == best_t && idx < best_prim)) { best_t = h.t; best_prim = idx; }
                }
            }
        } else if (sp + 2 <= 64) {
            stack[sp++] = right;
            stack[sp++] = left;
        }
    }
    *boxes = nb;
    return make_int4(0, best_prim, __float_as_int(best_t), tests);
}

// ---------------------------------------------------------------------------------------------
// Per-lane traversal state of the wavefront kernels.
//
// Conservative slab test: t' = fma(plane, 1/d_i, -o_i/d_i). Against the reference's
// RN(RN(plane - o_i) / d_i):
//     |t' - t| <= 4u |t| + 1.01u |o_i / d_i|                 (u = 2^-24)
// The second term is a per-axis constant of the ray, so it is folded into the addend: the value
// that can become the entry parameter of axis i uses -o_i/d_i - e_i, the one that can become the
// exit parameter -o_i/d_i + e_i, e_i = 2u |o_i/d_i| (n?l / n?h below; which of lo/hi is the entry
// plane follows from the sign of d_i). Only the relative term is left for the comparison:
//     slack = 1e-6 (|tn'| + |tf'|) + Q imax (|tn'| + |tf'|)^2
// Q is zero except for sphere culling boxes (scene.hpp / bvh.cpp cull_pad: the distance-squared
// rounding term of the sphere test; (|tn'|+|tf'|) bounds the distance to the shape, and
// imax = max_i |1/d_i| converts the spatial pad into parameter space).
//   pass   : tn' - tf' <= slack        and  tf' >= -slack       (never rejects what the reference accepts)
//   surely : tn' - tf' <= -slack - KS  and  tf' >=  slack + KS  (the reference's exact test passes;
//            KS = 4 max_i e_i undoes the padding of the addends)
// The GATE of a primitive (exact box of its reference leaf) that passes but not surely gets the exact test.
//
// Rays with a direction component |d_i| <= 1e-6: the reference's box test then only asks whether the
// origin lies inside the slab (its "parallel" rule), whatever the other axes say -- neither a superset nor
// a subset of the slab test. So for such rays the gate is always decided by the exact test (KS is set to
// infinity), while the culling boxes -- which bound where the primitive's routine can report a hit, i.e.
// geometry -- keep the slab test with the ray's true reciprocal, clamped to +-1e30 so that no product
// overflows (a component below 1e-30 moves the ray by less than 1e-24 over any scene: it is treated as
// zero, with an absolute e_i of 1e20 that turns the slab test into "origin inside the slab, to 1e-10").
// ---------------------------------------------------------------------------------------------
#define RT_CUR_NONE (-1)

// The traversal stack lives in shared memory, one column per thread (entry e of thread t at
// base + (e * blockDim + t) * 4): TravState::sp is the 32-bit shared-space ADDRESS of the next free
// entry, so a push is one STS and one add.
// Closest-hit queries store (node, entry parameter) pairs -- RT_STACK_WORDS = 2, one 64-bit
// access -- so that a popped sub-tree that starts beyond the current best hit is dropped without
// visiting it; occlusion queries have a fixed limit and store the node only.
// RT_ANY_SORTED: occlusion queries also visit children near-first (occluders tend to sit close to
// the surface the shadow ray leaves) instead of in slot order.
#ifndef RT_ANY_SORTED
#define RT_ANY_SORTED 1
#endif
#ifndef RT_STACK_ENTRY_T
#define RT_STACK_ENTRY_T 1
#endif
#define RT_STACK_WORDS (RT_STACK_ENTRY_T ? 2 : 1)
// RT_CHECKS=1 (debug builds, scripts: make variant NAME=checks NVCC_EXTRA=-DRT_CHECKS=1): trap on a
// traversal-stack overflow instead of corrupting a neighbour's stack.
#ifndef RT_CHECKS
#define RT_CHECKS 0
#endif
RT_DEV void stack_push(unsigned int& sp, unsigned int stride_bytes, int v) {
    asm volatile("st.shared.b32 [%0], %1;" :: "r"(sp), "r"(v) : "memory");
    sp += stride_bytes;
}
RT_DEV int stack_pop(unsigned int& sp, unsigned int stride_bytes) {
    int v;
    sp -= stride_bytes;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(sp) : "memory");
    return v;
}
RT_DEV void stack_push2(unsigned int& sp, unsigned int stride_bytes, int v, int t) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" :: "r"(sp), "r"(v), "r"(t) : "memory");
    sp += stride_bytes;
}
RT_DEV int stack_pop2(unsigned int& sp, unsigned int stride_bytes, int& t) {
    int v;
    sp -= stride_bytes;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v), "=r"(t) : "r"(sp) : "memory");
    return v;
}

struct TravState {
    Ray r;
    float ix, iy, iz;     // 1/d
    float nxl, nyl, nzl;  // -o/d -+ e: addend for the lo planes
    float nxh, nyh, nzh;  // addend for the hi planes
    float KS, imax;
    float max_t;       // ANY: light distance
    float best_t;
    float lim;
    int best_prim;     // closest primitive so far; ANY: >= 0 means occluded
    int cur;           // wide node to visit next, RT_CUR_NONE when there is none
    unsigned int sp;   // shared-space address of the next free stack entry
    unsigned int sp0;  // ... of entry 0 (stack empty when sp == sp0)
    unsigned int sp_end;  // RT_CHECKS: address one past the last entry
    unsigned int pend; // bit k: child k of node pend_node is a primitive waiting for its test
    int pend_node;     //        (RT_PEND_SLOTS == 2: bits 4-7 belong to pend_node2 -- a lane may keep stepping
    int pend_node2;    //        past one node with candidates before the warp's primitive phase has run)
    unsigned int staged;  // RT_STAGE_TOP: shared-space address of the staged top nodes
};

// Sets up a ray. Returns true when the ray is already finished (empty scene).
template <bool ANY>
RT_DEV bool trav_begin(const BvhView& b, TravState& s, const Ray& r, float max_t) {
    s.r = r;
    s.max_t = max_t;
    s.best_t = FLT_MAX;
    s.best_prim = -1;
    s.sp = s.sp0;
    s.cur = RT_CUR_NONE;
    s.pend = 0u;
    s.pend_node = 0;
    s.pend_node2 = 0;
    if (b.n_prims == 0) return true;
    const float BIG = 1e30f;
    s.ix = 1.0f / r.dx; s.iy = 1.0f / r.dy; s.iz = 1.0f / r.dz;
    const bool cx = !(fabsf(s.ix) <= BIG), cy = !(fabsf(s.iy) <= BIG), cz = !(fabsf(s.iz) <= BIG);
    if (cx) s.ix = copysignf(BIG, r.dx);
    if (cy) s.iy = copysignf(BIG, r.dy);
    if (cz) s.iz = copysignf(BIG, r.dz);
    const float nx = -(r.ox * s.ix), ny = -(r.oy * s.iy), nz = -(r.oz * s.iz);
    float ex = 1.2e-7f * fabsf(nx) + 1e-35f, ey = 1.2e-7f * fabsf(ny) + 1e-35f, ez = 1.2e-7f * fabsf(nz) + 1e-35f;
    if (cx) ex = fmaxf(ex, 1e20f);
    if (cy) ey = fmaxf(ey, 1e20f);
    if (cz) ez = fmaxf(ez, 1e20f);
    // d_i > 0: the lo plane is the entry plane (pad it towards -inf), the hi plane the exit plane
    s.nxl = s.ix > 0.0f ? nx - ex : nx + ex; s.nxh = s.ix > 0.0f ? nx + ex : nx - ex;
    s.nyl = s.iy > 0.0f ? ny - ey : ny + ey; s.nyh = s.iy > 0.0f ? ny + ey : ny - ey;
    s.nzl = s.iz > 0.0f ? nz - ez : nz + ez; s.nzh = s.iz > 0.0f ? nz + ez : nz - ez;
    const bool parallel = fabsf(r.dx) <= 1e-6f || fabsf(r.dy) <= 1e-6f || fabsf(r.dz) <= 1e-6f;
    s.KS = parallel ? __int_as_float(0x7f800000) : 4.0f * fmaxf(fmaxf(ex, ey), ez);
    s.imax = fmaxf(fmaxf(fabsf(s.ix), fabsf(s.iy)), fabsf(s.iz));
    s.lim = ANY ? prune_limit(max_t) : FLT_MAX;
    s.cur = 0;
    return false;
}

// One child of a wide node: conservative slab test. m = min(tf' - tn', tf'): the box passes iff
// m >= -slack and surely passes iff m >= slack + KS. `ent` = lower bound of the entry parameter.
// qi = Q * imax of the node (zero unless the node holds spheres).
RT_DEV void wide_child_test(const TravState& s, float lox, float hix, float loy, float hiy, float loz, float hiz, float qi,
                            bool& pass, bool& surely, float& ent) {
    const float x1 = __fmaf_rn(lox, s.ix, s.nxl), x2 = __fmaf_rn(hix, s.ix, s.nxh);
    const float y1 = __fmaf_rn(loy, s.iy, s.nyl), y2 = __fmaf_rn(hiy, s.iy, s.nyh);
    const float z1 = __fmaf_rn(loz, s.iz, s.nzl), z2 = __fmaf_rn(hiz, s.iz, s.nzh);
    const float tn = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fminf(z1, z2));
    const float tf = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fmaxf(z1, z2));
    const float sum = fabsf(tn) + fabsf(tf);
    const float sl = __fmaf_rn(sum, __fmaf_rn(qi, sum, 1e-6f), 1e-30f);  // 1e-6 s + Q imax s^2
    const float m = fminf(tf - tn, tf);
    ent = tn - sl;
    pass = m >= -sl && ent <= s.lim;
    surely = m >= sl + s.KS;
}

#ifndef RT_PEND_SLOTS
#define RT_PEND_SLOTS 1
#endif
// Visits node s.cur: tests its (up to) four children. Node children are pushed far-to-near and
// the nearest becomes s.cur; primitive children whose culling box passes become s.pend. The
// caller guarantees s.cur != RT_CUR_NONE and s.pend == 0. `stride` = bytes between two stack entries
// of a thread. Written to compile to straight-line predicated code.
template <bool ANY, bool STATS>
RT_DEV void trav_step(const BvhView& b, TravState& s, unsigned int stride, TraceStats& st) {
    const int node = s.cur;
    const float* w = b.wide + (size_t)node * 32;
#if RT_STAGE_TOP > 0
    F8 X, Y, Z, C;
    if (node < b.n_staged) {
        const unsigned int a = s.staged + (unsigned int)node * 128u;
        X = lds256(a); Y = lds256(a + 32u); Z = lds256(a + 64u); C = lds256(a + 96u);
    } else {
        X = ldg256(w); Y = ldg256(w + 8); Z = ldg256(w + 16); C = ldg256(w + 24);
    }
#else
    const F8 X = ldg256(w), Y = ldg256(w + 8), Z = ldg256(w + 16), C = ldg256(w + 24);
#endif
    if (STATS) st.nodes += 4;
    const int first = __float_as_int(C.v[0]);
    const unsigned int meta = __float_as_uint(C.v[1]);
    const float qi = C.v[2] * s.imax;
    bool pass[4];
    float ent[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        bool sure;  // unused for culling boxes: dead code after inlining
        wide_child_test(s, X.v[k], X.v[4 + k], Y.v[k], Y.v[4 + k], Z.v[k], Z.v[4 + k], qi, pass[k], sure, ent[k]);
    }
    unsigned int pm = ((pass[0] ? 1u : 0u) | (pass[1] ? 2u : 0u) | (pass[2] ? 4u : 0u) | (pass[3] ? 8u : 0u)) & meta;
    int next = RT_CUR_NONE;
    unsigned int sp = s.sp;
    // primitive children whose culling box passes wait for the warp's next primitive phase
#if RT_PEND_SLOTS == 2
    {
        const unsigned int cand = pm & (meta >> 4) & 15u;
        if (cand != 0u) {
            if ((s.pend & 15u) == 0u) { s.pend |= cand; s.pend_node = node; }
            else { s.pend |= cand << 4; s.pend_node2 = node; }
        }
    }
#else
    s.pend = pm & (meta >> 4);
    s.pend_node = node;
#endif
    pm &= ~(meta >> 4);
    if (ANY && !RT_ANY_SORTED) {
        // occlusion query: order does not matter
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool p = (pm >> k) & 1u;
            if (p && next != RT_CUR_NONE) stack_push(sp, stride, next);
            if (p) next = first + k;
        }
    } else {
        // sort the passing children by entry distance: key = entry bits with the slot in the low 2
        // bits, compared as SIGNED ints (negative entries -- origin inside the box -- sort first, in
        // any order among themselves); children that do not pass get the largest key
        int key[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            key[k] = ((pm >> k) & 1u) ? ((__float_as_int(ent[k]) & ~3) | k) : 0x7fffffff;
#define RT_CSWAP(i, j) { const int lo_ = min(key[i], key[j]), hi_ = max(key[i], key[j]); key[i] = lo_; key[j] = hi_; }
        RT_CSWAP(0, 1) RT_CSWAP(2, 3) RT_CSWAP(0, 2) RT_CSWAP(1, 3) RT_CSWAP(1, 2)
#undef RT_CSWAP
        const int n = __popc(pm);
#pragma unroll
        for (int j = 3; j >= 1; --j) {
            if (j < n) {
                if (RT_STACK_ENTRY_T && !ANY) stack_push2(sp, stride, first + (key[j] & 3), key[j]);
                else stack_push(sp, stride, first + (key[j] & 3));
            }
        }
        if (n > 0) next = first + (key[0] & 3);
    }
    if (RT_CHECKS && sp > s.sp_end) __trap();
    if (!ANY && RT_STACK_ENTRY_T) {
        // drop popped sub-trees that start beyond the best hit (the key's low 2 bits are the slot:
        // clearing them only lowers the entry bound for positive entries; negative ones always pass)
        while (next == RT_CUR_NONE && sp != s.sp0) {
            int key;
            const int node2 = stack_pop2(sp, stride, key);
            if (key < 0 || __int_as_float(key & ~3) <= s.lim) next = node2;
        }
    } else if (next == RT_CUR_NONE && sp != s.sp0) {
        next = stack_pop(sp, stride);
    }
    s.sp = sp;
    s.cur = next;
}

// The gate of primitive `idx` (sorted position): does the exact box of its reference leaf pass the
// reference's AABB::intersect for this ray? (Asked only for a primitive whose routine reported a hit that would
// change the answer: "the reference tests the shape" AND "the shape is hit" commute.) Conservative test first; the reference's own arithmetic
// (out of line, rare) when that is too close to call, and always for rays with a component |d_i| <= 1e-6
// (KS = infinity, see the comment above TravState).
RT_DEV bool gate_passes(const BvhView& b, const TravState& s, int idx) {
    const float4 lo = ld_once(b.leafbox + 2 * (size_t)idx), hi = ld_once(b.leafbox + 2 * (size_t)idx + 1);
    if (!(s.KS > 3e38f)) {
        bool pass, sure;
        float ent;
        wide_child_test(s, lo.x, hi.x, lo.y, hi.y, lo.z, hi.z, 0.0f, pass, sure, ent);
        if (!pass) return false;  // the reference's test cannot pass (or the leaf starts beyond the best hit)
        if (sure) return true;    // ... cannot fail
    }
    return box_exact_call(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, s.r);
}

// The warp's primitive phase: every lane with pending primitives tests them, one class of
// primitive at a time across the warp (transformed shapes, then planes), so that the lanes of a
// warp run the same intersection routine together. MUST be called by all 32 lanes.
template <bool ANY, bool STATS>
RT_DEV void trav_prims(const BvhView& b, TravState& s, TraceStats& st) {
    bool occluded = false;
#pragma unroll 1
    for (int slot = 0; slot < RT_PEND_SLOTS; ++slot) {
        unsigned int m_x = 0u, m_p = 0u;
        const unsigned int mine = RT_PEND_SLOTS == 2 ? ((s.pend >> (4 * slot)) & 15u) : s.pend;
        const float* w = b.wide + (size_t)((RT_PEND_SLOTS == 2 && slot == 1) ? s.pend_node2 : s.pend_node) * 32;
        if (mine != 0u && !occluded) {
            const unsigned int t = __float_as_uint(__ldg(w + 25)) >> 16;  // 2 bits of type per child
            const unsigned int pl = ((t & 3u) == 3u ? 1u : 0u) | (((t >> 2) & 3u) == 3u ? 2u : 0u) | (((t >> 4) & 3u) == 3u ? 4u : 0u) |
                                    (((t >> 6) & 3u) == 3u ? 8u : 0u);
            m_p = mine & pl;
            m_x = mine & ~pl;
        }
        if (RT_PEND_SLOTS == 2 && !__any_sync(0xffffffffu, (m_x | m_p) != 0u)) continue;
        while (__any_sync(0xffffffffu, m_x != 0u)) {
            if (m_x != 0u) {
                const int k = __ffs(m_x) - 1;
                m_x &= m_x - 1u;
                const int idx = __float_as_int(__ldg(w + 27 + k));
                Hit h;
                if (STATS) st.prims++;
                // the routine first, the gate only for a hit that would change the answer (most candidates miss)
                if (intersect_prim<false, PRIM_XFORM>(b.prims, idx, s.r, h)) {
                    if (ANY) {
                        if (!(h.t > s.max_t) && gate_passes(b, s, idx)) { occluded = true; m_x = 0u; m_p = 0u; }
                    } else if ((h.t < s.best_t || (h.t == s.best_t && idx < s.best_prim)) && gate_passes(b, s, idx)) {
                        s.best_t = h.t; s.best_prim = idx; s.lim = prune_limit(s.best_t);
                    }
                }
            }
        }
        while (__any_sync(0xffffffffu, m_p != 0u)) {
            if (m_p != 0u) {
                const int k = __ffs(m_p) - 1;
                m_p &= m_p - 1u;
                const int idx = __float_as_int(__ldg(w + 27 + k));
                Hit h;
                if (STATS) st.prims++;
                // the routine first, the gate only for a hit that would change the answer (most candidates miss)
                if (intersect_prim<false, PRIM_PLANE>(b.prims, idx, s.r, h)) {
                    if (ANY) {
                        if (!(h.t > s.max_t) && gate_passes(b, s, idx)) { occluded = true; m_p = 0u; }
                    } else if ((h.t < s.best_t || (h.t == s.best_t && idx < s.best_prim)) && gate_passes(b, s, idx)) {
                        s.best_t = h.t; s.best_prim = idx; s.lim = prune_limit(s.best_t);
                    }
                }
            }
        }
    }
    s.pend = 0u;
    if (ANY && occluded) { s.best_prim = 0; s.cur = RT_CUR_NONE; s.sp = s.sp0; }
}

// Warp-level work distribution for persistent kernels: the warp owns a pool [pool_lo, pool_hi)
// of consecutive item indices (refilled with one atomic per RT_POOL items); every lane that needs
// an item gets one. Returns the item index or -1. `more` turns false when the global counter has
// passed n. All 32 lanes must call it together.
#define RT_POOL 128u
// Pool size for n items on this grid: 128 when there is plenty of work, down to 32 when there are
// few items, so that a small wave spreads over many warps instead of queueing behind a few.
RT_DEV unsigned int pool_size(unsigned long long n) {
    const unsigned long long warps = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    const unsigned long long per_warp = (n + warps - 1) / warps;
    return (unsigned int)min((unsigned long long)RT_POOL, max(32ull, (per_warp + 31ull) & ~31ull));
}
RT_DEV long long warp_take(unsigned int* counter, unsigned long long n, bool need, unsigned int& pool_lo, unsigned int& pool_hi,
                           bool& more, unsigned int pool = RT_POOL) {
    const int lane = threadIdx.x & 31;
    long long item = -1;
    unsigned int mask = __ballot_sync(0xffffffffu, need);
    while (mask != 0u && (pool_lo < pool_hi || more)) {
        if (pool_lo >= pool_hi) {  // warp-uniform: get the next pool
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(counter, pool);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((unsigned long long)base >= n) { more = false; break; }
            pool_lo = base;
            pool_hi = (unsigned int)min((unsigned long long)base + pool, n);
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
