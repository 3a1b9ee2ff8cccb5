// render.cu -- the hot path on the GPU (sm_100a), organised as a WAVEFRONT: rays of one kind are
// processed together by small kernels so that warps stay converged (the first version was a
// per-pixel megakernel: ncu showed 7.4 of 32 lanes active, see profiles/README.md).
//
// Replaces the reference frame loop raytracer.cpp:433-476 and everything it calls:
//   compute_pixel_color + Camera::pixelToRay_thin_lens (raytracer.cpp:18-70, camera.cpp:97-178)
//                                                     -> gen_kernel
//   BVH::get_intersection for view rays (acceleration.cpp:142-150)   -> trace_packet_kernel (level 0:
//                                                        one traversal per warp / pixel block),
//                                                        trace_kernel (deeper levels: per-ray loop)
//   Trace (raytracer.cpp:280-351): miss colour, reflection / refraction ray construction
//                                                     -> shade_kernel (emits the next wave)
//   shade (raytracer.cpp:180-274): shadow rays         -> shadow_packet_kernel / shadow_kernel
//                                  Blinn-Phong sum     -> light_kernel
//   gamma / clamp / quantise (raytracer.cpp:446-457)   -> finalize_kernel
//
// One batch = up to `batch_slots` (pixel, sample) pairs of this rank's screen tiles. For each
// recursion level d = 0..max_depth the four kernels run over the level's ray queue; kernels read
// their work size from device counters, so a whole frame is enqueued without host round trips;
// shadow + light of level d run on a second stream and overlap trace + shade of level d+1.
// The recursion tree of a sample is numbered (root 1, reflection 2k, refraction 2k+1); random
// numbers are keyed by (pixel, sample, node, light, shadow sample), so the image does not depend
// on queue order, batch size or tile sharding. Colour is accumulated per pixel in 64-bit fixed
// point (2^-40) with integer atomics: order-independent, hence run-to-run deterministic.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "api_internal.hpp"
#include "rt_device.cuh"

namespace rtb {

#define RT_MAX_DEPTH 16
#define RT_LVL_STRIDE 8
// Tunables of the traversal loop (wave_loop), overridable with -D for A/B runs:
//   RT_T_FETCH : idle lanes of a warp are given new rays once at least this many are idle
//                (refilling a single lane costs a warp-wide fetch + ray set-up for one ray;
//                waiting for all 32 leaves lanes idle behind the warp's longest ray);
//   RT_T_PRIM  : the warp runs its primitive phase once this many lanes wait with candidates;
//   RT_PHASE_MAJORITY : 1 = run the primitive phase whenever at least as many lanes wait for it
//                as can take a traversal step (instead of the fixed threshold).
#ifndef RT_T_FETCH
#define RT_T_FETCH 12
#endif
#ifndef RT_T_PRIM
#define RT_T_PRIM 12
#endif
#ifndef RT_PHASE_MAJORITY
#define RT_PHASE_MAJORITY 1
#endif
#ifndef RT_TRACE_THREADS
#define RT_TRACE_THREADS 128
#endif
#ifndef RT_TRACE_MINBLOCKS
#define RT_TRACE_MINBLOCKS 7  /* 72 registers: measured +2-3 % over 6 blocks / 80 registers on configs[1]-[3] (profiles/README.md) */
#endif
#ifndef RT_WAVE_MINBLOCKS
#define RT_WAVE_MINBLOCKS RT_TRACE_MINBLOCKS  /* resident blocks per SM of the per-ray kernels (trace_kernel, shadow_kernel) */
#endif
#ifndef RT_SHADE_MINBLOCKS
#define RT_SHADE_MINBLOCKS 4  /* resident 256-thread blocks per SM of shade_kernel / light_kernel (64 registers; 3 / 5 / 6 measured: shade -8 % at 4) */
#endif
#ifndef RT_ANY_SORTED_PACKET
#define RT_ANY_SORTED_PACKET 0
#endif
#ifndef RT_STEPS_PER_VOTE
#define RT_STEPS_PER_VOTE 3
#endif
// L_RAYS / L_RECS are queue SIZES (level 0 is not compacted: slot = packet * 32 + lane, with dead
// slots, so that a packet of 32 consecutive slots is one 8x4 pixel block); L_LIVE_* count the real
// rays / shade records for the statistics.
enum LevelCounter { L_RAYS = 0, L_RECS = 1, L_WORK_TRACE = 2, L_WORK_SHADOW = 3, L_LIVE_RAYS = 4, L_LIVE_RECS = 5 };
#define RT_DEAD 0xffffffffu  /* pixel field of a dead ray slot / invalid shade record */
#ifndef RT_SELF_OCCLUSION_SHADE
#define RT_SELF_OCCLUSION_SHADE 1
#endif
#define RT_VIS_SELF_OCCLUDED (-(1 << 30)) /* visibility counter of a (record, light) pair settled in shade_kernel */
enum Total { T_PRIMARY = 0, T_SHADOW = 1, T_SECONDARY = 2, T_NODES = 3, T_PRIMS = 4, T_OVERFLOW = 5 };

struct FrameParams {
    BvhView bvh;
    const float4* __restrict__ mats;    // 4 x float4 per material
    const float4* __restrict__ lights;  // 2 x float4 per light
    const DTexture* __restrict__ textures;
    const uint8_t* __restrict__ texels;
    int n_lights;
    // camera (camera.cpp)
    float cam_loc[3], xdir[3], ydir[3], zdir[3];
    float focal, half_sw, half_sh, aperture, focus_dist;
    int res_x, res_y;
    // sampling
    int samples_sqrt, spp, light_samples, max_depth;
    uint32_t seed_lo, seed_hi;
    float fixed_time;
    int shadow_per_rec;  // shadow rays per shaded hit = sum over lights of (radius > 0 ? light_samples : 1)
    // screen tiles of this rank (TilePlan): tiles[i] = index of the i-th tile this rank renders,
    // tile_slot[t] = i for those tiles and -1 for every other tile of the frame
    int tile_w, tile_h, tiles_x, n_tiles, rank, world, n_my_tiles;
    int sub_x, sub_per_tile;  // 8x4 pixel blocks per tile
    const int* __restrict__ tiles;
    const int* __restrict__ tile_slot;
    int win_x0, win_y0, win_x1, win_y1;  // only pixels inside this window are rendered (whole frame by default)
    int sort_emit;  // 1: shade_kernel orders the rays it emits by direction octant within its 256-ray runs
    int packed;  // 1: outputs are indexed tile-major over this rank's tiles (rt_render_multi) instead of frame row-major
    // wavefront buffers
    float4* q[2];                  // ray queues (ping-pong by level parity), 3 x float4 per ray
    int* hit_prim;                 // closest primitive per ray of the current level (-1 = miss)
    float4* recs[2];               // shade records, 5 x float4 per shaded hit; by level parity, so that
    int* vis[2];                   // shadow/light of level d can overlap trace/shade of level d+1
                                   // (vis = unoccluded shadow samples per (record, light))
    unsigned long long* accum;     // 3 per pixel, fixed point 2^-40
    int* hit_ids;                  // optional frame-sized output
    unsigned int* lvl;             // [RT_MAX_DEPTH + 2][RT_LVL_STRIDE] per-level counters
    unsigned long long* totals;    // [8] frame totals
    unsigned int* overflow_host;   // page-locked host word (mapped): set when a ray queue overflowed
    int capacity;                  // rays per queue
};

// Is pixel (x, y) rendered by this rank / inside the window, and where does its output go?
// Returns -1 for a pixel this rank does not render.
RT_DEV long long out_index(const FrameParams& p, int x, int y) {
    if (x < p.win_x0 || x >= p.win_x1 || y < p.win_y0 || y >= p.win_y1) return -1;
    const int tx = x / p.tile_w, ty = y / p.tile_h;
    const int slot = __ldg(p.tile_slot + ty * p.tiles_x + tx);
    if (slot < 0) return -1;
    if (!p.packed) return (long long)y * p.res_x + x;
    return ((long long)slot * p.tile_h + (y - ty * p.tile_h)) * p.tile_w + (x - tx * p.tile_w);
}

// Ray record: a = (origin, time)  b = (direction, weight)  c = bits(pixel, sample, node, 0)
// Shade record: r0 = (P, bits pixel) r1 = (N, weight * local share) r2 = (V, bits material)
//               r3 = (albedo rgb, 0) r4 = bits(sample, node, 0, 0)

#define FIXED_ONE 1099511627776.0f /* 2^40 */

RT_DEV void accumulate(const FrameParams& p, uint32_t pixel, float r, float g, float b) {
    const float lim = 4194304.0f;  // 2^22: keeps the 64-bit sum far from overflow
    r = fminf(fmaxf(r, -lim), lim); g = fminf(fmaxf(g, -lim), lim); b = fminf(fmaxf(b, -lim), lim);
    unsigned long long* a = p.accum + (size_t)pixel * 3;
    atomicAdd(a + 0, (unsigned long long)__float2ll_rn(r * FIXED_ONE));
    atomicAdd(a + 1, (unsigned long long)__float2ll_rn(g * FIXED_ONE));
    atomicAdd(a + 2, (unsigned long long)__float2ll_rn(b * FIXED_ONE));
}

// Reserves one slot for every lane with want == true, consecutive per warp, with one atomic.
RT_DEV unsigned int warp_reserve(unsigned int* counter, bool want) {
    const unsigned int mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0u) return 0u;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned int)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (unsigned int)__popc(mask & ((1u << lane) - 1u));
}

// Block-level version for 256-thread blocks: consecutive slots for every thread with want == true,
// ordered by thread index, with ONE atomic per block -- so that the queue keeps runs of up to 256
// rays from one 32x8 pixel region together (coherent pools for the next level's traversal).
// Optionally reserves a second class right behind the first (want2). `sh` = 20 words of shared
// memory. MUST be called by all threads of the block.
RT_DEV void block_reserve2(unsigned int* counter, bool want1, bool want2, unsigned int* sh, unsigned int& slot1, unsigned int& slot2) {
    const unsigned int m1 = __ballot_sync(0xffffffffu, want1), m2 = __ballot_sync(0xffffffffu, want2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[warp] = (unsigned int)__popc(m1); sh[8 + warp] = (unsigned int)__popc(m2); }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int w = 0; w < 16; ++w) { const unsigned int c = sh[w]; sh[w] = run; run += c; }  // exclusive prefix, class 1 then 2
        sh[16] = run ? atomicAdd(counter, run) : 0u;
    }
    __syncthreads();
    const unsigned int base = sh[16];
    const unsigned int lt = (1u << lane) - 1u;
    slot1 = base + sh[warp] + (unsigned int)__popc(m1 & lt);
    slot2 = base + sh[8 + warp] + (unsigned int)__popc(m2 & lt);
    __syncthreads();  // sh is reused by the next call
}

// Keyed version: the block's slots are ordered by (class, key, thread) -- class 1 before class 2, keys 0..7 within a
// class -- so that rays with the same key (direction octant) sit next to each other in the queue and end up in the
// same warps of the next level's traversal. `sh` = 132 words of shared memory. MUST be called by all threads of a
// 256-thread block.
RT_DEV void block_reserve_keyed(unsigned int* counter, bool want1, int key1, bool want2, int key2, unsigned int* sh,
                                unsigned int& slot1, unsigned int& slot2) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int lt = (1u << lane) - 1u;
    unsigned int rank1 = 0, rank2 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const unsigned int m1 = __ballot_sync(0xffffffffu, want1 && key1 == k), m2 = __ballot_sync(0xffffffffu, want2 && key2 == k);
        if (lane == 0) { sh[k * 8 + warp] = (unsigned int)__popc(m1); sh[64 + k * 8 + warp] = (unsigned int)__popc(m2); }
        if (want1 && key1 == k) rank1 = (unsigned int)__popc(m1 & lt);
        if (want2 && key2 == k) rank2 = (unsigned int)__popc(m2 & lt);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive prefix over the 128 (class, key, warp) counts: 4 per lane + a warp scan
        unsigned int c[4], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) { c[j] = sh[lane * 4 + j]; sum += c[j]; }
        unsigned int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
        unsigned int run = incl - sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) { sh[lane * 4 + j] = run; run += c[j]; }
        if (lane == 31) sh[128] = incl ? atomicAdd(counter, incl) : 0u;
    }
    __syncthreads();
    const unsigned int base = sh[128];
    slot1 = base + sh[(key1 & 7) * 8 + warp] + rank1;
    slot2 = base + sh[64 + (key2 & 7) * 8 + warp] + rank2;
    __syncthreads();  // sh is reused by the next call
}

RT_DEV int direction_octant(float x, float y, float z) { return (x < 0.0f ? 1 : 0) | (y < 0.0f ? 2 : 0) | (z < 0.0f ? 4 : 0); }

// ---------------------------------------------------------------------------------------------
// gen_kernel: primary rays of units [unit0, unit0 + n_units); a unit = (8x4 pixel block, sample),
// one lane per pixel. compute_pixel_color + Camera::pixelToRay_thin_lens.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gen_kernel(const __grid_constant__ FrameParams p, long long unit0, int n_units) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_units; w += warps) {
        const long long unit = unit0 + w;
        const int s = (int)(unit % p.spp);
        const long long blk = unit / p.spp;
        const int sub = (int)(blk % p.sub_per_tile);
        const int tile = __ldg(p.tiles + (int)(blk / p.sub_per_tile));
        const int lx = (sub % p.sub_x) * 8 + (lane & 7), ly = (sub / p.sub_x) * 4 + (lane >> 3);
        const int x = (tile % p.tiles_x) * p.tile_w + lx, y = (tile / p.tiles_x) * p.tile_h + ly;
        const bool valid = lx < p.tile_w && ly < p.tile_h && x < p.win_x1 && y < p.win_y1 && x >= p.win_x0 && y >= p.win_y0;
        const unsigned int slot = (unsigned int)w * 32u + (unsigned int)lane;  // one packet per unit, no compaction
        {
            const unsigned int live = __ballot_sync(0xffffffffu, valid);
            if (lane == 0) {
                if (live) atomicAdd(p.lvl + L_LIVE_RAYS, (unsigned int)__popc(live));
                if (w == 0) p.lvl[L_RAYS] = (unsigned int)n_units * 32u;
            }
        }
        if (!valid) {  // dead slot: zero direction, pixel = RT_DEAD
            float4* q = p.q[0] + (size_t)slot * 3;
            q[0] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            q[1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            q[2] = make_float4(__uint_as_float(RT_DEAD), 0.0f, 0.0f, 0.0f);
            continue;
        }
        const uint32_t pixel = (uint32_t)(y * p.res_x + x);
        const U4 u = rt_rng(pixel, p.seed_lo, p.seed_hi, (uint32_t)s, RNG_CAMERA, 0u, 0u, 0u);
        float fx, fy;
        if (p.samples_sqrt <= 1) {
            fx = (float)x + 0.5f;  // raytracer.cpp:33
            fy = (float)y + 0.5f;
        } else {
            // stratified jitter in double, narrowed to float by the tuple<float,float> (raytracer.cpp:50-58)
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
        float ox = p.cam_loc[0], oy = p.cam_loc[1], oz = p.cam_loc[2];
        if (p.aperture > 0.0f) {  // thin lens (camera.cpp:141-177)
            const float fpx = p.cam_loc[0] + dx * p.focus_dist;
            const float fpy = p.cam_loc[1] + dy * p.focus_dist;
            const float fpz = p.cam_loc[2] + dz * p.focus_dist;
            float rx = 0.0f, ry = 0.0f;
            for (uint32_t attempt = 0;; ++attempt) {  // random_in_unit_disk (camera.cpp:89-95)
                const U4 l = rt_rng(pixel, p.seed_lo, p.seed_hi, (uint32_t)s, RNG_LENS, 0u, 0u, attempt);
                rx = u32_to_unit_float(l.x) * 2.0f - 1.0f;
                ry = u32_to_unit_float(l.y) * 2.0f - 1.0f;
                if (rx * rx + ry * ry < 1.0f || attempt >= 63u) break;
            }
            const float lens_radius = p.aperture / 2.0f;
            rx *= lens_radius;
            ry *= lens_radius;
            ox = p.cam_loc[0] + (p.xdir[0] * rx + p.ydir[0] * ry);
            oy = p.cam_loc[1] + (p.xdir[1] * rx + p.ydir[1] * ry);
            oz = p.cam_loc[2] + (p.xdir[2] * rx + p.ydir[2] * ry);
            dx = fpx - ox; dy = fpy - oy; dz = fpz - oz;
            normalize3(dx, dy, dz);
        }
        const float time = (p.fixed_time >= 0.0f) ? p.fixed_time : u32_to_unit_float(u.z);  // raytracer.cpp:37,61
        float4* q = p.q[0] + (size_t)slot * 3;
        q[0] = make_float4(ox, oy, oz, time);
        q[1] = make_float4(dx, dy, dz, 1.0f);
        q[2] = make_float4(__uint_as_float(pixel), __uint_as_float((uint32_t)s), __uint_as_float(1u), 0.0f);
    }
}

// ---------------------------------------------------------------------------------------------
// wave_loop: the traversal loop shared by trace_kernel (closest hit) and shadow_kernel (any hit).
// Persistent warps. Every iteration the warp votes and does ONE of three things together:
//   fetch : lanes without a ray take the next rays of the warp's pool (one atomic per 128 rays);
//   step  : lanes with a node to visit test its four children (trav_step: uniform code, 128 B
//           node, four conservative slab tests, push / pop on the shared-memory stack);
//   prims : lanes that reached a reference leaf with candidate primitives run the exact
//           intersection routines, class by class (trav_prims).
// Lanes waiting for another phase idle for that iteration; the thresholds above bound how long.
// Src supplies the rays: load(item, ray, max_t) and store(item, state).
// ---------------------------------------------------------------------------------------------
extern __shared__ __align__(8) int rt_stack_smem[];

template <bool ANY, bool STATS, class Src>
RT_DEV void wave_loop(const BvhView& bvh, Src& src, unsigned int* counter, unsigned long long n, TraceStats& st) {
    const unsigned int FULL = 0xffffffffu;
    const unsigned int words = ANY ? 1u : (unsigned int)RT_STACK_WORDS;
    const unsigned int stride = blockDim.x * words * (unsigned int)sizeof(int);
    TravState s;
    s.cur = RT_CUR_NONE;
    s.pend = 0u;
    s.sp0 = (unsigned int)__cvta_generic_to_shared(rt_stack_smem + threadIdx.x * words);
    s.sp = s.sp0;
    s.sp_end = s.sp0 + (unsigned int)bvh.stack_depth * stride;
#if RT_STAGE_TOP > 0
    {   // the top of the tree, once per block, behind the stacks
        float4* staged = reinterpret_cast<float4*>(rt_stack_smem + (size_t)bvh.stack_depth * blockDim.x * words);
        for (int i = threadIdx.x; i < bvh.n_staged * 8; i += blockDim.x) staged[i] = __ldg(reinterpret_cast<const float4*>(bvh.wide) + i);
        s.staged = (unsigned int)__cvta_generic_to_shared(staged);
        __syncthreads();
    }
#endif
    long long item = -1;
    unsigned int pool_lo = 0, pool_hi = 0;
    bool more = true;
    const unsigned int pool = pool_size(n);
    while (true) {
        const unsigned int idle = __ballot_sync(FULL, item < 0);
        if (idle != 0u) {
            const bool work = more || pool_lo < pool_hi;  // warp-uniform
            if (!work) {
                if (idle == FULL) break;
            } else if (idle == FULL || __popc(idle) >= RT_T_FETCH) {
                const long long got = warp_take(counter, n, item < 0, pool_lo, pool_hi, more, pool);
                if (item < 0 && got >= 0) {
                    item = got;
                    Ray r;
                    float max_t;
                    if (!src.load(item, r, max_t)) item = -1;
                    else if (trav_begin<ANY>(bvh, s, r, max_t)) { src.store(item, s); item = -1; }
                }
                continue;
            }
        }
        const bool has_pend = s.pend != 0u;
#if RT_PEND_SLOTS == 2
        // a lane may step past ONE node with candidates: it only has to stop when both slots are taken (or it has no
        // node left); the primitive phase runs when at least as many lanes are stopped by it as can step
        const bool full = (s.pend & 15u) != 0u && (s.pend & 240u) != 0u;
        const bool can_step = item >= 0 && !full && s.cur != RT_CUR_NONE;
        const bool blocked = has_pend && !can_step;
        const unsigned int pm = __ballot_sync(FULL, blocked), sm = __ballot_sync(FULL, can_step);
        const bool do_prims = pm != 0u && (sm == 0u || __popc(pm) >= __popc(sm));
        if (do_prims) trav_prims<ANY, STATS>(bvh, s, st);
        else if (can_step) {
#pragma unroll 1
            for (int k = 0; k < RT_STEPS_PER_VOTE && !((s.pend & 15u) != 0u && (s.pend & 240u) != 0u) && s.cur != RT_CUR_NONE; ++k)
                trav_step<ANY, STATS>(bvh, s, stride, st);
        }
#else
        const bool can_step = item >= 0 && !has_pend && s.cur != RT_CUR_NONE;
        const unsigned int pm = __ballot_sync(FULL, has_pend), sm = __ballot_sync(FULL, can_step);
        const bool do_prims = pm != 0u && (sm == 0u || (RT_PHASE_MAJORITY ? __popc(pm) >= __popc(sm) : __popc(pm) >= RT_T_PRIM));
        if (do_prims) trav_prims<ANY, STATS>(bvh, s, st);
        else if (can_step) {
#pragma unroll 1
            for (int k = 0; k < RT_STEPS_PER_VOTE && s.pend == 0u && s.cur != RT_CUR_NONE; ++k)
                trav_step<ANY, STATS>(bvh, s, stride, st);
        }
#endif
        if (item >= 0 && s.pend == 0u && s.cur == RT_CUR_NONE) { src.store(item, s); item = -1; }
    }
}

// ---------------------------------------------------------------------------------------------
// packet_loop: traversal for COHERENT waves (the view rays of an 8x4 pixel block, and the shadow
// rays of its hit points towards one light). The 32 rays of a warp share ONE traversal: one
// stack (per warp, in shared memory), one node per step -- the node load is a broadcast, the
// stack and the order are warp-uniform, and at a leaf all lanes test the SAME primitive, so the
// intersection routine runs converged whatever the mix of primitive types in the scene.
// Every stack entry carries the mask of lanes that passed the child's box (and, for a gated
// child, the reference's leaf test), so each lane visits exactly the nodes its own per-ray
// traversal could visit and tests a primitive only if its own culling box and its own leaf gate
// passed: per lane the result is the same (t, shape) as wave_loop's. Lanes outside the mask idle;
// the packet pays off while the rays stay together (measured: see profiles/README.md).
// Stack entry = (node, lane mask, key): key = smallest entry parameter among the lanes (closest
// hit only): a popped sub-tree is dropped for lanes whose best hit is nearer.
// ---------------------------------------------------------------------------------------------
template <bool ANY, bool STATS, class Src>
RT_DEV void packet_loop(const BvhView& bvh, Src& src, unsigned int* counter, unsigned long long n, TraceStats& st) {
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // per-warp stack of 16-byte entries (node, lane mask, key, -): one STS.128 / one broadcast LDS.128
    const unsigned int stk0 = (unsigned int)__cvta_generic_to_shared(rt_stack_smem) + (threadIdx.x >> 5) * bvh.packet_stack_depth * 16u;
    TravState s;
    s.sp0 = 0u;
    s.imax = 0.0f;
    while (true) {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(counter, 32u);
        base = __shfl_sync(FULL, base, 0);
        if ((unsigned long long)base >= n) break;
        const long long item = (long long)base + lane;
        bool active = (unsigned long long)item < n;
        if (active) {
            Ray r;
            float max_t;
            if (!src.load(item, r, max_t)) active = false;
            else if (trav_begin<ANY>(bvh, s, r, max_t)) { src.store(item, s); active = false; }
        }
        const bool traversing = active;
        unsigned int alive = __ballot_sync(FULL, active);  // lanes still looking for an answer
        unsigned int sp = stk0;
        int cur = 0;
        unsigned int curmask = alive;
        while (alive != 0u) {
            // ---- visit node `cur` with the lanes of curmask ----
            const float* w = bvh.wide + (size_t)cur * 32;
            const F8 X = ldg256(w), Y = ldg256(w + 8), Z = ldg256(w + 16), C = ldg256(w + 24);
            const bool in = (curmask >> lane) & 1u;
            if (STATS && in) st.nodes += 4;
            const int first = __float_as_int(C.v[0]);
            const unsigned int meta = __float_as_uint(C.v[1]);
            const float qi = C.v[2] * s.imax;
            bool pass[4];
            float ent[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                bool sure;
                wide_child_test(s, X.v[k], X.v[4 + k], Y.v[k], Y.v[4 + k], Z.v[k], Z.v[4 + k], qi, pass[k], sure, ent[k]);
            }
            unsigned int pm = ((pass[0] ? 1u : 0u) | (pass[1] ? 2u : 0u) | (pass[2] ? 4u : 0u) | (pass[3] ? 8u : 0u)) & meta;
            if (!in) pm = 0u;
            unsigned int b0 = __ballot_sync(FULL, pm & 1u), b1 = __ballot_sync(FULL, pm & 2u);
            unsigned int b2 = __ballot_sync(FULL, pm & 4u), b3 = __ballot_sync(FULL, pm & 8u);
            const unsigned int primmask = (meta >> 4) & 15u;  // warp-uniform: which children are primitives
            bool descended = false;
            if (primmask != 0u) {
                // primitive children: every candidate is tested by all its lanes together, each lane behind its own gate
                // (gate first here: the lanes are converged anyway, and it spares the routine; the per-ray loop asks the
                // gate only after a hit -- measured both ways, profiles/ab_r2h_ab.jsonl)
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    if (!((primmask >> k) & 1u)) continue;
                    const unsigned int bk = (k == 0 ? b0 : (k == 1 ? b1 : (k == 2 ? b2 : b3))) & alive;
                    if (bk == 0u) continue;
                    const int idx = __float_as_int(k == 0 ? C.v[3] : (k == 1 ? C.v[4] : (k == 2 ? C.v[5] : C.v[6])));
                    const bool mine = ((bk >> lane) & 1u) && gate_passes(bvh, s, idx);
                    const unsigned int type = (meta >> (16 + 2 * k)) & 3u;  // warp-uniform
                    Hit h;
                    bool hit = false;
                    if (type == RT_PLANE) { if (mine) hit = intersect_prim<false, PRIM_PLANE>(bvh.prims, idx, s.r, h); }
                    else { if (mine) hit = intersect_prim<false, PRIM_XFORM>(bvh.prims, idx, s.r, h); }
                    if (STATS && mine) st.prims++;
                    if (hit) {
                        if (ANY) { if (!(h.t > s.max_t)) { s.best_prim = 0; active = false; } }
                        else if (h.t < s.best_t || (h.t == s.best_t && idx < s.best_prim)) {
                            s.best_t = h.t; s.best_prim = idx; s.lim = prune_limit(s.best_t);
                        }
                    }
                    if (ANY) alive = __ballot_sync(FULL, active);
                }
                if (primmask & 1u) b0 = 0u;
                if (primmask & 2u) b1 = 0u;
                if (primmask & 4u) b2 = 0u;
                if (primmask & 8u) b3 = 0u;
            }
            if (ANY && !RT_ANY_SORTED_PACKET) {
                // occlusion packets take the inner children in SLOT order (bvh.cpp chooses it at build time): no keys, no
                // sorting network -- the first child with lanes is visited next, the others are pushed so that they
                // pop in slot order
                if (((b0 | b1 | b2 | b3) & alive) != 0u) {
                    const int fs = b0 != 0u ? 0 : (b1 != 0u ? 1 : (b2 != 0u ? 2 : 3));  // warp-uniform
#pragma unroll
                    for (int k = 3; k >= 1; --k) {
                        const unsigned int bk = k == 1 ? b1 : (k == 2 ? b2 : b3);
                        if (bk != 0u && k != fs) {
                            if (lane == 0)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %3};" :: "r"(sp), "r"(first + k), "r"(bk), "r"(0) : "memory");
                            sp += 16u;
                            if (RT_CHECKS && sp > stk0 + (unsigned int)bvh.packet_stack_depth * 16u) __trap();
                        }
                    }
                    cur = first + fs;
                    curmask = (fs == 0 ? b0 : (fs == 1 ? b1 : (fs == 2 ? b2 : b3))) & alive;
                    descended = curmask != 0u;
                }
            } else if (((b0 | b1 | b2 | b3) & alive) != 0u) {
                // inner children, ordered by the packet's smallest entry parameter
                int key[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned int bk = k == 0 ? b0 : (k == 1 ? b1 : (k == 2 ? b2 : b3));
                    int v = 0x7fffffff;
                    if (bk != 0u) {
                        if (ANY && !RT_ANY_SORTED_PACKET) v = k;  // slot order
                        else {
                            const int mine = (((pm >> k) & 1u) != 0u) ? __float_as_int(fmaxf(ent[k], 0.0f)) : 0x7fffffff;
                            v = (__reduce_min_sync(FULL, mine) & ~3) | k;
                        }
                    }
                    key[k] = v;
                }
#define RT_CSWAP(i, j) { const int lo_ = min(key[i], key[j]), hi_ = max(key[i], key[j]); key[i] = lo_; key[j] = hi_; }
                RT_CSWAP(0, 1) RT_CSWAP(2, 3) RT_CSWAP(0, 2) RT_CSWAP(1, 3) RT_CSWAP(1, 2)
#undef RT_CSWAP
#pragma unroll
                for (int j = 3; j >= 1; --j) {
                    if (key[j] != 0x7fffffff) {
                        const int slot = key[j] & 3;
                        const unsigned int bk = slot == 0 ? b0 : (slot == 1 ? b1 : (slot == 2 ? b2 : b3));
                        if (lane == 0)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %3};" :: "r"(sp), "r"(first + slot), "r"(bk), "r"(key[j]) : "memory");
                        sp += 16u;
                        if (RT_CHECKS && sp > stk0 + (unsigned int)bvh.packet_stack_depth * 16u) __trap();
                    }
                }
                const int slot = key[0] & 3;
                cur = first + slot;
                curmask = (slot == 0 ? b0 : (slot == 1 ? b1 : (slot == 2 ? b2 : b3))) & alive;
                descended = curmask != 0u;
            }
            // ---- pop until an entry still has lanes that need it ----
            __syncwarp();
            while (!descended && sp != stk0) {
                sp -= 16u;
                int node, key, pad;
                unsigned int mask;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(node), "=r"(mask), "=r"(key), "=r"(pad) : "r"(sp) : "memory");
                mask &= alive;
                if (!ANY) mask &= __ballot_sync(FULL, __int_as_float(key & ~3) <= s.lim);
                if (mask != 0u) { cur = node; curmask = mask; descended = true; }
            }
            if (!descended) break;
        }
        if (traversing) src.store(item, s);
    }
}

// ---------------------------------------------------------------------------------------------
// Literal mode (rt_render_params.reserved[0] bit 0, and -bvh off): the reference's own traversal, one thread
// per ray, nothing shared with the production loops but the intersection routines and the ray sources --
// BVH::intersect over the reference's binary tree with exact box tests (traverse_reference_impl), or
// BVH::intersect_linear (traverse_linear_impl). The parity tests hold the production path to this one bit
// for bit at full benchmark sizes.
// ---------------------------------------------------------------------------------------------
template <bool ANY, bool STATS, class Src>
RT_DEV void literal_loop(const BvhView& bvh, Src& src, unsigned long long n, TraceStats& st) {
    for (unsigned long long item = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; item < n; item += (unsigned long long)gridDim.x * blockDim.x) {
        Ray r;
        float max_t;
        if (!src.load((long long)item, r, max_t)) continue;
        TravState s;
        s.best_prim = -1;
        s.best_t = FLT_MAX;
        if (bvh.n_prims > 0) {
            unsigned int boxes = 0;
            const int4 v = bvh.use_bvh ? traverse_reference_impl<ANY>(bvh.prims, bvh.ref_tree, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, r.time, max_t, &boxes)
                                       : traverse_linear_impl<ANY>(bvh.prims, bvh.n_prims, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, r.time, max_t);
            s.best_prim = ANY ? (v.x ? 0 : -1) : v.y;
            s.best_t = __int_as_float(v.z);
            if (STATS) { st.prims += (unsigned int)v.w; st.nodes += boxes; }
        }
        src.store((long long)item, s);
    }
}

RT_DEV void flush_stats(const FrameParams& p, const TraceStats& st) {
    unsigned long long a = st.nodes, b = st.prims;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(p.totals + T_NODES, a); atomicAdd(p.totals + T_PRIMS, b); }
}

// ---------------------------------------------------------------------------------------------
// trace_kernel: closest hit for every ray of the level (BVH::get_intersection for view rays).
// ---------------------------------------------------------------------------------------------
struct ViewRays {
    const float4* __restrict__ q;
    int* hit_prim;
    // false: a dead slot (padding of a level-0 packet)
    RT_DEV bool load(long long item, Ray& r, float& max_t) const {
        const float4 a = ld_once_rw(q + (size_t)item * 3 + 0), b = ld_once_rw(q + (size_t)item * 3 + 1);
        r.ox = a.x; r.oy = a.y; r.oz = a.z; r.time = a.w;
        r.dx = b.x; r.dy = b.y; r.dz = b.z;
        max_t = 0.0f;
        return !(b.x == 0.0f && b.y == 0.0f && b.z == 0.0f);
    }
    RT_DEV void store(long long item, const TravState& s) const { hit_prim[item] = s.best_prim; }
};

template <bool STATS>
__global__ void __launch_bounds__(RT_TRACE_THREADS, RT_WAVE_MINBLOCKS) trace_kernel(const __grid_constant__ FrameParams p, int level) {
    unsigned int* lv = p.lvl + level * RT_LVL_STRIDE;
    const unsigned int n = min(lv[L_RAYS], (unsigned int)p.capacity);
    TraceStats st = {0u, 0u};
    ViewRays src = {p.q[level & 1], p.hit_prim};
    wave_loop<false, STATS>(p.bvh, src, lv + L_WORK_TRACE, n, st);
    if (STATS) flush_stats(p, st);
}

template <bool STATS>
__global__ void __launch_bounds__(128) trace_literal_kernel(const __grid_constant__ FrameParams p, int level) {
    unsigned int* lv = p.lvl + level * RT_LVL_STRIDE;
    const unsigned int n = min(lv[L_RAYS], (unsigned int)p.capacity);
    TraceStats st = {0u, 0u};
    ViewRays src = {p.q[level & 1], p.hit_prim};
    literal_loop<false, STATS>(p.bvh, src, n, st);
    if (STATS) flush_stats(p, st);
}

// Material::getDiffuseColor (material.hpp:99-134)
RT_DEV void diffuse_color(const FrameParams& p, const float4 m0, int tex, float u, float v, float& r, float& g, float& b) {
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

// ---------------------------------------------------------------------------------------------
// shade_kernel: per ray of the level -- background for misses; for hits the shade record (for the
// shadow and light kernels) and the reflection / refraction rays of the next level.
// Trace() raytracer.cpp:280-351, createReflectionRay :101-115, createRefractionRay :118-150.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, RT_SHADE_MINBLOCKS) shade_kernel(const __grid_constant__ FrameParams p, int level) {
    unsigned int* lv = p.lvl + level * RT_LVL_STRIDE;
    unsigned int* lv_next = p.lvl + (level + 1) * RT_LVL_STRIDE;
    const unsigned int n = min(lv[L_RAYS], (unsigned int)p.capacity);
    const float4* __restrict__ q = p.q[level & 1];
    float4* __restrict__ qn = p.q[(level + 1) & 1];
    __shared__ unsigned int sh_reserve[132];
    const unsigned int n_round = (n + 255u) & ~255u;  // whole 256-thread blocks take part in block_reserve2
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool live = i < n;
        int prim = -1;
        float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
        if (live) {
            a = q[(size_t)i * 3 + 0]; b = q[(size_t)i * 3 + 1]; c = q[(size_t)i * 3 + 2];
            prim = p.hit_prim[i];
        }
        const uint32_t pixel = __float_as_uint(c.x), sample = __float_as_uint(c.y), node = __float_as_uint(c.z);
        const float weight = b.w;
        const bool dead = pixel == RT_DEAD;  // padding slot of a level-0 packet
        if (live && !dead && prim < 0) {
            const float bg = weight * 0.1f;  // background {0.1,0.1,0.1} (raytracer.cpp:297)
            accumulate(p, pixel, bg, bg, bg);
            if (node == 1u && sample == 0u && p.hit_ids) p.hit_ids[out_index(p, (int)(pixel % (uint32_t)p.res_x), (int)(pixel / (uint32_t)p.res_x))] = -1;
        }
        const bool hit = live && !dead && prim >= 0;
        Ray r;
        r.ox = a.x; r.oy = a.y; r.oz = a.z; r.time = a.w;
        r.dx = b.x; r.dy = b.y; r.dz = b.z;
        Hit h;
        h.px = h.py = h.pz = h.nx = h.ny = h.nz = h.u = h.v = h.t = 0.0f;
        float4 m2 = make_float4(0, 0, 0, 0), m3 = m2;
        int mat = 0;
        if (hit) {
            intersect_prim<true>(p.bvh.prims, prim, r, h);
            const uint32_t tag = __float_as_uint(__ldg(p.bvh.prims + (size_t)prim * 8).w);
            mat = (int)(tag >> 2);
            m2 = __ldg(p.mats + 4 * mat + 2);
            m3 = __ldg(p.mats + 4 * mat + 3);
            if (node == 1u && sample == 0u && p.hit_ids)
                p.hit_ids[out_index(p, (int)(pixel % (uint32_t)p.res_x), (int)(pixel / (uint32_t)p.res_x))] =
                    __float_as_int(__ldg(p.bvh.prims + (size_t)prim * 8 + 7).x);
        }
        const float roughness = m2.z, reflectivity = m2.w, transparency = m3.x, ior = m3.y;

        // shade record. Level 0 keeps record i for ray i (misses leave an invalid record) so that the
        // shadow rays of a packet belong to one pixel block; deeper levels compact.
        unsigned int rec;
        if (level == 0) {
            rec = i;
            const unsigned int hits = __ballot_sync(0xffffffffu, hit);
            if ((threadIdx.x & 31) == 0 && hits) atomicAdd(lv + L_LIVE_RECS, (unsigned int)__popc(hits));
            if (i == 0) lv[L_RECS] = n;
            if (live && !hit) p.recs[0][(size_t)rec * 5] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(RT_DEAD));
        } else {
            unsigned int unused;
            block_reserve2(lv + L_RECS, hit, false, sh_reserve, rec, unused);  // compact: the queue size is the live count
        }
        if (hit) {
            const float4 m0 = __ldg(p.mats + 4 * mat + 0);
            float br, bg, bb;
            diffuse_color(p, m0, __float_as_int(m3.z), h.u, h.v, br, bg, bb);
            float vx = r.ox - h.px, vy = r.oy - h.py, vz = r.oz - h.pz;  // view vector (raytracer.cpp:197)
            normalize3(vx, vy, vz);
            const float local_share = fmaxf(0.0f, 1.0f - reflectivity - transparency);  // raytracer.cpp:346
            float4* o = p.recs[level & 1] + (size_t)rec * 5;
            o[0] = make_float4(h.px, h.py, h.pz, __uint_as_float(pixel));
            o[1] = make_float4(h.nx, h.ny, h.nz, weight * local_share);
            o[2] = make_float4(vx, vy, vz, __int_as_float(mat));
            o[3] = make_float4(br, bg, bb, __int_as_float(prim));  // .w: the shape the point lies on (sorted position)
            o[4] = make_float4(__uint_as_float(sample), __uint_as_float(node), 0.0f, 0.0f);
            for (int l = 0; l < p.n_lights; ++l) {
                int v0 = 0;
#if RT_SELF_OCCLUSION_SHADE
                // Point lights: a shadow ray that leaves the surface inwards (N . L < 0) is almost always
                // stopped by the shape it starts on. Decide that HERE, where one thread per hit runs
                // converged: the exact routine on that shape + the exact box test of its reference leaf
                // (the reference tests the shape iff that box passes). A hit closer than the light
                // marks the (record, light) pair as occluded (negative counter) and the shadow kernel
                // never sees the ray; a miss changes nothing. Same ray arithmetic as ShadowRaysT::load.
                const float4 l0 = __ldg(p.lights + 2 * l), l1 = __ldg(p.lights + 2 * l + 1);
                if (l1.w <= 0.0f && p.bvh.prune) {
                    float lx = l0.x - h.px, ly = l0.y - h.py, lz = l0.z - h.pz;
                    const float light_dist = sqrtf(dot3(lx, ly, lz, lx, ly, lz));
                    normalize3(lx, ly, lz);
                    if (dot3(h.nx, h.ny, h.nz, lx, ly, lz) < 0.0f) {
                        Ray sr;
                        sr.ox = h.px + h.nx * 1e-4f; sr.oy = h.py + h.ny * 1e-4f; sr.oz = h.pz + h.nz * 1e-4f;
                        sr.dx = lx; sr.dy = ly; sr.dz = lz;
                        sr.time = 0.0f;
                        Hit sh;
                        if (intersect_prim<false>(p.bvh.prims, prim, sr, sh) && !(sh.t > light_dist)) {
                            const float4 blo = __ldg(p.bvh.leafbox + 2 * (size_t)prim), bhi = __ldg(p.bvh.leafbox + 2 * (size_t)prim + 1);
                            if (!p.bvh.use_bvh || box_exact_call(blo.x, blo.y, blo.z, bhi.x, bhi.y, bhi.z, sr)) v0 = RT_VIS_SELF_OCCLUDED;
                        }
                    }
                }
#endif
                p.vis[level & 1][(size_t)rec * p.n_lights + l] = v0;
            }
        }

        const bool deeper = level + 1 <= p.max_depth;
        // reflection (glossy when roughness > 0, raytracer.cpp:308-333)
        bool want_refl = false;
        float rx = 0, ry = 0, rz = 0;
        if (hit && deeper && reflectivity > 0.0f) {
            const float k = 2.0f * dot3(r.dx, r.dy, r.dz, h.nx, h.ny, h.nz);
            rx = r.dx - h.nx * k; ry = r.dy - h.ny * k; rz = r.dz - h.nz * k;
            if (roughness > 0.0f) {
                RngCtx g = {pixel, p.seed_lo, p.seed_hi, sample};
                float fx, fy, fz;
                random_in_unit_sphere(g, RNG_GLOSSY, node, 0u, fx, fy, fz);
                rx = rx + fx * roughness; ry = ry + fy * roughness; rz = rz + fz * roughness;
                normalize3(rx, ry, rz);
                if (dot3(rx, ry, rz, h.nx, h.ny, h.nz) < 0.0f) { rx = 0.0f; ry = 0.0f; rz = 0.0f; }
            }
            want_refl = dot3(rx, ry, rz, rx, ry, rz) > 0.001f;
        }
        // refraction (raytracer.cpp:336-344)
        bool want_refr = false;
        float tx = 0, ty = 0, tz = 0, fnx = 0, fny = 0, fnz = 0;
        if (hit && deeper && transparency > 0.0f) {
            fnx = h.nx; fny = h.ny; fnz = h.nz;
            float n_in = 1.0f, n_out = ior;
            const float cos_i = dot3(r.dx, r.dy, r.dz, fnx, fny, fnz);
            if (cos_i > 0.0f) { const float tmp = n_in; n_in = n_out; n_out = tmp; fnx = fnx * -1.0f; fny = fny * -1.0f; fnz = fnz * -1.0f; }
            const float eta = n_in / n_out;
            const float cos_abs = fabsf(cos_i);
            const float disc = 1.0f - eta * eta * (1.0f - cos_abs * cos_abs);
            if (!(disc < 0.0f)) {
                const float cos_t = sqrtf(disc);
                const float k = eta * cos_abs - cos_t;
                tx = r.dx * eta + fnx * k; ty = r.dy * eta + fny * k; tz = r.dz * eta + fnz * k;
                normalize3(tx, ty, tz);
                want_refr = dot3(tx, ty, tz, tx, ty, tz) > 1e-6f;
            }
        }
        unsigned int s_refl, s_refr;  // the block's reflection rays, then its refraction rays
        if (p.sort_emit) block_reserve_keyed(lv_next + L_RAYS, want_refl, direction_octant(rx, ry, rz), want_refr, direction_octant(tx, ty, tz), sh_reserve, s_refl, s_refr);
        else block_reserve2(lv_next + L_RAYS, want_refl, want_refr, sh_reserve, s_refl, s_refr);
        if (want_refl) {
            if (s_refl < (unsigned int)p.capacity) {
                float4* o = qn + (size_t)s_refl * 3;  // secondary rays carry the default time 0 (shapes.hpp:28)
                o[0] = make_float4(h.px + h.nx * 1e-4f, h.py + h.ny * 1e-4f, h.pz + h.nz * 1e-4f, 0.0f);
                o[1] = make_float4(rx, ry, rz, weight * reflectivity);
                o[2] = make_float4(c.x, c.y, __uint_as_float(node * 2u), 0.0f);
            } else {
                p.totals[T_OVERFLOW] = 1ull;
                *p.overflow_host = 1u;
            }
        }
        if (want_refr) {
            if (s_refr < (unsigned int)p.capacity) {
                float4* o = qn + (size_t)s_refr * 3;
                o[0] = make_float4(h.px + fnx * -1e-4f, h.py + fny * -1e-4f, h.pz + fnz * -1e-4f, 0.0f);
                o[1] = make_float4(tx, ty, tz, weight * transparency);
                o[2] = make_float4(c.x, c.y, __uint_as_float(node * 2u + 1u), 0.0f);
            } else {
                p.totals[T_OVERFLOW] = 1ull;
                *p.overflow_host = 1u;
            }
        }
    }
}

template <bool STATS>
__global__ void __launch_bounds__(RT_TRACE_THREADS, RT_TRACE_MINBLOCKS) trace_packet_kernel(const __grid_constant__ FrameParams p, int level) {
    unsigned int* lv = p.lvl + level * RT_LVL_STRIDE;
    const unsigned int n = min(lv[L_RAYS], (unsigned int)p.capacity);
    TraceStats st = {0u, 0u};
    ViewRays src = {p.q[level & 1], p.hit_prim};
    packet_loop<false, STATS>(p.bvh, src, lv + L_WORK_TRACE, n, st);
    if (STATS) flush_stats(p, st);
}

// ---------------------------------------------------------------------------------------------
// shadow_kernel: one lane per shadow ray = (light, shade record, light sample), LIGHT-MAJOR: the
// rays of a warp go to the same light from neighbouring surface points (or, for an area light,
// from one point to neighbouring targets), so they traverse the same part of the tree.
// shade() raytracer.cpp:201-236.
// ---------------------------------------------------------------------------------------------
// RT_SELF_OCCLUSION: a shadow ray that leaves its surface towards the inside (N . L < 0) is almost
// always stopped by the very shape it starts on. That shape is tested FIRST, with the exact routine
// and the exact box test of its reference leaf (the reference tests it iff that box passes): a hit
// closer than the light settles the query without any traversal. A miss changes nothing -- the
// ray is traversed as usual -- so the result is the reference's either way. Not used in the
// literal validation mode (prune = 0), which therefore also validates this shortcut.
// Used for AREA lights in the packet kernel only (>= 8 samples per light: the samples of one point
// take the branch together and whole packets vanish: +16 % on configs[2]); for point lights the
// test diverges inside the fetch phase and costs more than the short traversal it saves
// (1862 vs 1964 Mrays/s on configs[1]).
#ifndef RT_SELF_OCCLUSION
#define RT_SELF_OCCLUSION 1
#endif
template <bool SELF>
struct ShadowRaysT {
    const FrameParams& p;
    const float4* __restrict__ recs;  // this level's shade records
    int* vis;
    unsigned int n_recs;
    int* vis_slot;
    // false: the record is invalid (level 0: the view ray missed)
    RT_DEV bool load(long long item, Ray& sr, float& max_t) {
        unsigned long long rest = (unsigned long long)item;
        int li = 0, cnt = 1;
        float4 l0, l1;
        while (true) {  // which light's block of n_recs * cnt items
            l0 = __ldg(p.lights + 2 * li);
            l1 = __ldg(p.lights + 2 * li + 1);
            cnt = (l1.w > 0.0f) ? p.light_samples : 1;
            const unsigned long long block = (unsigned long long)n_recs * (unsigned int)cnt;
            if (rest < block) break;
            rest -= block;
            ++li;
        }
        const unsigned int rec = (unsigned int)(rest / (unsigned int)cnt);
        const int k = (int)(rest % (unsigned int)cnt);
        const float4 r0 = ld_once_rw(recs + (size_t)rec * 5 + 0);
        if (__float_as_uint(r0.w) == RT_DEAD) return false;
        if (RT_SELF_OCCLUSION_SHADE && vis[(size_t)rec * p.n_lights + li] < 0) return false;  // settled in shade_kernel
        const float4 r1 = ld_once_rw(recs + (size_t)rec * 5 + 1);
        float tx = l0.x, ty = l0.y, tz = l0.z;
        const float radius = l1.w;
        if (radius > 0.0f) {
            const float4 r4 = recs[(size_t)rec * 5 + 4];
            RngCtx g = {__float_as_uint(r0.w), p.seed_lo, p.seed_hi, __float_as_uint(r4.x)};
            float rx, ry, rz;
            random_in_unit_sphere(g, RNG_LIGHT, __float_as_uint(r4.y), ((uint32_t)li << 16) | (uint32_t)k, rx, ry, rz);
            tx = tx + rx * radius; ty = ty + ry * radius; tz = tz + rz * radius;
        }
        float lx = tx - r0.x, ly = ty - r0.y, lz = tz - r0.z;
        max_t = sqrtf(dot3(lx, ly, lz, lx, ly, lz));  // light_dist
        normalize3(lx, ly, lz);
        sr.ox = r0.x + r1.x * 1e-4f; sr.oy = r0.y + r1.y * 1e-4f; sr.oz = r0.z + r1.z * 1e-4f;
        sr.dx = lx; sr.dy = ly; sr.dz = lz;
        sr.time = 0.0f;  // `Ray shadowRay;` keeps the default time (shapes.hpp:28)
        if (SELF && p.bvh.prune && cnt >= 8 && dot3(r1.x, r1.y, r1.z, lx, ly, lz) < 0.0f) {
            const int prim = __float_as_int(recs[(size_t)rec * 5 + 3].w);
            Hit h;
            if (intersect_prim<false>(p.bvh.prims, prim, sr, h) && !(h.t > max_t)) {
                const float4 blo = __ldg(p.bvh.leafbox + 2 * (size_t)prim), bhi = __ldg(p.bvh.leafbox + 2 * (size_t)prim + 1);
                if (!p.bvh.use_bvh || box_exact_call(blo.x, blo.y, blo.z, bhi.x, bhi.y, bhi.z, sr)) return false;  // occluded by its own shape
            }
        }
        vis_slot = vis + (size_t)rec * p.n_lights + li;
        return true;
    }
    // nothing closer than the light: this sample is lit (raytracer.cpp:233-235)
    RT_DEV void store(long long, const TravState& s) const { if (s.best_prim < 0) atomicAdd(vis_slot, 1); }
};
typedef ShadowRaysT<false> ShadowRays;
typedef ShadowRaysT<RT_SELF_OCCLUSION != 0> ShadowRaysPacket;

template <bool STATS>
__global__ void __launch_bounds__(RT_TRACE_THREADS, RT_WAVE_MINBLOCKS) shadow_kernel(const __grid_constant__ FrameParams p, int level) {
    unsigned int* lv = p.lvl + level * RT_LVL_STRIDE;
    const unsigned long long n = (unsigned long long)lv[L_RECS] * (unsigned long long)p.shadow_per_rec;
    TraceStats st = {0u, 0u};
    ShadowRays src = {p, p.recs[level & 1], p.vis[level & 1], lv[L_RECS], nullptr};
    wave_loop<true, STATS>(p.bvh, src, lv + L_WORK_SHADOW, n, st);
    if (STATS) flush_stats(p, st);
}

template <bool STATS>
__global__ void __launch_bounds__(RT_TRACE_THREADS, RT_TRACE_MINBLOCKS) shadow_packet_kernel(const __grid_constant__ FrameParams p, int level) {
    unsigned int* lv = p.lvl + level * RT_LVL_STRIDE;
    const unsigned long long n = (unsigned long long)lv[L_RECS] * (unsigned long long)p.shadow_per_rec;
    TraceStats st = {0u, 0u};
    ShadowRaysPacket src = {p, p.recs[level & 1], p.vis[level & 1], lv[L_RECS], nullptr};
    packet_loop<true, STATS>(p.bvh, src, lv + L_WORK_SHADOW, n, st);
    if (STATS) flush_stats(p, st);
}

template <bool STATS>
__global__ void __launch_bounds__(128) shadow_literal_kernel(const __grid_constant__ FrameParams p, int level) {
    unsigned int* lv = p.lvl + level * RT_LVL_STRIDE;
    const unsigned long long n = (unsigned long long)lv[L_RECS] * (unsigned long long)p.shadow_per_rec;
    TraceStats st = {0u, 0u};
    ShadowRays src = {p, p.recs[level & 1], p.vis[level & 1], lv[L_RECS], nullptr};
    literal_loop<true, STATS>(p.bvh, src, n, st);
    if (STATS) flush_stats(p, st);
}

// ---------------------------------------------------------------------------------------------
// light_kernel: Blinn-Phong sum of one shade record with the visibilities from shadow_kernel.
// shade() raytracer.cpp:191-273, then Trace()'s local_contribution * localColor.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, RT_SHADE_MINBLOCKS) light_kernel(const __grid_constant__ FrameParams p, int level) {
    const unsigned int n = p.lvl[level * RT_LVL_STRIDE + L_RECS];
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4* rc = p.recs[level & 1] + (size_t)i * 5;
        const float4 r0 = rc[0];
        if (__float_as_uint(r0.w) == RT_DEAD) continue;  // level 0: ray i missed
        const float4 r1 = rc[1], r2 = rc[2], r3 = rc[3];
        const int mat = __float_as_int(r2.w);
        const float4 m0 = __ldg(p.mats + 4 * mat + 0);
        const float4 m1 = __ldg(p.mats + 4 * mat + 1);
        const float4 m2 = __ldg(p.mats + 4 * mat + 2);
        const float ka = m0.w, kd = m1.w, ks = m2.x, shininess = m2.y;
        const float br = r3.x, bg = r3.y, bb = r3.z;
        float fr = br * ka, fg = bg * ka, fb = bb * ka;
        for (int li = 0; li < p.n_lights; ++li) {
            const float4 l0 = __ldg(p.lights + 2 * li);
            const float4 l1 = __ldg(p.lights + 2 * li + 1);
            const int shadow_samples = (l1.w > 0.0f) ? p.light_samples : 1;
            // visibility += 1.0f per unoccluded sample, then /= samples: the count is exact in float
            float visibility = (float)p.vis[level & 1][(size_t)i * p.n_lights + li];
            visibility /= (float)shadow_samples;
            if (visibility <= 0.0f) continue;
            float cx = l0.x - r0.x, cy = l0.y - r0.y, cz = l0.z - r0.z;
            const float dist_sq = dot3(cx, cy, cz, cx, cy, cz);
            const float light_distance = sqrtf(dist_sq);
            normalize3(cx, cy, cz);
            const float ndl = fmaxf(0.0f, dot3(r1.x, r1.y, r1.z, cx, cy, cz));
            float hx = cx + r2.x, hy = cy + r2.y, hz = cz + r2.z;
            normalize3(hx, hy, hz);
            const float ndh = fmaxf(0.0f, dot3(r1.x, r1.y, r1.z, hx, hy, hz));
            const float spec = powf(ndh, shininess);
            const float att = 10.0f * l0.w / (25.0f + 10.0f * light_distance + 150.0f * dist_sq);
            const float cr = l1.x * ((br * ndl) * kd + (m1.x * spec) * ks) * att;
            const float cg = l1.y * ((bg * ndl) * kd + (m1.y * spec) * ks) * att;
            const float cb = l1.z * ((bb * ndl) * kd + (m1.z * spec) * ks) * att;
            fr = fr + cr * visibility;
            fg = fg + cg * visibility;
            fb = fb + cb * visibility;
        }
        const float w = r1.w;
        accumulate(p, __float_as_uint(r0.w), w * fr, w * fg, w * fb);
    }
}

// ---------------------------------------------------------------------------------------------
// Self-tests of the exactness machinery (called by tests/test_gpu_properties.py through the C ABI).
// ---------------------------------------------------------------------------------------------
RT_DEV float st_uniform(uint32_t a, uint32_t b, uint32_t c, uint32_t lane_word) {
    const U4 u = philox4x32_10(U4{a, b, c, 0x51u}, 0xA5A5A5A5u, 0x0F0F0F0Fu);
    const uint32_t w = lane_word == 0 ? u.x : (lane_word == 1 ? u.y : (lane_word == 2 ? u.z : u.w));
    return u32_to_unit_float(w);
}

// out[0] tests, out[1] exact passes, out[2] conservative passes, out[3] surely,
// out[4] VIOLATION exact && !conservative, out[5] VIOLATION surely && !exact, out[6] rays skipped (|d_i| <= 1e-6)
__global__ void selftest_box_kernel(uint32_t seed, long long n, unsigned long long* out) {
    unsigned long long c_exact = 0, c_pass = 0, c_sure = 0, v1 = 0, v2 = 0, skipped = 0, tests = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t lo32 = (uint32_t)i, hi32 = (uint32_t)(i >> 32) ^ seed;
        auto rnd = [&](uint32_t k) { return st_uniform(lo32, hi32, k >> 2, k & 3u); };
        const int mode = (int)(rnd(0) * 6.0f);  // 0,1 generic; 2 tiny direction component; 3 huge coordinates; 4,5 grazing
        Ray r;
        const float span = mode == 3 ? 2.0e4f : 100.0f;
        r.ox = (rnd(1) * 2.0f - 1.0f) * span; r.oy = (rnd(2) * 2.0f - 1.0f) * span; r.oz = (rnd(3) * 2.0f - 1.0f) * span;
        r.dx = rnd(4) * 2.0f - 1.0f; r.dy = rnd(5) * 2.0f - 1.0f; r.dz = rnd(6) * 2.0f - 1.0f;
        if (mode == 2) {  // one component between 1e-6 and 1e-2 (log-uniform)
            const float tiny = exp2f(-20.0f + 13.0f * rnd(7)) * (rnd(8) < 0.5f ? -1.0f : 1.0f);
            const int ax = (int)(rnd(9) * 3.0f);
            if (ax == 0) r.dx = tiny; else if (ax == 1) r.dy = tiny; else r.dz = tiny;
        }
        normalize3(r.dx, r.dy, r.dz);
        r.time = 0.0f;
        // box: log-uniform size, centre near a point on the ray (or behind it, or around the origin)
        const float t0 = (rnd(10) * 1.3f - 0.3f) * 150.0f;
        const float sx = exp2f(-10.0f + 16.0f * rnd(11)), sy = exp2f(-10.0f + 16.0f * rnd(12)), sz = exp2f(-10.0f + 16.0f * rnd(13));
        float cx = r.ox + t0 * r.dx, cy = r.oy + t0 * r.dy, cz = r.oz + t0 * r.dz;
        float offx, offy, offz;
        if (mode >= 4) {
            // grazing: the point of the ray sits on a face / edge / corner of the box, up to a few ulps
            const float e = (rnd(14) * 2.0f - 1.0f) * 4e-7f;
            offx = (rnd(15) < 0.5f ? -1.0f : 1.0f) * (1.0f + e);
            offy = rnd(16) < 0.6f ? (rnd(17) < 0.5f ? -1.0f : 1.0f) * (1.0f + e) : (rnd(17) * 2.0f - 1.0f);
            offz = rnd(18) < 0.3f ? (rnd(19) < 0.5f ? -1.0f : 1.0f) * (1.0f - e) : (rnd(19) * 2.0f - 1.0f);
        } else {
            offx = (rnd(15) * 2.0f - 1.0f) * 1.5f; offy = (rnd(16) * 2.0f - 1.0f) * 1.5f; offz = (rnd(17) * 2.0f - 1.0f) * 1.5f;
        }
        cx += offx * sx; cy += offy * sy; cz += offz * sz;
        const float lox = cx - sx, hix = cx + sx, loy = cy - sy, hiy = cy + sy, loz = cz - sz, hiz = cz + sz;
        if (fabsf(r.dx) <= 1e-6f || fabsf(r.dy) <= 1e-6f || fabsf(r.dz) <= 1e-6f) { skipped++; continue; }
        TravState s;
        s.r = r;
        s.ix = 1.0f / r.dx; s.iy = 1.0f / r.dy; s.iz = 1.0f / r.dz;
        {   // the same set-up as trav_begin
            const float nx = -(r.ox * s.ix), ny = -(r.oy * s.iy), nz = -(r.oz * s.iz);
            const float ex = 1.2e-7f * fabsf(nx) + 1e-35f, ey = 1.2e-7f * fabsf(ny) + 1e-35f, ez = 1.2e-7f * fabsf(nz) + 1e-35f;
            s.nxl = s.ix > 0.0f ? nx - ex : nx + ex; s.nxh = s.ix > 0.0f ? nx + ex : nx - ex;
            s.nyl = s.iy > 0.0f ? ny - ey : ny + ey; s.nyh = s.iy > 0.0f ? ny + ey : ny - ey;
            s.nzl = s.iz > 0.0f ? nz - ez : nz + ez; s.nzh = s.iz > 0.0f ? nz + ez : nz - ez;
            s.KS = 4.0f * fmaxf(fmaxf(ex, ey), ez);
            s.imax = fmaxf(fmaxf(fabsf(s.ix), fabsf(s.iy)), fabsf(s.iz));
        }
        s.lim = FLT_MAX;
        bool pass, sure;
        float ent, tn;
        wide_child_test(s, lox, hix, loy, hiy, loz, hiz, 0.0f, pass, sure, ent);
        const bool exact = box_exact(lox, loy, loz, hix, hiy, hiz, r, tn);
        tests++;
        c_exact += exact; c_pass += pass; c_sure += sure;
        v1 += exact && !pass;
        v2 += sure && !exact;
    }
    atomicAdd(out + 0, tests); atomicAdd(out + 1, c_exact); atomicAdd(out + 2, c_pass); atomicAdd(out + 3, c_sure);
    atomicAdd(out + 4, v1); atomicAdd(out + 5, v2); atomicAdd(out + 6, skipped);
}

// For every primitive child of every node of the wide tree and `per_prim` rays aimed at / around
// it: if the exact intersection routine reports a hit, the conservative test of its culling box must
// pass. out[0] tests, out[1] exact hits, out[2] culling passes, out[3] VIOLATION hit && !pass.
__global__ void selftest_cull_kernel(BvhView b, int n_nodes, int per_prim, uint32_t seed, float scene_span, unsigned long long* out) {
    unsigned long long tests = 0, hits = 0, passes = 0, viol = 0;
    const long long total = (long long)n_nodes * 4 * per_prim;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int node = (int)(i / (4 * per_prim)), k = (int)((i / per_prim) & 3);
        const float* w = b.wide + (size_t)node * 32;
        const unsigned int meta = __float_as_uint(__ldg(w + 25));
        if (!((meta >> k) & 1u) || !((meta >> (4 + k)) & 1u)) continue;  // primitive children only
        const int idx = __float_as_int(__ldg(w + 27 + k));
        const float lox = __ldg(w + k), hix = __ldg(w + 4 + k), loy = __ldg(w + 8 + k), hiy = __ldg(w + 12 + k), loz = __ldg(w + 16 + k), hiz = __ldg(w + 20 + k);
        if (!(hix - lox < 1e29f)) continue;  // unbounded culling box: never skipped
        const uint32_t lo32 = (uint32_t)i, hi32 = (uint32_t)(i >> 32) ^ seed;
        auto rnd = [&](uint32_t q) { return st_uniform(lo32, hi32, 16u + (q >> 2), q & 3u); };
        // origin: near, mid or far (up to the scene span) from the box; target: a point in the box
        // stretched by 1.2 (so that many rays graze or just miss)
        const float cx = 0.5f * (lox + hix), cy = 0.5f * (loy + hiy), cz = 0.5f * (loz + hiz);
        const float ex = 0.5f * (hix - lox), ey = 0.5f * (hiy - loy), ez = 0.5f * (hiz - loz);
        const float dist = exp2f(-3.0f + rnd(0) * (log2f(scene_span) + 3.0f));
        float ux = rnd(1) * 2.0f - 1.0f, uy = rnd(2) * 2.0f - 1.0f, uz = rnd(3) * 2.0f - 1.0f;
        normalize3(ux, uy, uz);
        Ray r;
        r.ox = cx + ux * dist; r.oy = cy + uy * dist; r.oz = cz + uz * dist;
        const float tx = cx + (rnd(4) * 2.0f - 1.0f) * 1.2f * ex, ty = cy + (rnd(5) * 2.0f - 1.0f) * 1.2f * ey, tz = cz + (rnd(6) * 2.0f - 1.0f) * 1.2f * ez;
        r.dx = tx - r.ox; r.dy = ty - r.oy; r.dz = tz - r.oz;
        normalize3(r.dx, r.dy, r.dz);
        r.time = rnd(7);
        if (fabsf(r.dx) <= 1e-6f || fabsf(r.dy) <= 1e-6f || fabsf(r.dz) <= 1e-6f) continue;
        Hit h;
        const bool hit = intersect_prim<false>(b.prims, idx, r, h);
        TravState s;
        s.r = r;
        s.ix = 1.0f / r.dx; s.iy = 1.0f / r.dy; s.iz = 1.0f / r.dz;
        {
            const float nx = -(r.ox * s.ix), ny = -(r.oy * s.iy), nz = -(r.oz * s.iz);
            const float e0 = 1.2e-7f * fabsf(nx) + 1e-35f, e1 = 1.2e-7f * fabsf(ny) + 1e-35f, e2 = 1.2e-7f * fabsf(nz) + 1e-35f;
            s.nxl = s.ix > 0.0f ? nx - e0 : nx + e0; s.nxh = s.ix > 0.0f ? nx + e0 : nx - e0;
            s.nyl = s.iy > 0.0f ? ny - e1 : ny + e1; s.nyh = s.iy > 0.0f ? ny + e1 : ny - e1;
            s.nzl = s.iz > 0.0f ? nz - e2 : nz + e2; s.nzh = s.iz > 0.0f ? nz + e2 : nz - e2;
            s.KS = 4.0f * fmaxf(fmaxf(e0, e1), e2);
            s.imax = fmaxf(fmaxf(fabsf(s.ix), fabsf(s.iy)), fabsf(s.iz));
        }
        s.lim = FLT_MAX;
        bool pass, sure;
        float ent;
        wide_child_test(s, lox, hix, loy, hiy, loz, hiz, __ldg(w + 26) * s.imax, pass, sure, ent);
        tests++; hits += hit; passes += pass; viol += hit && !pass;
    }
    atomicAdd(out + 0, tests); atomicAdd(out + 1, hits); atomicAdd(out + 2, passes); atomicAdd(out + 3, viol);
}

// Folds the per-level counters of a finished batch into the frame totals and clears them.
__global__ void fold_kernel(const __grid_constant__ FrameParams p) {
    const int level = threadIdx.x;
    if (level > RT_MAX_DEPTH + 1) return;
    unsigned int* lv = p.lvl + level * RT_LVL_STRIDE;
    // level 0 queues have dead slots (packets = pixel blocks): its live counts are kept separately
    const unsigned long long rays = level == 0 ? lv[L_LIVE_RAYS] : min(lv[L_RAYS], (unsigned int)p.capacity);
    const unsigned long long recs = level == 0 ? lv[L_LIVE_RECS] : lv[L_RECS];
    if (rays) atomicAdd(p.totals + (level == 0 ? T_PRIMARY : T_SECONDARY), rays);
    if (recs) atomicAdd(p.totals + T_SHADOW, recs * (unsigned long long)p.shadow_per_rec);
    for (int k = 0; k < RT_LVL_STRIDE; ++k) lv[k] = 0u;
}

__global__ void clear_accum_kernel(const __grid_constant__ FrameParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = p.res_x * p.res_y;
    if (idx >= n) return;
    if (out_index(p, idx % p.res_x, idx / p.res_x) < 0) return;
    p.accum[(size_t)idx * 3 + 0] = 0ull; p.accum[(size_t)idx * 3 + 1] = 0ull; p.accum[(size_t)idx * 3 + 2] = 0ull;
}

// Average, gamma 1.1, clamp, * 255.999 (raytracer.cpp:69, 446-457).
__global__ void finalize_kernel(const __grid_constant__ FrameParams p, uint8_t* rgb8, float* linear) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = p.res_x * p.res_y;
    if (idx >= n) return;
    const long long o = out_index(p, idx % p.res_x, idx / p.res_x);
    if (o < 0) return;
    const double inv = 1.0 / 1099511627776.0;
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c[k] = (float)((double)(long long)p.accum[(size_t)idx * 3 + k] * inv);
        if (p.samples_sqrt > 1) c[k] = c[k] / (float)p.spp;
    }
    if (linear) { linear[3 * (size_t)o + 0] = c[0]; linear[3 * (size_t)o + 1] = c[1]; linear[3 * (size_t)o + 2] = c[2]; }
    if (rgb8) {
        const float inv_gamma = 1.0f / 1.1f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v = powf(c[k], inv_gamma);
            v = (v < 1.0f) ? v : 1.0f;  // std::max(0.0f, std::min(1.0f, v)): NaN -> 1
            v = (0.0f < v) ? v : 0.0f;
            int q = (int)((double)v * 255.999);
            q = max(0, min(q, 255));
            rgb8[3 * (size_t)o + k] = (uint8_t)q;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host side: device scenes (one per CUDA device), buffers, launches
// ---------------------------------------------------------------------------------------------

// Which screen tiles a rank renders, as device arrays (gen_kernel / out_index read them). Cached per
// (frame geometry, tile size, rank, world, block, window) in the DeviceScene.
struct TilePlan {
    int key[12] = {0};
    std::vector<int> tiles;  // this rank's tiles, increasing
    int* d_tiles = nullptr;
    int* d_slot = nullptr;
    int64_t pixels = 0;      // pixels this rank renders (inside the window)
};

struct DeviceScene {
    int device = -1;
    // The scene lives in ONE device arena filled from ONE page-locked host staging buffer (shared by
    // all devices, DeviceSet::staging), so making the scene resident is a single asynchronous H2D
    // copy. Allocated once per scene and device; rt_scene_evict only marks the device copy stale.
    uint8_t* arena = nullptr;
    bool resident = false;
    cudaEvent_t ev_upload = nullptr;  // recorded behind the arena copy: every frame waits for it
    float4* prims = nullptr;
    float* wide = nullptr;
    float4* leafbox = nullptr;
    float* ref_tree = nullptr;   // literal mode only: the reference's binary tree (HostScene::tree), uploaded on first use
    float4* mats = nullptr;
    float4* lights = nullptr;
    DTexture* textures = nullptr;
    uint8_t* texels = nullptr;
    // wavefront buffers
    float4* q[2] = {nullptr, nullptr};
    int* hit_prim = nullptr;
    float4* recs[2] = {nullptr, nullptr};
    int* vis[2] = {nullptr, nullptr};
    cudaStream_t aux = nullptr;                  // shadow + light kernels run here, overlapping the next level
    cudaStream_t own = nullptr;                  // rt_render_multi: this device's frame stream
    std::vector<cudaEvent_t> ev_shade, ev_light;  // per level
    long long capacity = 0;
    int vis_lights = 0;
    unsigned long long* accum = nullptr;
    size_t accum_pixels = 0;
    unsigned int* lvl = nullptr;
    unsigned long long* totals = nullptr;
    unsigned int* overflow_host = nullptr;  // page-locked, mapped: set by shade_kernel when a queue overflows
    bool async_pending = false;             // the last frame was asynchronous: its overflow flag has not been looked at
    // (pixel, sample) pairs per batch; halved on queue overflow. Large batches keep the waves of the
    // deeper recursion levels big (a level of a batch is one launch); RT_B200_BATCH_SLOTS overrides.
    long long batch_slots = [] { const char* e = std::getenv("RT_B200_BATCH_SLOTS"); return e ? std::atoll(e) : (8ll << 20); }();
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    int sm_count = 0;
    int trace_blocks = 1, shadow_blocks = 1;  // resident blocks per SM: per-ray kernels ...
    int trace_packet_blocks = 1, shadow_packet_blocks = 1;  // ... and packet kernels (their stacks are per warp: less shared memory)
    int stack_depth = 4;         // entries per thread of the per-ray kernels' stacks
    int packet_stack_depth = 4;  // entries per warp of the packet kernels' stacks
    int n_staged = 0;            // RT_STAGE_TOP builds: top nodes staged in shared memory by the per-ray kernels
    size_t stack_bytes = 0;
    bool timed = false;
    int last_launches = 0;
    unsigned long long total_launches = 0;  // kernels launched on this device for this scene, ever
    // optional per-kernel-class timing (rt_render_params.reserved[1] & 1): event pairs around the
    // trace / shadow / shade / light launches of the most recent frame
    std::vector<cudaEvent_t> class_ev;
    std::vector<int> class_of;  // class of pair i: 0 trace, 1 shadow, 2 shade, 3 light
    int class_launches[4] = {0, 0, 0, 0};  // all launches of the frame per class (timed or not)
    std::vector<TilePlan*> plans;
    // device outputs of rt_render / rt_render_multi, kept between calls
    uint8_t* out_rgb = nullptr;
    int32_t* out_ids = nullptr;
    float* out_lin = nullptr;
    size_t out_pixels = 0;
    // rt_render_multi: page-locked host landing zone of this device's packed tiles
    uint8_t* host_rgb = nullptr;
    int32_t* host_ids = nullptr;
    float* host_lin = nullptr;
    size_t host_pixels = 0;
};

struct MultiGpu;

// Everything a scene owns on the CUDA side (HostScene::dev).
struct DeviceSet {
    std::mutex mu;
    std::vector<DeviceScene*> devs;  // by CUDA device ordinal
    uint8_t* staging = nullptr;      // cudaHostAlloc(portable): the packed scene, source of every upload
    size_t arena_bytes = 0;
    size_t o_prims = 0, o_wide = 0, o_leafbox = 0, o_mats = 0, o_lights = 0, o_tex = 0, o_texels = 0;
    uint64_t bytes = 0;  // payload bytes (what an upload copies)
    MultiGpu* multi = nullptr;
};

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            rtb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                \
            return RT_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

static void free_plan(TilePlan* t) {
    if (!t) return;
    cudaFree(t->d_tiles);
    cudaFree(t->d_slot);
    delete t;
}

static void free_device(DeviceScene* d) {
    if (!d) return;
    int cur = 0;
    cudaGetDevice(&cur);
    if (d->device >= 0) cudaSetDevice(d->device);
    cudaFree(d->arena);
    cudaFree(d->ref_tree);
    cudaFree(d->q[0]); cudaFree(d->q[1]); cudaFree(d->hit_prim);
    for (int i = 0; i < 2; ++i) { cudaFree(d->recs[i]); cudaFree(d->vis[i]); }
    if (d->aux) cudaStreamDestroy(d->aux);
    if (d->own) cudaStreamDestroy(d->own);
    for (cudaEvent_t e : d->ev_shade) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : d->ev_light) if (e) cudaEventDestroy(e);
    cudaFree(d->accum); cudaFree(d->lvl); cudaFree(d->totals);
    if (d->overflow_host) cudaFreeHost(d->overflow_host);
    cudaFree(d->out_rgb); cudaFree(d->out_ids); cudaFree(d->out_lin);
    if (d->host_rgb) cudaFreeHost(d->host_rgb);
    if (d->host_ids) cudaFreeHost(d->host_ids);
    if (d->host_lin) cudaFreeHost(d->host_lin);
    for (cudaEvent_t e : d->ev) if (e) cudaEventDestroy(e);
    if (d->ev_upload) cudaEventDestroy(d->ev_upload);
    for (cudaEvent_t e : d->class_ev) if (e) cudaEventDestroy(e);
    for (TilePlan* t : d->plans) free_plan(t);
    delete d;
    cudaSetDevice(cur);
    cudaGetLastError();
}

static void multi_shutdown(MultiGpu* m);

void device_release(HostScene& h) {
    DeviceSet* set = h.dev;
    if (!set) return;
    multi_shutdown(set->multi);
    for (DeviceScene* d : set->devs) free_device(d);
    if (set->staging) cudaFreeHost(set->staging);
    cudaGetLastError();
    delete set;
    h.dev = nullptr;
}

void device_invalidate(HostScene& h) {
    if (!h.dev) return;
    for (DeviceScene* d : h.dev->devs) if (d) d->resident = false;
}

template <typename T>
static size_t place(const std::vector<T>& v, size_t& offset) {
    const size_t at = offset;
    offset += (std::max<size_t>(v.size(), 1) * sizeof(T) + 255) & ~(size_t)255;
    return at;
}
template <typename T>
static void stage(uint8_t* staging, size_t at, const std::vector<T>& v, uint64_t& bytes) {
    if (!v.empty()) std::memcpy(staging + at, v.data(), v.size() * sizeof(T));
    bytes += v.size() * sizeof(T);
}

// The scene packed into page-locked host memory, once per scene (whatever the number of devices).
static int ensure_staging(HostScene& h) {
    if (!h.dev) h.dev = new DeviceSet();
    DeviceSet* set = h.dev;
    std::lock_guard<std::mutex> lock(set->mu);
    if (set->staging) return RT_OK;
    size_t off = 0;
    set->o_prims = place(h.dprims, off); set->o_wide = place(h.dwide, off); set->o_leafbox = place(h.dleafbox, off);
    set->o_mats = place(h.dmaterials, off); set->o_lights = place(h.dlights, off); set->o_tex = place(h.dtextures, off);
    set->o_texels = place(h.texels, off);
    set->arena_bytes = off;
    uint8_t* st = nullptr;
    CUDA_TRY(cudaHostAlloc((void**)&st, set->arena_bytes, cudaHostAllocPortable));
    std::memset(st, 0, set->arena_bytes);
    set->bytes = 0;
    stage(st, set->o_prims, h.dprims, set->bytes); stage(st, set->o_wide, h.dwide, set->bytes);
    stage(st, set->o_leafbox, h.dleafbox, set->bytes);
    stage(st, set->o_mats, h.dmaterials, set->bytes); stage(st, set->o_lights, h.dlights, set->bytes);
    stage(st, set->o_tex, h.dtextures, set->bytes); stage(st, set->o_texels, h.texels, set->bytes);
    set->staging = st;
    return RT_OK;
}

// The DeviceScene of the CURRENT CUDA device (created on first use).
static int device_scene(HostScene& h, DeviceScene** out) {
    int rc = ensure_staging(h);
    if (rc != RT_OK) return rc;
    DeviceSet* set = h.dev;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(set->mu);
        if ((int)set->devs.size() <= dev) set->devs.resize((size_t)dev + 1, nullptr);
        if (set->devs[dev]) { *out = set->devs[dev]; return RT_OK; }
    }
    DeviceScene* d = new DeviceScene();
    d->device = dev;
    auto init = [&]() -> int {
        CUDA_TRY(cudaMalloc((void**)&d->arena, set->arena_bytes));
        d->prims = (float4*)(d->arena + set->o_prims); d->wide = (float*)(d->arena + set->o_wide);
        d->leafbox = (float4*)(d->arena + set->o_leafbox);
        d->mats = (float4*)(d->arena + set->o_mats); d->lights = (float4*)(d->arena + set->o_lights);
        d->textures = (DTexture*)(d->arena + set->o_tex); d->texels = d->arena + set->o_texels;
        CUDA_TRY(cudaMalloc((void**)&d->lvl, (RT_MAX_DEPTH + 2) * RT_LVL_STRIDE * sizeof(unsigned int)));
        CUDA_TRY(cudaMalloc((void**)&d->totals, 8 * sizeof(unsigned long long)));
        CUDA_TRY(cudaHostAlloc((void**)&d->overflow_host, 64, cudaHostAllocMapped));
        *d->overflow_host = 0u;
        for (auto& e : d->ev) CUDA_TRY(cudaEventCreate(&e));
        CUDA_TRY(cudaEventCreateWithFlags(&d->ev_upload, cudaEventDisableTiming));
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, d->device));
        d->sm_count = prop.multiProcessorCount;
        // Traversal stacks in shared memory: h.stack_need entries (bvh.cpp keeps the tree at <= 32 where it can),
        // per thread in the per-ray kernels, per warp (16-byte entries) in the packet kernels.
        d->stack_depth = d->packet_stack_depth = std::max(4, h.stack_need);
        d->stack_bytes = (size_t)d->stack_depth * RT_TRACE_THREADS * sizeof(int) * RT_STACK_WORDS;
        d->n_staged = std::min<int>(RT_STAGE_TOP, (int)h.dwide.size());
        d->stack_bytes += (size_t)d->n_staged * sizeof(DWide);
        if (d->stack_bytes > 200 * 1024) { set_error("BVH too deep for the shared-memory traversal stack"); return RT_ERR_SCENE; }
        if (d->stack_bytes > 48 * 1024) {
            CUDA_TRY(cudaFuncSetAttribute(trace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->stack_bytes));
            CUDA_TRY(cudaFuncSetAttribute(trace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->stack_bytes));
            CUDA_TRY(cudaFuncSetAttribute(shadow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->stack_bytes));
            CUDA_TRY(cudaFuncSetAttribute(shadow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->stack_bytes));
        }
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->trace_blocks, trace_kernel<false>, RT_TRACE_THREADS, d->stack_bytes));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->shadow_blocks, shadow_kernel<false>, RT_TRACE_THREADS, d->stack_bytes));
        const size_t packet_smem0 = (size_t)(RT_TRACE_THREADS / 32) * d->packet_stack_depth * 16;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->trace_packet_blocks, trace_packet_kernel<false>, RT_TRACE_THREADS, packet_smem0));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->shadow_packet_blocks, shadow_packet_kernel<false>, RT_TRACE_THREADS, packet_smem0));
        d->trace_blocks = std::max(1, d->trace_blocks);
        d->shadow_blocks = std::max(1, d->shadow_blocks);
        d->trace_packet_blocks = std::max(1, d->trace_packet_blocks);
        d->shadow_packet_blocks = std::max(1, d->shadow_packet_blocks);
        return RT_OK;
    };
    rc = init();
    if (rc != RT_OK) { free_device(d); return rc; }  // never keep a half-initialised device scene
    std::lock_guard<std::mutex> lock(set->mu);
    set->devs[dev] = d;
    *out = d;
    return RT_OK;
}

// Makes the scene resident on the current device: one H2D copy of the arena on `stream`, ordered
// BEHIND the last frame that still reads the old copy and AHEAD of every later frame (ev_upload),
// whatever streams those frames use.
static int ensure_uploaded(HostScene& h, cudaStream_t stream, uint64_t* bytes_out, DeviceScene** out = nullptr) {
    if (bytes_out) *bytes_out = 0;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device: the renderer has no CPU fallback");
        return RT_ERR_CUDA;
    }
    DeviceScene* d = nullptr;
    const int rc = device_scene(h, &d);
    if (rc != RT_OK) return rc;
    if (!d->resident) {
        if (d->timed) CUDA_TRY(cudaStreamWaitEvent(stream, d->ev[2], 0));
        CUDA_TRY(cudaMemcpyAsync(d->arena, h.dev->staging, h.dev->arena_bytes, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaEventRecord(d->ev_upload, stream));
        d->resident = true;
        if (bytes_out) *bytes_out = h.dev->bytes;
    }
    if (out) *out = d;
    return RT_OK;
}

// rank that renders tile (tx, ty): tiles are dealt round-robin, singly (block <= 1) or in
// block x block groups (a rank then touches a compact part of the scene per group)
static inline int tile_owner(int tx, int ty, int tiles_x, int world, int block) {
    if (block <= 1) return (ty * tiles_x + tx) % world;
    const int bx = (tiles_x + block - 1) / block;
    return ((ty / block) * bx + tx / block) % world;
}

static int fill_params(const HostScene& h, const rt_render_params& rp, FrameParams& k) {
    if (h.cam.res_x <= 0 || h.cam.res_y <= 0) { set_error("Camera resolution is 0. Check scene.json."); return RT_ERR_SCENE; }
    if (h.cam.res_x > 65535 || h.cam.res_y > 65535) { set_error("resolution above 65535 is not supported"); return RT_ERR_SCENE; }
    if (rp.world < 1 || rp.rank < 0 || rp.rank >= rp.world) { set_error("rank/world out of range"); return RT_ERR_INVALID; }
    if (rp.tile_w < 8 || rp.tile_h < 4 || rp.tile_w % 8 || rp.tile_h % 4) { set_error("tile_w must be a multiple of 8 and tile_h of 4"); return RT_ERR_INVALID; }
    if (rp.max_depth < 0 || rp.max_depth > RT_MAX_DEPTH) { set_error("max_depth must be in [0,16]"); return RT_ERR_INVALID; }
    if (rp.light_samples < 1 || rp.light_samples > 65535) { set_error("light_samples must be in [1,65535]"); return RT_ERR_INVALID; }
    if (rp.samples_sqrt > 1024) { set_error("samples_sqrt must be <= 1024"); return RT_ERR_INVALID; }
    if (rp.reserved[5] < 0 || rp.reserved[5] > 1024) { set_error("tile block must be in [0,1024]"); return RT_ERR_INVALID; }
    std::memset(&k, 0, sizeof(k));
    k.bvh.n_prims = (int)h.dprims.size();
    k.bvh.use_bvh = rp.use_bvh ? 1 : 0;
    k.bvh.prune = (rp.reserved[0] & 1) ? 0 : 1;  // reserved[0] bit 0: literal reference traversal (test hook)
    k.n_lights = (int)h.dlights.size();
    for (int i = 0; i < 3; ++i) {
        k.cam_loc[i] = h.cam.location[i]; k.xdir[i] = h.xdir[i]; k.ydir[i] = h.ydir[i]; k.zdir[i] = h.zdir[i];
    }
    k.focal = h.cam.focal_length;
    k.half_sw = (float)h.cam.sensor_width / 2.0f;  // camera.cpp:106-107
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
    k.shadow_per_rec = 0;
    for (const rt_light_desc& l : h.lights) k.shadow_per_rec += (l.radius > 0.0f) ? rp.light_samples : 1;
    k.tile_w = rp.tile_w;
    k.tile_h = rp.tile_h;
    k.tiles_x = (k.res_x + k.tile_w - 1) / k.tile_w;
    const int tiles_y = (k.res_y + k.tile_h - 1) / k.tile_h;
    k.n_tiles = k.tiles_x * tiles_y;
    k.rank = rp.rank;
    k.world = rp.world;
    k.sub_x = k.tile_w / 8;
    k.sub_per_tile = k.sub_x * (k.tile_h / 4);
    // reserved[3] / [4]: render window x0 | x1 << 16, y0 | y1 << 16 (0 = whole frame)
    k.win_x0 = 0; k.win_y0 = 0; k.win_x1 = k.res_x; k.win_y1 = k.res_y;
    if (rp.reserved[3] != 0 || rp.reserved[4] != 0) {
        const uint32_t wx = (uint32_t)rp.reserved[3], wy = (uint32_t)rp.reserved[4];
        k.win_x0 = (int)(wx & 0xffffu); k.win_x1 = std::min(k.res_x, (int)(wx >> 16));
        k.win_y0 = (int)(wy & 0xffffu); k.win_y1 = std::min(k.res_y, (int)(wy >> 16));
        if (k.win_x0 >= k.win_x1 || k.win_y0 >= k.win_y1) { set_error("empty render window"); return RT_ERR_INVALID; }
    }
    return RT_OK;
}

// The tiles of (rank, world, block) that intersect the window, in increasing order.
static void plan_tiles(const FrameParams& k, int block, std::vector<int>& tiles, int64_t& pixels) {
    tiles.clear();
    pixels = 0;
    for (int t = 0; t < k.n_tiles; ++t) {
        const int tx = t % k.tiles_x, ty = t / k.tiles_x;
        if (tile_owner(tx, ty, k.tiles_x, k.world, block) != k.rank) continue;
        const int x0 = std::max(tx * k.tile_w, k.win_x0), x1 = std::min((tx + 1) * k.tile_w, k.win_x1);
        const int y0 = std::max(ty * k.tile_h, k.win_y0), y1 = std::min((ty + 1) * k.tile_h, k.win_y1);
        if (x0 >= x1 || y0 >= y1) continue;
        tiles.push_back(t);
        pixels += (int64_t)(x1 - x0) * (y1 - y0);
    }
}

static int ensure_plan(DeviceScene* d, FrameParams& k, int block, cudaStream_t stream, TilePlan** out) {
    const int key[12] = {k.res_x, k.res_y, k.tile_w, k.tile_h, k.rank, k.world, block, k.win_x0, k.win_y0, k.win_x1, k.win_y1, 0};
    TilePlan* plan = nullptr;
    for (TilePlan* t : d->plans) if (std::memcmp(t->key, key, sizeof(key)) == 0) { plan = t; break; }
    if (!plan) {
        if (d->plans.size() >= 64) {  // frames still in flight may read the old plans: drain before dropping them
            CUDA_TRY(cudaDeviceSynchronize());
            for (TilePlan* t : d->plans) free_plan(t);
            d->plans.clear();
        }
        plan = new TilePlan();
        std::memcpy(plan->key, key, sizeof(key));
        plan_tiles(k, block, plan->tiles, plan->pixels);
        std::vector<int> slot((size_t)k.n_tiles, -1);
        for (size_t i = 0; i < plan->tiles.size(); ++i) slot[plan->tiles[i]] = (int)i;
        cudaError_t e = cudaMalloc((void**)&plan->d_tiles, std::max<size_t>(1, plan->tiles.size()) * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc((void**)&plan->d_slot, (size_t)k.n_tiles * sizeof(int));
        // pageable sources: the copies are staged before the calls return, so `slot` may go out of scope
        if (e == cudaSuccess && !plan->tiles.empty())
            e = cudaMemcpyAsync(plan->d_tiles, plan->tiles.data(), plan->tiles.size() * sizeof(int), cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(plan->d_slot, slot.data(), slot.size() * sizeof(int), cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { free_plan(plan); set_error(std::string("tile plan: ") + cudaGetErrorString(e)); return RT_ERR_CUDA; }
        d->plans.push_back(plan);
    }
    k.tiles = plan->d_tiles;
    k.tile_slot = plan->d_slot;
    k.n_my_tiles = (int)plan->tiles.size();
    *out = plan;
    return RT_OK;
}

// Rays of a level per (pixel, sample) of the batch: a hit spawns at most one reflection and one
// refraction ray, so a scene without a material that is reflective AND transparent never has more
// rays at any level than at level 0.
static int ray_tree_branching(const HostScene& h) {
    int b = 0;
    for (const rt_material_desc& m : h.materials) b = std::max(b, (m.reflectivity > 0.0f ? 1 : 0) + (m.transparency > 0.0f ? 1 : 0));
    return b;
}

// Queue capacity (rays per level of a batch). Branching <= 1: the batch itself bounds every level.
// Branching 2 (reflective AND transparent materials): ray trees can double per level, so the queues
// get room for 16x the batch, but at most 6M rays unless the batch itself is larger -- then 2x the
// batch; if a level still overflows, the frame is re-rendered with half the batch.
static long long wanted_capacity(long long batch_slots, int branching) {
    if (branching <= 1) return std::max<long long>(batch_slots, 65536);
    return std::max<long long>(std::max<long long>(2 * batch_slots, 65536), std::min<long long>(16 * batch_slots, 6ll << 20));
}

static int ensure_buffers(DeviceScene* d, const FrameParams& k, long long batch_slots, int branching) {
    const long long want_cap = wanted_capacity(batch_slots, branching);
    if (want_cap > (1ll << 30)) { set_error("batch too large"); return RT_ERR_INVALID; }
    if (want_cap > d->capacity || k.n_lights > d->vis_lights) {
        const long long cap = std::max(want_cap, d->capacity);
        const int nl = std::max(1, std::max(k.n_lights, d->vis_lights));
        CUDA_TRY(cudaDeviceSynchronize());  // frames in flight on other streams still use the old buffers
        cudaFree(d->q[0]); cudaFree(d->q[1]); cudaFree(d->hit_prim);
        d->q[0] = d->q[1] = nullptr; d->hit_prim = nullptr;
        for (int i = 0; i < 2; ++i) { cudaFree(d->recs[i]); cudaFree(d->vis[i]); d->recs[i] = nullptr; d->vis[i] = nullptr; }
        d->capacity = 0;
        CUDA_TRY(cudaMalloc((void**)&d->q[0], (size_t)cap * 3 * sizeof(float4)));
        CUDA_TRY(cudaMalloc((void**)&d->q[1], (size_t)cap * 3 * sizeof(float4)));
        CUDA_TRY(cudaMalloc((void**)&d->hit_prim, (size_t)cap * sizeof(int)));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaMalloc((void**)&d->recs[i], (size_t)cap * 5 * sizeof(float4)));
            CUDA_TRY(cudaMalloc((void**)&d->vis[i], (size_t)cap * nl * sizeof(int)));
        }
        d->capacity = cap;
        d->vis_lights = nl;
    }
    const size_t pixels = (size_t)k.res_x * k.res_y;
    if (pixels > d->accum_pixels) {
        CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(d->accum);
        d->accum = nullptr;
        d->accum_pixels = 0;
        CUDA_TRY(cudaMalloc((void**)&d->accum, pixels * 3 * sizeof(unsigned long long)));
        d->accum_pixels = pixels;
    }
    return RT_OK;
}

// Enqueues one frame on `stream` (no host synchronisation).
static int enqueue_frame(const HostScene& h, DeviceScene* d, FrameParams& k, bool collect, bool time_classes, bool serial, uint8_t* rgb8,
                         float* linear, cudaStream_t stream) {
    const long long total_units = (long long)k.n_my_tiles * k.sub_per_tile * k.spp;  // unit = 32 (pixel, sample) slots
    // work counters are 32-bit: keep (slots of a batch) x (shadow rays per shaded hit) below 2^31 ...
    const long long spr = std::max(1, k.shadow_per_rec);
    const long long slot_cap = std::max<long long>(32, (1ll << 31) / spr);
    const long long batch_units = std::max<long long>(1, std::min<long long>(std::max<long long>(total_units, 1), std::min(d->batch_slots, slot_cap) / 32));
    const int branching = ray_tree_branching(h);
    int rc = ensure_buffers(d, k, batch_units * 32, branching);
    if (rc != RT_OK) return rc;
    k.q[0] = d->q[0]; k.q[1] = d->q[1];
    k.hit_prim = d->hit_prim;
    for (int i = 0; i < 2; ++i) { k.recs[i] = d->recs[i]; k.vis[i] = d->vis[i]; }
    k.accum = d->accum; k.lvl = d->lvl; k.totals = d->totals;
    k.overflow_host = d->overflow_host;
    const int grid_trace = d->sm_count * d->trace_blocks, grid_trace_packet = d->sm_count * d->trace_packet_blocks;
    const int grid_shadow = d->sm_count * d->shadow_blocks, grid_shadow_packet = d->sm_count * d->shadow_packet_blocks;
    const int grid_wide = d->sm_count * 8;
    // ... and at deeper levels, where the record count can grow up to the queue capacity, keep
    // capacity x (shadow rays per hit) + (what the persistent warps over-fetch past the end) below
    // 2^32. All of the allocated queue space otherwise, also after a retry with smaller batches (the
    // allocation never shrinks, so halving the batch really halves the pressure on the queues).
    const long long overshoot = (long long)std::max(std::max(grid_trace, grid_shadow), std::max(grid_trace_packet, grid_shadow_packet)) * (RT_TRACE_THREADS / 32) * 128 + 65536;
    const long long work_cap = ((1ll << 32) - overshoot) / spr;
    k.capacity = (int)std::min<long long>(std::min<long long>(d->capacity, (1ll << 30)), work_cap);

    int launches = 0;
    CUDA_TRY(cudaMemsetAsync(d->lvl, 0, (RT_MAX_DEPTH + 2) * RT_LVL_STRIDE * sizeof(unsigned int), stream));
    CUDA_TRY(cudaMemsetAsync(d->totals, 0, 8 * sizeof(unsigned long long), stream));
    const int n_pix = k.res_x * k.res_y;
    clear_accum_kernel<<<(n_pix + 255) / 256, 256, 0, stream>>>(k);
    ++launches;
    d->class_of.clear();
    const long long n_batches = total_units > 0 ? (total_units + batch_units - 1) / batch_units : 0;
    const size_t max_pairs = time_classes ? (size_t)std::min<long long>(n_batches * (k.max_depth + 1) * 4, 1ll << 20) : 0;
    if (time_classes && d->class_ev.size() < 2 * max_pairs) {
        const size_t have = d->class_ev.size();
        d->class_ev.resize(2 * max_pairs, nullptr);
        for (size_t i = have; i < d->class_ev.size(); ++i) CUDA_TRY(cudaEventCreate(&d->class_ev[i]));
    }
    // returns the pair index (or -1) at the begin mark; the end mark takes it back
    for (int& c : d->class_launches) c = 0;
    auto mark_begin = [&](int cls, cudaStream_t on) -> int {
        d->class_launches[cls]++;
        if (!time_classes || d->class_of.size() >= max_pairs) return -1;
        const int pair = (int)d->class_of.size();
        cudaEventRecord(d->class_ev[2 * pair], on);
        d->class_of.push_back(cls | 0x100);  // 0x100: waiting for its end event
        return pair;
    };
    auto mark_end = [&](int pair, cudaStream_t on) {
        if (pair < 0) return;
        cudaEventRecord(d->class_ev[2 * pair + 1], on);
        d->class_of[pair] &= 0xff;
    };
    // Two streams: trace + shade of level d+1 (main) overlap shadow + light of level d (aux) -- both
    // pairs only depend on shade(d). The persistent kernels fill the GPU, so the overlap mostly
    // hides each kernel's tail behind the other's start. RT_B200_OVERLAP=0 serialises everything.
    static const bool overlap_enabled = [] { const char* e = std::getenv("RT_B200_OVERLAP"); return !(e && e[0] == '0'); }();
    static const bool packet_enabled = [] { const char* e = std::getenv("RT_B200_PACKET"); return !(e && e[0] == '0'); }();
    const size_t packet_smem = (size_t)(RT_TRACE_THREADS / 32) * d->packet_stack_depth * 16;  // per-warp stacks only
    static const bool area_packets_enabled = [] { const char* e = std::getenv("RT_B200_AREA_PACKETS"); return !(e && e[0] == '0'); }();
    const bool area_light_packets = area_packets_enabled && k.light_samples >= 8 && k.shadow_per_rec >= k.light_samples;
    const bool overlap = overlap_enabled && !serial;
    if (overlap && !d->aux) CUDA_TRY(cudaStreamCreateWithFlags(&d->aux, cudaStreamNonBlocking));
    const cudaStream_t aux = overlap ? d->aux : stream;
    if (d->ev_shade.size() < RT_MAX_DEPTH + 2) {
        const size_t have = d->ev_shade.size();
        d->ev_shade.resize(RT_MAX_DEPTH + 2, nullptr);
        d->ev_light.resize(RT_MAX_DEPTH + 2, nullptr);
        for (size_t i = have; i < d->ev_shade.size(); ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&d->ev_shade[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&d->ev_light[i], cudaEventDisableTiming));
        }
    }
    for (long long u0 = 0; u0 < total_units; u0 += batch_units) {
        const int n_units = (int)std::min<long long>(batch_units, total_units - u0);
        gen_kernel<<<std::min(grid_wide, (n_units + 7) / 8), 256, 0, stream>>>(k, u0, n_units);
        ++launches;
        for (int level = 0; level <= k.max_depth; ++level) {
            // level 0 is coherent (pixel blocks): packet traversal; deeper levels: per-ray traversal
            static const int packet_levels = [] { const char* e = std::getenv("RT_B200_PACKET_LEVELS"); return e ? std::atoi(e) : 0; }();
            const bool packets = packet_enabled && level <= packet_levels;
            // the samples of an area light leave one point: coherent at every level
            const bool shadow_packets = packets || (packet_enabled && area_light_packets);
            const bool literal = !k.bvh.prune || !k.bvh.use_bvh;  // the reference's own traversal / its linear scan
            int pr = mark_begin(0, stream);
            if (literal) {
                if (collect) trace_literal_kernel<true><<<grid_wide, 128, 0, stream>>>(k, level);
                else trace_literal_kernel<false><<<grid_wide, 128, 0, stream>>>(k, level);
            } else if (packets) {
                if (collect) trace_packet_kernel<true><<<grid_trace_packet, RT_TRACE_THREADS, packet_smem, stream>>>(k, level);
                else trace_packet_kernel<false><<<grid_trace_packet, RT_TRACE_THREADS, packet_smem, stream>>>(k, level);
            } else {
                if (collect) trace_kernel<true><<<grid_trace, RT_TRACE_THREADS, d->stack_bytes, stream>>>(k, level);
                else trace_kernel<false><<<grid_trace, RT_TRACE_THREADS, d->stack_bytes, stream>>>(k, level);
            }
            mark_end(pr, stream);
            // shade(level) overwrites the record buffers that shadow/light of level - 2 read
            if (aux != stream && level >= 2) CUDA_TRY(cudaStreamWaitEvent(stream, d->ev_light[level - 2], 0));
            pr = mark_begin(2, stream);
            shade_kernel<<<grid_wide, 256, 0, stream>>>(k, level);
            mark_end(pr, stream);
            if (aux != stream) {
                CUDA_TRY(cudaEventRecord(d->ev_shade[level], stream));
                CUDA_TRY(cudaStreamWaitEvent(aux, d->ev_shade[level], 0));
            }
            if (k.shadow_per_rec > 0) {
                pr = mark_begin(1, aux);
                if (literal) {
                    if (collect) shadow_literal_kernel<true><<<grid_wide, 128, 0, aux>>>(k, level);
                    else shadow_literal_kernel<false><<<grid_wide, 128, 0, aux>>>(k, level);
                } else if (shadow_packets) {
                    if (collect) shadow_packet_kernel<true><<<grid_shadow_packet, RT_TRACE_THREADS, packet_smem, aux>>>(k, level);
                    else shadow_packet_kernel<false><<<grid_shadow_packet, RT_TRACE_THREADS, packet_smem, aux>>>(k, level);
                } else {
                    if (collect) shadow_kernel<true><<<grid_shadow, RT_TRACE_THREADS, d->stack_bytes, aux>>>(k, level);
                    else shadow_kernel<false><<<grid_shadow, RT_TRACE_THREADS, d->stack_bytes, aux>>>(k, level);
                }
                mark_end(pr, aux);
                ++launches;
            }
            pr = mark_begin(3, aux);
            light_kernel<<<grid_wide, 256, 0, aux>>>(k, level);
            mark_end(pr, aux);
            if (aux != stream) CUDA_TRY(cudaEventRecord(d->ev_light[level], aux));
            launches += 3;
        }
        // aux is ordered: the last light kernel finishes after all earlier ones
        if (aux != stream) CUDA_TRY(cudaStreamWaitEvent(stream, d->ev_light[k.max_depth], 0));
        fold_kernel<<<1, 32, 0, stream>>>(k);
        ++launches;
    }
    finalize_kernel<<<(n_pix + 255) / 256, 256, 0, stream>>>(k, rgb8, linear);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    d->last_launches = launches;
    d->total_launches += (unsigned long long)launches;
    return RT_OK;
}

// A queue overflow of an asynchronous frame cannot be answered by re-rendering (nobody waits for
// the frame), so it is STICKY: the next call on the scene that looks (rt_render*, rt_scene_last_timing)
// halves the batch, reports RT_ERR_SCENE once, and the caller renders again.
static int report_async_overflow(DeviceScene* d) {
    if (!d->async_pending || !*(volatile unsigned int*)d->overflow_host) return RT_OK;
    if (cudaEventSynchronize(d->ev[2]) != cudaSuccess) cudaGetLastError();
    *d->overflow_host = 0u;
    d->async_pending = false;
    d->batch_slots = std::max<long long>(32, d->batch_slots / 2);
    set_error("an earlier asynchronous frame on this scene overflowed a ray queue and dropped rays; the batch size has been "
              "halved: render the frame again (a synchronous call, stats != NULL, adapts the batch size by itself)");
    return RT_ERR_SCENE;
}

static int render_impl(HostScene& h, const rt_render_params& rp, uint8_t* rgb8, int32_t* hit_ids, float* linear,
                       cudaStream_t stream, rt_render_stats* stats, bool packed = false, TilePlan** plan_out = nullptr) {
    FrameParams k;
    int rc = fill_params(h, rp, k);
    if (rc != RT_OK) return rc;
    DeviceScene* d = nullptr;
    if ((rc = ensure_uploaded(h, stream, nullptr, &d)) != RT_OK) return rc;
    if ((rc = report_async_overflow(d)) != RT_OK) return rc;
    // one frame in flight per scene and device: the wavefront buffers are shared, so a frame on
    // another stream waits for the previous one; every frame waits for the scene upload
    if (d->timed) CUDA_TRY(cudaStreamWaitEvent(stream, d->ev[2], 0));
    CUDA_TRY(cudaStreamWaitEvent(stream, d->ev_upload, 0));
    TilePlan* plan = nullptr;
    if ((rc = ensure_plan(d, k, rp.reserved[5], stream, &plan)) != RT_OK) return rc;
    if (plan_out) *plan_out = plan;
    k.packed = packed ? 1 : 0;
    static const bool sort_emit = [] { const char* e = std::getenv("RT_B200_SORT_EMIT"); return e && e[0] == '1'; }();
    k.sort_emit = sort_emit ? 1 : 0;
    if (!k.bvh.prune && k.bvh.use_bvh && !d->ref_tree && !h.tree.empty()) {
        static_assert(sizeof(TreeNode) == 40, "RefNode layout of traverse_reference_impl: 10 words per node");
        CUDA_TRY(cudaMalloc((void**)&d->ref_tree, h.tree.size() * sizeof(TreeNode)));
        CUDA_TRY(cudaMemcpyAsync(d->ref_tree, h.tree.data(), h.tree.size() * sizeof(TreeNode), cudaMemcpyHostToDevice, stream));
    }
    k.bvh.ref_tree = d->ref_tree;
    k.bvh.prims = d->prims; k.bvh.wide = d->wide; k.bvh.leafbox = d->leafbox; k.bvh.stack_depth = d->stack_depth;
    k.bvh.packet_stack_depth = d->packet_stack_depth; k.bvh.n_staged = d->n_staged;
    k.mats = d->mats; k.lights = d->lights; k.textures = d->textures; k.texels = d->texels;
    k.hit_ids = hit_ids;

    for (int attempt = 0;; ++attempt) {
        CUDA_TRY(cudaEventRecord(d->ev[0], stream));
        CUDA_TRY(cudaEventRecord(d->ev[1], stream));
        if ((rc = enqueue_frame(h, d, k, rp.collect_stats != 0, (rp.reserved[1] & 1) != 0, (rp.reserved[2] & 1) != 0, rgb8, linear, stream)) != RT_OK) return rc;
        CUDA_TRY(cudaEventRecord(d->ev[2], stream));
        d->timed = true;
        if (!stats) { d->async_pending = true; return RT_OK; }  // asynchronous: an overflow is reported by the next call

        std::memset(stats, 0, sizeof(*stats));
        CUDA_TRY(cudaEventSynchronize(d->ev[2]));
        d->async_pending = false;
        *d->overflow_host = 0u;
        unsigned long long c[8];
        CUDA_TRY(cudaMemcpy(c, d->totals, sizeof(c), cudaMemcpyDeviceToHost));
        if (c[T_OVERFLOW] && d->batch_slots > 32 && attempt < 40) {
            // a ray queue overflowed (heavily branching ray trees): smaller batches, same capacity
            const long long total_slots = (long long)k.n_my_tiles * k.sub_per_tile * k.spp * 32;
            d->batch_slots = std::max<long long>(32, std::min(d->batch_slots, total_slots) / 2);
            continue;
        }
        if (c[T_OVERFLOW]) { set_error("ray queue overflow even with the smallest batch"); return RT_ERR_CUDA; }
        stats->primary_rays = c[T_PRIMARY];
        stats->shadow_rays = c[T_SHADOW];
        stats->secondary_rays = c[T_SECONDARY];
        stats->rays = c[T_PRIMARY] + c[T_SHADOW] + c[T_SECONDARY];
        stats->node_visits = c[T_NODES];
        stats->prim_tests = c[T_PRIMS];
        CUDA_TRY(cudaEventElapsedTime(&stats->kernel_ms, d->ev[1], d->ev[2]));
        CUDA_TRY(cudaEventElapsedTime(&stats->total_ms, d->ev[0], d->ev[2]));
        stats->launches = d->last_launches;
        stats->pixels = (int32_t)plan->pixels;
        return RT_OK;
    }
}

static bool partial_frame(const rt_render_params& rp) { return rp.world > 1 || rp.reserved[3] != 0 || rp.reserved[4] != 0; }

// rt_render: host output buffers. Upload (if the device copy is stale) -> render -> copy back, all
// on the default stream; frame-sized device buffers are kept between calls. total_ms = CUDA-event
// time of everything from the upload to the last copy.
static int render_host(HostScene& h, const rt_render_params& rp, uint8_t* rgb8, int32_t* hit_ids, float* linear,
                       rt_render_stats* stats) {
    if (h.cam.res_x <= 0 || h.cam.res_y <= 0) { set_error("Camera resolution is 0. Check scene.json."); return RT_ERR_SCENE; }
    const cudaStream_t stream = 0;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device: the renderer has no CPU fallback");
        return RT_ERR_CUDA;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CUDA_TRY(cudaEventCreate(&e0));
    cudaError_t ee = cudaEventCreate(&e1);
    if (ee != cudaSuccess) { cudaEventDestroy(e0); set_error(std::string("cudaEventCreate: ") + cudaGetErrorString(ee)); return RT_ERR_CUDA; }
    auto body = [&]() -> int {
        CUDA_TRY(cudaEventRecord(e0, stream));
        DeviceScene* d = nullptr;
        int rc = ensure_uploaded(h, stream, nullptr, &d);
        if (rc != RT_OK) return rc;
        const size_t n = (size_t)h.cam.res_x * h.cam.res_y;
        if (n > d->out_pixels) {
            CUDA_TRY(cudaDeviceSynchronize());
            cudaFree(d->out_rgb); cudaFree(d->out_ids); cudaFree(d->out_lin);
            d->out_rgb = nullptr; d->out_ids = nullptr; d->out_lin = nullptr; d->out_pixels = 0;
            CUDA_TRY(cudaMalloc((void**)&d->out_rgb, n * 3));
            CUDA_TRY(cudaMalloc((void**)&d->out_ids, n * sizeof(int32_t)));
            CUDA_TRY(cudaMalloc((void**)&d->out_lin, n * 3 * sizeof(float)));
            d->out_pixels = n;
        }
        if (partial_frame(rp)) {  // pixels outside this rank's tiles / the window stay zero / -1 in the host buffers
            if (rgb8) CUDA_TRY(cudaMemsetAsync(d->out_rgb, 0, n * 3, stream));
            if (hit_ids) CUDA_TRY(cudaMemsetAsync(d->out_ids, 0xff, n * sizeof(int32_t), stream));
            if (linear) CUDA_TRY(cudaMemsetAsync(d->out_lin, 0, n * 3 * sizeof(float), stream));
        }
        rt_render_stats local;
        rc = render_impl(h, rp, rgb8 ? d->out_rgb : nullptr, hit_ids ? d->out_ids : nullptr, linear ? d->out_lin : nullptr, stream, &local);
        if (rc != RT_OK) return rc;
        if (rgb8) CUDA_TRY(cudaMemcpyAsync(rgb8, d->out_rgb, n * 3, cudaMemcpyDeviceToHost, stream));
        if (hit_ids) CUDA_TRY(cudaMemcpyAsync(hit_ids, d->out_ids, n * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        if (linear) CUDA_TRY(cudaMemcpyAsync(linear, d->out_lin, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaEventRecord(e1, stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        local.total_ms = ms;
        if (stats) *stats = local;
        return RT_OK;
    };
    const int rc = body();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// rt_render_multi: one process, N GPUs. The reference's frame loop (raytracer.cpp:433-476) is one
// serial loop over pixels; pixels are independent, so the frame is dealt to the GPUs by screen
// tiles (scene + BVH replicated, no inter-GPU traffic in the render loop). One persistent host
// thread per GPU: upload (all GPUs copy from the same page-locked staging buffer, in parallel) ->
// render its tiles into a PACKED buffer (tile-major over the tiles it owns) -> one D2H copy per
// output into page-locked memory -> scatter the tile rows into the caller's frame.
// ---------------------------------------------------------------------------------------------
struct MultiJob {
    HostScene* h = nullptr;
    rt_render_params rp{};
    uint8_t* rgb8 = nullptr;
    int32_t* hit_ids = nullptr;
    float* linear = nullptr;
};

struct MultiWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    int device = 0;
    bool has_job = false, done = false, quit = false;
    MultiJob job;
    int rc = RT_OK;
    std::string err;
    rt_render_stats stats{};
};

struct MultiGpu {
    std::vector<MultiWorker*> workers;
};

// Copies this device's packed tiles from the landing zone into the caller's frame.
template <typename T>
static void scatter_tiles(const FrameParams& k, const TilePlan& plan, const T* packed, T* frame, int channels) {
    for (size_t i = 0; i < plan.tiles.size(); ++i) {
        const int t = plan.tiles[i], tx = t % k.tiles_x, ty = t / k.tiles_x;
        const int x0 = std::max(tx * k.tile_w, k.win_x0), x1 = std::min((tx + 1) * k.tile_w, k.win_x1);
        const int y0 = std::max(ty * k.tile_h, k.win_y0), y1 = std::min((ty + 1) * k.tile_h, k.win_y1);
        for (int y = y0; y < y1; ++y) {
            const T* src = packed + (((size_t)i * k.tile_h + (y - ty * k.tile_h)) * k.tile_w + (x0 - tx * k.tile_w)) * channels;
            std::memcpy(frame + ((size_t)y * k.res_x + x0) * channels, src, (size_t)(x1 - x0) * channels * sizeof(T));
        }
    }
}

static int multi_render_shard(const MultiJob& job, rt_render_stats* stats) {
    HostScene& h = *job.h;
    FrameParams k;
    int rc = fill_params(h, job.rp, k);
    if (rc != RT_OK) return rc;
    DeviceScene* d = nullptr;
    if ((rc = device_scene(h, &d)) != RT_OK) return rc;
    if (!d->own) CUDA_TRY(cudaStreamCreateWithFlags(&d->own, cudaStreamNonBlocking));
    const cudaStream_t stream = d->own;
    if ((rc = ensure_uploaded(h, stream, nullptr, &d)) != RT_OK) return rc;
    // worst case of a shard: every tile of the frame (world = 1)
    const size_t cap = (size_t)k.n_tiles * k.tile_w * k.tile_h;
    if (cap > d->out_pixels) {
        CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(d->out_rgb); cudaFree(d->out_ids); cudaFree(d->out_lin);
        d->out_rgb = nullptr; d->out_ids = nullptr; d->out_lin = nullptr; d->out_pixels = 0;
        CUDA_TRY(cudaMalloc((void**)&d->out_rgb, cap * 3));
        CUDA_TRY(cudaMalloc((void**)&d->out_ids, cap * sizeof(int32_t)));
        CUDA_TRY(cudaMalloc((void**)&d->out_lin, cap * 3 * sizeof(float)));
        d->out_pixels = cap;
    }
    if (cap > d->host_pixels) {
        if (d->host_rgb) cudaFreeHost(d->host_rgb);
        if (d->host_ids) cudaFreeHost(d->host_ids);
        if (d->host_lin) cudaFreeHost(d->host_lin);
        d->host_rgb = nullptr; d->host_ids = nullptr; d->host_lin = nullptr; d->host_pixels = 0;
        CUDA_TRY(cudaHostAlloc((void**)&d->host_rgb, cap * 3, cudaHostAllocDefault));
        CUDA_TRY(cudaHostAlloc((void**)&d->host_ids, cap * sizeof(int32_t), cudaHostAllocDefault));
        CUDA_TRY(cudaHostAlloc((void**)&d->host_lin, cap * 3 * sizeof(float), cudaHostAllocDefault));
        d->host_pixels = cap;
    }
    TilePlan* plan = nullptr;
    rt_render_stats local;
    rc = render_impl(h, job.rp, job.rgb8 ? d->out_rgb : nullptr, job.hit_ids ? d->out_ids : nullptr, job.linear ? d->out_lin : nullptr,
                     stream, &local, true, &plan);
    if (rc != RT_OK) return rc;
    const size_t n = plan->tiles.size() * (size_t)k.tile_w * k.tile_h;
    if (n > 0) {
        if (job.rgb8) CUDA_TRY(cudaMemcpyAsync(d->host_rgb, d->out_rgb, n * 3, cudaMemcpyDeviceToHost, stream));
        if (job.hit_ids) CUDA_TRY(cudaMemcpyAsync(d->host_ids, d->out_ids, n * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        if (job.linear) CUDA_TRY(cudaMemcpyAsync(d->host_lin, d->out_lin, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (job.rgb8) scatter_tiles<uint8_t>(k, *plan, d->host_rgb, job.rgb8, 3);
        if (job.hit_ids) scatter_tiles<int32_t>(k, *plan, d->host_ids, job.hit_ids, 1);
        if (job.linear) scatter_tiles<float>(k, *plan, d->host_lin, job.linear, 3);
    }
    *stats = local;
    return RT_OK;
}

static void multi_worker_main(MultiWorker* w) {
    cudaSetDevice(w->device);
    std::unique_lock<std::mutex> lock(w->mu);
    while (true) {
        w->cv.wait(lock, [w] { return w->has_job || w->quit; });
        if (w->quit) return;
        MultiJob job = w->job;
        lock.unlock();
        rt_render_stats st;
        std::memset(&st, 0, sizeof(st));
        const int rc = multi_render_shard(job, &st);
        lock.lock();
        w->rc = rc;
        w->err = rc == RT_OK ? std::string() : std::string(rt_last_error());
        w->stats = st;
        w->has_job = false;
        w->done = true;
        w->cv.notify_all();
    }
}

static void multi_shutdown(MultiGpu* m) {
    if (!m) return;
    for (MultiWorker* w : m->workers) {
        { std::lock_guard<std::mutex> lock(w->mu); w->quit = true; }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    delete m;
}

static int render_multi(HostScene& h, const rt_render_params& rp_in, int n_devices, const int32_t* devices, uint8_t* rgb8,
                        int32_t* hit_ids, float* linear, rt_render_stats* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device: the renderer has no CPU fallback");
        return RT_ERR_CUDA;
    }
    if (n_devices < 1 || n_devices > count) { set_error("rt_render_multi: n_devices must be in [1, rt_device_count()]"); return RT_ERR_INVALID; }
    if (rp_in.world != 1 || rp_in.rank != 0) { set_error("rt_render_multi: rank/world must be 0/1 (the call deals the tiles itself)"); return RT_ERR_INVALID; }
    std::vector<int> ids((size_t)n_devices);
    for (int i = 0; i < n_devices; ++i) {
        ids[i] = devices ? devices[i] : i;
        if (ids[i] < 0 || ids[i] >= count) { set_error("rt_render_multi: device ordinal out of range"); return RT_ERR_INVALID; }
        for (int j = 0; j < i; ++j) if (ids[j] == ids[i]) { set_error("rt_render_multi: a device is listed twice"); return RT_ERR_INVALID; }
    }
    {
        FrameParams probe;
        const int rc = fill_params(h, rp_in, probe);
        if (rc != RT_OK) return rc;
    }
    int rc = ensure_staging(h);
    if (rc != RT_OK) return rc;
    DeviceSet* set = h.dev;
    if (!set->multi) set->multi = new MultiGpu();
    MultiGpu* m = set->multi;
    // workers are bound to a device for life; find or start the one of each requested device
    std::vector<MultiWorker*> use;
    for (int dev : ids) {
        MultiWorker* w = nullptr;
        for (MultiWorker* c : m->workers) if (c->device == dev) w = c;
        if (!w) {
            w = new MultiWorker();
            w->device = dev;
            w->th = std::thread(multi_worker_main, w);
            m->workers.push_back(w);
        }
        use.push_back(w);
    }
    // pixels no device renders (outside the window) keep the "nothing rendered" values of rt_render
    if (rp_in.reserved[3] != 0 || rp_in.reserved[4] != 0) {
        const size_t n = (size_t)h.cam.res_x * h.cam.res_y;
        if (rgb8) std::memset(rgb8, 0, n * 3);
        if (hit_ids) std::memset(hit_ids, 0xff, n * sizeof(int32_t));
        if (linear) std::memset(linear, 0, n * 3 * sizeof(float));
    }
    for (int i = 0; i < n_devices; ++i) {
        MultiWorker* w = use[i];
        std::lock_guard<std::mutex> lock(w->mu);
        w->job.h = &h;
        w->job.rp = rp_in;
        w->job.rp.rank = i;
        w->job.rp.world = n_devices;
        w->job.rgb8 = rgb8; w->job.hit_ids = hit_ids; w->job.linear = linear;
        w->done = false;
        w->has_job = true;
        w->cv.notify_all();
    }
    rt_render_stats total;
    std::memset(&total, 0, sizeof(total));
    int first_rc = RT_OK;
    std::string first_err;
    for (MultiWorker* w : use) {
        std::unique_lock<std::mutex> lock(w->mu);
        w->cv.wait(lock, [w] { return w->done; });
        if (w->rc != RT_OK && first_rc == RT_OK) { first_rc = w->rc; first_err = "device " + std::to_string(w->device) + ": " + w->err; }
        total.rays += w->stats.rays; total.primary_rays += w->stats.primary_rays; total.shadow_rays += w->stats.shadow_rays;
        total.secondary_rays += w->stats.secondary_rays; total.node_visits += w->stats.node_visits; total.prim_tests += w->stats.prim_tests;
        total.kernel_ms = std::max(total.kernel_ms, w->stats.kernel_ms);
        total.launches += w->stats.launches;
        total.pixels += w->stats.pixels;
    }
    if (first_rc != RT_OK) { set_error(first_err); return first_rc; }
    total.total_ms = (float)(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3);
    if (stats) *stats = total;
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------
// Traversal ceiling: trav_step -- the very code of the traversal loop -- with all 32 lanes of every
// warp busy on nodes that stay in L1 (`n_nodes` synthetic nodes whose children all contain the scene,
// visited round robin; the stack is reset after every step). Box tests per second of that loop on this
// chip is what the traversal kernels would reach without divergence, cache misses, fetch / primitive
// phases: the denominator of bench.py's roofline.
// ---------------------------------------------------------------------------------------------
template <bool ANY>
__global__ void __launch_bounds__(RT_TRACE_THREADS, RT_TRACE_MINBLOCKS) trav_peak_kernel(BvhView bvh, int n_nodes, int steps, float3 eye, float3 lo, float3 hi,
                                                                                            unsigned long long* out) {
    const unsigned int words = ANY ? 1u : (unsigned int)RT_STACK_WORDS;
    const unsigned int stride = blockDim.x * words * (unsigned int)sizeof(int);
    TravState s;
    s.sp0 = (unsigned int)__cvta_generic_to_shared(rt_stack_smem + threadIdx.x * words);
    s.sp_end = s.sp0 + (unsigned int)bvh.stack_depth * stride;
    const unsigned int gid = blockIdx.x * blockDim.x + threadIdx.x;
    // a ray from the eye into the scene box (never axis-parallel)
    const U4 u = philox4x32_10(U4{gid, 0x77u, 0u, 0u}, 0x9E3779B9u, 0xBB67AE85u);
    Ray r;
    r.ox = eye.x; r.oy = eye.y; r.oz = eye.z;
    r.dx = lo.x + (hi.x - lo.x) * u32_to_unit_float(u.x) - eye.x;
    r.dy = lo.y + (hi.y - lo.y) * u32_to_unit_float(u.y) - eye.y;
    r.dz = lo.z + (hi.z - lo.z) * u32_to_unit_float(u.z) - eye.z;
    if (fabsf(r.dx) < 1e-3f) r.dx = 1e-3f;
    if (fabsf(r.dy) < 1e-3f) r.dy = 1e-3f;
    if (fabsf(r.dz) < 1e-3f) r.dz = 1e-3f;
    normalize3(r.dx, r.dy, r.dz);
    r.time = 0.0f;
    TraceStats st = {0u, 0u};
    BvhView b = bvh;
    b.prune = 1;
    b.use_bvh = 1;
    trav_begin<ANY>(b, s, r, 1e30f);
    unsigned long long acc = 0;
    int node = (int)(gid % (unsigned int)n_nodes);
#pragma unroll 1
    for (int i = 0; i < steps; ++i) {
        s.cur = node;
        s.sp = s.sp0;
        s.pend = 0u;
        trav_step<ANY, false>(b, s, stride, st);
        acc += (unsigned int)s.cur + s.pend + (s.sp - s.sp0);
        node = node + 1 == n_nodes ? 0 : node + 1;
    }
    if (acc == 0x1234567887654321ull) out[1] = acc;  // keeps the loop alive
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)gridDim.x * blockDim.x * (unsigned long long)steps * 4ull;
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
    return rtb::ensure_uploaded(*rtb::host_of(scene), 0, bytes);
}

int rt_scene_evict(rt_scene* scene) {
    if (!scene) { rtb::set_error("null scene"); return RT_ERR_INVALID; }
    rtb::device_invalidate(*rtb::host_of(scene));
    return RT_OK;
}

static rtb::DeviceScene* current_device_scene(rt_scene* scene) {
    rtb::DeviceSet* set = rtb::host_of(scene)->dev;
    int dev = 0;
    if (!set || cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lock(set->mu);
    return dev < (int)set->devs.size() ? set->devs[dev] : nullptr;
}

int rt_scene_last_timing(rt_scene* scene, float* kernel_ms, float* total_ms) {
    if (!scene) { rtb::set_error("null scene"); return RT_ERR_INVALID; }
    rtb::DeviceScene* d = current_device_scene(scene);
    if (!d || !d->timed) { rtb::set_error("rt_scene_last_timing: no render has been recorded on this scene"); return RT_ERR_INVALID; }
    cudaError_t e = cudaEventSynchronize(d->ev[2]);
    float k = 0.0f, t = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&k, d->ev[1], d->ev[2]);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, d->ev[0], d->ev[2]);
    if (e != cudaSuccess) { rtb::set_error(std::string("rt_scene_last_timing: ") + cudaGetErrorString(e)); return RT_ERR_CUDA; }
    // an asynchronous frame cannot re-render itself when a ray queue overflows: report it here
    const int rc = rtb::report_async_overflow(d);
    if (rc != RT_OK) return rc;
    d->async_pending = false;  // finished, looked at, clean
    if (kernel_ms) *kernel_ms = k;
    if (total_ms) *total_ms = t;
    return RT_OK;
}

int rt_scene_last_kernel_times(rt_scene* scene, float* ms4, int32_t* launches4, int32_t* frame_launches4) {
    if (!scene) { rtb::set_error("null scene"); return RT_ERR_INVALID; }
    rtb::DeviceScene* d = current_device_scene(scene);
    if (!d || !d->timed) { rtb::set_error("rt_scene_last_kernel_times: no render has been recorded on this scene"); return RT_ERR_INVALID; }
    float ms[4] = {0, 0, 0, 0};
    int32_t n[4] = {0, 0, 0, 0};
    cudaError_t e = cudaEventSynchronize(d->ev[2]);
    for (size_t i = 0; e == cudaSuccess && i < d->class_of.size(); ++i) {
        if (d->class_of[i] & 0x100) continue;
        float t = 0.0f;
        e = cudaEventElapsedTime(&t, d->class_ev[2 * i], d->class_ev[2 * i + 1]);
        ms[d->class_of[i] & 3] += t;
        n[d->class_of[i] & 3] += 1;
    }
    if (e != cudaSuccess) { rtb::set_error(std::string("rt_scene_last_kernel_times: ") + cudaGetErrorString(e)); return RT_ERR_CUDA; }
    for (int i = 0; i < 4; ++i) {
        if (ms4) ms4[i] = ms[i];
        if (launches4) launches4[i] = n[i];
        if (frame_launches4) frame_launches4[i] = d->class_launches[i];
    }
    return RT_OK;
}

int rt_scene_launch_count(rt_scene* scene, uint64_t* launches) {
    if (!scene || !launches) { rtb::set_error("null argument"); return RT_ERR_INVALID; }
    *launches = 0;
    rtb::DeviceSet* set = rtb::host_of(scene)->dev;
    if (!set) return RT_OK;
    std::lock_guard<std::mutex> lock(set->mu);
    for (rtb::DeviceScene* d : set->devs) if (d) *launches += d->total_launches;
    return RT_OK;
}

static int selftest_grid() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return 1184; }
    return std::max(1, sms) * 8;
}

int rt_selftest_boxes(uint64_t seed, int64_t n, uint64_t* out8) {
    if (!out8 || n <= 0) { rtb::set_error("rt_selftest_boxes: bad argument"); return RT_ERR_INVALID; }
    unsigned long long* d = nullptr;
    if (cudaMalloc((void**)&d, 8 * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); rtb::set_error("no CUDA device"); return RT_ERR_CUDA; }
    cudaMemset(d, 0, 8 * sizeof(unsigned long long));
    rtb::selftest_box_kernel<<<selftest_grid(), 256>>>((uint32_t)seed, (long long)n, d);
    const cudaError_t e = cudaMemcpy(out8, d, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { rtb::set_error(std::string("rt_selftest_boxes: ") + cudaGetErrorString(e)); return RT_ERR_CUDA; }
    return RT_OK;
}

int rt_selftest_cull(rt_scene* scene, uint64_t seed, int32_t rays_per_primitive, uint64_t* out8) {
    if (!scene || !out8 || rays_per_primitive <= 0) { rtb::set_error("rt_selftest_cull: bad argument"); return RT_ERR_INVALID; }
    rtb::HostScene& h = *rtb::host_of(scene);
    rtb::DeviceScene* ds = nullptr;
    int rc = rtb::ensure_uploaded(h, 0, nullptr, &ds);
    if (rc != RT_OK) return rc;
    rtb::BvhView b;
    std::memset(&b, 0, sizeof(b));
    b.prims = ds->prims; b.wide = ds->wide; b.leafbox = ds->leafbox; b.n_prims = (int)h.dprims.size(); b.use_bvh = 1; b.prune = 1;
    float span = 1.0f;
    if (!h.tree.empty())
        for (int a = 0; a < 3; ++a) span = std::max(span, h.tree[0].box.hi[a] - h.tree[0].box.lo[a]);
    for (int a = 0; a < 3; ++a) span = std::max(span, 2.0f * std::fabs(h.cam.location[a]));
    unsigned long long* d = nullptr;
    if (cudaMalloc((void**)&d, 8 * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); rtb::set_error("cudaMalloc failed"); return RT_ERR_CUDA; }
    cudaMemset(d, 0, 8 * sizeof(unsigned long long));
    rtb::selftest_cull_kernel<<<selftest_grid(), 256>>>(b, (int)h.dwide.size(), rays_per_primitive, (uint32_t)seed, span, d);
    const cudaError_t e = cudaMemcpy(out8, d, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { rtb::set_error(std::string("rt_selftest_cull: ") + cudaGetErrorString(e)); return RT_ERR_CUDA; }
    return RT_OK;
}

int rt_traversal_peak(rt_scene* scene, int32_t any_hit, int32_t n_nodes, int32_t steps, int32_t repeats, double* box_tests_per_s, float* ms) {
    if (!scene || !box_tests_per_s || n_nodes < 1 || steps < 1 || repeats < 1) { rtb::set_error("rt_traversal_peak: bad argument"); return RT_ERR_INVALID; }
    rtb::HostScene& h = *rtb::host_of(scene);
    if (h.dwide.empty()) { rtb::set_error("rt_traversal_peak: the scene has no shapes"); return RT_ERR_SCENE; }
    rtb::DeviceScene* ds = nullptr;
    int rc = rtb::ensure_uploaded(h, 0, nullptr, &ds);
    if (rc != RT_OK) return rc;
    rtb::BvhView b;
    std::memset(&b, 0, sizeof(b));
    b.prims = ds->prims; b.wide = ds->wide; b.leafbox = ds->leafbox; b.n_prims = (int)h.dprims.size(); b.use_bvh = 1; b.prune = 1;
    b.stack_depth = ds->stack_depth;
    // FULL-WORK steps on synthetic nodes: every node has four node children whose boxes all contain the scene, so
    // every ray passes all four -- four slab tests, the whole sorting network, three pushes, one descent: every
    // instruction of the step does work, which no real traversal exceeds per box test. (The top of a real tree would
    // make the number depend on the scene: steps whose children all miss skip the pushes warp-wide.)
    const int n = std::max(4, std::min(n_nodes, 4096));
    const rtb::Box& box = h.tree[0].box;
    std::vector<rtb::DWide> synth((size_t)n);
    for (int i = 0; i < n; ++i) {
        rtb::DWide w;
        for (float& x : w.f) x = 0.0f;
        for (int k = 0; k < 4; ++k) {
            for (int a = 0; a < 3; ++a) {
                const float ext = box.hi[a] - box.lo[a];
                w.f[8 * a + k] = box.lo[a] - (1.0f + 0.25f * k) * ext - 1.0f;      // lo: slightly different entry distances per child
                w.f[8 * a + 4 + k] = box.hi[a] + (1.0f + 0.25f * k) * ext + 1.0f;  // hi
            }
        }
        const uint32_t first = (uint32_t)((i * 4 + 1) % (n - 3)), meta = 0xFu;
        std::memcpy(&w.f[24], &first, 4);
        std::memcpy(&w.f[25], &meta, 4);
        synth[(size_t)i] = w;
    }
    float* d_synth = nullptr;
    if (cudaMalloc((void**)&d_synth, synth.size() * sizeof(rtb::DWide)) != cudaSuccess ||
        cudaMemcpy(d_synth, synth.data(), synth.size() * sizeof(rtb::DWide), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(d_synth);
        rtb::set_error("rt_traversal_peak: cudaMalloc / cudaMemcpy failed");
        return RT_ERR_CUDA;
    }
    b.wide = d_synth;
    const float3 eye = make_float3(h.cam.location[0], h.cam.location[1], h.cam.location[2]);
    const float3 lo = make_float3(box.lo[0], box.lo[1], box.lo[2]), hi = make_float3(box.hi[0], box.hi[1], box.hi[2]);
    unsigned long long* d = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto body = [&]() -> int {
        CUDA_TRY(cudaMalloc((void**)&d, 2 * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(d, 0, 2 * sizeof(unsigned long long)));
        CUDA_TRY(cudaEventCreate(&e0));
        CUDA_TRY(cudaEventCreate(&e1));
        const int grid = ds->sm_count * (any_hit ? ds->shadow_blocks : ds->trace_blocks);
        double best = 0.0;
        float best_ms = 0.0f;
        for (int r = 0; r < repeats + 1; ++r) {  // first launch = warm-up
            CUDA_TRY(cudaEventRecord(e0, 0));
            if (any_hit) rtb::trav_peak_kernel<true><<<grid, RT_TRACE_THREADS, ds->stack_bytes>>>(b, n, steps, eye, lo, hi, d);
            else rtb::trav_peak_kernel<false><<<grid, RT_TRACE_THREADS, ds->stack_bytes>>>(b, n, steps, eye, lo, hi, d);
            CUDA_TRY(cudaEventRecord(e1, 0));
            CUDA_TRY(cudaEventSynchronize(e1));
            float t = 0.0f;
            CUDA_TRY(cudaEventElapsedTime(&t, e0, e1));
            unsigned long long tests = 0;
            CUDA_TRY(cudaMemcpy(&tests, d, sizeof(tests), cudaMemcpyDeviceToHost));
            if (r > 0 && t > 0.0f && (double)tests / (t * 1e-3) > best) { best = (double)tests / (t * 1e-3); best_ms = t; }
        }
        *box_tests_per_s = best;
        if (ms) *ms = best_ms;
        return RT_OK;
    };
    rc = body();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d);
    cudaFree(d_synth);
    return rc;
}

int rt_shard_pixels(const rt_scene* scene, const rt_render_params* p, int64_t* n_pixels) {
    if (!scene || !p || !n_pixels) { rtb::set_error("null argument"); return RT_ERR_INVALID; }
    rtb::FrameParams k;
    int rc = rtb::fill_params(*rtb::host_of(scene), *p, k);
    if (rc != RT_OK) return rc;
    std::vector<int> tiles;
    int64_t pixels = 0;
    rtb::plan_tiles(k, p->reserved[5], tiles, pixels);
    *n_pixels = pixels;
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
    return rtb::render_host(*rtb::host_of(scene), *p, rgb8, hit_ids, linear, stats);
}

int rt_render_multi(rt_scene* scene, const rt_render_params* p, int32_t n_devices, const int32_t* devices, uint8_t* rgb8,
                    int32_t* hit_ids, float* linear, rt_render_stats* stats) {
    if (!scene || !p) { rtb::set_error("null argument"); return RT_ERR_INVALID; }
    return rtb::render_multi(*rtb::host_of(scene), *p, n_devices, devices, rgb8, hit_ids, linear, stats);
}

}  // extern "C"
