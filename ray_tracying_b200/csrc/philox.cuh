// philox.cuh -- Philox4x32-10 counter-based RNG (Salmon et al., SC'11), usable on host and device.
//
// The reference draws every random number from ONE serial std::mt19937 stream
// (raytracer.cpp:425-427), which cannot be reproduced in parallel. We key a counter-based
// generator by what the number is FOR, so any thread can produce any draw independently:
//     key     = (pixel index, seed low)
//     counter = (sample index, purpose << 28 | ray-tree node id, light << 16 | shadow sample, try)
// with seed high folded into the key. The distributions drawn from it are the reference's
// (uniform [0,1), rejection-sampled unit disk / unit ball).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

namespace rtb {

struct U4 { uint32_t x, y, z, w; };

RT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

RT_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
        U4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += W0;
        k1 += W1;
    }
    return c;
}

enum RngPurpose : uint32_t { RNG_CAMERA = 0, RNG_LENS = 1, RNG_LIGHT = 2, RNG_GLOSSY = 3 };

RT_HD U4 rt_rng(uint32_t pixel, uint32_t seed_lo, uint32_t seed_hi, uint32_t sample, uint32_t purpose, uint32_t node,
                uint32_t sub, uint32_t attempt) {
    U4 c;
    c.x = sample;
    c.y = (purpose << 28) | node;
    c.z = sub;
    c.w = attempt;
    return philox4x32_10(c, pixel ^ (seed_hi * 0x9E3779B1u), seed_lo);
}

// uniform float in [0,1) with 24 random bits
RT_HD float u32_to_unit_float(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// uniform double in [0,1) with 32 random bits
RT_HD double u32_to_unit_double(uint32_t x) { return (double)x * (1.0 / 4294967296.0); }

}  // namespace rtb
