// rng.cuh -- Philox4x32-10 counter RNG and uniform/normal conversions (mirrored bit-for-bit on the
// integer/uniform level by oracle/philox.py).  Replaces the jax.random (threefry) draws of the
// reference, e.g. transport/smc.py:82, ssm/filtering.py:294, mcmc/standard_mcmc.py:54-56,120-122.
#pragma once
#include <stdint.h>

#define MB_P_INIT     0u
#define MB_P_MOVE     1u
#define MB_P_RESAMPLE 2u
#define MB_P_SIM      3u

struct Philox4 { uint32_t x, y, z, w; };

template <int ROUNDS = 10>
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        if (r > 0) { k0 += W0; k1 += W1; }
        // __umulhi + 32-bit product fuse into ONE IMAD.WIDE each; the (uint64_t) form costs two stray adds per round
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    }
    return Philox4{c0, c1, c2, c3};
}

template <int ROUNDS = 10>
__device__ __forceinline__ Philox4 philox_raw(uint64_t seed, uint64_t gid, uint32_t step, uint32_t purpose,
                                              uint32_t index) {
    return philox4x32_10<ROUNDS>((uint32_t)gid, (uint32_t)(gid >> 32), step, (purpose << 20) | index,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// [0,1) with 24 bits, exact in fp32
__device__ __forceinline__ float u24(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }
// (0,1]: (float(x) + 0.5f) * 2^-32, round-to-nearest at each step
__device__ __forceinline__ float u_open(uint32_t x) {
    return __fmul_rn(__fadd_rn(__uint2float_rn(x), 0.5f), 2.3283064365386963e-10f);
}
// [0,1) fp64 with 53 bits from two words
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * 1.1102230246251565e-16;
}

// Box-Muller: z0 = r cos(2 pi u2), z1 = r sin(2 pi u2), r = sqrt(-2 ln u1).
// cos/sin evaluated on phi = 2 pi u2 - pi in [-pi, pi) where the MUFU approximations are accurate:
// cos(2 pi u2) = -cos(phi), sin(2 pi u2) = -sin(phi).
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& z0, float& z1) {
    const float u1 = u_open(xa);                       // in [2^-33, 1]: never denormal, so the raw MUFU ops are safe
    const float u2 = u24(xb);
    float l2, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l2 * -1.3862943611198906f));   // sqrt(-2 ln u1), -2 ln 2 folded
    const float phi = fmaf(u2, 6.283185307179586f, -3.141592653589793f);
    float s, c;
    __sincosf(phi, &s, &c);
    z0 = -r * c;
    z1 = -r * s;
}

// D standard normals for particle gid: normal j <- slot index0 + j/4, words (0,1) / (2,3).
template <int D>
__device__ __forceinline__ void philox_normals(float (&z)[D], uint64_t seed, uint64_t gid, uint32_t step,
                                               uint32_t purpose, uint32_t index0) {
#pragma unroll
    for (int s = 0; s < (D + 3) / 4; ++s) {
        const Philox4 r = philox_raw(seed, gid, step, purpose, index0 + s);
        float a, b, c, d;
        box_muller(r.x, r.y, a, b);
        box_muller(r.z, r.w, c, d);
        if (4 * s + 0 < D) z[4 * s + 0] = a;
        if (4 * s + 1 < D) z[4 * s + 1] = b;
        if (4 * s + 2 < D) z[4 * s + 2] = c;
        if (4 * s + 3 < D) z[4 * s + 3] = d;
    }
}

// D standard normals AND the accept uniform of one Metropolised move.  When D mod 4 is 1 or 2 the last normal slot has two
// unused words: the uniform is word 2 of that slot (no extra Philox call); otherwise it is word 0 of the next slot.
// Slots consumed per move: MB_MOVE_SLOTS(D).  Mirrored by oracle/smc.py.
#define MB_MOVE_SPARE(D) ((D) % 4 == 1 || (D) % 4 == 2)
#define MB_MOVE_SLOTS(D) (((D) + 3) / 4 + (MB_MOVE_SPARE(D) ? 0 : 1))
template <int D>
__device__ __forceinline__ void philox_normals_accept(float (&z)[D], float& uacc, uint64_t seed, uint64_t gid, uint32_t step,
                                                      uint32_t purpose, uint32_t index0) {
    constexpr int NZ = (D + 3) / 4;
#pragma unroll
    for (int s = 0; s < NZ; ++s) {
        const Philox4 r = philox_raw(seed, gid, step, purpose, index0 + s);
        float a, b, c, d;
        box_muller(r.x, r.y, a, b);
        if (4 * s + 0 < D) z[4 * s + 0] = a;
        if (4 * s + 1 < D) z[4 * s + 1] = b;
        if (4 * s + 2 < D) {
            box_muller(r.z, r.w, c, d);
            z[4 * s + 2] = c;
            if (4 * s + 3 < D) z[4 * s + 3] = d;
        } else if (MB_MOVE_SPARE(D) && s == NZ - 1) {
            uacc = u24(r.z);
        }
    }
    if (!MB_MOVE_SPARE(D)) uacc = u24(philox_raw(seed, gid, step, purpose, index0 + NZ).x);
}
