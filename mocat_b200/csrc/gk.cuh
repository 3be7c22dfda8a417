// gk.cuh -- the g-and-k simulator shared by the SMC-ABC kernels (abc.cu) and tempered EKI (teki.cu).
#pragma once
#include "common.cuh"
#include "rng.cuh"

#define GK_DIM 4

// likelihood_sample of GKTransformedUniformPrior (abc/scenarios/gk.py:38-41,68-96): constrain, M quantile draws from the
// particle's Philox stream (purpose MB_P_SIM, slots index0 ...), summary = the draws sorted ascending
template <int M>
__device__ __forceinline__ void gk_simulate(const mb_gk& g, const float (&x)[GK_DIM], uint64_t seed, uint64_t gid,
                                            uint32_t step, uint32_t index0, float (&y)[M]) {
    // constrain (gk.py:70-72): theta = min + Phi(x) (max - min)
    float th[GK_DIM];
#pragma unroll
    for (int k = 0; k < GK_DIM; ++k) th[k] = fmaf(normcdff(x[k]), g.prior_max - g.prior_min, g.prior_min);
#pragma unroll
    for (int s = 0; s < (M + 3) / 4; ++s) {
        const Philox4 r = philox_raw(seed, gid, step, MB_P_SIM, index0 + s);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (4 * s + j < M) {
                const float u = fmaf(u24(w[j]), 1.f - 2.f * g.buffer, g.buffer);     // U(buffer, 1-buffer), gk.py:83
                const float z = normcdfinvf(u);                                       // norm.ppf, :84
                const float e = __expf(-th[2] * z);                                   // :85
                y[4 * s + j] = fmaf(th[1] * (1.f + g.c * (1.f - e) / (1.f + e)) * z, __powf(fmaf(z, z, 1.f), th[3]), th[0]);
            }
        }
    }
    // bitonic sorting network, fully unrolled (summary statistic = order statistics)
#pragma unroll
    for (int k = 2; k <= M; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = ((i & k) == 0);
                    const float a = y[i], b = y[l];
                    const bool sw = up ? (a > b) : (a < b);
                    y[i] = sw ? b : a;
                    y[l] = sw ? a : b;
                }
            }
        }
    }
}
