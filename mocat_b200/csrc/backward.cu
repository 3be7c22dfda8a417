// backward.cu -- backward simulation (FFBSi) for the Gaussian-transition state-space models: SURVEY 8 (f1).
//
// Replaces `full_resampling` / `backward_simulation_full` (ssm/backward.py:20-40, 241-272): for every backward sample
// x1_j at time t+1 draw an index i ~ Cat(lw_i - transition_potential(x0_i -> x1_j)) over the n_pf filter particles at
// time t (random.categorical = Gumbel-max) and take x0_i.  An n_samples x n_pf contraction with a categorical-sample
// epilogue; the transition is Gaussian (linear_gaussian.py:73-84, nonlinear_gaussian.py:98-105), so after whitening
// (a_i = L_Q^-1 mean(x0_i), b_j = L_Q^-1 x1_j) the potential is |a_i - b_j|^2 / 2 + const and the constant cancels.
//
//   bs_means_kernel    a_i for every filter particle: F x0 (linear-Gaussian, forward substitution with L_Q) or the RK4
//                      flow of Lorenz-96 scaled by 1 / q_std; one thread per particle, row-major (n, d).
//   bs_sample_kernel   one thread per backward sample j, tiles of 128 filter particles staged in shared memory;
//                      s_ij = lw_i - |a_i - b_j|^2 / 2 (fp32) + Gumbel noise from Philox (counter: gid = j, step = time
//                      index, purpose MB_P_BACKWARD, slot i / 4, word i mod 4); running arg-max (first index wins ties).
// oracle/backward.py restates it in NumPy; exact fp32-vs-fp64 agreement of the arg-max fails only for near-ties.
#include "common.cuh"
#include "rng.cuh"

#define MB_P_BACKWARD 4u
#define BS_THREADS 128
#define BS_TILE 128

struct BsArgs {
    mb_ssm ssm; float dt;
    const float* x0; const float* lw0; int64_t n_pf;      // filter particles at time t, row-major (n_pf, d), log-weights
    const float* x1; int64_t n_s;                         // backward samples at time t+1, row-major (n_s, d); NULL: none
    float* a;                                             // workspace (n_pf, d): whitened predicted means
    uint64_t seed; uint32_t step;
    int32_t* idx; float* x_out;                           // chosen index and state x0[idx] (n_s, d)
};

template <int D>
__device__ __forceinline__ void bs_l96_rhs(const float (&x)[D], float F, float (&k)[D]) {
#pragma unroll
    for (int r = 0; r < D; ++r) k[r] = (x[(r + 1) % D] - x[(r + D - 2) % D]) * x[(r + D - 1) % D] + (F - x[r]);
}

template <int D>
__global__ void __launch_bounds__(BS_THREADS) bs_means_kernel(BsArgs a) {
    const mb_ssm& m = a.ssm;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_pf; i += (int64_t)gridDim.x * blockDim.x) {
        float x[D], v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = a.x0[i * D + k];
        if (m.kind == MB_SSM_LINEAR_GAUSSIAN) {
#pragma unroll
            for (int r = 0; r < D; ++r) {                              // mean = F x0
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < D; ++c) acc = fmaf(m.F[r * MB_MAX_SMALL_DIM + c], x[c], acc);
                v[r] = acc;
            }
#pragma unroll
            for (int r = 0; r < D; ++r) {                              // L_Q z = mean (forward substitution)
                float acc = v[r];
#pragma unroll
                for (int c = 0; c < r; ++c) acc = fmaf(-m.LQ[r * MB_MAX_SMALL_DIM + c], v[c], acc);
                v[r] = acc / m.LQ[r * MB_MAX_SMALL_DIM + r];
            }
        } else {                                                       // Lorenz-96: `substeps` RK4 steps, then / q_std
            const float h = a.dt / (float)m.substeps;
            for (int s = 0; s < m.substeps; ++s) {
                float k1[D], k2[D], k3[D], k4[D], t[D];
                bs_l96_rhs<D>(x, m.forcing, k1);
#pragma unroll
                for (int r = 0; r < D; ++r) t[r] = fmaf(0.5f * h, k1[r], x[r]);
                bs_l96_rhs<D>(t, m.forcing, k2);
#pragma unroll
                for (int r = 0; r < D; ++r) t[r] = fmaf(0.5f * h, k2[r], x[r]);
                bs_l96_rhs<D>(t, m.forcing, k3);
#pragma unroll
                for (int r = 0; r < D; ++r) t[r] = fmaf(h, k3[r], x[r]);
                bs_l96_rhs<D>(t, m.forcing, k4);
#pragma unroll
                for (int r = 0; r < D; ++r) x[r] = fmaf(h * (1.f / 6.f), (k1[r] + 2.f * k2[r]) + (2.f * k3[r] + k4[r]), x[r]);
            }
            const float iq = 1.f / m.q_std;
#pragma unroll
            for (int r = 0; r < D; ++r) v[r] = x[r] * iq;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) a.a[i * D + k] = v[k];
    }
}

template <int D>
__global__ void __launch_bounds__(BS_THREADS) bs_sample_kernel(BsArgs a) {
    __shared__ float sa[BS_TILE * D];
    __shared__ float slw[BS_TILE];
    const mb_ssm& m = a.ssm;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = j < a.n_s;
    float b[D];
#pragma unroll
    for (int k = 0; k < D; ++k) b[k] = 0.f;
    if (valid && a.x1) {                                               // whiten the backward sample like the means
        float v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = a.x1[j * D + k];
        if (m.kind == MB_SSM_LINEAR_GAUSSIAN) {
#pragma unroll
            for (int r = 0; r < D; ++r) {
                float acc = v[r];
#pragma unroll
                for (int c = 0; c < r; ++c) acc = fmaf(-m.LQ[r * MB_MAX_SMALL_DIM + c], v[c], acc);
                v[r] = acc / m.LQ[r * MB_MAX_SMALL_DIM + r];
            }
        } else {
            const float iq = 1.f / m.q_std;
#pragma unroll
            for (int r = 0; r < D; ++r) v[r] *= iq;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) b[k] = v[k];
    }
    float best = -INFINITY;
    int64_t arg = 0;
    for (int64_t i0 = 0; i0 < a.n_pf; i0 += BS_TILE) {
        const int cnt = (int)min((int64_t)BS_TILE, a.n_pf - i0);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * D; e += BS_THREADS) sa[e] = a.x1 ? a.a[i0 * D + e] : 0.f;
        for (int e = threadIdx.x; e < cnt; e += BS_THREADS) slw[e] = a.lw0[i0 + e];
        __syncthreads();
        if (!valid) continue;
        for (int q = 0; q < cnt; q += 4) {                             // one Philox call feeds four candidates
            const Philox4 r = philox_raw(a.seed, (uint64_t)j, a.step, MB_P_BACKWARD, (uint32_t)((i0 + q) >> 2));
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (q + c >= cnt) break;
                float quad = 0.f;
                if (a.x1) {
#pragma unroll
                    for (int k = 0; k < D; ++k) { const float df = sa[(q + c) * D + k] - b[k]; quad = fmaf(df, df, quad); }
                }
                const float gum = -__logf(-__logf(u_open(w[c])));       // Gumbel(0, 1): random.categorical is Gumbel-max
                const float s = fmaf(-0.5f, quad, slw[q + c]) + gum;
                if (s > best) { best = s; arg = i0 + q + c; }
            }
        }
    }
    if (valid) {
        a.idx[j] = (int32_t)arg;
#pragma unroll
        for (int k = 0; k < D; ++k) a.x_out[j * D + k] = a.x0[arg * D + k];
    }
}

// x0 (n_pf, d), lw0 (n_pf), x1 (n_s, d) or NULL (no transition term: a plain categorical draw from the weights, the
// final-time draw of backward_simulation), work (n_pf * d floats), idx (n_s), x_out (n_s, d): all device, row-major
extern "C" int mb_backward_sample(mb_ctx* ctx, const mb_ssm* ssm, float dt, const float* x0, const float* lw0, int64_t n_pf,
                                  const float* x1, int64_t n_s, float* work, uint64_t seed, uint32_t step, int32_t* idx,
                                  float* x_out, mb_stream_t stream) {
    MB_REQUIRE(ctx && ssm && x0 && lw0 && idx && x_out && n_pf > 0 && n_s > 0 && n_pf < 0x7fffffffll && (x1 == nullptr || work),
               "mb_backward_sample: bad arguments");
    MB_REQUIRE(n_pf % 4 == 0 || true, "mb_backward_sample: internal");
    BsArgs a{};
    a.ssm = *ssm; a.dt = dt; a.x0 = x0; a.lw0 = lw0; a.n_pf = n_pf; a.x1 = x1; a.n_s = n_s; a.a = work;
    a.seed = seed; a.step = step; a.idx = idx; a.x_out = x_out;
    cudaStream_t st = mb_s(stream);
    const unsigned g0 = (unsigned)((n_pf + BS_THREADS - 1) / BS_THREADS), g1 = (unsigned)((n_s + BS_THREADS - 1) / BS_THREADS);
#define BS_CASE(DD)                                                                        \
    if (ssm->dim == DD) {                                                                  \
        if (x1) bs_means_kernel<DD><<<g0, BS_THREADS, 0, st>>>(a);                         \
        bs_sample_kernel<DD><<<g1, BS_THREADS, 0, st>>>(a);                                \
        MB_CHECK_LAUNCH();                                                                 \
        return MB_OK;                                                                      \
    }
    if (ssm->kind == MB_SSM_LINEAR_GAUSSIAN) { BS_CASE(1) BS_CASE(2) BS_CASE(3) BS_CASE(4) BS_CASE(5) BS_CASE(6) BS_CASE(8) }
    else if (ssm->kind == MB_SSM_LORENZ96) { BS_CASE(8) BS_CASE(16) BS_CASE(40) }
    mb_set_error("mb_backward_sample: model kind %d with dimension %d is not built", ssm->kind, ssm->dim);
    return MB_ERR_UNSUPPORTED;
}
