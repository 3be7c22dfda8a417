// backward.cu -- backward simulation (FFBSi) for the Gaussian-transition state-space models: SURVEY 8 (f1).
//
// Replaces `full_resampling` / `backward_simulation_full` (ssm/backward.py:20-40, 241-272): for every backward sample
// x1_j at time t+1 draw an index i ~ Cat(lw_i - transition_potential(x0_i -> x1_j)) over the n_pf filter particles at
// time t (random.categorical = Gumbel-max) and take x0_i.  An n_samples x n_pf contraction with a categorical-sample
// epilogue; the transition is Gaussian (linear_gaussian.py:73-84, nonlinear_gaussian.py:98-105), so after whitening
// (a_i = mean(x0_i) L_Q^-1, b_j = x1_j L_Q^-1: the reference's row-vector convention, see bs_whiten_lg) the potential is
// |a_i - b_j|^2 / 2 + const and the constant cancels.
//
//   bs_means_kernel    a_i for every filter particle: F x0 (linear-Gaussian, times L_Q^-1 from the right) or the RK4
//                      flow of Lorenz-96 scaled by 1 / q_std; one thread per particle, row-major (n, d).
//   bs_sample_kernel   one thread per backward sample j, tiles of 128 filter particles staged in shared memory;
//                      s_ij = lw_i - |a_i - b_j|^2 / 2 (fp32) + Gumbel noise from Philox (counter: gid = j, step = time
//                      index, purpose MB_P_BACKWARD, slot i / 4, word i mod 4); running arg-max (first index wins ties).
// oracle/backward.py restates it in NumPy; exact fp32-vs-fp64 agreement of the arg-max fails only for near-ties.
//
// Fixed-lag stitching (ssm/online_smoothing.py:21-44 full_stitch, :167-207 fixed_lag_stitching) is the same contraction
// with the roles swapped: for every FIXED trajectory end x0_i draw j ~ Cat(lw1_j - transition_potential(x0_i -> x1_j))
// over the candidate continuations x1_j.  The queries carry the flow (bs_means_kernel on x0), the candidates are only
// whitened (bs_whiten_kernel); bs_sample_kernel reads both pre-whitened (`qw`).  Gumbel stream: purpose MB_P_STITCH.
// mb_transition_potential evaluates the potential of matched pairs (x0_i -> x1_i), normalising constant included
// (utils.py:49-79), for the non-interacting weights of :182-184.
#include "common.cuh"
#include "rng.cuh"

#define MB_P_BACKWARD 4u
#define MB_P_STITCH 5u
#define BS_THREADS 128
#define BS_TILE 128

struct BsArgs {
    mb_ssm ssm; float dt;
    const float* x0; const float* lw0; int64_t n_pf;      // filter particles at time t, row-major (n_pf, d), log-weights
    const float* x1; int64_t n_s;                         // backward samples at time t+1, row-major (n_s, d); NULL: none
    float* a;                                             // workspace (n_pf, d): whitened predicted means
    const float* qw;                                      // stitching: pre-whitened queries (n_s, d); NULL: whiten x1 here
    uint32_t purpose;
    uint64_t seed; uint32_t step;
    int32_t* idx; float* x_out;                           // chosen index and state x0[idx] (n_s, d)
};

template <int D>
__device__ __forceinline__ void bs_l96_rhs(const float (&x)[D], float F, float (&k)[D]) {
#pragma unroll
    for (int r = 0; r < D; ++r) k[r] = (x[(r + 1) % D] - x[(r + D - 2) % D]) * x[(r + D - 1) % D] + (F - x[r]);
}

// Whitening of the linear-Gaussian transition AS THE REFERENCE DOES IT: gaussian_potential evaluates
// 0.5 |(x - mean) @ sqrt_prec|^2 with sqrt_prec = inv(chol(Q)) (utils.py:26-30, reset_covariance utils.py:257-258) -- the
// ROW vector times L^-1, i.e. the precision (L^T L)^-1, which is Q^-1 for diagonal Q only.  Mirrored exactly (the
// likelihood potential of the filter kernels does the same with R): z = v L^-1 by back substitution with L^T.
template <int D>
__device__ __forceinline__ void bs_whiten_lg(const mb_ssm& m, float (&v)[D]) {
#pragma unroll
    for (int j = D - 1; j >= 0; --j) {
        float acc = v[j];
#pragma unroll
        for (int i = j + 1; i < D; ++i) acc = fmaf(-m.LQ[i * MB_MAX_SMALL_DIM + j], v[i], acc);
        v[j] = acc / m.LQ[j * MB_MAX_SMALL_DIM + j];
    }
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
            bs_whiten_lg<D>(m, v);                                     // z = mean L_Q^-1 (the reference's convention)
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
    if (valid && a.qw) {
#pragma unroll
        for (int k = 0; k < D; ++k) b[k] = a.qw[j * D + k];
    } else if (valid && a.x1) {                                        // whiten the backward sample like the means
        float v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = a.x1[j * D + k];
        if (m.kind == MB_SSM_LINEAR_GAUSSIAN) {
            bs_whiten_lg<D>(m, v);
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
        const bool pot = a.x1 || a.qw;
        for (int e = threadIdx.x; e < cnt * D; e += BS_THREADS) sa[e] = pot ? a.a[i0 * D + e] : 0.f;
        for (int e = threadIdx.x; e < cnt; e += BS_THREADS) slw[e] = a.lw0[i0 + e];
        __syncthreads();
        if (!valid) continue;
        for (int q = 0; q < cnt; q += 4) {                             // one Philox call feeds four candidates
            const Philox4 r = philox_raw(a.seed, (uint64_t)j, a.step, a.purpose, (uint32_t)((i0 + q) >> 2));
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (q + c >= cnt) break;
                float quad = 0.f;
                if (pot) {
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
        if (a.x_out) {
#pragma unroll
            for (int k = 0; k < D; ++k) a.x_out[j * D + k] = a.x0[arg * D + k];
        }
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
    a.seed = seed; a.step = step; a.idx = idx; a.x_out = x_out; a.qw = nullptr; a.purpose = MB_P_BACKWARD;
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

// ---- fixed-lag stitching and matched-pair transition potentials ------------------------------------------------------
template <int D>
__device__ __forceinline__ void bs_whiten(const mb_ssm& m, float (&v)[D]) {
    if (m.kind == MB_SSM_LINEAR_GAUSSIAN) {
        bs_whiten_lg<D>(m, v);
    } else {
        const float iq = 1.f / m.q_std;
#pragma unroll
        for (int r = 0; r < D; ++r) v[r] *= iq;
    }
}

// out = L_Q^-1 in (rows, no flow); pot != NULL: pot_i = |mean_w_i - out_i|^2 / 2 + cst instead (matched pairs)
template <int D>
__global__ void __launch_bounds__(BS_THREADS) bs_whiten_kernel(mb_ssm m, const float* __restrict__ in, float* __restrict__ out,
                                                               int64_t n, const float* __restrict__ mean_w,
                                                               float* __restrict__ pot, float cst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = in[i * D + k];
        bs_whiten<D>(m, v);
        if (pot) {
            float quad = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) { const float df = mean_w[i * D + k] - v[k]; quad = fmaf(df, df, quad); }
            pot[i] = fmaf(0.5f, quad, cst);
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) out[i * D + k] = v[k];
        }
    }
}

static float bs_potential_constant(const mb_ssm* ssm) {
    // (d log 2 pi - log det prec) / 2 = d log(2 pi) / 2 + log det L_Q (utils.py:79)
    double c = 0.5 * ssm->dim * 1.8378770664093453;
    if (ssm->kind == MB_SSM_LINEAR_GAUSSIAN) for (int r = 0; r < ssm->dim; ++r) c += log((double)ssm->LQ[r * MB_MAX_SMALL_DIM + r]);
    else c += ssm->dim * log((double)ssm->q_std);
    return (float)c;
}

// x0 (n_s, d): the fixed trajectory ends; x1 (n_c, d), lw1 (n_c): candidate continuations and their log-weights;
// work: (n_s + n_c) * d floats; idx (n_s): the chosen candidate of every fixed end.  All device, row-major.
extern "C" int mb_stitch_sample(mb_ctx* ctx, const mb_ssm* ssm, float dt, const float* x0, int64_t n_s, const float* x1,
                                const float* lw1, int64_t n_c, float* work, uint64_t seed, uint32_t step, int32_t* idx,
                                mb_stream_t stream) {
    MB_REQUIRE(ctx && ssm && x0 && x1 && lw1 && work && idx && n_s > 0 && n_c > 0 && n_c < 0x7fffffffll,
               "mb_stitch_sample: bad arguments");
    float* wq = work;
    float* wc = work + n_s * (int64_t)ssm->dim;
    BsArgs q{};                                                        // flow + whitening of the fixed ends -> wq
    q.ssm = *ssm; q.dt = dt; q.x0 = x0; q.n_pf = n_s; q.a = wq;
    BsArgs a{};
    a.ssm = *ssm; a.dt = dt; a.x0 = x1; a.lw0 = lw1; a.n_pf = n_c; a.x1 = nullptr; a.n_s = n_s; a.a = wc; a.qw = wq;
    a.purpose = MB_P_STITCH; a.seed = seed; a.step = step; a.idx = idx; a.x_out = nullptr;
    cudaStream_t st = mb_s(stream);
    const unsigned g0 = (unsigned)((n_s + BS_THREADS - 1) / BS_THREADS), g1 = (unsigned)((n_c + BS_THREADS - 1) / BS_THREADS);
#define ST_CASE(DD)                                                                                       \
    if (ssm->dim == DD) {                                                                                 \
        bs_means_kernel<DD><<<g0, BS_THREADS, 0, st>>>(q);                                                \
        bs_whiten_kernel<DD><<<g1, BS_THREADS, 0, st>>>(*ssm, x1, wc, n_c, nullptr, nullptr, 0.f);        \
        bs_sample_kernel<DD><<<g0, BS_THREADS, 0, st>>>(a);                                               \
        MB_CHECK_LAUNCH();                                                                                \
        return MB_OK;                                                                                     \
    }
    if (ssm->kind == MB_SSM_LINEAR_GAUSSIAN) { ST_CASE(1) ST_CASE(2) ST_CASE(3) ST_CASE(4) ST_CASE(5) ST_CASE(6) ST_CASE(8) }
    else if (ssm->kind == MB_SSM_LORENZ96) { ST_CASE(8) ST_CASE(16) ST_CASE(40) }
    mb_set_error("mb_stitch_sample: model kind %d with dimension %d is not built", ssm->kind, ssm->dim);
    return MB_ERR_UNSUPPORTED;
}

// pot_i = transition_potential(x0_i -> x1_i) (linear_gaussian.py:73-84, nonlinear_gaussian.py:98-105) of n matched
// pairs; work: n * d floats
extern "C" int mb_transition_potential(mb_ctx* ctx, const mb_ssm* ssm, float dt, const float* x0, const float* x1, int64_t n,
                                       float* work, float* pot, mb_stream_t stream) {
    MB_REQUIRE(ctx && ssm && x0 && x1 && work && pot && n > 0, "mb_transition_potential: bad arguments");
    BsArgs q{};
    q.ssm = *ssm; q.dt = dt; q.x0 = x0; q.n_pf = n; q.a = work;
    const float cst = bs_potential_constant(ssm);
    cudaStream_t st = mb_s(stream);
    const unsigned g0 = (unsigned)((n + BS_THREADS - 1) / BS_THREADS);
#define TP_CASE(DD)                                                                                       \
    if (ssm->dim == DD) {                                                                                 \
        bs_means_kernel<DD><<<g0, BS_THREADS, 0, st>>>(q);                                                \
        bs_whiten_kernel<DD><<<g0, BS_THREADS, 0, st>>>(*ssm, x1, nullptr, n, work, pot, cst);            \
        MB_CHECK_LAUNCH();                                                                                \
        return MB_OK;                                                                                     \
    }
    if (ssm->kind == MB_SSM_LINEAR_GAUSSIAN) { TP_CASE(1) TP_CASE(2) TP_CASE(3) TP_CASE(4) TP_CASE(5) TP_CASE(6) TP_CASE(8) }
    else if (ssm->kind == MB_SSM_LORENZ96) { TP_CASE(8) TP_CASE(16) TP_CASE(40) }
    mb_set_error("mb_transition_potential: model kind %d with dimension %d is not built", ssm->kind, ssm->dim);
    return MB_ERR_UNSUPPORTED;
}
