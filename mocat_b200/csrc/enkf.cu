// enkf.cu -- analysis step of the ensemble Kalman filter (EnsembleKalmanFilter.propose_and_intermediate_weight_vectorised,
// ssm/nonlinear_gaussian.py:325-350) for the device family H = I, R = r^2 I on the ROW-MAJOR (n, d) population of the
// Lorenz-96 engine (csrc/pf_l96.cu).  The forecast  mx = f(x) + q z  is the bootstrap step kernel; here:
//   spread_matrix = (mx - mean(mx))^T / sqrt(n - 1)                          (:339)   -> P = spread spread^T (d x d)
//   prop_kalman_gain = P H^T (H P H^T + R)^-1 = P (P + r^2 I)^-1             (:341-343, utils.py:477-484)
//   y_prop = mx + r z2                                                       (:345-346)
//   x_new = mx + (y - y_prop) K^T,  log-weights zero                         (:348-350)
// Kernels: enkf_cov_kernel (shifted first and second moments of the rows: fp32 products over 32-row tiles, fp64 across
// tiles, per-block partials merged in fixed order), enkf_gain_kernel (one block, fp64 Cholesky of P + r^2 I and the d
// triangular solves), enkf_apply_kernel (one thread per particle, K broadcast from shared memory, observation noise z2
// from the particle's own Philox stream: purpose MB_P_SIM, step t).
#include <math.h>
#include "common.cuh"
#include "rng.cuh"

#define EK_THREADS 256
#define EK_ROWS 32

// partial[block][1 + d + d*d]: count is implicit (n); [0] unused, [1 + k] = sum v_k, [1 + d + i d + j] = sum v_i v_j, v = x - shift
template <int D>
__global__ void __launch_bounds__(EK_THREADS) enkf_cov_kernel(const float* __restrict__ x, int64_t n, double* partials) {
    __shared__ float xs[EK_ROWS][D + 1];
    __shared__ float shift[D];
    constexpr int NE = D * D, PER = (NE + D + EK_THREADS - 1) / EK_THREADS;
    if (threadIdx.x < D) shift[threadIdx.x] = x[threadIdx.x];         // particle 0: conditions the second moments
    double acc[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) acc[q] = 0.0;
    const int64_t ntiles = (n + EK_ROWS - 1) / EK_ROWS;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        for (int e = threadIdx.x; e < EK_ROWS * D; e += EK_THREADS) {
            const int r = e / D, k = e - r * D;
            const int64_t row = tile * EK_ROWS + r;
            xs[r][k] = (row < n) ? x[row * D + k] - shift[k] : 0.f;   // rows beyond n contribute nothing
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int e = threadIdx.x + q * EK_THREADS;               // e < D: first moment of column e; else entry (i, j)
            if (e >= NE + D) continue;
            float s = 0.f;
            if (e < D) {
#pragma unroll 8
                for (int r = 0; r < EK_ROWS; ++r) s += xs[r][e];
            } else {
                const int i = (e - D) / D, j = (e - D) - i * D;
#pragma unroll 8
                for (int r = 0; r < EK_ROWS; ++r) s = fmaf(xs[r][i], xs[r][j], s);
            }
            acc[q] += (double)s;
        }
    }
    double* mine = partials + (size_t)blockIdx.x * (1 + D + NE);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int e = threadIdx.x + q * EK_THREADS;
        if (e < NE + D) mine[1 + e] = acc[q];
    }
}

// mean[d], cov[d][d] (unbiased, / (n - 1)) from the block partials (fixed order), then
// K = P (P + r^2 I)^-1 by Cholesky (P + r^2 I = L L^T) and two triangular solves per column (K is symmetric: P commutes
// with any function of itself, so solving (P + r^2 I) K = P column by column gives it)
__global__ void __launch_bounds__(1024) enkf_gain_kernel(const double* partials, int nblocks, int64_t n, int d,
                                                         const float* __restrict__ x, float r_std, double* mean,
                                                         double* cov, float* gain) {
    extern __shared__ double sm[];                    // A[d][d] | P[d][d] | m[d]
    double* A = sm; double* P = sm + d * d; double* m = P + d * d;
    const int ne = d * d, stride = 1 + d + ne;
    for (int e = threadIdx.x; e < d + ne; e += blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * stride + 1 + e];
        if (e < d) m[e] = s / (double)n; else P[e - d] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
        const int i = e / d, j = e - i * d;
        const double c = (P[e] - (double)n * m[i] * m[j]) / (double)(n - 1);
        P[e] = c;
        A[e] = c + (i == j ? (double)r_std * (double)r_std : 0.0);
        if (cov) cov[e] = c;
    }
    if (mean) for (int k = threadIdx.x; k < d; k += blockDim.x) mean[k] = m[k] + (double)x[k];
    __syncthreads();
    // in-place Cholesky of A (lower triangle), right-looking
    for (int k = 0; k < d; ++k) {
        if (threadIdx.x == 0) A[k * d + k] = sqrt(A[k * d + k]);
        __syncthreads();
        for (int i = k + 1 + threadIdx.x; i < d; i += blockDim.x) A[i * d + k] /= A[k * d + k];
        __syncthreads();
        for (int e = threadIdx.x; e < (d - k - 1) * (d - k - 1); e += blockDim.x) {
            const int i = k + 1 + e / (d - k - 1), j = k + 1 + e % (d - k - 1);
            if (j <= i) A[i * d + j] -= A[i * d + k] * A[j * d + k];
        }
        __syncthreads();
    }
    // column c of K: L L^T k_c = p_c  (thread c; P is overwritten column by column)
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        for (int i = 0; i < d; ++i) {                 // forward
            double s = P[i * d + c];
            for (int j = 0; j < i; ++j) s -= A[i * d + j] * P[j * d + c];
            P[i * d + c] = s / A[i * d + i];
        }
        for (int i = d - 1; i >= 0; --i) {            // backward
            double s = P[i * d + c];
            for (int j = i + 1; j < d; ++j) s -= A[j * d + i] * P[j * d + c];
            P[i * d + c] = s / A[i * d + i];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ne; e += blockDim.x) gain[e] = (float)P[e];
}

struct EnkfApplyArgs {
    float* x; int64_t n; const float* gain; const float* y; float r_std;
    uint64_t seed; uint32_t t; int64_t gid0, n_total;
    float* lw; mb_control* ctl; mb_hist* hist;
};

template <int D>
__global__ void __launch_bounds__(128) enkf_apply_kernel(EnkfApplyArgs a) {
    __shared__ __align__(16) float K[D * D];
    __shared__ float ys[D];
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) K[e] = a.gain[e];
    if (threadIdx.x < D) ys[threadIdx.x] = a.y[threadIdx.x];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) {
        float* row = a.x + i * D;
        float xr[D], inn[D];
#pragma unroll
        for (int k = 0; k < D; k += 4) {
            const float4 v = *reinterpret_cast<const float4*>(row + k);
            xr[k] = v.x; xr[k + 1] = v.y; xr[k + 2] = v.z; xr[k + 3] = v.w;
        }
        philox_normals<D>(inn, a.seed, (uint64_t)(a.gid0 + i), a.t, MB_P_SIM, 0u);
#pragma unroll
        for (int k = 0; k < D; ++k) inn[k] = ys[k] - xr[k] - a.r_std * inn[k];     // y - y_prop
#pragma unroll 4
        for (int r = 0; r < D; ++r) {
            float s = row[r];                                          // still the forecast value: row r is written below
#pragma unroll
            for (int c = 0; c < D; ++c) s = fmaf(K[r * D + c], inn[c], s);
            row[r] = s;
        }
        a.lw[i] = 0.f;                                                 // log-weights zero (:350)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {                         // equal weights: ess = n, evidence untouched
        mb_control c = *a.ctl;
        const double nd = (double)a.n_total;
        c.wmax = 0.0; c.s1 = nd; c.s2 = nd; c.lse = log(nd); c.lse2 = log(nd); c.log_ess = log(nd); c.ess = nd;
        c.log_z = 0.0; c.resample = 0; c.resampled = 0; c.iter = (int32_t)a.t; c.done = 0;
        *a.ctl = c;
        if (a.hist && a.t < MB_HIST_MAX) {
            mb_hist h;
            h.beta = 0.0; h.ess = nd; h.log_z = 0.0; h.alpha_mean = 0.0; h.lse = c.lse; h.resampled = 0; h.search_iters = 0;
            a.hist[a.t] = h;
        }
    }
}

extern "C" int mb_rows_mean_cov(mb_ctx* ctx, const float* x, int64_t n, int d, float r_std, double* mean, double* cov,
                                float* gain, mb_stream_t stream) {
    MB_REQUIRE(ctx && x && gain && n > 1 && (d == 8 || d == 16 || d == 40), "mb_rows_mean_cov: bad arguments (d in {8, 16, 40}, n > 1)");
    cudaStream_t st = mb_s(stream);
    const int64_t ntiles = (n + EK_ROWS - 1) / EK_ROWS;
    int64_t grid = ntiles < (int64_t)ctx->sms * 4 ? ntiles : (int64_t)ctx->sms * 4;
    const size_t stride = 1 + (size_t)d + (size_t)d * d;
    if (mb_ensure_scratch(ctx, (size_t)grid * stride * sizeof(double)) != MB_OK) return MB_ERR_CUDA;
    double* partials = (double*)ctx->scratch;
    if (d == 8) enkf_cov_kernel<8><<<(unsigned)grid, EK_THREADS, 0, st>>>(x, n, partials);
    else if (d == 16) enkf_cov_kernel<16><<<(unsigned)grid, EK_THREADS, 0, st>>>(x, n, partials);
    else enkf_cov_kernel<40><<<(unsigned)grid, EK_THREADS, 0, st>>>(x, n, partials);
    MB_CHECK_LAUNCH();
    const size_t smem = (size_t)(2 * d * d + d) * sizeof(double);
    enkf_gain_kernel<<<1, 1024, smem, st>>>(partials, (int)grid, n, d, x, r_std, mean, cov, gain);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

extern "C" int mb_enkf_analysis(mb_ctx* ctx, const mb_ssm* ssm, float* x_rows, int64_t n, const float* y, float* lw,
                                uint64_t seed, uint32_t t, int64_t gid0, mb_control* ctl, mb_hist* hist, float* gain,
                                double* mean, double* cov, mb_stream_t stream) {
    MB_REQUIRE(ctx && ssm && x_rows && y && lw && ctl && gain && n > 1, "mb_enkf_analysis: bad arguments");
    MB_REQUIRE(ssm->kind == MB_SSM_LORENZ96 && ssm->dim_obs == ssm->dim, "mb_enkf_analysis: Lorenz-96 with H = I only");
    const int d = ssm->dim;
    int rc = mb_rows_mean_cov(ctx, x_rows, n, d, ssm->r_std, mean, cov, gain, stream);
    if (rc != MB_OK) return rc;
    cudaStream_t st = mb_s(stream);
    EnkfApplyArgs a{x_rows, n, gain, y, ssm->r_std, seed, t, gid0, n, lw, ctl, hist};
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (d == 8) enkf_apply_kernel<8><<<grid, 128, 0, st>>>(a);
    else if (d == 16) enkf_apply_kernel<16><<<grid, 128, 0, st>>>(a);
    else enkf_apply_kernel<40><<<grid, 128, 0, st>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}
