// teki.cu -- tempered ensemble Kalman inversion (TemperedEKI / AdaptiveTemperedEKI, transport/teki.py:38-185) for the
// device simulator family (g-and-k, abc/scenarios/gk.py:68-96: d_x = 4 unconstrained parameters, d_y = M sorted draws).
//
// One update (teki.py:117-150) on the ROW-MAJOR ensemble x (n, 4), simulated_data (n, M):
//   teki_moments_kernel   shifted first / second moments of [x, simulated_data, constrain(x)]: fp32 products over 32-row
//                         tiles, fp64 across tiles, per-block partials merged in fixed order (calculate_covariances :20-35)
//   teki_solve_kernel     one block, fp64: termination_criterion (:104-111) on the CURRENT ensemble, then
//                         cov_y_given_x = cov_y - cov_xy^T (cov_x + nugget I)^-1 cov_xy, its Cholesky factor and precision
//                         (:125-127); the next temperature of a schedule / the default geometric rule (:88-92), or the
//                         set-up of the adaptive search
//   teki_ppot_kernel      adaptive only: pseudo_likelihood_potential_i = diff_i^T prec diff_i / 2 (:174-175); the root of
//                         log_ess(-(x - temperature) potential) = log(n ess_threshold) (:176-183) is found by the
//                         tempering search of the SMC sampler (mb_temper_adapt: the same `bisect`, utils.py:205-237)
//   teki_gain_kernel      alph = 1 / (new - prev), cov_alph = cov_y + (alph - 1) cov_y_given_x,
//                         kalman_gain = cov_xy (cov_alph + nugget I)^-1 (:129-135)
//   teki_update_kernel    one thread per particle: perturbs = sqrt(alph - 1) z chol^T with z from the particle's Philox
//                         stream (purpose MB_P_MOVE, step = iter), NaN -> 0 (:137-139); value += (data - simulated_data +
//                         perturbs) kalman_gain^T (:141-142); simulated_data = likelihood_sample(value) (:144-145;
//                         purpose MB_P_SIM, step = iter)
// Every kernel is predicated on state->done, so the host enqueues updates without looking and polls the state record.
#include <math.h>
#include "common.cuh"
#include "rng.cuh"
#include "gk.cuh"

#define TK_THREADS 256
#define TK_ROWS 32
#define TK_MAX_BLOCKS 296
#define TK_S MB_TEKI_MAX_DY                     // row stride of every d_y-sized matrix in mb_teki

__device__ __forceinline__ float tk_constrain(const mb_gk& g, float v) {
    return fmaf(normcdff(v), g.prior_max - g.prior_min, g.prior_min);
}

// entries of one partial record: [0, Z) sums of the Z = 8 + M shifted columns, then D * D products of the first
// D = 4 + M columns, then the 4 squares of the constrained columns
template <int M> struct TkLayout {
    static constexpr int D = 4 + M, Z = 8 + M, NE = Z + D * D + 4;
};

template <int M>
__global__ void __launch_bounds__(TK_THREADS) teki_moments_kernel(mb_gk g, const float* __restrict__ x,
                                                                  const float* __restrict__ sim, int64_t n,
                                                                  double* partials, const mb_teki* state) {
    using L = TkLayout<M>;
    if (state->done) return;
    __shared__ float zs[TK_ROWS][L::Z + 1];
    __shared__ float shift[L::Z];
    constexpr int PER = (L::NE + TK_THREADS - 1) / TK_THREADS;
    if (threadIdx.x < L::Z) {                                         // particle 0 conditions the second moments
        const int k = threadIdx.x;
        shift[k] = k < 4 ? x[k] : (k < L::D ? sim[k - 4] : tk_constrain(g, x[k - L::D]));
    }
    double acc[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) acc[q] = 0.0;
    const int64_t ntiles = (n + TK_ROWS - 1) / TK_ROWS;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        for (int e = threadIdx.x; e < TK_ROWS * L::Z; e += TK_THREADS) {
            const int r = e / L::Z, k = e - r * L::Z;
            const int64_t row = tile * TK_ROWS + r;
            float v = 0.f;
            if (row < n) {
                v = k < 4 ? x[row * 4 + k] : (k < L::D ? sim[row * M + (k - 4)] : tk_constrain(g, x[row * 4 + (k - L::D)]));
                v -= shift[k];
            }
            zs[r][k] = v;                                             // rows beyond n contribute nothing
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int e = threadIdx.x + q * TK_THREADS;
            if (e >= L::NE) continue;
            float s = 0.f;
            if (e < L::Z) {
#pragma unroll 8
                for (int r = 0; r < TK_ROWS; ++r) s += zs[r][e];
            } else if (e < L::Z + L::D * L::D) {
                const int i = (e - L::Z) / L::D, j = (e - L::Z) - i * L::D;
#pragma unroll 8
                for (int r = 0; r < TK_ROWS; ++r) s = fmaf(zs[r][i], zs[r][j], s);
            } else {
                const int k = L::D + (e - L::Z - L::D * L::D);
#pragma unroll 8
                for (int r = 0; r < TK_ROWS; ++r) s = fmaf(zs[r][k], zs[r][k], s);
            }
            acc[q] += (double)s;
        }
    }
    double* mine = partials + (size_t)blockIdx.x * L::NE;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int e = threadIdx.x + q * TK_THREADS;
        if (e < L::NE) mine[e] = acc[q];
    }
}

// ---- small dense fp64 linear algebra on one thread (matrices of at most 16 x 16, row stride TK_S) ---------------------
__device__ void tk_inverse(const double* a, int m, double* inv) {        // Gauss-Jordan with partial pivoting
    double w[TK_S][2 * TK_S];
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) { w[i][j] = a[i * TK_S + j]; w[i][m + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < m; ++c) {
        int p = c;
        for (int r = c + 1; r < m; ++r) if (fabs(w[r][c]) > fabs(w[p][c])) p = r;
        if (p != c) for (int j = 0; j < 2 * m; ++j) { const double t = w[c][j]; w[c][j] = w[p][j]; w[p][j] = t; }
        const double piv = 1.0 / w[c][c];
        for (int j = 0; j < 2 * m; ++j) w[c][j] *= piv;
        for (int r = 0; r < m; ++r) {
            if (r == c) continue;
            const double f = w[r][c];
            if (f != 0.0) for (int j = 0; j < 2 * m; ++j) w[r][j] -= f * w[c][j];
        }
    }
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) inv[i * TK_S + j] = w[i][m + j];
}

__device__ void tk_cholesky(const double* a, int m, double* l) {         // lower factor; NaN when not positive definite
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) l[i * TK_S + j] = 0.0;
    for (int j = 0; j < m; ++j) {
        double s = a[j * TK_S + j];
        for (int k = 0; k < j; ++k) s -= l[j * TK_S + k] * l[j * TK_S + k];
        const double dj = sqrt(s);                                       // jnp.linalg.cholesky: NaN, no exception
        l[j * TK_S + j] = dj;
        for (int i = j + 1; i < m; ++i) {
            double t = a[i * TK_S + j];
            for (int k = 0; k < j; ++k) t -= l[i * TK_S + k] * l[j * TK_S + k];
            l[i * TK_S + j] = t / dj;
        }
    }
}

struct TkSolveArgs {
    mb_teki_prm prm;
    const double* partials; int blocks; int64_t n;
    mb_teki* state; mb_control* search_ctl; double* temp_hist;
    int init;
};

__device__ void tk_gain(const mb_teki_prm& P, mb_teki& s, int M, double new_temp, double* temp_hist) {
    // teki.py:129-135
    s.prev_temperature = s.temperature;
    s.alph = 1.0 / (new_temp - s.temperature);
    s.temperature = new_temp;
    double ca[TK_S * TK_S], inv[TK_S * TK_S];
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j)
            ca[i * TK_S + j] = s.cov_y[i * TK_S + j] + (s.alph - 1.0) * s.cov_y_given_x[i * TK_S + j] + (i == j ? P.nugget : 0.0);
    tk_inverse(ca, M, inv);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < M; ++j) {
            double t = 0.0;
            for (int k = 0; k < M; ++k) t += s.cov_xy[i * TK_S + k] * inv[k * TK_S + j];
            s.gain[i * TK_S + j] = t;
        }
    if (temp_hist && s.iter >= 0) temp_hist[s.iter] = new_temp;
}

// merge the moment partials, decide termination, condition the covariances; non-adaptive modes also finish the gain
template <int M>
__global__ void teki_solve_kernel(TkSolveArgs a) {
    using L = TkLayout<M>;
    __shared__ double tot[L::NE];
    mb_teki& s = *a.state;
    if (!a.init && s.done) return;
    for (int e = threadIdx.x; e < L::NE; e += blockDim.x) {
        double t = 0.0;
        for (int b = 0; b < a.blocks; ++b) t += a.partials[(size_t)b * L::NE + e];   // fixed order: deterministic
        tot[e] = t;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const mb_teki_prm& P = a.prm;
    const double nd = (double)a.n, inv_n1 = 1.0 / (nd - 1.0);
    // covariance of the shifted columns: (sum v_i v_j - sum v_i sum v_j / n) / (n - 1)  (teki.py:28-33)
    auto cov = [&](int i, int j) { return (tot[L::Z + i * L::D + j] - tot[i] * tot[j] / nd) * inv_n1; };
    double stds_c[4], stds_x[4];
    bool nan_value = false;
    for (int k = 0; k < 4; ++k) {
        stds_x[k] = sqrt(cov(k, k));
        const double sc = tot[L::D + k];
        stds_c[k] = sqrt((tot[L::Z + L::D * L::D + k] - sc * sc / nd) * inv_n1);
        nan_value = nan_value || (tot[k] != tot[k]);
    }
    if (a.init) {                                                     // startup (teki.py:94-101)
        memset(&s, 0, sizeof(mb_teki));
        for (int k = 0; k < 4; ++k) { s.prior_stds[k] = stds_x[k]; s.stds[k] = stds_c[k]; }
        for (int i = 0; i < M; ++i) s.prec[i * TK_S + i] = 1.0;
        s.ess = nd;
        if (a.temp_hist) a.temp_hist[0] = 0.0;
        return;
    }
    for (int k = 0; k < 4; ++k) s.stds[k] = stds_c[k];
    s.value_nan = nan_value ? 1 : 0;
    bool all_small = true;                                            // teki.py:108-109 (strict <)
    for (int k = 0; k < 4; ++k) all_small = all_small && (stds_c[k] < P.term_std * s.prior_stds[k]);
    if (s.temperature >= P.max_temperature || s.iter >= P.max_iter || all_small || nan_value) {   // :104-111
        s.done = 1;
        if (a.search_ctl) a.search_ctl->done = 1;
        return;
    }
    s.iter += 1;                                                      // :120
    s.perturb_nan = 0;
    // the shift moves the means only
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) s.cov_x[i * 4 + j] = cov(i, j);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < M; ++j) s.cov_xy[i * TK_S + j] = cov(i, 4 + j);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) s.cov_y[i * TK_S + j] = cov(4 + i, 4 + j);
    double cx[TK_S * TK_S], cxi[TK_S * TK_S];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) cx[i * TK_S + j] = s.cov_x[i * 4 + j] + (i == j ? P.nugget : 0.0);
    tk_inverse(cx, 4, cxi);
    double cyx[TK_S * TK_S];
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) {                                 // cov_y - cov_xy^T inv cov_xy   (:125)
            double t = 0.0;
            for (int p = 0; p < 4; ++p)
                for (int q = 0; q < 4; ++q) t += s.cov_xy[p * TK_S + i] * cxi[p * TK_S + q] * s.cov_xy[q * TK_S + j];
            s.cov_y_given_x[i * TK_S + j] = s.cov_y[i * TK_S + j] - t;
            cyx[i * TK_S + j] = s.cov_y_given_x[i * TK_S + j] + (i == j ? P.nugget : 0.0);
        }
    tk_cholesky(cyx, M, s.chol);                                      // :126
    tk_inverse(cyx, M, s.prec);                                       // :127
    if (P.mode == 2) {                                                // adaptive: hand the search its control block
        mb_control c;
        memset(&c, 0, sizeof(c));
        c.s1 = nd; c.s2 = nd; c.lse = log(nd); c.lse2 = log(nd); c.log_ess = log(nd); c.ess = nd;
        c.beta = s.temperature;
        *a.search_ctl = c;
        return;
    }
    double new_temp;
    if (P.mode == 0) {                                                // schedule[iter], index clamped as jnp does (:66)
        const int idx = s.iter < P.schedule_len - 1 ? s.iter : P.schedule_len - 1;
        new_temp = P.schedule[idx];
    } else {                                                          // round(2^(iter / 50) - 1, 4)   (:91-92)
        new_temp = rint((exp2((double)s.iter / 50.0) - 1.0) * 1e4) / 1e4;
    }
    tk_gain(P, s, M, new_temp, a.temp_hist);
}

template <int M>
__global__ void __launch_bounds__(TK_THREADS) teki_ppot_kernel(mb_gk g, const float* __restrict__ sim, int64_t n,
                                                               float* __restrict__ ppot, float* __restrict__ lwz,
                                                               const mb_teki* state) {
    if (state->done) return;
    __shared__ float prec[M * M];
    for (int e = threadIdx.x; e < M * M; e += blockDim.x) prec[e] = (float)state->prec[(e / M) * TK_S + (e % M)];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float df[M];
#pragma unroll
        for (int j = 0; j < M; ++j) df[j] = sim[i * M + j] - g.data[j];
        float q = 0.f;
#pragma unroll
        for (int r = 0; r < M; ++r) {
            float t = 0.f;
#pragma unroll
            for (int c = 0; c < M; ++c) t = fmaf(prec[r * M + c], df[c], t);
            q = fmaf(df[r], t, q);
        }
        ppot[i] = 0.5f * q;                                            // teki.py:174-175
        lwz[i] = 0.f;
    }
}

template <int M>
__global__ void teki_gain_kernel(TkSolveArgs a) {
    mb_teki& s = *a.state;
    if (s.done || threadIdx.x != 0) return;
    s.ess = a.search_ctl->ess;
    s.search_iters = a.search_ctl->search_iters;
    tk_gain(a.prm, s, M, a.search_ctl->beta, a.temp_hist);
}

template <int M>
__global__ void __launch_bounds__(TK_THREADS) teki_update_kernel(mb_gk g, float* __restrict__ x, float* __restrict__ sim,
                                                                 int64_t n, uint64_t seed, int64_t gid0, mb_teki* state) {
    if (state->done) return;
    __shared__ float gain[4 * M];
    __shared__ float chol[M * M];
    __shared__ int nan_block;
    for (int e = threadIdx.x; e < 4 * M; e += blockDim.x) gain[e] = (float)state->gain[(e / M) * TK_S + (e % M)];
    for (int e = threadIdx.x; e < M * M; e += blockDim.x) chol[e] = (float)state->chol[(e / M) * TK_S + (e % M)];
    if (threadIdx.x == 0) nan_block = 0;
    __syncthreads();
    const float scale = (float)sqrt(state->alph - 1.0);               // NaN for alph < 1: perturbs -> 0 (:137-139)
    const uint32_t step = (uint32_t)state->iter;
    int nans = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t gid = (uint64_t)(gid0 + i);
        float z[M];
        philox_normals<M>(z, seed, gid, step, MB_P_MOVE, 0u);
        float v[GK_DIM];
        const float4 xv = reinterpret_cast<const float4*>(x)[i];
        v[0] = xv.x; v[1] = xv.y; v[2] = xv.z; v[3] = xv.w;
        float inn[M];
#pragma unroll
        for (int j = 0; j < M; ++j) {
            float p = 0.f;
#pragma unroll
            for (int k = 0; k <= j; ++k) p = fmaf(chol[j * M + k], z[k], p);   // z @ chol^T
            p *= scale;
            if (p != p) { p = 0.f; ++nans; }
            inn[j] = g.data[j] - sim[i * M + j] + p;
        }
#pragma unroll
        for (int r = 0; r < GK_DIM; ++r) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < M; ++j) t = fmaf(gain[r * M + j], inn[j], t);
            v[r] += t;
        }
        reinterpret_cast<float4*>(x)[i] = make_float4(v[0], v[1], v[2], v[3]);
        float y[M];
        gk_simulate<M>(g, v, seed, gid, step, 0u, y);
#pragma unroll
        for (int j = 0; j < M; ++j) sim[i * M + j] = y[j];
    }
    if (nans) atomicAdd(&nan_block, nans);
    __syncthreads();
    if (threadIdx.x == 0 && nan_block)                                // cleared by the solve kernel of this update
        atomicAdd((unsigned long long*)&state->perturb_nan, (unsigned long long)nan_block);
}

template <int M>
__global__ void __launch_bounds__(TK_THREADS) teki_init_kernel(mb_gk g, float* __restrict__ x, float* __restrict__ sim, int64_t n,
                                                               int sample_prior, uint64_t seed, int64_t gid0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t gid = (uint64_t)(gid0 + i);
        float v[GK_DIM];
        if (sample_prior) {                                           // prior_sample N(0, I), gk.py:93-95
            philox_normals<GK_DIM>(v, seed, gid, 0u, MB_P_INIT, 0u);
            reinterpret_cast<float4*>(x)[i] = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            const float4 xv = reinterpret_cast<const float4*>(x)[i];
            v[0] = xv.x; v[1] = xv.y; v[2] = xv.z; v[3] = xv.w;
        }
        float y[M];
        gk_simulate<M>(g, v, seed, gid, 0u, 0u, y);                   // teki.py:94-99
#pragma unroll
        for (int j = 0; j < M; ++j) sim[i * M + j] = y[j];
    }
}

static unsigned tk_grid(int64_t n, int per) {
    const int64_t b = (n + per - 1) / per;
    return (unsigned)(b < 1 ? 1 : (b > TK_MAX_BLOCKS ? TK_MAX_BLOCKS : b));
}

extern "C" int mb_teki_workspace_doubles(int m) { return TK_MAX_BLOCKS * (8 + m + (4 + m) * (4 + m) + 4); }

extern "C" int mb_teki_init(mb_ctx* ctx, const mb_gk* gk, float* x, float* sim, int64_t n, int sample_prior, uint64_t seed,
                            int64_t gid0, double* partials, double* temp_hist, mb_teki* state, mb_stream_t stream) {
    MB_REQUIRE(ctx && gk && x && sim && partials && state && n > 1, "mb_teki_init: bad arguments");
    cudaStream_t st = mb_s(stream);
    TkSolveArgs a{};
    a.partials = partials; a.n = n; a.state = state; a.temp_hist = temp_hist; a.init = 1;
    a.blocks = (int)tk_grid(n, TK_ROWS);
#define TK_INIT(MM)                                                                                         \
    if (gk->m == MM) {                                                                                      \
        teki_init_kernel<MM><<<tk_grid(n, TK_THREADS), TK_THREADS, 0, st>>>(*gk, x, sim, n, sample_prior, seed, gid0); \
        MB_CUDA(cudaMemsetAsync(state, 0, sizeof(mb_teki), st));                                            \
        teki_moments_kernel<MM><<<a.blocks, TK_THREADS, 0, st>>>(*gk, x, sim, n, partials, state);         \
        teki_solve_kernel<MM><<<1, 128, 0, st>>>(a);                                                       \
        MB_CHECK_LAUNCH();                                                                                  \
        return MB_OK;                                                                                       \
    }
    TK_INIT(4) TK_INIT(8) TK_INIT(16)
    mb_set_error("mb_teki_init: %d summary draws are not built (4, 8 or 16)", gk->m);
    return MB_ERR_UNSUPPORTED;
}

extern "C" int mb_teki_update(mb_ctx* ctx, const mb_gk* gk, const mb_teki_prm* prm, float* x, float* sim, int64_t n,
                              uint64_t seed, int64_t gid0, double* partials, float* scratch /*2 * roundup(n, 32) floats*/,
                              mb_control* search_ctl, double* temp_hist, mb_teki* state, mb_stream_t stream) {
    MB_REQUIRE(ctx && gk && prm && x && sim && partials && state && n > 1, "mb_teki_update: bad arguments");
    MB_REQUIRE(prm->mode != 2 || (scratch && search_ctl), "mb_teki_update: the adaptive search needs scratch and search_ctl");
    MB_REQUIRE(prm->mode != 0 || (prm->schedule && prm->schedule_len > 0), "mb_teki_update: empty temperature schedule");
    cudaStream_t st = mb_s(stream);
    const int64_t npad = (n + 31) & ~(int64_t)31;                     // keeps both halves of `scratch` 16-byte aligned
    TkSolveArgs a{};
    a.prm = *prm; a.partials = partials; a.n = n; a.state = state; a.search_ctl = search_ctl; a.temp_hist = temp_hist;
    a.blocks = (int)tk_grid(n, TK_ROWS);
    mb_temper tp{};
    tp.max_temperature = prm->max_temperature; tp.ess_retain = prm->ess_threshold; tp.ess_resample = 0.0; tp.tol = prm->tol;
    tp.max_search_iter = prm->max_search_iter; tp.max_iter = 0x7fffffff; tp.schedule = nullptr; tp.schedule_len = 0;
#define TK_UPDATE(MM)                                                                                       \
    if (gk->m == MM) {                                                                                      \
        teki_moments_kernel<MM><<<a.blocks, TK_THREADS, 0, st>>>(*gk, x, sim, n, partials, state);         \
        teki_solve_kernel<MM><<<1, 128, 0, st>>>(a);                                                       \
        if (prm->mode == 2) {                                                                               \
            teki_ppot_kernel<MM><<<tk_grid(n, TK_THREADS), TK_THREADS, 0, st>>>(*gk, sim, n, scratch, scratch + npad, state); \
            MB_CHECK_LAUNCH();                                                                              \
            const int rc = mb_temper_adapt(ctx, scratch + npad, scratch, n, &tp, 0, 0, n, search_ctl, nullptr, nullptr, stream); \
            if (rc != MB_OK) return rc;                                                                     \
            teki_gain_kernel<MM><<<1, 32, 0, st>>>(a);                                                     \
        }                                                                                                   \
        teki_update_kernel<MM><<<tk_grid(n, TK_THREADS), TK_THREADS, 0, st>>>(*gk, x, sim, n, seed, gid0, state); \
        MB_CHECK_LAUNCH();                                                                                  \
        return MB_OK;                                                                                       \
    }
    TK_UPDATE(4) TK_UPDATE(8) TK_UPDATE(16)
    mb_set_error("mb_teki_update: %d summary draws are not built (4, 8 or 16)", gk->m);
    return MB_ERR_UNSUPPORTED;
}
