// propagate.cu -- K1: per-particle move + potential/gradient evaluation + log-weight increment.
//
//   smc_move   : vmap(forward_proposal) of MetropolisedSMCSampler (transport/smc.py:91-95,337-365) with
//                MALA/HMC (mcmc/standard_mcmc.py:72-153, utils.py:108-146) or random walk (:21-65) and the
//                Metropolis correction (mcmc/metropolis.py:48-70), fused with the ancestor gather.
//   pf_step    : bootstrap filter body (ssm/filtering.py:280-311, 154-170) for the linear-Gaussian model
//                (ssm/linear_gaussian/linear_gaussian.py:86-94,118-128; Lorenz-96: pf_l96.cu), fused with
//                the ancestor gather, the LSE/ESS reduction and the log-evidence / resample bookkeeping.
//
// Layout: SoA, one thread per particle, every column access is a fully coalesced 128 B/warp request;
// the state of a particle lives in registers (D is a template parameter).
#include "common.cuh"
#include "rng.cuh"
#include "comm.cuh"
#include "pf_common.cuh"

const MbCommDev* mb_comm_dev(const mb_comm* c);

#define MV_THREADS 256
#define TWO_PI_F 6.283185307179586f

// =================================================================================================
// potentials of the static targets (core.py:190-194: U = U_prior + beta * U_lik)
template <int D>
__device__ __forceinline__ void prior_eval(const mb_target& t, const float (&x)[D], float& up, float (&gp)[D]) {
    up = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const float r = (x[k] - t.prior_mean) * t.prior_pscale;
        up = fmaf(0.5f * r, r, up);
        gp[k] = r * t.prior_pscale;
    }
}

template <int LIK, int D>
__device__ __forceinline__ void lik_eval(const mb_target& t, const float (&x)[D], float& ul, float (&gl)[D]) {
    if (LIK == MB_LIK_RASTRIGIN) {            // toy_examples.py:146-149
        float acc = t.a * (float)D;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float f = x[k] - rintf(x[k]);            // cos(2 pi x) = cos(2 pi frac), |2 pi frac| <= pi
            float s, c;
            __sincosf(TWO_PI_F * f, &s, &c);
            acc += fmaf(x[k], x[k], -t.a * c);
            gl[k] = fmaf(TWO_PI_F * t.a, s, 2.f * x[k]);
        }
        ul = acc;
    } else if (LIK == MB_LIK_GAUSSIAN) {      // toy_examples.py:40-44: y = (x - mean) S^T
        float y[D];
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) acc = fmaf(t.prec_sqrt[r * MB_MAX_SMALL_DIM + c], x[c] - t.mean[c], acc);
            y[r] = acc;
        }
        ul = 0.f;
#pragma unroll
        for (int r = 0; r < D; ++r) ul = fmaf(0.5f * y[r], y[r], ul);
#pragma unroll
        for (int c = 0; c < D; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < D; ++r) acc = fmaf(y[r], t.prec_sqrt[r * MB_MAX_SMALL_DIM + c], acc);
            gl[c] = acc;
        }
    } else {
        ul = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) gl[k] = 0.f;
    }
}

template <int LIK, int D>
__device__ __forceinline__ void target_eval(const mb_target& t, float beta, const float (&x)[D], float& up, float& ul,
                                            float (&g)[D]) {
    float gp[D], gl[D];
    prior_eval<D>(t, x, up, gp);
    lik_eval<LIK, D>(t, x, ul, gl);
#pragma unroll
    for (int k = 0; k < D; ++k) g[k] = fmaf(beta, gl[k], gp[k]);
}

// =================================================================================================
struct SmcArgs {
    mb_target tgt;
    mb_move mv;
    const float* x_in; float* x_out; int64_t ld; int64_t n;
    const int32_t* anc;
    float* lw; float* up_out; float* lik_out; float* alpha_out;
    uint64_t seed; int64_t gid0;
    mb_control* ctl;
    int sample_prior;
    mb_shard sh; int sharded;
};

// initial population: x ~ prior (transport/sampler.py:24-30), potentials, lw = 0, control block reset
// (transport/smc.py:128-164: temperature 0, log_weight 0, ess n, log_norm_constant LSE(0, b=1/n) = 0).
template <int LIK, int D>
__global__ void __launch_bounds__(MV_THREADS) smc_init_kernel(SmcArgs a, int64_t n_total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        float x[D];
        if (a.sample_prior) {
            float z[D];
            philox_normals<D>(z, a.seed, (uint64_t)(a.gid0 + i), 0u, MB_P_INIT, 0u);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                x[k] = fmaf(a.tgt.prior_std, z[k], a.tgt.prior_mean);
                a.x_out[(int64_t)k * a.ld + i] = x[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) x[k] = a.x_out[(int64_t)k * a.ld + i];
        }
        float up, ul, g[D];
        target_eval<LIK, D>(a.tgt, 0.f, x, up, ul, g);
        if (a.up_out) a.up_out[i] = up;
        a.lik_out[i] = ul;
        a.lw[i] = 0.f;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        mb_control c;
        memset(&c, 0, sizeof(c));
        const double nd = (double)n_total;
        c.wmax = 0.0; c.s1 = nd; c.s2 = nd;
        c.lse = log(nd); c.lse2 = log(nd); c.log_ess = log(nd); c.ess = nd;
        c.alpha_mean = 1.0;
        c.seed = a.seed;
        *a.ctl = c;
    }
}

template <int LIK, int D, int MOVE>
__global__ void __launch_bounds__(MV_THREADS, (D <= 6 ? 4 : (D <= 10 ? 3 : 2))) smc_move_kernel(SmcArgs a) {
    const mb_control* ctl = a.ctl;
    if (ctl->done) return;
    const bool resample = ctl->resample != 0;
    const float beta = (float)ctl->beta;
    const uint32_t step = (uint32_t)(ctl->iter + 1);
    const uint64_t seed = ctl->seed;                   // device-resident key: the captured graph is seed-agnostic
    // stepsize <= 0: the device-resident stepsize of the Robbins-Monro adaptation (mb_rm_adapt keeps it in ctl->aux1)
    const float eps = a.mv.stepsize > 0.f ? a.mv.stepsize : (float)ctl->aux1;
    constexpr uint32_t S = MB_MOVE_SLOTS(D);            // Philox slots per Metropolised move (normals + accept uniform)
    __shared__ double red[MV_THREADS / 32];
    long long nan_local = 0;
    double alpha_local = 0.0;

    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t src = resample ? (int64_t)a.anc[i] : i;                // fused ancestor gather (core.py:46-56)
        const float* xb = a.x_in;
        if (a.sharded && resample) {                                   // ancestor lives on rank `owner`: read it over NVLink
            const int owner = (int)(src / a.sh.n_local);
            src -= (int64_t)owner * a.sh.n_local;
            xb = a.sh.x_peers[owner];
        }
        float x[D];
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = __ldg(xb + (int64_t)k * a.ld + src);
        float up, ul, g[D];
        target_eval<LIK, D>(a.tgt, beta, x, up, ul, g);                // MCMC startup, standard_mcmc.py:94-102
        float U = fmaf(beta, ul, up);
        float alpha_sum = 0.f;
        const uint64_t gid = (uint64_t)(a.gid0 + i);
        for (int s = 0; s < a.mv.mcmc_steps; ++s) {
            float z[D], uacc;
            philox_normals_accept<D>(z, uacc, seed, gid, step, MB_P_MOVE, (uint32_t)s * S);
            float xp[D], gpn[D], upn, uln, Un, alpha;
            if (MOVE == MB_MOVE_MALA) {
                // always(): p = z (friction = inf, :116-122); leapfrog (utils.py:117-134); p' = -p' (:141)
                float p[D], kin0 = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) { p[k] = z[k]; kin0 = fmaf(0.5f * z[k], z[k], kin0); xp[k] = x[k]; gpn[k] = g[k]; }
                for (int l = 0; l < a.mv.leapfrog_steps; ++l) {
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        p[k] = fmaf(-0.5f * eps, gpn[k], p[k]);        // p_half
                        xp[k] = fmaf(eps, p[k], xp[k]);
                    }
                    target_eval<LIK, D>(a.tgt, beta, xp, upn, uln, gpn);
#pragma unroll
                    for (int k = 0; k < D; ++k) p[k] = fmaf(-0.5f * eps, gpn[k], p[k]);
                }
                Un = fmaf(beta, uln, upn);
                float kin1 = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) kin1 = fmaf(0.5f * p[k], p[k], kin1);
                alpha = fminf(1.f, __expf(-Un + U - kin1 + kin0));     // standard_mcmc.py:145-153
            } else {
                const float sq = sqrtf(eps);                           // standard_mcmc.py:57
#pragma unroll
                for (int k = 0; k < D; ++k) xp[k] = fmaf(sq, z[k], x[k]);
                target_eval<LIK, D>(a.tgt, beta, xp, upn, uln, gpn);
                Un = fmaf(beta, uln, upn);
                alpha = fminf(1.f, __expf(-Un + U));                   // :61-65
            }
            if (alpha != alpha) alpha = 0.f;                           // metropolis.py:58
            if (uacc < alpha) {                                        // :61-66
#pragma unroll
                for (int k = 0; k < D; ++k) { x[k] = xp[k]; g[k] = gpn[k]; }
                up = upn; ul = uln; U = Un;
            }
            alpha_sum += alpha;
        }
        const float alpha_mean = alpha_sum / (float)a.mv.mcmc_steps;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            a.x_out[(int64_t)k * a.ld + i] = x[k];
            nan_local += (x[k] != x[k]);
        }
        if (a.up_out) a.up_out[i] = up;
        a.lik_out[i] = ul;
        if (a.alpha_out) a.alpha_out[i] = alpha_mean;
        if (resample) a.lw[i] = 0.f;                                   // smc.py:69
        alpha_local += (double)alpha_mean;
    }
    // deterministic (integer) accumulation of NaN count and mean acceptance.  One particle per thread, so this tail runs
    // once per particle: fixed point (24 fractional bits) + the hardware integer warp reduction (REDUX) instead of two
    // fp64 shuffle trees (~200 -> ~15 instructions per particle).  A thread that looped over several particles
    // (grids beyond 2^31 blocks) is still exact up to 255 particles per thread.
    const unsigned long long afx = (unsigned long long)llrint(alpha_local * 16777216.0);
    const unsigned wa_lo = __reduce_add_sync(MB_FULL, (unsigned)(afx & 0xffffffu));        // 32 x 2^24 < 2^32
    const unsigned wa_hi = __reduce_add_sync(MB_FULL, (unsigned)(afx >> 24));
    const unsigned wn = __reduce_add_sync(MB_FULL, (unsigned)nan_local);
    unsigned long long* redu = reinterpret_cast<unsigned long long*>(red);
    __shared__ unsigned redn[MV_THREADS / 32];
    if ((threadIdx.x & 31) == 0) {
        redu[threadIdx.x >> 5] = (unsigned long long)wa_lo + ((unsigned long long)wa_hi << 24);
        redn[threadIdx.x >> 5] = wn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long asum = 0, nsum = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { asum += redu[w]; nsum += redn[w]; }
        atomicAdd((unsigned long long*)&a.ctl->alpha_fx, asum << 8);                      // units of 2^-32
        if (nsum > 0) atomicAdd((unsigned long long*)&a.ctl->nan_count, nsum);
        if (blockIdx.x == 0) a.ctl->resampled = resample ? 1 : 0;
    }
}

// dispatch over (likelihood kind, dimension, move)
#define SMC_DIMS(X) X(1) X(2) X(3) X(4) X(5) X(6) X(8) X(10) X(16)

template <int LIK, int D>
static int smc_launch(const SmcArgs& a, int init, int64_t n_total, int grid, cudaStream_t st) {
    if (init) { smc_init_kernel<LIK, D><<<grid, MV_THREADS, 0, st>>>(a, n_total); }
    else if (a.mv.kind == MB_MOVE_MALA) { smc_move_kernel<LIK, D, MB_MOVE_MALA><<<grid, MV_THREADS, 0, st>>>(a); }
    else if (a.mv.kind == MB_MOVE_RW) { smc_move_kernel<LIK, D, MB_MOVE_RW><<<grid, MV_THREADS, 0, st>>>(a); }
    else { mb_set_error("smc: unknown move kind %d", a.mv.kind); return MB_ERR_ARG; }
    MB_CHECK_LAUNCH();
    return MB_OK;
}

static int smc_dispatch(mb_ctx* ctx, const SmcArgs& a, int init, int64_t n_total, cudaStream_t st) {
    int64_t grid = (a.n + MV_THREADS - 1) / MV_THREADS;      // one particle per thread: no grid-stride tail
    if (grid > 0x7fffffffll) grid = 0x7fffffffll;
    const int d = a.tgt.dim;
#define CASE_R(DD) if (d == DD) return smc_launch<MB_LIK_RASTRIGIN, DD>(a, init, n_total, (int)grid, st);
#define CASE_G(DD) if (d == DD && DD <= MB_MAX_SMALL_DIM) return smc_launch<MB_LIK_GAUSSIAN, (DD <= MB_MAX_SMALL_DIM ? DD : 1)>(a, init, n_total, (int)grid, st);
#define CASE_N(DD) if (d == DD) return smc_launch<MB_LIK_NONE, DD>(a, init, n_total, (int)grid, st);
    if (a.tgt.kind == MB_LIK_RASTRIGIN) { SMC_DIMS(CASE_R) }
    else if (a.tgt.kind == MB_LIK_GAUSSIAN) { SMC_DIMS(CASE_G) }
    else if (a.tgt.kind == MB_LIK_NONE) { SMC_DIMS(CASE_N) }
    mb_set_error("smc: unsupported target kind %d / dim %d (built-in device scenarios only; no CPU fallback)",
                 a.tgt.kind, d);
    return MB_ERR_UNSUPPORTED;
}

extern "C" int mb_smc_init(mb_ctx* ctx, const mb_target* tgt, float* x, int64_t ld, int64_t n, int64_t n_total,
                           int sample_prior, float* up, float* lik, float* lw, uint64_t seed, int64_t gid0,
                           mb_control* ctl, mb_stream_t stream) {
    MB_REQUIRE(ctx && tgt && x && lik && lw && ctl && n > 0 && ld >= n, "mb_smc_init: bad arguments");
    SmcArgs a{};
    a.tgt = *tgt; a.x_in = x; a.x_out = x; a.ld = ld; a.n = n; a.lw = lw; a.up_out = up; a.lik_out = lik;
    a.seed = seed; a.gid0 = gid0; a.ctl = ctl; a.sample_prior = sample_prior;
    return smc_dispatch(ctx, a, 1, n_total, mb_s(stream));
}

extern "C" int mb_smc_move(mb_ctx* ctx, const mb_target* tgt, const mb_move* mv, const float* x_in, float* x_out,
                           int64_t ld, int64_t n, const int32_t* anc, float* lw, float* up_out, float* lik_out,
                           float* alpha_out, uint64_t seed, int64_t gid0, mb_control* ctl, const mb_shard* sh,
                           mb_stream_t stream) {
    MB_REQUIRE(ctx && tgt && mv && x_in && x_out && anc && lw && lik_out && ctl && n > 0 && ld >= n && x_in != x_out,
               "mb_smc_move: bad arguments");
    MB_REQUIRE(mv->mcmc_steps >= 1 && mv->leapfrog_steps >= 1, "mb_smc_move: mcmc_steps/leapfrog_steps must be >= 1");
    SmcArgs a{};
    a.tgt = *tgt; a.mv = *mv; a.x_in = x_in; a.x_out = x_out; a.ld = ld; a.n = n; a.anc = anc; a.lw = lw;
    a.up_out = up_out; a.lik_out = lik_out; a.alpha_out = alpha_out; a.seed = seed; a.gid0 = gid0; a.ctl = ctl;
    if (sh) { a.sh = *sh; a.sharded = sh->world > 1; }
    return smc_dispatch(ctx, a, 0, 0, mb_s(stream));
}

// potential and gradient of a static target at row-major points (SVGD: vmap(potential_and_grad),
// transport/svgd.py:99,141-142).  One thread per particle.
template <int LIK, int D>
__global__ void __launch_bounds__(MV_THREADS)
target_pg_kernel(mb_target tgt, float beta, const float* __restrict__ X, int n, float* U, float* G) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[D], g[D], up, ul;
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = X[(int64_t)i * D + k];
    target_eval<LIK, D>(tgt, beta, x, up, ul, g);
    if (U) U[i] = fmaf(beta, ul, up);
#pragma unroll
    for (int k = 0; k < D; ++k) G[(int64_t)i * D + k] = g[k];
}

extern "C" int mb_target_potential_grad(mb_ctx* ctx, const mb_target* tgt, double beta, const float* X, int n,
                                        float* U, float* G, mb_stream_t stream) {
    MB_REQUIRE(ctx && tgt && X && G && n > 0, "mb_target_potential_grad: bad arguments");
    const int grid = (n + MV_THREADS - 1) / MV_THREADS;
    const int d = tgt->dim;
    cudaStream_t st = mb_s(stream);
#define PG_R(DD) if (tgt->kind == MB_LIK_RASTRIGIN && d == DD) { target_pg_kernel<MB_LIK_RASTRIGIN, DD><<<grid, MV_THREADS, 0, st>>>(*tgt, (float)beta, X, n, U, G); MB_CHECK_LAUNCH(); return MB_OK; }
#define PG_G(DD) if (tgt->kind == MB_LIK_GAUSSIAN && d == DD && DD <= MB_MAX_SMALL_DIM) { target_pg_kernel<MB_LIK_GAUSSIAN, (DD <= MB_MAX_SMALL_DIM ? DD : 1)><<<grid, MV_THREADS, 0, st>>>(*tgt, (float)beta, X, n, U, G); MB_CHECK_LAUNCH(); return MB_OK; }
    SMC_DIMS(PG_R)
    SMC_DIMS(PG_G)
    mb_set_error("mb_target_potential_grad: unsupported target kind %d / dim %d", tgt->kind, d);
    return MB_ERR_UNSUPPORTED;
}

// =================================================================================================
// bootstrap particle filter
struct PfArgs {
    mb_ssm ssm;
    const float* x_in; float* x_out; int64_t ld; int64_t n; int64_t n_total;
    const int32_t* anc;
    const float* y;            // device, dim_obs floats for this time step
    float* lw;
    uint64_t seed; uint32_t t; int64_t gid0;
    double ess_threshold;
    mb_control* ctl; mb_hist* hist;
    double* partials; uint32_t* counter;
    int init;
    mb_shard sh; int sharded;
    MbCommDev comm; int has_comm;
};

// normals of one particle, produced four at a time where they are consumed (keeps them out of the register budget
// of the Lorenz-96 flow): slot s holds normals 4s..4s+3 (rng.cuh)
struct PfNormals {
    uint64_t seed, gid; uint32_t step, purpose;
    __device__ __forceinline__ void slot(int s, float (&z4)[4]) const {
        const Philox4 r = philox_raw(seed, gid, step, purpose, (uint32_t)s);
        box_muller(r.x, r.y, z4[0], z4[1]);
        box_muller(r.z, r.w, z4[2], z4[3]);
    }
};

template <int KIND, int D>
__device__ __forceinline__ float pf_particle(const mb_ssm& m, float (&x)[D], const PfNormals& rng, const float* ys,
                                             bool init) {
    // returns the log-weight increment -likelihood_potential(x', y)
    if (KIND == MB_SSM_LINEAR_GAUSSIAN) {
        float z[D];
#pragma unroll
        for (int s = 0; s < (D + 3) / 4; ++s) {
            float z4[4];
            rng.slot(s, z4);
#pragma unroll
            for (int c = 0; c < 4; ++c) if (4 * s + c < D) z[4 * s + c] = z4[c];
        }
        float xn[D];
        if (init) {                                     // linear_gaussian.py:45-50: L0 z + m0
#pragma unroll
            for (int r = 0; r < D; ++r) {
                float acc = m.m0[r];
#pragma unroll
                for (int c = 0; c < D; ++c) acc = fmaf(m.L0[r * MB_MAX_SMALL_DIM + c], z[c], acc);
                xn[r] = acc;
            }
        } else {                                        // :86-94: F x + LQ z
#pragma unroll
            for (int r = 0; r < D; ++r) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < D; ++c) acc = fmaf(m.F[r * MB_MAX_SMALL_DIM + c], x[c], acc);
#pragma unroll
                for (int c = 0; c < D; ++c) acc = fmaf(m.LQ[r * MB_MAX_SMALL_DIM + c], z[c], acc);
                xn[r] = acc;
            }
        }
#pragma unroll
        for (int r = 0; r < D; ++r) x[r] = xn[r];
        float diff[D];                                  // :118-128 with utils.py:26-30 (diff @ sqrt_prec)
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float acc = ys[r];
#pragma unroll
            for (int c = 0; c < D; ++c) acc = fmaf(-m.H[r * MB_MAX_SMALL_DIM + c], x[c], acc);
            diff[r] = acc;
        }
        float quad = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < D; ++r) acc = fmaf(diff[r], m.Rps[r * MB_MAX_SMALL_DIM + c], acc);
            quad = fmaf(0.5f * acc, acc, quad);
        }
        return -(quad + m.lik_const);
    } else {
        return 0.f;                                     // Lorenz-96 lives in pf_l96.cu (row-major layout, lane-split kernel)
    }
}

#define PF_THREADS(D) MV_THREADS
template <int KIND, int D, bool INIT>
__global__ void __launch_bounds__(PF_THREADS(D)) pf_step_kernel(PfArgs a) {
    mb_control* ctl = a.ctl;
    constexpr bool init = INIT;                       // compile-time: the initial-sample variant is a separate (small) kernel
    if (!init && ctl->done) return;
    const bool resample = !init && ctl->resample != 0;
    __shared__ Lse3 smem[MV_THREADS / 32];
    __shared__ float ys[D];
    if (threadIdx.x < D) ys[threadIdx.x] = (threadIdx.x < a.ssm.dim_obs) ? a.y[threadIdx.x] : 0.f;
    __syncthreads();

    // per-thread online LSE accumulator (same association rules as reduce.cu)
    float am = -INFINITY;
    double as1 = 0.0, as2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t src = resample ? (int64_t)a.anc[i] : i;
        const float* xb = a.x_in;
        if (a.sharded && resample) {                                   // ancestor on another rank: NVLink peer read
            const int owner = (int)(src / a.sh.n_local);
            src -= (int64_t)owner * a.sh.n_local;
            xb = a.sh.x_peers[owner];
        }
        float x[D];
        if (!init) {
#pragma unroll
            for (int k = 0; k < D; ++k) x[k] = __ldg(xb + (int64_t)k * a.ld + src);
        }
        const PfNormals rng{a.seed, (uint64_t)(a.gid0 + i), a.t, init ? MB_P_INIT : MB_P_MOVE};
        const float incr = pf_particle<KIND, D>(a.ssm, x, rng, ys, init);
#pragma unroll
        for (int k = 0; k < D; ++k) a.x_out[(int64_t)k * a.ld + i] = x[k];
        const float w = ((init || resample) ? 0.f : a.lw[i]) + incr;           // filtering.py:292,303
        a.lw[i] = w;
        // online (max, sum, sumsq)
        if (w > am) {
            if (am != -INFINITY) { const double f = exp((double)am - (double)w); as1 *= f; as2 *= f * f; }
            am = w;
        }
        if (am != -INFINITY || w != w) {
            const float e = __expf(w - ((am == -INFINITY) ? 0.f : am));
            as1 += (double)e; as2 += (double)e * (double)e;
        }
    }
    PfTail tl{};
    tl.n_total = a.n_total; tl.t = a.t; tl.ess_threshold = a.ess_threshold; tl.seed = a.seed; tl.ctl = a.ctl; tl.hist = a.hist;
    tl.partials = a.partials; tl.counter = a.counter; tl.comm = a.comm; tl.has_comm = a.has_comm;
    pf_finish<INIT>(tl, Lse3{(double)am, as1, as2}, resample, smem);
}

static int64_t pf_grid(mb_ctx* ctx, int64_t n, int threads) {
    int64_t grid = (n + threads - 1) / threads;
    int64_t cap = (int64_t)ctx->sms * (threads == 128 ? 16 : 8);
    if (cap > MB_MAX_PARTIAL_BLOCKS) cap = MB_MAX_PARTIAL_BLOCKS;
    return grid > cap ? cap : grid;
}

#define PF_LAUNCH(K, DD) (a.init ? pf_step_kernel<K, DD, true> : pf_step_kernel<K, DD, false>)
static int pf_dispatch(mb_ctx* ctx, PfArgs& a, cudaStream_t st) {
    a.partials = ctx->partials;
    a.counter = ctx->counters + MB_CNT_MOVE;
    const int d = a.ssm.dim;
#define PF_LG(DD) if (a.ssm.kind == MB_SSM_LINEAR_GAUSSIAN && d == DD) { PF_LAUNCH(MB_SSM_LINEAR_GAUSSIAN, DD)<<<(unsigned)pf_grid(ctx, a.n, PF_THREADS(DD)), PF_THREADS(DD), 0, st>>>(a); MB_CHECK_LAUNCH(); return MB_OK; }
    PF_LG(1) PF_LG(2) PF_LG(3) PF_LG(4) PF_LG(5) PF_LG(6) PF_LG(8)
    if (a.ssm.kind == MB_SSM_LORENZ96) { mb_set_error("pf: Lorenz-96 runs through mb_pf_l96_init / mb_pf_l96_step (row-major layout)"); return MB_ERR_UNSUPPORTED; }
    mb_set_error("pf: unsupported ssm kind %d / dim %d (built-in device models only; no CPU fallback)", a.ssm.kind, d);
    return MB_ERR_UNSUPPORTED;
}

extern "C" int mb_pf_init(mb_ctx* ctx, const mb_ssm* ssm, float* x, int64_t ld, int64_t n, int64_t n_total,
                          const float* y0, float* lw, uint64_t seed, int64_t gid0, double ess_threshold,
                          mb_control* ctl, mb_hist* hist, mb_comm* comm, mb_stream_t stream) {
    MB_REQUIRE(ctx && ssm && x && y0 && lw && ctl && n > 0 && ld >= n, "mb_pf_init: bad arguments");
    MB_REQUIRE(ssm->dim_obs <= ssm->dim, "mb_pf_init: dim_obs must be <= dim");
    PfArgs a{};
    a.ssm = *ssm; a.x_in = x; a.x_out = x; a.ld = ld; a.n = n; a.n_total = n_total; a.y = y0; a.lw = lw;
    a.seed = seed; a.t = 0; a.gid0 = gid0; a.ess_threshold = ess_threshold; a.ctl = ctl; a.hist = hist; a.init = 1;
    if (comm) { a.comm = *mb_comm_dev(comm); a.has_comm = a.comm.world > 1; }
    return pf_dispatch(ctx, a, mb_s(stream));
}

extern "C" int mb_pf_step(mb_ctx* ctx, const mb_ssm* ssm, const float* x_in, float* x_out, int64_t ld, int64_t n,
                          int64_t n_total, const int32_t* anc, const float* y, float* lw, uint64_t seed, uint32_t t,
                          int64_t gid0, double ess_threshold, mb_control* ctl, mb_hist* hist, const mb_shard* sh,
                          mb_comm* comm, mb_stream_t stream) {
    MB_REQUIRE(ctx && ssm && x_in && x_out && anc && y && lw && ctl && n > 0 && ld >= n && x_in != x_out,
               "mb_pf_step: bad arguments");
    MB_REQUIRE(ssm->dim_obs <= ssm->dim && ssm->substeps >= 1, "mb_pf_step: bad model");
    PfArgs a{};
    a.ssm = *ssm; a.x_in = x_in; a.x_out = x_out; a.ld = ld; a.n = n; a.n_total = n_total; a.anc = anc; a.y = y;
    a.lw = lw; a.seed = seed; a.t = t; a.gid0 = gid0; a.ess_threshold = ess_threshold; a.ctl = ctl; a.hist = hist;
    a.init = 0;
    if (sh) { a.sh = *sh; a.sharded = sh->world > 1; }
    if (comm) { a.comm = *mb_comm_dev(comm); a.has_comm = a.comm.world > 1; }
    return pf_dispatch(ctx, a, mb_s(stream));
}

// prior sample of any dimension at row-major points (transport/sampler.py:24-30 vmap(prior_sample)); the same
// Philox stream as smc_init_kernel (purpose INIT, step 0): one thread per (particle, 4-normal slot).
__global__ void prior_sample_kernel(float mean, float std, int d, int64_t n, uint64_t seed, int64_t gid0, float* X) {
    const int nz = (d + 3) / 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nz) return;
    const int64_t i = t / nz;
    const int s = (int)(t - i * nz);
    const Philox4 r = philox_raw(seed, (uint64_t)(gid0 + i), 0u, MB_P_INIT, (uint32_t)s);
    float z[4];
    box_muller(r.x, r.y, z[0], z[1]);
    box_muller(r.z, r.w, z[2], z[3]);
    for (int c = 0; c < 4; ++c)
        if (4 * s + c < d) X[i * d + 4 * s + c] = fmaf(std, z[c], mean);
}

extern "C" int mb_prior_sample(mb_ctx* ctx, float prior_mean, float prior_std, int d, int64_t n, uint64_t seed,
                               int64_t gid0, float* X, mb_stream_t stream) {
    MB_REQUIRE(ctx && X && d > 0 && n > 0, "mb_prior_sample: bad arguments");
    const int64_t total = n * ((d + 3) / 4);
    prior_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, mb_s(stream)>>>(prior_mean, prior_std, d, n, seed, gid0, X);
    MB_CHECK_LAUNCH();
    return MB_OK;
}


// =================================================================================================
// Kalman filter of the time-homogeneous linear-Gaussian model (ssm/linear_gaussian/kalman.py:16-57): exact
// filtering means / covariances and the innovation log-likelihood the particle filter's log-evidence is checked
// against.  O(T d^3) sequential algebra on matrices of at most 8 x 8: ONE warp, fp64, lane l owns the matrix entries
// l and l + 32; nothing here is bandwidth- or throughput-relevant, it exists so that the cross-check runs where the
// filter runs.  cov_0 = L0 L0^T (the reference passes the Cholesky factor, kalman.py:20 -- identical when P0 = I).
#define KF_D MB_MAX_SMALL_DIM
struct KfArgs { mb_ssm ssm; const float* y; int T; double* means; double* covs; double* loglik; };

__device__ __forceinline__ void kf_sync() { __syncwarp(); }

// C = A B (or A B^T) for KF_D x KF_D matrices in shared memory, rows/cols beyond (r, c) ignored; lane-parallel over entries
__device__ __forceinline__ void kf_mul(const double* A, const double* B, double* Cm, int r, int k, int c, bool bt) {
    const int lane = threadIdx.x;
    for (int e = lane; e < KF_D * KF_D; e += 32) {
        const int i = e / KF_D, j = e % KF_D;
        double acc = 0.0;
        if (i < r && j < c)
            for (int q = 0; q < k; ++q) acc += A[i * KF_D + q] * (bt ? B[j * KF_D + q] : B[q * KF_D + j]);
        Cm[e] = acc;
    }
    kf_sync();
}

__global__ void __launch_bounds__(32) kalman_kernel(KfArgs a) {
    __shared__ double F[KF_D * KF_D], H[KF_D * KF_D], Q[KF_D * KF_D], R[KF_D * KF_D], P[KF_D * KF_D], T1[KF_D * KF_D],
        T2[KF_D * KF_D], S[KF_D * KF_D], Si[KF_D * KF_D], K[KF_D * KF_D], mu[KF_D], innov[KF_D], tmp[KF_D];
    __shared__ double ll, logdet;
    const int lane = threadIdx.x, d = a.ssm.dim, dy = a.ssm.dim_obs;
    for (int e = lane; e < KF_D * KF_D; e += 32) { F[e] = a.ssm.F[e]; H[e] = a.ssm.H[e]; T1[e] = a.ssm.L0[e]; T2[e] = a.ssm.LQ[e]; K[e] = a.ssm.Rps[e]; }
    if (lane < KF_D) mu[lane] = a.ssm.m0[lane];
    if (lane == 0) ll = 0.0;
    kf_sync();
    kf_mul(T1, T1, P, d, d, d, true);                                   // P0 = L0 L0^T
    kf_mul(T2, T2, Q, d, d, d, true);                                   // Q = LQ LQ^T
    // R = (Rps^T Rps)^-1 with Rps = inv(chol(R)):  R = L L^T, L = Rps^-1 (lower triangular): forward substitution
    for (int e = lane; e < KF_D * KF_D; e += 32) T1[e] = 0.0;
    kf_sync();
    if (lane < dy) {                                                    // column `lane` of L = Rps^-1
        for (int i = 0; i < dy; ++i) {
            double v = (i == lane) ? 1.0 : 0.0;
            for (int q = 0; q < i; ++q) v -= K[i * KF_D + q] * T1[q * KF_D + lane];
            T1[i * KF_D + lane] = v / K[i * KF_D + i];
        }
    }
    kf_sync();
    kf_mul(T1, T1, R, dy, dy, dy, true);
    for (int t = 0; t < a.T; ++t) {
        if (t > 0) {                                                    // predict, kalman.py:34-35
            if (lane < d) { double v = 0.0; for (int q = 0; q < d; ++q) v += F[lane * KF_D + q] * mu[q]; tmp[lane] = v; }
            kf_sync();
            if (lane < d) mu[lane] = tmp[lane];
            kf_mul(F, P, T1, d, d, d, false);
            kf_mul(T1, F, T2, d, d, d, true);
            for (int e = lane; e < KF_D * KF_D; e += 32) P[e] = T2[e] + Q[e];
            kf_sync();
        }
        kf_mul(H, P, T1, dy, d, d, false);                              // H P
        kf_mul(T1, H, S, dy, d, dy, true);                              // S = H P H^T + R
        for (int e = lane; e < KF_D * KF_D; e += 32) S[e] += R[e];
        if (lane < dy) {
            double v = (double)a.y[(int64_t)t * dy + lane];
            for (int q = 0; q < d; ++q) v -= H[lane * KF_D + q] * mu[q];
            innov[lane] = v;
        }
        kf_sync();
        // S^-1 and log det S by Gauss-Jordan without pivoting (S is symmetric positive definite); serial on lane 0
        if (lane == 0) {
            double A[KF_D][2 * KF_D];
            for (int i = 0; i < dy; ++i)
                for (int j = 0; j < dy; ++j) { A[i][j] = S[i * KF_D + j]; A[i][dy + j] = (i == j) ? 1.0 : 0.0; }
            double ld = 0.0;
            for (int c = 0; c < dy; ++c) {
                const double piv = A[c][c];
                ld += log(piv);
                for (int j = 0; j < 2 * dy; ++j) A[c][j] /= piv;
                for (int i = 0; i < dy; ++i) {
                    if (i == c) continue;
                    const double f = A[i][c];
                    for (int j = 0; j < 2 * dy; ++j) A[i][j] -= f * A[c][j];
                }
            }
            for (int i = 0; i < KF_D; ++i)
                for (int j = 0; j < KF_D; ++j) Si[i * KF_D + j] = (i < dy && j < dy) ? A[i][dy + j] : 0.0;
            logdet = ld;
        }
        kf_sync();
        if (lane == 0) {
            double quad = 0.0;
            for (int i = 0; i < dy; ++i) for (int j = 0; j < dy; ++j) quad += innov[i] * Si[i * KF_D + j] * innov[j];
            ll += -0.5 * (quad + logdet + (double)dy * 1.8378770664093453);
        }
        kf_mul(P, H, T1, d, d, dy, true);                               // P H^T
        kf_mul(T1, Si, K, d, dy, dy, false);                            // K = P H^T S^-1, kalman.py:42
        if (lane < d) { double v = mu[lane]; for (int q = 0; q < dy; ++q) v += K[lane * KF_D + q] * innov[q]; tmp[lane] = v; }
        kf_sync();
        if (lane < d) { mu[lane] = tmp[lane]; a.means[(int64_t)t * d + lane] = tmp[lane]; }
        kf_mul(K, H, T1, d, dy, d, false);                              // K H
        kf_mul(T1, P, T2, d, d, d, false);                              // K H P
        for (int e = lane; e < KF_D * KF_D; e += 32) {
            P[e] -= T2[e];
            const int i = e / KF_D, j = e % KF_D;
            if (i < d && j < d) a.covs[((int64_t)t * d + i) * d + j] = P[e];
        }
        kf_sync();
    }
    if (lane == 0 && a.loglik) a.loglik[0] = ll;
}

extern "C" int mb_kalman_filter(mb_ctx* ctx, const mb_ssm* ssm, const float* y, int T, double* means, double* covs,
                                double* loglik, mb_stream_t stream) {
    MB_REQUIRE(ctx && ssm && y && means && covs && T > 0, "mb_kalman_filter: bad arguments");
    MB_REQUIRE(ssm->kind == MB_SSM_LINEAR_GAUSSIAN && ssm->dim >= 1 && ssm->dim <= KF_D && ssm->dim_obs >= 1 && ssm->dim_obs <= ssm->dim,
               "mb_kalman_filter: time-homogeneous linear-Gaussian model with dim_obs <= dim <= 8");
    KfArgs a{*ssm, y, T, means, covs, loglik};
    kalman_kernel<<<1, 32, 0, mb_s(stream)>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}
