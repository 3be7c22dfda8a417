// abc.cu -- K1c/K7: SMC-ABC population step for the g-and-k model.
//
//   abc_init / abc_move : ABCSMCSampler.startup (abc/smc.py:44-79) and vmap(forward_proposal)
//                         (abc/smc.py:184-219) with RandomWalkABC (abc/mcmc.py:40-76) on
//                         GKTransformedUniformPrior (abc/scenarios/gk.py:68-96), summary = the m simulated
//                         draws sorted, distance = L2 to data (abc/abc.py:36-38);
//                         fused with the ancestor gather of (value, prior_potential, distance, alpha).
//   abc_adapt           : MetropolisedABCSMCSampler.adapt (abc/smc.py:228-245): quantile threshold,
//                         0/-inf weights, ess = #alive, alpha_mean over previously-alive particles,
//                         RW scale = per-dimension variance * 2.38^2/d (abc/smc.py:94-98), termination
//                         (:157-161) and the resample decision for the next update (:152-155, strict <).
#include "common.cuh"
#include "rng.cuh"

#define ABC_THREADS 256

int mb_quantile_impl(mb_ctx* ctx, const float* v, int64_t n, const double* q_dev, double q_host, double* out3,
                     cudaStream_t st);

#include "gk.cuh"

template <int M>
__device__ __forceinline__ float gk_distance(const mb_gk& g, const float (&x)[GK_DIM], uint64_t seed, uint64_t gid,
                                             uint32_t step, uint32_t index0) {
    float y[M];
    gk_simulate<M>(g, x, seed, gid, step, index0, y);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < M; ++i) { const float dlt = y[i] - g.data[i]; acc = fmaf(dlt, dlt, acc); }
    return sqrtf(acc);                                                                // abc.py:36-38
}

struct AbcArgs {
    mb_gk gk;
    int mcmc_steps;
    const float* x_in; float* x_out; int64_t ld; int64_t n; int64_t n_total;
    const int32_t* anc;
    const float* up_in; float* up_out; const float* dist_in; float* dist_out;
    float* lw; const float* alpha_in; float* alpha_out;
    const float* stepsize;
    uint64_t seed; int64_t gid0;
    mb_control* ctl;
    int sample_prior;
    // population sharded over GPUs: ancestors are GLOBAL ids, the ancestor's state is read from its owner (peer mapped)
    int sharded; int64_t n_local;
    const float* x_peers[MB_MAX_WORLD]; const float* up_peers[MB_MAX_WORLD]; const float* dist_peers[MB_MAX_WORLD];
    const float* alpha_peers[MB_MAX_WORLD];
};

template <int M>
__global__ void __launch_bounds__(ABC_THREADS) abc_init_kernel(AbcArgs a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t gid = (uint64_t)(a.gid0 + i);
        float x[GK_DIM];
        if (a.sample_prior) {                                        // prior_sample N(0, I), gk.py:93-95
            philox_normals<GK_DIM>(x, a.seed, gid, 0u, MB_P_INIT, 0u);
#pragma unroll
            for (int k = 0; k < GK_DIM; ++k) a.x_out[(int64_t)k * a.ld + i] = x[k];
        } else {
#pragma unroll
            for (int k = 0; k < GK_DIM; ++k) x[k] = a.x_out[(int64_t)k * a.ld + i];
        }
        float up = 0.f;
#pragma unroll
        for (int k = 0; k < GK_DIM; ++k) up = fmaf(0.5f * x[k], x[k], up);            // gk.py:88-91
        a.up_out[i] = up;
        a.dist_out[i] = gk_distance<M>(a.gk, x, a.seed, gid, 0u, 0u);
        a.lw[i] = 0.f;
        a.alpha_out[i] = 1.f;                                        // metropolis.py:45
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        mb_control c;
        memset(&c, 0, sizeof(c));
        const double nd = (double)a.n_total;
        c.s1 = nd; c.s2 = nd; c.lse = log(nd); c.lse2 = log(nd); c.log_ess = log(nd); c.ess = nd;
        c.beta = INFINITY;                                           // threshold = inf, abc/smc.py:65-67
        c.alpha_mean = 1.0;
        c.seed = a.seed;
        *a.ctl = c;
    }
}

template <int M>
__global__ void __launch_bounds__(ABC_THREADS) abc_move_kernel(AbcArgs a) {
    const mb_control* ctl = a.ctl;
    if (ctl->done) return;
    const bool resample = ctl->resample != 0;
    const float thr = (float)ctl->beta;
    const uint32_t step = (uint32_t)(ctl->iter + 1);
    const uint64_t seed = ctl->seed;
    float sq[GK_DIM];
#pragma unroll
    for (int k = 0; k < GK_DIM; ++k) sq[k] = sqrtf(a.stepsize[k]);
    constexpr uint32_t MS = (M + 3) / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t src = resample ? (int64_t)a.anc[i] : i;
        const uint64_t gid = (uint64_t)(a.gid0 + i);
        const float* xi = a.x_in; const float* upi = a.up_in; const float* di = a.dist_in; const float* ai = a.alpha_in;
        if (a.sharded && resample) {                                 // the ancestor lives on rank o: read it over NVLink
            const int o = (int)(src / a.n_local);
            src -= (int64_t)o * a.n_local;
            xi = a.x_peers[o]; upi = a.up_peers[o]; di = a.dist_peers[o]; ai = a.alpha_peers[o];
        }
        float x[GK_DIM];
#pragma unroll
        for (int k = 0; k < GK_DIM; ++k) x[k] = __ldg(xi + (int64_t)k * a.ld + src);
        float up = __ldg(upi + src), dist = __ldg(di + src), alpha = __ldg(ai + src);
        const float lw = resample ? 0.f : a.lw[i];                   // abc/smc.py:69 via SMCSampler.resample
        if (lw > -INFINITY) {                                        // only alive particles move, :210-219
            float asum = 0.f;
            for (int s = 0; s < a.mcmc_steps; ++s) {
                float z[GK_DIM], xp[GK_DIM];
                philox_normals<GK_DIM>(z, seed, gid, step, MB_P_MOVE, (uint32_t)s * 2u);
                const float uacc = u24(philox_raw(seed, gid, step, MB_P_MOVE, (uint32_t)s * 2u + 1u).x);
                float upn = 0.f;
#pragma unroll
                for (int k = 0; k < GK_DIM; ++k) {                   // abc/mcmc.py:71
                    xp[k] = fmaf(sq[k], z[k], x[k]);
                    upn = fmaf(0.5f * xp[k], xp[k], upn);
                }
                const float dn = gk_distance<M>(a.gk, xp, seed, gid, step, (uint32_t)s * MS);
                float al = fminf(1.f, __expf(-upn + up) * ((dn < thr) ? 1.f : 0.f));   // abc/mcmc.py:59-62
                if (al != al) al = 0.f;
                if (uacc < al) {
#pragma unroll
                    for (int k = 0; k < GK_DIM; ++k) x[k] = xp[k];
                    up = upn; dist = dn;
                }
                asum += al;
            }
            alpha = asum / (float)a.mcmc_steps;
        }
#pragma unroll
        for (int k = 0; k < GK_DIM; ++k) a.x_out[(int64_t)k * a.ld + i] = x[k];
        a.up_out[i] = up; a.dist_out[i] = dist; a.alpha_out[i] = alpha;
        if (resample) a.lw[i] = 0.f;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ctl->resampled = resample ? 1 : 0;
}

static int abc_grid(mb_ctx* ctx, int64_t n) {
    int64_t grid = (n + ABC_THREADS - 1) / ABC_THREADS;
    if (grid > (int64_t)ctx->sms * 16) grid = (int64_t)ctx->sms * 16;
    return (int)grid;
}

#define ABC_DISPATCH(KERNEL)                                                                       \
    if (a.gk.m == 4) KERNEL<4><<<grid, ABC_THREADS, 0, st>>>(a);                                   \
    else if (a.gk.m == 8) KERNEL<8><<<grid, ABC_THREADS, 0, st>>>(a);                              \
    else if (a.gk.m == 16) KERNEL<16><<<grid, ABC_THREADS, 0, st>>>(a);                            \
    else { mb_set_error("abc: g-and-k summary size m=%d not built (4, 8, 16)", a.gk.m); return MB_ERR_UNSUPPORTED; }

extern "C" int mb_abc_init(mb_ctx* ctx, const mb_gk* gk, float* x, int64_t ld, int64_t n, int64_t n_total,
                           int sample_prior, float* up, float* dist, float* lw, float* alpha, uint64_t seed,
                           int64_t gid0, mb_control* ctl, mb_stream_t stream) {
    MB_REQUIRE(ctx && gk && x && up && dist && lw && alpha && ctl && n > 0 && ld >= n, "mb_abc_init: bad arguments");
    AbcArgs a{};
    a.gk = *gk; a.x_out = x; a.ld = ld; a.n = n; a.n_total = n_total; a.up_out = up; a.dist_out = dist; a.lw = lw;
    a.alpha_out = alpha; a.seed = seed; a.gid0 = gid0; a.ctl = ctl; a.sample_prior = sample_prior;
    const int grid = abc_grid(ctx, n);
    cudaStream_t st = mb_s(stream);
    ABC_DISPATCH(abc_init_kernel)
    MB_CHECK_LAUNCH();
    return MB_OK;
}

extern "C" int mb_abc_move(mb_ctx* ctx, const mb_gk* gk, int mcmc_steps, const float* x_in, float* x_out, int64_t ld,
                           int64_t n, const int32_t* anc, const float* up_in, float* up_out, const float* dist_in,
                           float* dist_out, float* lw, const float* alpha_in, float* alpha_out, const float* stepsize,
                           uint64_t seed, int64_t gid0, mb_control* ctl, mb_stream_t stream) {
    MB_REQUIRE(ctx && gk && x_in && x_out && anc && up_in && up_out && dist_in && dist_out && lw && alpha_in &&
                   alpha_out && stepsize && ctl && n > 0 && ld >= n && mcmc_steps >= 1 && x_in != x_out,
               "mb_abc_move: bad arguments");
    AbcArgs a{};
    a.gk = *gk; a.mcmc_steps = mcmc_steps; a.x_in = x_in; a.x_out = x_out; a.ld = ld; a.n = n; a.anc = anc;
    a.up_in = up_in; a.up_out = up_out; a.dist_in = dist_in; a.dist_out = dist_out; a.lw = lw; a.alpha_in = alpha_in;
    a.alpha_out = alpha_out; a.stepsize = stepsize; a.seed = seed; a.gid0 = gid0; a.ctl = ctl;
    const int grid = abc_grid(ctx, n);
    cudaStream_t st = mb_s(stream);
    ABC_DISPATCH(abc_move_kernel)
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// the same move for one shard of a population spread over `world` GPUs: anc holds GLOBAL ancestor ids and the four
// per-particle arrays of every rank are reachable through peer-mapped pointers ([world] each, this step's input parity)
extern "C" int mb_abc_move_sharded(mb_ctx* ctx, const mb_gk* gk, int mcmc_steps, float* x_out, int64_t ld, int64_t n,
                                   const int32_t* anc, float* up_out, float* dist_out, float* lw, float* alpha_out,
                                   const float* stepsize, uint64_t seed, int rank, int world, const void* const* x_peers,
                                   const void* const* up_peers, const void* const* dist_peers,
                                   const void* const* alpha_peers, mb_control* ctl, mb_stream_t stream) {
    MB_REQUIRE(ctx && gk && x_out && anc && up_out && dist_out && lw && alpha_out && stepsize && ctl && n > 0 && ld >= n &&
                   mcmc_steps >= 1 && world >= 1 && world <= MB_MAX_WORLD && rank >= 0 && rank < world && x_peers && up_peers &&
                   dist_peers && alpha_peers, "mb_abc_move_sharded: bad arguments");
    AbcArgs a{};
    a.gk = *gk; a.mcmc_steps = mcmc_steps; a.x_out = x_out; a.ld = ld; a.n = n; a.anc = anc;
    a.up_out = up_out; a.dist_out = dist_out; a.lw = lw; a.alpha_out = alpha_out; a.stepsize = stepsize; a.seed = seed;
    a.gid0 = (int64_t)rank * n; a.ctl = ctl; a.sharded = 1; a.n_local = n;
    for (int r = 0; r < world; ++r) {
        a.x_peers[r] = (const float*)x_peers[r]; a.up_peers[r] = (const float*)up_peers[r];
        a.dist_peers[r] = (const float*)dist_peers[r]; a.alpha_peers[r] = (const float*)alpha_peers[r];
        MB_REQUIRE(a.x_peers[r] && a.up_peers[r] && a.dist_peers[r] && a.alpha_peers[r], "mb_abc_move_sharded: peer pointer missing");
    }
    a.x_in = a.x_peers[rank]; a.up_in = a.up_peers[rank]; a.dist_in = a.dist_peers[rank]; a.alpha_in = a.alpha_peers[rank];
    MB_REQUIRE(a.x_in != x_out, "mb_abc_move_sharded: in-place move");
    const int grid = abc_grid(ctx, n);
    cudaStream_t st = mb_s(stream);
    ABC_DISPATCH(abc_move_kernel)
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// ------------------------------------------------------------------------------------------------ adapt
struct AdaptArgs {
    int64_t n; int64_t n_total;
    const float* dist; float* lw; const float* alpha; float* stepsize;
    double ess_retain, ess_resample, termination_alpha;
    int max_iter; const double* schedule; int advance_iter; int d;
    mb_control* ctl; mb_hist* hist;
    double* q_dev;          // scratch: [0] quantile level
    double* thr3;           // scratch: quantile output (value, lo, hi)
    const double* var;      // scratch: per-dimension variance
    unsigned long long* counts;   // [0] alive_new, [1] alive_prev, [2] alpha_fx
};

__global__ void abc_q_kernel(AdaptArgs a) {
    const mb_control* c = a.ctl;
    const double ess = c->resampled ? (double)a.n_total : c->ess;    // ess[0] of the post-resample state
    a.q_dev[0] = a.ess_retain * ess / (double)a.n_total;             // abc/smc.py:166
    a.counts[0] = 0; a.counts[1] = 0; a.counts[2] = 0;
}

__global__ void __launch_bounds__(ABC_THREADS) abc_weight_kernel(AdaptArgs a) {
    if (a.ctl->done) return;
    const int iter_new = a.ctl->iter + (a.advance_iter ? 1 : 0);
    const float thr = a.schedule ? (float)a.schedule[iter_new] : (float)a.thr3[0];
    unsigned long long alive_new = 0, alive_prev = 0;
    double asum = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const float lw_old = a.lw[i];
        const bool ap = lw_old > -INFINITY;                          // alive_inds, abc/smc.py:240
        alive_prev += ap;
        if (ap) asum += (double)a.alpha[i];
        const bool dead = a.dist[i] > thr;                           // :168-173
        a.lw[i] = dead ? -INFINITY : 0.f;
        alive_new += !dead;
    }
    __shared__ double red[ABC_THREADS / 32];
    const double s_new = block_sum_d((double)alive_new, red);
    const double s_prev = block_sum_d((double)alive_prev, red);
    const double s_alpha = block_sum_d(asum, red);
    if (threadIdx.x == 0) {
        atomicAdd(&a.counts[0], (unsigned long long)s_new);
        atomicAdd(&a.counts[1], (unsigned long long)s_prev);
        atomicAdd(&a.counts[2], (unsigned long long)llrint(s_alpha * 4294967296.0));
    }
}

__global__ void abc_finish_kernel(AdaptArgs a) {
    mb_control c = *a.ctl;
    if (c.done) return;
    const int iter_new = c.iter + (a.advance_iter ? 1 : 0);
    const double thr = a.schedule ? a.schedule[iter_new] : a.thr3[0];
    const double alive = (double)a.counts[0];
    c.beta = thr;
    c.wmax = 0.0; c.s1 = alive; c.s2 = alive;
    c.lse = log(alive); c.lse2 = log(alive); c.log_ess = c.lse; c.ess = alive;   // ess = #alive (Appendix A.1)
    c.alpha_mean = ((double)a.counts[2] / 4294967296.0) / (double)a.counts[1];   // :240-241
    c.iter = iter_new;
    c.resample = (c.ess < a.ess_resample * (double)a.n_total) ? 1 : 0;           // :152-155 strict
    c.done = (c.alpha_mean <= a.termination_alpha || iter_new >= a.max_iter) ? 1 : 0;   // :157-161
    for (int k = 0; k < a.d; ++k) a.stepsize[k] = (float)(a.var[k] / (double)a.d * 2.38 * 2.38);   // :94-98
    *a.ctl = c;
    if (a.hist && iter_new < MB_HIST_MAX) {
        mb_hist h;
        h.beta = thr; h.ess = c.ess; h.log_z = 0.0; h.alpha_mean = c.alpha_mean; h.lse = c.lse;
        h.resampled = c.resampled; h.search_iters = 0;
        a.hist[iter_new] = h;
    }
}

extern "C" int mb_abc_adapt(mb_ctx* ctx, const float* x, int64_t ld, int64_t n, int64_t n_total, int d,
                            const float* dist, float* lw, const float* alpha, float* stepsize, double ess_retain,
                            double ess_resample, double termination_alpha, int max_iter, const double* schedule,
                            int advance_iter, mb_control* ctl, mb_hist* hist, mb_stream_t stream) {
    MB_REQUIRE(ctx && x && dist && lw && alpha && stepsize && ctl && n > 1 && d > 0 && d <= 16,
               "mb_abc_adapt: bad arguments");
    cudaStream_t st = mb_s(stream);
    // scratch layout (ctx->scratch, 8 MiB): [0,1M) colstats partials | [1M, 1M+4K) select state | [2M, ...) adapt
    if (mb_ensure_scratch(ctx, 4u << 20) != MB_OK) return MB_ERR_CUDA;
    char* base = (char*)ctx->scratch + (2u << 20);
    AdaptArgs a{};
    a.n = n; a.n_total = n_total; a.dist = dist; a.lw = lw; a.alpha = alpha; a.stepsize = stepsize;
    a.ess_retain = ess_retain; a.ess_resample = ess_resample; a.termination_alpha = termination_alpha;
    a.max_iter = max_iter; a.schedule = schedule; a.advance_iter = advance_iter; a.d = d; a.ctl = ctl; a.hist = hist;
    a.q_dev = (double*)base;
    a.thr3 = (double*)(base + 64);
    double* mean = (double*)(base + 128);
    double* var = (double*)(base + 128 + 16 * sizeof(double));
    a.var = var;
    a.counts = (unsigned long long*)(base + 512);
    abc_q_kernel<<<1, 1, 0, st>>>(a);
    MB_CHECK_LAUNCH();
    if (!schedule) {
        int rc = mb_quantile_impl(ctx, dist, n, a.q_dev, 0.0, a.thr3, st);
        if (rc != MB_OK) return rc;
    }
    int rc = mb_colstats(ctx, x, ld, n, d, mean, var, stream);
    if (rc != MB_OK) return rc;
    const int grid = abc_grid(ctx, n);
    abc_weight_kernel<<<grid, ABC_THREADS, 0, st>>>(a);
    MB_CHECK_LAUNCH();
    abc_finish_kernel<<<1, 1, 0, st>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}


// ------------------------------------------------------------------------------------------------ adapt, staged
// The same adaptation (abc/smc.py:163-166 quantile, :94-98 column variances, :228-245) for ONE SHARD of a population
// spread over several GPUs: every stage works on the local particles and leaves small integer / fp64 records in the
// caller's workspace `ws`; between the stages the caller adds the ranks' records (any all-reduce -- the host layer uses
// NCCL), so every rank takes identical decisions.  Workspace (MB_ABC_WS_BYTES, device):
//   [0] q  [8] thr3[3]  [64] counts[3] u64  [128] colsums[1 + 2 d] fp64  [512] select state  [576] frac[2]
//   [640] select exchange: count_le (i64), min key above (i64)   [704] variances[d]   [1024] 2048 u32 histogram counters
// Stages:  0 begin (q, select rank, column sums -> all-reduce colsums)
//          1..3 histogram of radix pass p (-> all-reduce hist)     11..13 pick of pass p
//          4 count / next-larger pass (-> all-reduce SUM of exch[0], MIN of exch[1])
//          5 threshold + weight update (-> all-reduce counts)       6 control block, step sizes
#include "select.cuh"

__global__ void __launch_bounds__(256) abc_colsums_kernel(const float* __restrict__ x, int64_t ld, int64_t n, int d, double* out) {
    __shared__ double red[256 / 32];
    const int col = blockIdx.y;
    const float* xc = x + (int64_t)col * ld;
    double s1 = 0, s2 = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)xc[i];
        s1 += v; s2 += v * v;
    }
    s1 = block_sum_d(s1, red); s2 = block_sum_d(s2, red);
    if (threadIdx.x == 0) {                        // fp64 atomics: the order of the adds differs run to run at the last bit
        atomicAdd(out + 1 + col, s1);              // of a 1e8-term sum; the variances only scale the random-walk proposal
        atomicAdd(out + 1 + d + col, s2);
        if (col == 0 && blockIdx.x == 0) out[0] = (double)n;
    }
}

__global__ void abc_select_export_kernel(const SelectState* st, long long* exch) {
    exch[0] = (long long)st->count_le;
    exch[1] = (long long)st->min_gt_key;
}
__global__ void abc_select_import_kernel(SelectState* st, const long long* exch) {
    st->count_le = exch[0];
    st->min_gt_key = (uint32_t)exch[1];
}
__global__ void abc_var_kernel(const double* colsums, int d, double* var) {
    const int k = threadIdx.x;
    if (k >= d) return;
    const double n = colsums[0], m = colsums[1 + k] / n;
    var[k] = (colsums[1 + d + k] - n * m * m) / (n - 1.0);             // ddof = 1 (vmap(jnp.cov), abc/smc.py:97)
}

extern "C" int mb_abc_adapt_stage(mb_ctx* ctx, int stage, const float* x, int64_t ld, int64_t n, int64_t n_total, int d,
                                  const float* dist, float* lw, const float* alpha, float* stepsize, double ess_retain,
                                  double ess_resample, double termination_alpha, int max_iter, const double* schedule,
                                  int advance_iter, void* ws, mb_control* ctl, mb_hist* hist, mb_stream_t stream) {
    MB_REQUIRE(ctx && x && dist && lw && alpha && stepsize && ctl && ws && n > 1 && d > 0 && d <= 16, "mb_abc_adapt_stage: bad arguments");
    cudaStream_t st = mb_s(stream);
    char* base = (char*)ws;
    AdaptArgs a{};
    a.n = n; a.n_total = n_total; a.dist = dist; a.lw = lw; a.alpha = alpha; a.stepsize = stepsize;
    a.ess_retain = ess_retain; a.ess_resample = ess_resample; a.termination_alpha = termination_alpha;
    a.max_iter = max_iter; a.schedule = schedule; a.advance_iter = advance_iter; a.d = d; a.ctl = ctl; a.hist = hist;
    a.q_dev = (double*)base;
    a.thr3 = (double*)(base + 8);
    a.counts = (unsigned long long*)(base + 64);
    double* colsums = (double*)(base + 128);
    double* var = (double*)(base + 704);
    a.var = var;
    SelectState* state = (SelectState*)(base + 512);
    double* frac = (double*)(base + 576);
    long long* exch = (long long*)(base + 640);
    uint32_t* hst = (uint32_t*)(base + 1024);
    const int grid = abc_grid(ctx, n);
    int sgrid = (int)((n + (int64_t)SEL_THREADS * 16 - 1) / ((int64_t)SEL_THREADS * 16));
    if (sgrid > ctx->sms * 8) sgrid = ctx->sms * 8;
    if (sgrid < 1) sgrid = 1;
    switch (stage) {
    case 0:
        abc_q_kernel<<<1, 1, 0, st>>>(a);
        select_rank_kernel<<<1, 1, 0, st>>>(state, a.q_dev, 0.0, n_total, frac);
        MB_CUDA(cudaMemsetAsync(hst, 0, 2048 * sizeof(uint32_t), st));
        MB_CUDA(cudaMemsetAsync(colsums, 0, 40 * sizeof(double), st));
        abc_colsums_kernel<<<dim3(sgrid > 64 ? 64 : sgrid, d), 256, 0, st>>>(x, ld, n, d, colsums);
        break;
    case 1: select_hist_kernel<21, 11, 0><<<sgrid, SEL_THREADS, 0, st>>>(dist, n, state, hst); break;
    case 2: select_hist_kernel<10, 11, 11><<<sgrid, SEL_THREADS, 0, st>>>(dist, n, state, hst); break;
    case 3: select_hist_kernel<0, 10, 22><<<sgrid, SEL_THREADS, 0, st>>>(dist, n, state, hst); break;
    case 11: select_pick_kernel<21, 11><<<1, 256, 0, st>>>(state, hst); break;
    case 12: select_pick_kernel<10, 11><<<1, 256, 0, st>>>(state, hst); break;
    case 13: select_pick_kernel<0, 10><<<1, 256, 0, st>>>(state, hst); break;
    case 4:
        select_next_kernel<<<sgrid, SEL_THREADS, 0, st>>>(dist, n, state);
        abc_select_export_kernel<<<1, 1, 0, st>>>(state, exch);
        break;
    case 5:
        if (!schedule) {
            abc_select_import_kernel<<<1, 1, 0, st>>>(state, exch);
            select_finish_dev_kernel<<<1, 1, 0, st>>>(state, frac, a.thr3);
        }
        abc_weight_kernel<<<grid, ABC_THREADS, 0, st>>>(a);
        break;
    case 6:
        abc_var_kernel<<<1, 32, 0, st>>>(colsums, d, var);
        abc_finish_kernel<<<1, 1, 0, st>>>(a);
        break;
    default:
        mb_set_error("mb_abc_adapt_stage: unknown stage %d", stage);
        return MB_ERR_ARG;
    }
    MB_CHECK_LAUNCH();
    return MB_OK;
}
