/* mocat_b200.h -- C-ABI of libmocat_b200.so: the B200 (sm_100a) particle-population hot path
 * behind mocat's Python API.
 *
 * The reference (SamDuffield/mocat v0.2.6) is pure Python on JAX and has NO FFI boundary; its seam
 * is the Python method protocol (Sampler.startup/update, run_particle_filter_for_marginals ...).
 * Each entry point below names the reference code (path:line under /root/reference/mocat/src/) whose
 * device work it replaces; INTEGRATION.md shows the ctypes binding a mocat maintainer would add.
 *
 * Conventions
 *  - every pointer is a CUDA DEVICE pointer owned by the caller unless the name ends in _host;
 *  - particle state is SoA: column k of a (d x n) block lives at base + k*ld  (ld >= n, elements);
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and returns
 *    MB_OK or an error code; mb_last_error() gives the message of the last failure in this thread;
 *  - a context is bound to one device and one host thread; calls on one context are not thread safe;
 *  - random numbers: Philox4x32-10, counter (gid_lo, gid_hi, step, (purpose<<20)|index), key = seed;
 *    gid is the GLOBAL particle index = gid0 + local index, so results do not depend on sharding.
 */
#ifndef MOCAT_B200_H
#define MOCAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_ABI_VERSION 1

#define MB_OK              0
#define MB_ERR_CUDA        1
#define MB_ERR_ARG         2
#define MB_ERR_UNSUPPORTED 3

#define MB_MAX_SMALL_DIM   8    /* dense d x d models (Gaussian target, linear-Gaussian SSM) */
#define MB_HIST_MAX        16384
#define MB_MAX_WORLD       8    /* GPUs of one NVSwitch domain sharing a population */

typedef struct mb_ctx mb_ctx;
typedef struct mb_comm mb_comm;
typedef void* mb_stream_t;

/* ---- device-resident control block: every per-iteration scalar the reference keeps replicated per
 *      particle (temperature, ess, log_norm_constant: transport/smc.py:157-162,177-184; threshold:
 *      abc/smc.py:86-91) lives here once, is updated by the kernels, and is read back by the host
 *      only when it wants to look.  All kernels that take a control block early-exit when done != 0. */
typedef struct {
    double wmax, s1, s2;        /* max(lw), sum exp(lw-wmax), sum exp(2(lw-wmax)) of the CURRENT weights   */
    double lse, lse2, log_ess;  /* logsumexp(lw), logsumexp(2lw), 2*lse-lse2   (metrics.py:69-74)          */
    double ess;                 /* exp(log_ess)                                                            */
    double log_z;               /* running log normalising constant (transport/smc.py:160,212-215)         */
    double beta;                /* temperature (transport/smc.py:157) or ABC threshold (abc/smc.py:67)     */
    double alpha_mean;          /* mean acceptance prob. of the last move (abc/smc.py:240-241)             */
    double aux0, aux1;          /* kernel specific (alive count, sum alpha ...)                            */
    int64_t nan_count;          /* NaN entries in the value block after the last move (smc.py:174-175)     */
    int64_t alpha_fx;           /* sum of per-particle acceptance probabilities, fixed point 2^-32          */
    uint64_t seed;              /* Philox key of this population (set by the *_init kernels; read by the     */
                                /* move / ancestor kernels so that a captured CUDA graph is seed-agnostic)  */
    int32_t iter;               /* extra.iter (sample.py:64-65)                                            */
    int32_t resample;           /* resample_criterion evaluated for the NEXT update (smc.py:298-301)       */
    int32_t done;               /* termination_criterion (smc.py:171-175, abc/smc.py:157-161)              */
    int32_t search_iters;       /* regula-falsi iterations of the last temperature search                  */
    int32_t resampled;          /* whether the last update resampled                                       */
    int32_t pad0;
} mb_control;

/* per-iteration record appended by the adapt kernels (what clean_chain keeps: smc.py:177-184) */
typedef struct {
    double beta, ess, log_z, alpha_mean, lse;
    int32_t resampled, search_iters;
} mb_hist;

/* ---- scenario descriptions (POD; replaces the per-particle Python callables of core.py:146-261) ---- */
enum { MB_LIK_RASTRIGIN = 0,   /* scenarios/toy_examples.py:135-149 */
       MB_LIK_GAUSSIAN  = 1,   /* scenarios/toy_examples.py:17-51   */
       MB_LIK_NONE      = 2 };
enum { MB_MOVE_MALA = 0,       /* mcmc/standard_mcmc.py:72-153 Underdamped(friction=inf), utils.py:108-146 */
       MB_MOVE_RW   = 1 };     /* mcmc/standard_mcmc.py:21-65  RandomWalk                                   */
enum { MB_RESAMPLE_SYSTEMATIC = 0, MB_RESAMPLE_MULTINOMIAL = 1 };

typedef struct {
    int32_t kind, dim;
    float prior_mean, prior_std, prior_pscale; /* prior_sample = mean + std z; U_prior = .5 sum(((x-mean) pscale)^2) */
    float a;                                   /* Rastrigin a                                                        */
    float mean[MB_MAX_SMALL_DIM];              /* Gaussian target mean                                               */
    float prec_sqrt[MB_MAX_SMALL_DIM * MB_MAX_SMALL_DIM]; /* row-major S: y = (x-mean) S^T, U = .5|y|^2              */
} mb_target;

typedef struct {
    int32_t kind;            /* MB_MOVE_*                                  */
    int32_t mcmc_steps;      /* transport/smc.py:255                       */
    int32_t leapfrog_steps;  /* mcmc/standard_mcmc.py:78                   */
    float   stepsize;
} mb_move;

typedef struct {             /* adaptive likelihood tempering, transport/smc.py:231-259,311-326 */
    double  max_temperature, ess_retain, ess_resample, tol;
    int32_t max_search_iter, max_iter;
    const double* schedule;  /* device, or NULL for adaptive (transport/smc.py:115-126) */
    int32_t schedule_len, pad;
} mb_temper;

enum { MB_SSM_LINEAR_GAUSSIAN = 0,  /* ssm/linear_gaussian/linear_gaussian.py:142-261 */
       MB_SSM_LORENZ96 = 1 };       /* ssm/scenarios/lorenz96.py:14-44 on ssm/nonlinear_gaussian.py:19-131 */
enum { MB_PROPOSAL_BOOTSTRAP = 0,   /* ssm/filtering.py:142-170 */
       MB_PROPOSAL_OPTIMAL = 1,     /* ssm/nonlinear_gaussian.py:134-276 (H = I, diagonal noise) */
       MB_PROPOSAL_ENKF = 2 };      /* ssm/nonlinear_gaussian.py:279-350: conditioned initial sample as OPTIMAL, forecast =
                                       the bootstrap step, analysis = mb_enkf_analysis */

typedef struct {
    int32_t kind, dim, dim_obs, substeps;
    /* linear Gaussian (dim, dim_obs <= MB_MAX_SMALL_DIM), all row-major */
    float m0[MB_MAX_SMALL_DIM];
    float L0[MB_MAX_SMALL_DIM * MB_MAX_SMALL_DIM];      /* chol(P0)               */
    float F[MB_MAX_SMALL_DIM * MB_MAX_SMALL_DIM];
    float LQ[MB_MAX_SMALL_DIM * MB_MAX_SMALL_DIM];      /* chol(Q)                */
    float H[MB_MAX_SMALL_DIM * MB_MAX_SMALL_DIM];
    float Rps[MB_MAX_SMALL_DIM * MB_MAX_SMALL_DIM];     /* inv(chol(R)) used as in utils.py:26-30: .5|(y-Hx) Rps|^2 */
    float lik_const;                                    /* .5 (d_y log 2pi - log det R^-1) */
    /* Lorenz-96 with diagonal noise */
    float forcing, dt, q_std, r_std, init_mean, init_std;
    /* proposal of the filter (Lorenz-96 kernels): MB_PROPOSAL_BOOTSTRAP = transition (ssm/filtering.py:142-170),
     * MB_PROPOSAL_OPTIMAL = OptimalNonLinearGaussianParticleFilter (ssm/nonlinear_gaussian.py:134-276) */
    int32_t proposal;
} mb_ssm;

typedef struct {              /* g-and-k, abc/scenarios/gk.py:68-96 */
    int32_t m;                /* simulated draws per particle, sorted = summary (<= 16) */
    float c, prior_min, prior_max, buffer;
    float data[16];
} mb_gk;

/* ---- sharded population: rank r owns the contiguous global index range [r*n_local, (r+1)*n_local).  The
 *      peer tables hold CUDA-IPC mapped pointers into every rank's HBM so that a kernel can read an
 *      ancestor's state (or CDF) directly over NVLink: the redistribution after resampling is fused into
 *      the move kernel's gather instead of being a separate all-to-all. */
typedef struct {
    int32_t rank, world;
    int64_t n_local, n_total;
    const float*  x_peers[MB_MAX_WORLD];     /* input value block (d x ld) of every rank for this step   */
    const double* cdf_peers[MB_MAX_WORLD];   /* rank-relative exact fp64 CDF of every rank               */
    const double* totals;                    /* device [world]: quantised weight total of every rank     */
    int32_t*      anc_peers[MB_MAX_WORLD];   /* ancestor array (n_local int32) of every rank: the fused   */
                                             /* resampler writes an output's ancestor where the output lives */
    const float*  lw_peers[MB_MAX_WORLD];    /* log-weights (n_local) of every rank: heavy source tiles are */
                                             /* re-derived by every rank whose outputs they feed            */
    const void*   ws_peers[MB_MAX_WORLD];    /* mb_rs_* workspace of every rank (its heavy-tile records)    */
} mb_shard;

/* ---- context ---------------------------------------------------------------------------------- */
const char* mb_last_error(void);
int         mb_abi_version(void);
mb_ctx*     mb_create(int device);
void        mb_destroy(mb_ctx* ctx);
int         mb_sm_count(mb_ctx* ctx);
/* Generation of the context-owned workspaces (scratch, scan status): bumped whenever one of them is reallocated.
 * A CUDA graph captured from these entry points embeds the old addresses and must be re-captured after a change. */
unsigned long long mb_workspace_generation(mb_ctx* ctx);

/* ---- K2: log-sum-exp / ESS reduction.  Replaces metrics.py:69-78 (log_ess_log_weight), the logsumexp
 *      calls at transport/smc.py:160,214-215,288 and MetropolisedSMCSampler.log_ess (smc.py:303-309).
 *      out6 (device) = {wmax, s1, s2, lse, lse2, log_ess}. lik==NULL: plain; else weights lw - dbeta*lik. */
int mb_lse_ess(mb_ctx* ctx, const float* lw, const float* lik, double dbeta, int64_t n,
               double* out6, mb_stream_t stream);

/* ---- K3: adaptive-temperature search + weight update, entirely on device.  Replaces utils.py:205-237
 *      (bisect), transport/smc.py:311-326 (next_temperature_adaptive), :193-217 (adapt), :171-175
 *      (termination), :298-301 (resample criterion for the next update).  Reads/writes ctl; lw updated
 *      in place: lw += -(beta' - beta) * lik.  Appends one mb_hist record at hist[ctl->iter] if hist. */
int mb_temper_adapt(mb_ctx* ctx, float* lw, const float* lik, int64_t n, const mb_temper* prm,
                    int advance_iter, int64_t nan_denominator, int64_t n_total, mb_control* ctl,
                    mb_hist* hist, mb_comm* comm /*NULL: single GPU*/, mb_stream_t stream);

/* ---- K4: inclusive fp64 CDF of the normalised weights (implicit in random.categorical at
 *      transport/smc.py:65, ssm/filtering.py:199).  Exact-fp64 convention: q_i = rint(w_i*scale)*2^-52,
 *      cdf = min(cumsum(q), 1), cdf[n-1] = 1.   _lw: w_i = exp(lw_i - ctl->wmax), scale from ctl->s1
 *      (predicated on ctl->resample unless force).   _f32: caller supplies linear weights and scale. */
int mb_cumsum_lw(mb_ctx* ctx, const float* lw, int64_t n, const mb_control* ctl, int flags /*1 force, 2 raw*/,
                 double* cdf, mb_stream_t stream);
int mb_cumsum_f32(mb_ctx* ctx, const float* w, int64_t n, double scale, double* cdf, mb_stream_t stream);

/* ---- K5: ancestors a_i = min{ j : cdf[j] > u_i }.  mode systematic: u_i = (i + u0)/n_out;
 *      multinomial: n_out iid u_i.  u (device, fp64) may be NULL => Philox(seed, step, P_RESAMPLE).
 *      ctl may be NULL (always run) else predicated on ctl->resample. */
int mb_ancestors(mb_ctx* ctx, const double* cdf, int64_t n, int mode, const double* u,
                 uint64_t seed, uint32_t step, int64_t gid0, int32_t* anc, int64_t n_out,
                 const mb_control* ctl, mb_stream_t stream);

/* sharded variant: this rank's n_out output slots are the global slots sh->rank*n_local + [0, n_out); the
 * ancestors are GLOBAL particle indices found in the concatenation of the ranks' relative CDFs
 * (offset_r + cdf_r[j], exact in fp64), i.e. bit-identical to the single-GPU result. */
int mb_ancestors_sharded(mb_ctx* ctx, const mb_shard* sh, int mode, uint64_t seed, uint32_t step, int32_t* anc,
                         int64_t n_out, const mb_control* ctl, mb_stream_t stream);

/* ---- sorted-uniform resampling (production path of the engines).  Systematic, or STRATIFIED-EXACT multinomial:
 *      the n iid uniforms of multinomial resampling are generated as (counts of first-stage uniforms in B equal
 *      strata) + (fresh second-stage uniforms inside each stratum) -- identical in law to Cat(softmax(w))^n
 *      (transport/smc.py:65-67) but sorted by stratum, so ancestors are nearly sorted, the CDF window of a block
 *      of outputs is staged in shared memory and the fused gather stays coalesced.  B = mb_strata_count(n).
 *      mb_strata_hist: first stage (integer histogram, deterministic); mb_strata_reduce: sum over the ranks of
 *      a sharded population (peer reads after a mailbox barrier); mb_ancestors_sorted: scan of the counts +
 *      ancestor search (sh == NULL: cdf is the materialised CDF of n particles; else the sharded global CDF). */
int mb_strata_count(int64_t n_total_out);
int mb_strata_hist(mb_ctx* ctx, int64_t n_out, int64_t gid0, int B, uint64_t seed, uint32_t step,
                   const mb_control* ctl, uint32_t* hist, int clear /*1: zero hist first; 0: it is already zero --
                   mb_ancestors_sorted leaves the counts it consumed zeroed*/, mb_stream_t stream);
int mb_strata_reduce(mb_ctx* ctx, mb_comm* comm, const void* const* hist_peers, int world, int B,
                     uint32_t* hist_out, int barrier /*0: caller exchanged after its histogram kernel*/,
                     const mb_control* ctl, mb_stream_t stream);
int mb_ancestors_sorted(mb_ctx* ctx, const double* cdf, int64_t n, const mb_shard* sh, int mode,
                        uint32_t* hist /*consumed: zero on return*/, uint32_t* offsets, int B, uint64_t seed, uint32_t step, int64_t gid0,
                        int64_t n_total_out, int32_t* anc, int64_t n_out, const mb_control* ctl, mb_stream_t stream);

/* ---- K6: gather of SoA state columns by ancestor.  Replaces cdict.__getitem__ (core.py:46-56) as
 *      used at transport/smc.py:68 and ssm/filtering.py:199. */
int mb_gather_state(mb_ctx* ctx, const int32_t* anc, int64_t n_out, int ncols,
                    const float* src, int64_t ld_src, float* dst, int64_t ld_dst, mb_stream_t stream);

/* ---- K1a: tempered-SMC population move.  Replaces vmap(forward_proposal) (transport/smc.py:91-95,
 *      337-365): MCMC startup (potentials re-evaluated at ctl->beta), mcmc_steps Metropolised moves
 *      (mcmc/sampler.py:92-111, mcmc/metropolis.py:48-70), prior/likelihood potentials of the new state.
 *      If ctl->resample: particle i starts from x_in[:, anc[i]] and lw[i] is reset to 0 (smc.py:61-71).
 *      Outputs x_out (d x ld), up_out, lik_out, alpha_out (any of up/alpha may be NULL). */
int mb_smc_init(mb_ctx* ctx, const mb_target* tgt, float* x, int64_t ld, int64_t n, int64_t n_total,
                int sample_prior, float* up, float* lik, float* lw, uint64_t seed, int64_t gid0,
                mb_control* ctl, mb_stream_t stream);
int mb_smc_move(mb_ctx* ctx, const mb_target* tgt, const mb_move* mv, const float* x_in, float* x_out,
                int64_t ld, int64_t n, const int32_t* anc, float* lw, float* up_out, float* lik_out,
                float* alpha_out, uint64_t seed, int64_t gid0, mb_control* ctl, const mb_shard* sh /*or NULL*/,
                mb_stream_t stream);

/* Robbins-Monro stepsize adaptation (RMMetropolisedSMCSampler.adapt, transport/smc.py:406-421) on the device:
 * alpha_mean = sum softmax(lw)_i alpha_i, log stepsize += rm_stepsize (alpha_mean - target).  The stepsize lives in
 * ctl->aux1 and is read by mb_smc_move when mb_move.stepsize <= 0; init_stepsize > 0 sets it (after mb_smc_init) instead
 * of adapting.  stepsize_hist (device doubles [MB_HIST_MAX], or NULL): stepsize_hist[ctl->iter] after every adaptation. */
int mb_rm_adapt(mb_ctx* ctx, const float* alpha, const float* lw, int64_t n, double rm_stepsize, double target,
                double init_stepsize, mb_control* ctl, double* stepsize_hist /*or NULL*/, mb_stream_t stream);

/* ---- K1b: bootstrap particle filter.  Replaces initiate_particles (ssm/filtering.py:173-193) and one
 *      body of the scan in run_particle_filter_for_marginals (:280-311): optional ancestor gather,
 *      transition_sample, log-weight increment -likelihood_potential, ESS, log-evidence, and the
 *      resample decision for the next step (ess < ess_threshold*n, strict).  t is the time index. */
int mb_pf_init(mb_ctx* ctx, const mb_ssm* ssm, float* x, int64_t ld, int64_t n, int64_t n_total,
               const float* y0, float* lw, uint64_t seed, int64_t gid0, double ess_threshold,
               mb_control* ctl, mb_hist* hist, mb_comm* comm /*or NULL*/, mb_stream_t stream);
int mb_pf_step(mb_ctx* ctx, const mb_ssm* ssm, const float* x_in, float* x_out, int64_t ld, int64_t n,
               int64_t n_total, const int32_t* anc, const float* y, float* lw, uint64_t seed, uint32_t t,
               int64_t gid0, double ess_threshold, mb_control* ctl, mb_hist* hist, const mb_shard* sh /*or NULL*/,
               mb_comm* comm /*or NULL*/, mb_stream_t stream);

/* Kalman filter of the time-homogeneous linear-Gaussian model (ssm/linear_gaussian/kalman.py:16-57): exact filtering
 * means [T][dim] and covariances [T][dim][dim] (device fp64) and the innovation log-likelihood (loglik, may be NULL)
 * the particle filter's log-evidence is checked against (config C1).  y: device float [T][dim_obs].
 * cov_0 = L0 L0^T (kalman.py:20 passes the Cholesky factor instead; identical when P0 = I). */
int mb_kalman_filter(mb_ctx* ctx, const mb_ssm* ssm, const float* y, int T, double* means, double* covs,
                     double* loglik /*or NULL*/, mb_stream_t stream);

/* Backward simulation (FFBSi; ssm/backward.py:20-40 full_resampling, :241-272 backward_simulation_full) for the
 * Gaussian-transition models: for every backward sample x1_j (n_s, d) draw idx_j ~ Cat(lw0_i - transition_potential(
 * x0_i -> x1_j)) over the n_pf filter particles x0 (n_pf, d) by Gumbel-max (Philox counter: gid = j, step, slot i / 4)
 * and return x_out_j = x0[idx_j].  x1 == NULL: no transition term (the final-time categorical draw from the weights).
 * All arrays device, row-major; work: n_pf * d floats. */
int mb_backward_sample(mb_ctx* ctx, const mb_ssm* ssm, float dt, const float* x0, const float* lw0, int64_t n_pf,
                       const float* x1 /*or NULL*/, int64_t n_s, float* work, uint64_t seed, uint32_t step, int32_t* idx,
                       float* x_out, mb_stream_t stream);
/* Fixed-lag stitching (ssm/online_smoothing.py:21-44 full_stitch, :167-207 fixed_lag_stitching): for every fixed
 * trajectory end x0_i (n_s, d) draw idx_i ~ Cat(lw1_j - transition_potential(x0_i -> x1_j)) over the n_c candidate
 * continuations x1 (n_c, d) by Gumbel-max (Philox counter: gid = i, step, purpose 5, slot j / 4).  work: (n_s + n_c) * d
 * floats.  mb_transition_potential: pot_i = transition_potential(x0_i -> x1_i) of n matched pairs, normalising constant
 * included (linear_gaussian.py:73-84, nonlinear_gaussian.py:98-105, utils.py:49-79); work: n * d floats. */
int mb_stitch_sample(mb_ctx* ctx, const mb_ssm* ssm, float dt, const float* x0, int64_t n_s, const float* x1,
                     const float* lw1, int64_t n_c, float* work, uint64_t seed, uint32_t step, int32_t* idx,
                     mb_stream_t stream);
int mb_transition_potential(mb_ctx* ctx, const mb_ssm* ssm, float dt, const float* x0, const float* x1, int64_t n,
                            float* work, float* pot, mb_stream_t stream);

/* ---- K1b', Lorenz-96 (config C3): same contract as mb_pf_init / mb_pf_step for MB_SSM_LORENZ96 (dim 8, 16 or 40,
 *      H = I, diagonal noise, `substeps` RK4 steps per observation interval; ssm/scenarios/lorenz96.py:14-44 on
 *      ssm/nonlinear_gaussian.py:107-121) on the ROW-MAJOR layout of the reference's `value` array: particle i,
 *      coordinate k lives at x[i * dim + k] (allocate ceil(n/32)*32 rows; lw padded to a multiple of 32).  Source rows
 *      are staged by the TMA engine (one span, or one row per scattered / remote ancestor), finished rows leave by
 *      bulk store.  A particle pair is spread over four lanes and advanced with packed fp32x2 arithmetic;
 *      normals: Philox counter (gid >> 1, t, purpose << 20 | k / 2), words (2 (k & 1), 2 (k & 1) + 1), Box-Muller cos
 *      branch -> even particle, sin branch -> odd particle (gid0 must be even).  anc holds GLOBAL ancestor ids. */
int mb_pf_l96_init(mb_ctx* ctx, const mb_ssm* ssm, float* x_rows, int64_t n, int64_t n_total, const float* y0,
                   float* lw, uint64_t seed, int64_t gid0, double ess_threshold, mb_control* ctl, mb_hist* hist,
                   mb_comm* comm /*or NULL*/, mb_stream_t stream);
int mb_pf_l96_step(mb_ctx* ctx, const mb_ssm* ssm, const float* x_in_rows, float* x_out_rows, int64_t n,
                   int64_t n_total, const int32_t* anc, const float* y, float* lw, uint64_t seed, uint32_t t,
                   int64_t gid0, double ess_threshold, mb_control* ctl, mb_hist* hist, const mb_shard* sh /*or NULL*/,
                   mb_comm* comm /*or NULL*/, mb_stream_t stream);
/* ---- kernelised Stein discrepancy (metrics.ksd, metrics.py:88-130) under the Gaussian kernel (kernels.py:82-116):
 *      out3[0] = sqrt(sum_ij k0(x_i, x_j) w_i w_j) / sum_i w_i, out3[1] = the double sum, out3[2] = sum w (device fp64);
 *      w = exp(log_weight) or 1 (log_weight NULL).  k0 as metrics.py:116-124; reference_sign != 0 contracts the kernel
 *      gradients with grad_potential exactly as the reference does, 0 uses the score -grad_potential (the Stein kernel
 *      whose KSD vanishes for an exact sample).  X, grad_potential: device float (n, d) row-major, d <= 128. */
int mb_ksd(mb_ctx* ctx, const float* X, const float* grad_potential, const float* log_weight /*or NULL*/, int n, int d,
           float bandwidth, int reference_sign, double* out3, mb_stream_t stream);
/* ---- ensemble Kalman filter (EnsembleKalmanFilter, ssm/nonlinear_gaussian.py:279-350; H = I, R = r_std^2 I) on the
 *      row-major population of the Lorenz-96 engine.
 *      mb_rows_mean_cov: ensemble mean[d] and unbiased covariance cov[d][d] (device fp64; either may be NULL) of the
 *      rows (spread_matrix spread_matrix^T, :339) and the Kalman gain K = P (P + r_std^2 I)^-1 (:341-343,
 *      utils.py:477-484) as device float gain[d][d] (fp64 Cholesky on one block).
 *      mb_enkf_analysis: the same, then x_i <- x_i + K (y - x_i - r_std z_i) for every row (:345-348; z_i: Philox
 *      normals of particle gid0 + i, purpose MB_P_SIM = 3, step t), log-weights zero (:350), control block / history
 *      record t set to equal weights (ess = n).  x_rows holds the FORECAST ensemble f(x) + q z, i.e. the output of
 *      mb_pf_l96_step without resampling.  d in {8, 16, 40}. */
int mb_rows_mean_cov(mb_ctx* ctx, const float* x_rows, int64_t n, int d, float r_std, double* mean /*or NULL*/,
                     double* cov /*or NULL*/, float* gain, mb_stream_t stream);
int mb_enkf_analysis(mb_ctx* ctx, const mb_ssm* ssm, float* x_rows, int64_t n, const float* y, float* lw, uint64_t seed,
                     uint32_t t, int64_t gid0, mb_control* ctl, mb_hist* hist /*or NULL*/, float* gain,
                     double* mean /*or NULL*/, double* cov /*or NULL*/, mb_stream_t stream);
/* gather of a row-major (n, d) population by ancestor (cdict.__getitem__, core.py:46-56; resample_particles,
 * filtering.py:202-217).  staged != 0: source rows fetched by the TMA engine -- one cp.async.bulk of the span when the
 * 32 ancestors of a warp are close, else one d*4-byte copy per ancestor -- and the gathered rows leave with one bulk
 * store; staged == 0: per-element loads (comparison path). */
int mb_gather_rows(mb_ctx* ctx, const int32_t* anc, int64_t n_out, int d, const float* src_rows, int64_t n_src,
                   float* dst_rows, int staged, mb_stream_t stream);
/* weighted moments of a row-major population (per-step diagnostics instead of the stacked history, filtering.py:317-322) */
int mb_weighted_moments_rows(mb_ctx* ctx, const float* x_rows, int64_t n, int d, const float* lw,
                             const mb_control* ctl, double* mean, double* var, mb_stream_t stream);
/* the same sums left un-normalised, for one shard of a sharded population (ctl->wmax is the global maximum):
 * sums[0] = sum e_i, sums[1+k] = sum e_i (x_ik - shift_k), sums[1+d+k] = sum e_i (x_ik - shift_k)^2, e_i = exp(lw_i - wmax);
 * the caller adds the ranks' records and normalises (filtering.py:317-322 diagnostics over all GPUs). */
int mb_weighted_moment_sums_rows(mb_ctx* ctx, const float* x_rows, int64_t n, int d, const float* lw,
                                 const mb_control* ctl, const float* shift /*[d]*/, double* sums /*[1+2d]*/,
                                 mb_stream_t stream);

/* ---- K4+K5 fused: systematic resampling without a materialised CDF (transport/smc.py:61-71,
 *      ssm/filtering.py:196-199).  Integer weights e_i = rint(w_i 2^K) (w_i = exp(lw_i - ctl->wmax) in log mode, the
 *      caller's weight <= 1 in linear mode; K = min(40, 63 - ceil(log2 n_total))), exact uint64 cumulative sums C_j,
 *      total S, u0 = k0 / 2^32, and in exact rational arithmetic  a_i = min{ j : (i + u0)/n_total < C_j / S }.
 *      mb_rs_tile_sums: per-4096-particle sums + their exclusive scan into the caller's workspace `ws`
 *      (mb_rs_workspace_bytes(n) bytes; its first 8 bytes are the shard total, the word to exchange between ranks);
 *      mb_rs_ancestors: ancestors (GLOBAL particle ids) of the outputs fed by this shard's particles, written to anc
 *      (single shard) or to sh->anc_peers[owner of the output] (sharded; totals = device [world] uint64 shard totals,
 *      exchanged after every rank's mb_rs_tile_sums).  Source tiles with more than
 *      18432 outputs (collapsed weights) are only RECORDED; mb_rs_heavy, called after a barrier over the ranks, lets
 *      every rank fill its own share of their outputs (a single shard does both in mb_rs_ancestors).  k0 >= 0: caller's
 *      u0 bits; k0 < 0: Philox(ctl->seed, step ctl->iter + 1, P_RESAMPLE).x.  Predicated on ctl->resample unless force. */
size_t mb_rs_workspace_bytes(int64_t n);
int mb_rs_tile_sums(mb_ctx* ctx, void* ws, const float* in, int64_t n, int64_t n_total, int log_mode,
                    const mb_control* ctl, int force, mb_stream_t stream);
int mb_rs_ancestors(mb_ctx* ctx, void* ws, const float* in, int64_t n, int64_t n_total, int log_mode,
                    const mb_control* ctl, int force, int64_t k0, const unsigned long long* totals,
                    const mb_shard* sh /*or NULL*/, int32_t* anc, mb_stream_t stream);
int mb_rs_heavy(mb_ctx* ctx, void* ws, const float* in, int64_t n, int64_t n_total, int log_mode,
                const mb_control* ctl, int force, int64_t k0, const unsigned long long* totals,
                const mb_shard* sh, int32_t* anc, mb_stream_t stream);

/* weighted mean / variance of every column under weights exp(lw - ctl->wmax)/s1  (diagnostics) */
int mb_weighted_moments(mb_ctx* ctx, const float* x, int64_t ld, int64_t n, int d, const float* lw,
                        const mb_control* ctl, double* mean, double* var, mb_stream_t stream);

/* ---- K7 / K1c: SMC-ABC.  mb_quantile: linear-interpolated quantile (jnp.quantile, abc/smc.py:166);
 *      mb_colstats: per-dimension mean and ddof=1 variance (vmap(jnp.cov), abc/smc.py:97);
 *      mb_abc_init / mb_abc_move: abc/smc.py:44-79,210-219 + abc/mcmc.py:56-76 for the g-and-k model;
 *      mb_abc_adapt: abc/smc.py:228-245 (threshold, 0/-inf weights, ess, alpha_mean, RW scale). */
int mb_quantile(mb_ctx* ctx, const float* v, int64_t n, double q, double* out3 /*value, v[lo], v[hi]*/,
                mb_stream_t stream);
int mb_colstats(mb_ctx* ctx, const float* x, int64_t ld, int64_t n, int d, double* mean, double* var,
                mb_stream_t stream);
int mb_abc_init(mb_ctx* ctx, const mb_gk* gk, float* x, int64_t ld, int64_t n, int64_t n_total, int sample_prior,
                float* up, float* dist, float* lw, float* alpha, uint64_t seed, int64_t gid0, mb_control* ctl,
                mb_stream_t stream);
int mb_abc_move(mb_ctx* ctx, const mb_gk* gk, int mcmc_steps, const float* x_in, float* x_out, int64_t ld,
                int64_t n, const int32_t* anc, const float* up_in, float* up_out, const float* dist_in,
                float* dist_out, float* lw, const float* alpha_in, float* alpha_out,
                const float* stepsize /*device [4]*/, uint64_t seed, int64_t gid0, mb_control* ctl,
                mb_stream_t stream);
int mb_abc_adapt(mb_ctx* ctx, const float* x, int64_t ld, int64_t n, int64_t n_total, int d, const float* dist,
                 float* lw, const float* alpha, float* stepsize /*device [d]*/, double ess_retain, double ess_resample,
                 double termination_alpha, int max_iter, const double* schedule, int advance_iter,
                 mb_control* ctl, mb_hist* hist, mb_stream_t stream);
/* SMC-ABC population sharded over GPUs.  mb_abc_move_sharded: anc holds GLOBAL ancestor ids; the
 * value block and the prior-potential / distance / acceptance arrays of every rank (this step's input buffers) are read
 * through peer-mapped pointers ([world] each; entry `rank` is this rank's own input).  mb_abc_adapt_stage: the
 * adaptation of mb_abc_adapt split into stages that work on the local shard and leave small records in `ws`
 * (MB_ABC_WS_BYTES of device memory) for the caller to add over the ranks between stages:
 *   0 begin (-> all-reduce ws+128: 1 + 2d fp64 column sums)       1..3 radix histogram of pass p (-> all-reduce
 *   ws+1024: 2048 u32)   11..13 pick of pass p   4 count / next-larger pass (-> all-reduce SUM of the i64 at ws+640, MIN of
 *   the i64 at ws+648)   5 threshold + weights (-> all-reduce ws+64: 3 u64 counts)   6 control block and step sizes.
 * With a threshold schedule stages 1-4 and 11-13 are skipped. */
#define MB_ABC_WS_BYTES (1024 + 2048 * 4)
int mb_abc_move_sharded(mb_ctx* ctx, const mb_gk* gk, int mcmc_steps, float* x_out, int64_t ld, int64_t n,
                        const int32_t* anc, float* up_out, float* dist_out, float* lw, float* alpha_out,
                        const float* stepsize, uint64_t seed, int rank, int world, const void* const* x_peers,
                        const void* const* up_peers, const void* const* dist_peers, const void* const* alpha_peers,
                        mb_control* ctl, mb_stream_t stream);
int mb_abc_adapt_stage(mb_ctx* ctx, int stage, const float* x, int64_t ld, int64_t n, int64_t n_total, int d,
                       const float* dist, float* lw, const float* alpha, float* stepsize, double ess_retain,
                       double ess_resample, double termination_alpha, int max_iter, const double* schedule,
                       int advance_iter, void* ws, mb_control* ctl, mb_hist* hist, mb_stream_t stream);

/* ---- tempered ensemble Kalman inversion (TemperedEKI / AdaptiveTemperedEKI, transport/teki.py:38-185) on the g-and-k
 *      simulator: x (n, 4) and simulated_data sim (n, m) ROW-MAJOR device floats.  mb_teki is the device-resident record
 *      of the ensemble (what the reference keeps in ensemble_state / extra); every kernel is predicated on done.
 *      mb_teki_init: prior sample (or the caller's x), first simulation, prior_stds (teki.py:94-101).
 *      mb_teki_update: termination_criterion (:104-111) on the current ensemble, then one update (:117-150):
 *      covariances, next temperature (mode 0: schedule[min(iter, len - 1)]; 1: round(2^(iter/50) - 1, 4), :91-92;
 *      2: AdaptiveTemperedEKI.next_temperature, :168-185, through mb_temper_adapt), Kalman gain, perturbed update,
 *      re-simulation.  partials: mb_teki_workspace_doubles(m) doubles; scratch: 2 * roundup(n, 32) floats and search_ctl (mode 2);
 *      temp_hist (or NULL): device doubles [max_iter + 1], temp_hist[iter] = temperature after update iter.
 *      Normals: Philox (gid0 + i, step = iter, purpose 1); simulator draws: purpose 3, step = iter (0 at init). */
#define MB_TEKI_MAX_DY 16
typedef struct {
    double temperature, prev_temperature, alph, ess;
    double cov_x[16], cov_xy[4 * MB_TEKI_MAX_DY], cov_y[MB_TEKI_MAX_DY * MB_TEKI_MAX_DY];  /* row stride 4 / 16 / 16 */
    double cov_y_given_x[MB_TEKI_MAX_DY * MB_TEKI_MAX_DY], chol[MB_TEKI_MAX_DY * MB_TEKI_MAX_DY];
    double prec[MB_TEKI_MAX_DY * MB_TEKI_MAX_DY], gain[4 * MB_TEKI_MAX_DY];
    double stds[4], prior_stds[4];  /* std (ddof 1) of constrain(value) now / of the initial unconstrained value */
    int64_t value_nan, perturb_nan;
    int32_t iter, done, search_iters, pad;
} mb_teki;
typedef struct {
    double max_temperature, nugget, term_std, ess_threshold, tol;
    int32_t max_search_iter, max_iter, mode, schedule_len;
    const double* schedule;   /* device, mode 0 */
} mb_teki_prm;
int mb_teki_workspace_doubles(int m);
int mb_teki_init(mb_ctx* ctx, const mb_gk* gk, float* x, float* sim, int64_t n, int sample_prior, uint64_t seed,
                 int64_t gid0, double* partials, double* temp_hist /*or NULL*/, mb_teki* state, mb_stream_t stream);
int mb_teki_update(mb_ctx* ctx, const mb_gk* gk, const mb_teki_prm* prm, float* x, float* sim, int64_t n, uint64_t seed,
                   int64_t gid0, double* partials, float* scratch, mb_control* search_ctl, double* temp_hist /*or NULL*/,
                   mb_teki* state, mb_stream_t stream);

/* ---- conditional section of a captured step.  Every resampling kernel is predicated on the control block, so
 *      enqueueing them unconditionally is always correct (reference: `cond(resample_bool, ...)`,
 *      transport/smc.py:76-78, ssm/filtering.py:287-293).  While `stream` is under CUDA-graph capture, launches made
 *      between mb_cond_begin and mb_cond_end on *body_stream go into the body of a graph IF node whose condition
 *      (ctl->resample && !ctl->done) is evaluated on the device; outside capture *body_stream = stream. */
int mb_cond_begin(mb_ctx* ctx, const mb_control* ctl, mb_stream_t stream, mb_stream_t* body_stream);
int mb_cond_end(mb_ctx* ctx, mb_stream_t stream);

/* ---- K8-K10: SVGD.  mb_svgd_phi replaces kernelised_grad_matrix (transport/svgd.py:18-32) with the
 *      Gaussian kernel (kernels.py:90-102); X, G, phi are row-major (n x d).  mb_pairdist_bandwidth
 *      replaces median/mean_bandwidth_update (kernels.py:220-229, utils.py:437-439); h on device.
 *      mb_adagrad replaces jax.example_libraries.optimizers.adagrad (svgd.py:104-106,138-139). */
int mb_svgd_phi(mb_ctx* ctx, const float* X, const float* G, int n, int d, const float* bandwidth,
                float* phi, int variant, mb_stream_t stream);
int mb_pairdist_bandwidth(mb_ctx* ctx, const float* X, int n, int d, int mode /*0 median, 1 mean*/,
                          float* h, int variant /*0 exact fp32 SIMT, 1 tcgen05 (bf16 coordinates)*/,
                          mb_stream_t stream);
/* Ensemble sharded over GPUs (tcgen05 variant): X, G hold the WHOLE all-gathered ensemble.
 * mb_svgd_phi_rows writes rows [row_begin, row_begin + row_count) of phi (row_begin a multiple of 128).
 * mb_pairdist_partial accumulates share `share` of `shares` (128-row tiles share, share + shares, ...: interleaved,
 * the symmetric evaluation gives tile row i only T - i tiles of work) into acc (MB_PAIRDIST_ACC_BYTES of device memory):
 * [0] bracket of the median (identical on every rank), [64] u64 entries below it, [72] fp64 sum of distances,
 * [1024] 2048 u32 histogram counters.  After the ranks' counters / sums have been added (any all-reduce),
 * mb_pairdist_finish turns them into the bandwidth h (device float). */
#define MB_PAIRDIST_ACC_BYTES (1024 + 2048 * 4)
int mb_svgd_phi_rows(mb_ctx* ctx, const float* X, const float* G, int n, int d, const float* bandwidth,
                     float* phi, int row_begin, int row_count, mb_stream_t stream);
int mb_pairdist_partial(mb_ctx* ctx, const float* X, int n, int d, int mode, int share, int shares, void* acc,
                        mb_stream_t stream);
int mb_pairdist_finish(mb_ctx* ctx, int mode, int n, const void* acc, float* h, mb_stream_t stream);
/* k(x, y) of one pair (Kernel.__call__, kernels.py:90-95); x, y device float[d], out device float */
int mb_gaussian_kernel(mb_ctx* ctx, const float* x, const float* y, int d, float bandwidth, float* out,
                       mb_stream_t stream);
int mb_adagrad(mb_ctx* ctx, float* X, float* gsq, float* mom, const float* phi, int64_t len, float step,
               float momentum, mb_stream_t stream);
/* x ~ N(prior_mean, prior_std^2 I) at row-major X (n x d), any d: transport/sampler.py:24-30 (vmap(prior_sample));
 * same Philox stream as mb_smc_init (purpose INIT, step 0). */
int mb_prior_sample(mb_ctx* ctx, float prior_mean, float prior_std, int d, int64_t n, uint64_t seed, int64_t gid0,
                    float* X, mb_stream_t stream);
/* Bayesian logistic regression target of config C4 (the reference has no such Scenario): features
 * (N x d), labels (N) on the device, isotropic Gaussian prior; W, G row-major (n x d). */
int mb_logistic_potential_grad(mb_ctx* ctx, const float* features, const float* labels, int N, int d,
                               float prior_mean, float prior_pscale, double beta, const float* W, int n, float* U,
                               float* G, int variant, mb_stream_t stream);
int mb_target_potential_grad(mb_ctx* ctx, const mb_target* tgt, double beta, const float* X /*n x d row-major*/,
                             int n, float* U, float* G, mb_stream_t stream);

/* ---- multi-GPU plumbing (one process per GPU; handles are exchanged by the host) ------------------------
 *      mb_alloc/mb_free: cudaMalloc'd (IPC-shareable) buffers; mb_ipc_*: 64-byte CUDA IPC handles;
 *      mb_comm_*: peer-mapped mailbox communicator used INSIDE kernels for the LSE/ESS allreduce and the
 *      allgather of the ranks' weight totals (the only collectives the path needs). */
int  mb_alloc(mb_ctx* ctx, size_t bytes, void** out);
int  mb_free(mb_ctx* ctx, void* p);
int  mb_ipc_get_handle(mb_ctx* ctx, void* dev_ptr, void* handle64_host);
int  mb_ipc_open(mb_ctx* ctx, const void* handle64_host, void** out);
int  mb_ipc_close(mb_ctx* ctx, void* p);
int  mb_comm_create(mb_ctx* ctx, int rank, int world, mb_comm** out, void* handle64_host);
int  mb_comm_connect(mb_comm* comm, const void* handles_host /*world x 64 bytes*/);
void mb_comm_destroy(mb_comm* comm);
/* ctl != NULL: skipped unless the replicated control block asks for a resampling step.  Doubles as a barrier:
 * every rank's kernels that precede the call in stream order are complete once it returns on any rank. */
int  mb_comm_allgather(mb_comm* comm, const double* in, int nd, double* out, const mb_control* ctl,
                       mb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MOCAT_B200_H */
