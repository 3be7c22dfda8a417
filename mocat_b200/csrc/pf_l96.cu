// pf_l96.cu -- bootstrap-filter step of the Lorenz-96 state-space model (config C3), the hot kernel of the bench.
//
// Replaces one body of the scan in run_particle_filter_for_marginals (ssm/filtering.py:280-311) for
// Lorenz96(NonLinearGaussian) (ssm/scenarios/lorenz96.py:14-44, ssm/nonlinear_gaussian.py:107-121) with diagonal noise,
// H = I and the fixed-step RK4 flow: ancestor gather (core.py:46-56), transition_sample, log-weight
// increment -likelihood_potential, (max, sum, sumsq) of the weights, log-evidence and the next resample decision.
//
// Why a kernel of its own (round 1 ran this model through the generic one-thread-per-particle kernel at 54 % of the
// HBM roofline, issue bound at 128 registers / 16 warps per SM, ~2100 SASS instructions per particle):
//   * ROW-MAJOR layout, the reference's own (n, d) `value` array: particle i is one contiguous row of D floats
//     (160 B at D = 40).  All traffic between HBM and the SM is done by the TMA engine in whole rows: a window of
//     consecutive rows, or -- when the ancestors of 32 outputs are scattered -- one 160-byte bulk copy per ancestor,
//     local or over NVLink alike; results leave through a bulk store of 16 finished rows.  (The first version of this
//     round used 32-particle AoSoA tiles: a scattered ancestor then costs D sectors of 32 B, 8x its size -- with the
//     collapsed weights of config C3 about 3 % of the outputs descend from such scattered light particles, 25 % of the
//     population read for them, and over NVLink that dominated the sharded step.)
//   * A particle PAIR (2m, 2m+1) is spread over FOUR lanes, D/4 coordinates each: 3*D/4 packed registers of RK4 state
//     per thread instead of 3*D (<= 64 registers, 32 warps per SM).  The cyclic stencil needs three neighbour values per
//     stage; they come from the adjacent lanes by warp shuffle.
//   * The two particles of a pair live in the two halves of 64-bit registers and every RK4 / noise / likelihood
//     operation is a packed fp32x2 instruction (FFMA2 / FADD2: Blackwell), i.e. half the issue slots per flop.
//   * Box-Muller naturally yields two normals per (u1, u2): the cos branch goes to the even particle of the pair, the
//     sin branch to the odd one, so they arrive packed.  Philox counter = (pair id, step, purpose<<20 | coordinate/2):
//     words (0,1) -> coordinate 2c, words (2,3) -> coordinate 2c+1  (mirrored by oracle/philox.py normals_pairwise).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "rng.cuh"
#include "comm.cuh"
#include "pf_common.cuh"

const MbCommDev* mb_comm_dev(const mb_comm* c);

typedef unsigned long long f2;      // two packed fp32: low half = even particle of the pair, high half = odd particle

__device__ __forceinline__ f2 f2_pack(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 f2_splat(float v) { return f2_pack(v, v); }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// Box-Muller pair scaled by `sd` (rng.cuh box_muller folded with the scaling and trimmed for issue slots):
//   (zc, zs) = sd * sqrt(-2 ln u1) * (cos, sin)(2 pi u2),  u1 = u_open(xa) = (float(xa) + .5) 2^-32,  u2 = u24(xb).
// u1 comes out of ONE FFMA (float(xa) 2^-32 + 2^-33 rounds exactly like u_open: scaling by a power of two commutes
// with rounding) and lg2 is taken on u1 in (0, 1], where lg2.approx has an ABSOLUTE error of 2^-22 near 1 -- taking it
// on float(xa) + .5 (result ~ 32, relative error 2^-22) loses the small radii: 7e-5 absolute on the normals.
// sd^2 and -2 ln 2 fold into the multiplier under the square root; the angle is evaluated on phi = 2 pi u2 - pi
// (cos(2 pi u2) = -cos(phi)), the sign is taken by the consumer's FMA.
// returns rq = sd * sqrt(-2 ln u1) and (c, s) = (cos, sin)(phi):  zc = -rq c,  zs = -rq s.
__device__ __forceinline__ void box_muller_scaled(uint32_t xa, uint32_t xb, float k1, float& rq, float& c, float& s) {
    const float u1 = fmaf(__uint2float_rn(xa), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rq) : "f"(l2 * k1));                 // k1 = -2 ln2 sd^2
    const float phi = fmaf((float)(xb >> 8), 3.7450703e-7f, -3.141592653589793f);  // 2 pi 2^-24
    __sincosf(phi, &s, &c);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define L96_THREADS 256
#define L96_WARPS (L96_THREADS / 32)

struct L96Consts {                   // packed constants of the flow: h/2 F, -h/2, -h, -h/6, 2
    f2 hhF, nhh, nhf, nh6, two;
};

// NEGATED slope of coordinate r without the forcing: n = (x[r-2] - x[r+1]) * x[r-1] + x[r] = F - dx_r/dt
// (lorenz96.py:18-19) -- one FADD2 + one FFMA2; indices outside [0, CPL) are the halo values received from the
// neighbouring lanes
#define L96_EXT(arr, r) ((r) == -2 ? hm2 : ((r) == -1 ? hm1 : ((r) == CPL ? hp1 : arr[((r) < 0 || (r) >= CPL) ? 0 : (r)])))
#define L96_NSLOPE(arr, r) f2_fma(f2_sub(L96_EXT(arr, (r) - 2), L96_EXT(arr, (r) + 1)), L96_EXT(arr, (r) - 1), arr[r])
#define L96_HALO(arr)                                                  \
    const f2 hm2 = __shfl_sync(MB_FULL, arr[CPL - 2], prev);           \
    const f2 hm1 = __shfl_sync(MB_FULL, arr[CPL - 1], prev);           \
    const f2 hp1 = __shfl_sync(MB_FULL, arr[0], next);

// one classical RK4 step (device definition of the L96 flow, SURVEY 8c / DESIGN.md).  With k = F - n the forcing leaves
// the slopes: the stage states are  x + c h k = (x + c h F) - c h n  and  x' = (x + h F) - h/6 (n1 + 2 n2 + 2 n3 + n4),
// so x is advanced IN PLACE to x + h/2 F (stage 1) and x + h F (stage 3) and every stage costs 4 packed instructions per
// coordinate (5 in stage 3) instead of 5.  Three arrays of CPL packed registers (state x, accumulator, stage state s);
// the in-place sweeps go upwards and the two old values a later slope still needs ride along in two temporaries.
template <int CPL>
__device__ __forceinline__ void l96_rk4(f2 (&x)[CPL], const L96Consts& c, int prev, int next) {
    f2 acc[CPL], s[CPL];
    {
        const f2 hm2 = __shfl_sync(MB_FULL, x[CPL - 2], prev);
        const f2 hp1 = __shfl_sync(MB_FULL, x[0], next);
        f2 o2 = hm2, o1 = __shfl_sync(MB_FULL, x[CPL - 1], prev);
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            const f2 cur = x[r];
            const f2 nn = f2_fma(f2_sub(o2, r + 1 < CPL ? x[r + 1] : hp1), o1, cur);
            acc[r] = nn;
            x[r] = f2_add(cur, c.hhF);                                 // x + h/2 F
            s[r] = f2_fma(c.nhh, nn, x[r]);
            o2 = o1; o1 = cur;
        }
    }
#pragma unroll
    for (int stage = 0; stage < 2; ++stage) {
        const f2 hm2 = __shfl_sync(MB_FULL, s[CPL - 2], prev);
        const f2 hp1 = __shfl_sync(MB_FULL, s[0], next);
        f2 o2 = hm2, o1 = __shfl_sync(MB_FULL, s[CPL - 1], prev);
#pragma unroll
        for (int r = 0; r < CPL; ++r) {
            const f2 cur = s[r];
            const f2 nn = f2_fma(f2_sub(o2, r + 1 < CPL ? s[r + 1] : hp1), o1, cur);
            acc[r] = f2_fma(c.two, nn, acc[r]);
            if (stage == 1) x[r] = f2_add(x[r], c.hhF);                // x + h F
            s[r] = f2_fma(stage == 0 ? c.nhh : c.nhf, nn, x[r]);
            o2 = o1; o1 = cur;
        }
    }
    {
        L96_HALO(s)
#pragma unroll
        for (int r = 0; r < CPL; ++r) { const f2 nn = L96_NSLOPE(s, r); x[r] = f2_fma(c.nh6, f2_add(acc[r], nn), x[r]); }
    }
}

struct L96Args {
    L96Consts c;                                     // packed constants of the flow (constant bank, not registers)
    f2 nir2, zmean2;                                 // (-1/r_std, -1/r_std), (initial mean, initial mean)
    f2 kgain2;                                       // optimal proposal: Kp sqrt(q^2 + r^2), packed
    float forcing, h, ir, lik_const, zmean, bm_k1;   // zmean: initial mean; bm_k1: folded Box-Muller scale
    float k0;                                        // optimal proposal: initial Kalman gain p0^2 / (p0^2 + r^2)
    int substeps;
    const float* x_in; float* x_out; int64_t n;      // (n, D) row-major
    const int32_t* anc; const float* y; float* lw;
    int64_t gid0;
    const float* x_peers[MB_MAX_WORLD]; int64_t n_local; int world; int sharded; int rank;
    int stagger;                                     // experiment: start-up skew between the warps of an SM sub-partition (cycles)
    PfTail tail;
};

// Staging (north_star (3): TMA-staged ancestor gather).  A warp advances 32 OUTPUT particles per iteration, in two halves
// of 16 (8 particle pairs x 4 lanes).  Their source rows are fetched into the warp's shared-memory window by the TMA
// engine (cp.async.bulk -> SASS UBLKCP, completion counted on the warp's mbarrier; no destination registers, no
// scoreboard), one iteration ahead, so the copy lands while the previous 32 particles are being integrated:
//   WINDOW  the 32 sources span at most L96_ROWS consecutive rows (no resampling; flat weights; a run of outputs that
//           descend from one heavy particle): ONE copy of the span.  A span that is already resident is not fetched
//           again -- with collapsed weights a heavy ancestor is read once per run of outputs, not once per output.
//   ROWS    scattered sources (the light tail of a collapsed population; ancestors on several GPUs): every lane copies
//           its own ancestor's row, D*4 contiguous bytes, from whichever GPU owns it.
// Finished rows are collected in shared memory and leave with one bulk store per 16 particles.
#define L96_ROWS 64                  // default window (rows) of a warp; a template parameter of the kernel

template <int D, int W, int ROWS = L96_ROWS>
struct L96Smem {
    static constexpr int IN_FLOATS = ROWS * D;                     // source window of one warp
    static constexpr int OUT_FLOATS = 16 * D;                          // 16 finished rows of one warp
    static constexpr size_t bytes = (size_t)W * (IN_FLOATS + OUT_FLOATS) * sizeof(float);
};

// OCC: resident blocks per SM the register allocation is tuned for; ROUNDS: Philox rounds (10 = production; the 7-round
// variant exists only to measure how much of the step is RNG, MB_L96_VARIANT=27, and is never the default)
// OPT: the locally optimal proposal of OptimalNonLinearGaussianParticleFilter (ssm/nonlinear_gaussian.py:134-276) for
// H = I and diagonal Q = q^2 I, R = r^2 I, P0 = p0^2 I, where every matrix of its `startup` is a scalar:
//   step  x' = mx + Kp (y - mx) + sd_p z,  Kp = q^2 / (q^2 + r^2),  sd_p^2 = q^2 r^2 / (q^2 + r^2)        (:257-266)
//         log w += log N(y; mx, (q^2 + r^2) I)   -- from the PREDICTION mx, not from the sampled state      (:268-271)
//   init  x0 = m0 + K0 (y0 - m0) + sd_0 z,  K0 = p0^2 / (p0^2 + r^2),  sd_0^2 = 1 / (1/p0^2 + 1/r^2),  log w = 0  (:191-214)
// Here a.ir = 1 / sqrt(q^2 + r^2), a.kgain2 = Kp / a.ir, a.bm_k1 carries sd_p (sd_0), a.lik_const the normaliser.
template <int D, bool INIT, int ROUNDS, int W, int ROWS = L96_ROWS, bool OPT = false>
__device__ __forceinline__ void l96_body(const L96Args& a) {
    static_assert(ROWS >= 32, "ROWS mode needs one slot per lane");
    static_assert(D % 8 == 0, "the lane split needs an even number of coordinates per lane");
    constexpr int CPL = D / 4;                       // coordinates per lane
    mb_control* ctl = a.tail.ctl;
    if (!INIT && ctl->done) return;
    const bool resample = !INIT && ctl->resample != 0;
    extern __shared__ __align__(128) float stage_all[];            // [W][L96_ROWS + 16][D]
    __shared__ __align__(8) unsigned long long bars[W];
    __shared__ Lse3 smem[W];
    __shared__ f2 ysm[D];
    __shared__ const float* peers[MB_MAX_WORLD];
    if (threadIdx.x < D)                             // OPT && INIT: the conditioned initial mean of the coordinate
        ysm[threadIdx.x] = f2_splat((OPT && INIT) ? fmaf(a.k0, a.y[threadIdx.x] - a.zmean, a.zmean) : a.y[threadIdx.x] * a.ir);
    const int own = a.sharded ? a.rank : 0;                         // index of this GPU's own buffer in peers[]
    if (threadIdx.x < MB_MAX_WORLD)
        peers[threadIdx.x] = (a.sharded && (int)threadIdx.x != a.rank) ? a.x_peers[threadIdx.x] : a.x_in;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, p = lane & 3;
    const uint32_t bar_a = smem_u32(&bars[warp]);
    if (!INIT && lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    if (!INIT) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int prev = (lane & ~3) | ((p + 3) & 3), next = (lane & ~3) | ((p + 1) & 3);
    const uint64_t seed = a.tail.seed;
    const L96Consts& c = a.c;
    const f2 nir = a.nir2, zmean = a.zmean2;
    float* const win = stage_all + (size_t)warp * (L96Smem<D, W, ROWS>::IN_FLOATS + L96Smem<D, W, ROWS>::OUT_FLOATS);
    float* const outb = win + L96Smem<D, W, ROWS>::IN_FLOATS;
    const uint32_t win_a = smem_u32(win), outb_a = smem_u32(outb);
    const int64_t gid0 = a.gid0;

    float am = -INFINITY, as1 = 0.f, as2 = 0.f;      // per-thread online (max, sum, sumsq): one weight per lane and group
    const int ntiles = (int)((a.n + 31) >> 5);       // groups of 32 outputs (n < 2^31 per GPU: ancestors are int32)
    const int stride = (int)gridDim.x * W;
    const int n_local = (int)a.n_local;

    // source of the 32 outputs of a group: `src` = this lane's source row (owner-relative), `owner` = the GPU that holds
    // it (per lane); warp-uniform: mode 0 = WINDOW [r0, r0 + nr) of the (common) owner, 2 = the same but already resident
    // (the buffer starts at row r0), 1 = ROWS (lane l's row sits in slot l).  Rows and GPUs are 32-bit quantities
    // (ancestors are int32), so the span of the group is four warp reductions (REDUX), not shuffle trees.
    struct Win { int src, owner, r0, mode, nr; };
    auto describe = [&](int tile) -> Win {
        Win w;
        const int64_t i = (int64_t)tile * 32 + lane;
        int s_ = (int)i, o = own;                    // not resampling / beyond n: the particle's own row
        if (resample && i < a.n) {
            s_ = __ldg(a.anc + i);                   // GLOBAL id of the ancestor
            if (a.sharded) { o = s_ / n_local; s_ -= o * n_local; }    // owner + owner-relative row
        }
        w.src = s_; w.owner = o; w.mode = 0;
        const int lo = __reduce_min_sync(MB_FULL, s_), hi = __reduce_max_sync(MB_FULL, s_);
        w.r0 = lo;
        w.nr = (int)min((unsigned)(hi - lo), (unsigned)ROWS) + 1;
        if (w.nr > ROWS) w.mode = 1;                               // scattered sources
        if (a.sharded && __reduce_min_sync(MB_FULL, o) != __reduce_max_sync(MB_FULL, o)) w.mode = 1;   // several GPUs
        return w;
    };
    int buf_owner = -1, buf_r0 = 0, buf_nr = 0;                        // what the window holds (WINDOW mode)
    auto fetch = [&](Win& w) {
        if (w.mode == 0) {
            if (w.owner == buf_owner && w.r0 >= buf_r0 && w.r0 + w.nr <= buf_r0 + buf_nr) { w.r0 = buf_r0; w.mode = 2; return; }
            buf_owner = w.owner; buf_r0 = w.r0; buf_nr = w.nr;
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)(w.nr * D * 4);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(win_a), "l"(peers[w.owner] + (int64_t)w.r0 * D), "r"(bytes), "r"(bar_a) : "memory");
            }
        } else {
            buf_owner = -1;                                            // the window no longer holds a span
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)(32 * D * 4)) : "memory");
            __syncwarp();
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(win_a + (uint32_t)(lane * D * 4)), "l"(peers[w.owner] + (int64_t)w.src * D), "r"((uint32_t)(D * 4)), "r"(bar_a) : "memory");
        }
    };

    if (!INIT && a.stagger > 0) {                                      // experiment (MB_L96_STAGGER): de-phase the warps
        const long long skew = (long long)((((warp >> 2) & 1) << 1) | (blockIdx.x >= (gridDim.x >> 1) ? 1 : 0)) * a.stagger;
        const long long t0 = clock64();
        while (clock64() - t0 < skew) {}
    }
    int tile = (int)blockIdx.x * W + warp;
    Win cur{}, nxtw{};
    uint32_t phase = 0;
    if (!INIT && tile < ntiles) { cur = describe(tile); fetch(cur); }
    for (; tile < ntiles; tile += stride) {
        const bool more = tile + stride < ntiles;
        // ancestors of the next group: the line is pulled into L1 now and read half a group later, right before the
        // window is refilled (read here, the span reductions waited for the load: 60 % of the long-scoreboard stalls of
        // the kernel sat on the first REDUX, profiles/ncu_c3_r2h.md)
        if (!INIT && more && resample) {
            const int64_t inext = (int64_t)(tile + stride) * 32 + lane;
            if (inext < a.n) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.anc + inext));
        }
        if (!INIT && cur.mode != 2) {
            uint32_t done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
            phase ^= 1;
        }
        float wq = 0.f;                                                // quadratic form of THIS lane's particle of the group
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
            const int64_t iP = (int64_t)tile * 32 + sub * 16 + 2 * g;           // even particle of the pair; the odd one is iP + 1
            f2 x[CPL];
            if (!INIT) {
                const int lP = sub * 16 + 2 * g;
                int rowP = lP, rowQ = lP + 1;                          // ROWS mode: the slot is the lane that fetched it
                if (cur.mode != 1) {
                    rowP = __shfl_sync(MB_FULL, cur.src, lP) - cur.r0;
                    rowQ = __shfl_sync(MB_FULL, cur.src, lP + 1) - cur.r0;
                }
                const float* bP = win + rowP * D + CPL * p;
                const float* bQ = win + rowQ * D + CPL * p;
#pragma unroll
                for (int r = 0; r < CPL; ++r) x[r] = f2_pack(bP[r], bQ[r]);
                if (sub == 1) {                                        // the window is in registers: refill it
                    __syncwarp();
                    if (more) { nxtw = describe(tile + stride); fetch(nxtw); }
                }
                for (int s = 0; s < a.substeps; ++s) l96_rk4<CPL>(x, c, prev, next);
            }
            // process noise (nonlinear_gaussian.py:112-113) / initial sample, and -likelihood_potential (:115-121, H = I)
            const uint64_t pair = (uint64_t)(gid0 + iP) >> 1;
            f2 quad = f2_pack(0.f, 0.f);
#pragma unroll
            for (int r = 0; r < CPL; r += 2) {
                const Philox4 w = philox_raw<ROUNDS>(seed, pair, a.tail.t, INIT ? MB_P_INIT : MB_P_MOVE, (uint32_t)((CPL * p + r) >> 1));
                float rq0, c0, s0, rq1, c1, s1;
                box_muller_scaled(w.x, w.y, a.bm_k1, rq0, c0, s0);
                box_muller_scaled(w.z, w.w, a.bm_k1, rq1, c1, s1);
                f2 b0 = INIT ? (OPT ? ysm[CPL * p + r] : zmean) : x[r], b1 = INIT ? (OPT ? ysm[CPL * p + r + 1] : zmean) : x[r + 1];
                if (OPT && !INIT) {                                    // weight from the prediction; proposal mean mx + Kp (y - mx)
                    const f2 d0 = f2_fma(b0, nir, ysm[CPL * p + r]);
                    const f2 d1 = f2_fma(b1, nir, ysm[CPL * p + r + 1]);
                    quad = f2_fma(d0, d0, quad);
                    quad = f2_fma(d1, d1, quad);
                    b0 = f2_fma(d0, a.kgain2, b0);
                    b1 = f2_fma(d1, a.kgain2, b1);
                }
                float xl, xh;
                f2_unpack(b0, xl, xh);
                x[r] = f2_pack(fmaf(-rq0, c0, xl), fmaf(-rq0, s0, xh));        // cos branch -> even, sin branch -> odd particle
                f2_unpack(b1, xl, xh);
                x[r + 1] = f2_pack(fmaf(-rq1, c1, xl), fmaf(-rq1, s1, xh));
                if (!OPT) {
                    const f2 d0 = f2_fma(x[r], nir, ysm[CPL * p + r]);
                    const f2 d1 = f2_fma(x[r + 1], nir, ysm[CPL * p + r + 1]);
                    quad = f2_fma(d0, d0, quad);
                    quad = f2_fma(d1, d1, quad);
                }
            }
            quad = f2_add(quad, __shfl_xor_sync(MB_FULL, quad, 1));
            quad = f2_add(quad, __shfl_xor_sync(MB_FULL, quad, 2));
            // finished rows -> shared memory -> one bulk store of 16 rows.  The previous store of this buffer (one half
            // group ago) must have READ it before it is overwritten.
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            {
                float* oP = outb + (2 * g) * D + CPL * p;
                float* oQ = oP + D;
#pragma unroll
                for (int r = 0; r < CPL; r += 2) {
                    float l0, h0, l1, h1;
                    f2_unpack(x[r], l0, h0);
                    f2_unpack(x[r + 1], l1, h1);
                    *reinterpret_cast<float2*>(oP + r) = make_float2(l0, l1);
                    *reinterpret_cast<float2*>(oQ + r) = make_float2(h0, h1);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) {
                const int64_t row0 = (int64_t)tile * 32 + sub * 16;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(a.x_out + row0 * D), "r"(outb_a), "r"((uint32_t)(16 * D * 4)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            {                                                          // lane p of the quad keeps particle 16 (p >> 1) + 2 g + (p & 1)
                float qP, qQ;
                f2_unpack(quad, qP, qQ);
                if ((p >> 1) == sub) wq = (p & 1) ? qQ : qP;
            }
        }
        {   // log-weights of the group, one per lane: -likelihood_potential (+ the carried weight, filtering.py:292,303)
            const int64_t iw = (int64_t)tile * 32 + ((p >> 1) << 4) + 2 * g + (p & 1);
            float w = (OPT && INIT) ? 0.f : -fmaf(0.5f, wq, a.lik_const);
            if (!INIT && !resample) w += a.lw[iw];                     // lw is padded to a multiple of 32
            if (iw >= a.n) w = -INFINITY;
            a.lw[iw] = w;
            // branch-free online (max, sum e, sum e^2): rescale by f = exp(old max - new max) (= 1 when unchanged)
            const float amn = fmaxf(am, w);
            const float ref = (amn == -INFINITY) ? 0.f : amn;
            const float f = __expf(am - ref);                          // am = -inf -> 0 (sums are still 0)
            const float e = __expf(w - ref);                           // a NaN weight propagates into the sums
            as1 = fmaf(as1, f, e);
            as2 = fmaf(as2, f * f, e * e);
            am = amn;
        }
        cur = nxtw;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the last rows have left shared memory
    pf_finish<INIT>(a.tail, Lse3{(double)am, (double)as1, (double)as2}, resample, smem);
}

template <int D, bool INIT, int OCC = 2, int ROUNDS = 10, int W = L96_WARPS, bool OPT = false>
__global__ void __launch_bounds__(W * 32, OCC) pf_l96_kernel(const __grid_constant__ L96Args a) { l96_body<D, INIT, ROUNDS, W, L96_ROWS, OPT>(a); }

// experiment (MB_L96_VARIANT=36x): explicit register cap instead of the occupancy hint
template <int D, int NREG, int W, int ROWS>
__global__ void __maxnreg__(NREG) pf_l96_kernel_r(const __grid_constant__ L96Args a) { l96_body<D, false, 10, W, ROWS>(a); }

template <int D, int NREG, int W, int OCC, int ROWS>
static int l96_launch_r(mb_ctx* ctx, const L96Args& a, cudaStream_t st) {
    const size_t smem = L96Smem<D, W, ROWS>::bytes;
    MB_CUDA(cudaFuncSetAttribute(pf_l96_kernel_r<D, NREG, W, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (a.n + 31) >> 5;
    int64_t grid = (ntiles + W - 1) / W;
    if (grid > (int64_t)ctx->sms * OCC) grid = (int64_t)ctx->sms * OCC;
    if (grid > MB_MAX_PARTIAL_BLOCKS) grid = MB_MAX_PARTIAL_BLOCKS;
    pf_l96_kernel_r<D, NREG, W, ROWS><<<(unsigned)grid, W * 32, smem, st>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

template <int D, bool INIT, int OCC, int ROUNDS, int W = L96_WARPS, bool OPT = false>
static int l96_launch(mb_ctx* ctx, const L96Args& a, cudaStream_t st) {
    const size_t smem = L96Smem<D, W>::bytes;
    static bool configured = false;                                    // per instantiation
    if (!configured) {
        MB_CUDA(cudaFuncSetAttribute(pf_l96_kernel<D, INIT, OCC, ROUNDS, W, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int64_t ntiles = (a.n + 31) >> 5;
    int64_t grid = (ntiles + W - 1) / W;
    const int64_t cap = (int64_t)ctx->sms * OCC;                       // persistent: exactly the resident blocks
    if (grid > cap) grid = cap;
    if (grid > MB_MAX_PARTIAL_BLOCKS) grid = MB_MAX_PARTIAL_BLOCKS;
    pf_l96_kernel<D, INIT, OCC, ROUNDS, W, OPT><<<(unsigned)grid, W * 32, smem, st>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

static int l96_dispatch(mb_ctx* ctx, const mb_ssm* ssm, L96Args& a, bool init, cudaStream_t st) {
    a.forcing = ssm->forcing; a.h = ssm->dt / (float)ssm->substeps; a.ir = 1.f / ssm->r_std;
    a.lik_const = ssm->lik_const; a.zmean = ssm->init_mean; a.substeps = ssm->substeps;
    double sd = init ? (double)ssm->init_std : (double)ssm->q_std;             // z * sd = sqrt(-2 ln u1 sd^2) * (cos, sin)
    // the ensemble Kalman filter draws its initial ensemble like the optimal filter (:313-323); its forecast is the bootstrap step
    const bool opt = ssm->proposal == MB_PROPOSAL_OPTIMAL || (init && ssm->proposal == MB_PROPOSAL_ENKF);
    if (opt) {                                                                 // scalars of the optimal proposal (see l96_body)
        const double q2 = (double)ssm->q_std * ssm->q_std, r2 = (double)ssm->r_std * ssm->r_std;
        const double p2 = (double)ssm->init_std * ssm->init_std, v = q2 + r2;
        a.ir = (float)(1.0 / sqrt(v));
        a.kgain2 = 0; { const float kg = (float)(q2 / sqrt(v)); uint32_t b; memcpy(&b, &kg, 4); a.kgain2 = (f2)b | ((f2)b << 32); }
        a.k0 = (float)(p2 / (p2 + r2));
        a.lik_const = (float)(0.5 * ssm->dim * log(2.0 * 3.14159265358979323846 * v));
        sd = init ? sqrt(1.0 / (1.0 / p2 + 1.0 / r2)) : sqrt(q2 * r2 / v);
    }
    a.bm_k1 = (float)(-2.0 * 0.6931471805599453 * sd * sd);
    auto splat = [](float v) { uint32_t b; memcpy(&b, &v, 4); return (f2)b | ((f2)b << 32); };
    a.c.hhF = splat(0.5f * a.h * a.forcing); a.c.nhh = splat(-0.5f * a.h); a.c.nhf = splat(-a.h);
    a.c.nh6 = splat(-a.h * (1.f / 6.f)); a.c.two = splat(2.f);
    a.nir2 = splat(-a.ir); a.zmean2 = splat(a.zmean);
    a.tail.partials = ctx->partials;
    a.tail.counter = ctx->counters + MB_CNT_MOVE;
    static int variant = -1, stagger = 0;                              // experiment switches (scratch/l96_variants.sh)
    if (variant < 0) {
        const char* v = getenv("MB_L96_VARIANT"); variant = v ? atoi(v) : 20;
        const char* g = getenv("MB_L96_STAGGER"); stagger = g ? atoi(g) : 0;
    }
    a.stagger = stagger;
    if (!init && !opt && ssm->dim == 40 && variant != 20) {
        if (variant == 10) return l96_launch<40, false, 1, 10>(ctx, a, st);
        if (variant == 27) return l96_launch<40, false, 2, 7>(ctx, a, st);
        if (variant == 28) return l96_launch<40, false, 2, 10, 8>(ctx, a, st);     // two blocks of 8 warps
        if (variant == 26) return l96_launch_r<40, 168, 6, 2, 64>(ctx, a, st);     // 12 warps
        if (variant == 36) return l96_launch_r<40, 112, 6, 3, 48>(ctx, a, st);     // 18 warps
        if (variant == 45) return l96_launch_r<40, 104, 5, 4, 40>(ctx, a, st);     // 20 warps
        if (variant == 37) return l96_launch_r<40, 96, 7, 3, 40>(ctx, a, st);      // 21 warps
        if (variant == 38) return l96_launch_r<40, 80, 8, 3, 40>(ctx, a, st);      // 24 warps

        mb_set_error("pf_l96: unknown MB_L96_VARIANT %d", variant);
        return MB_ERR_ARG;
    }
#define L96_CASE(DD)                                                                                   \
    if (ssm->dim == DD && opt)                                                                         \
        return init ? l96_launch<DD, true, 2, 10, L96_WARPS, true>(ctx, a, st) : l96_launch<DD, false, 2, 10, L96_WARPS, true>(ctx, a, st); \
    if (ssm->dim == DD) return init ? l96_launch<DD, true, 2, 10>(ctx, a, st) : l96_launch<DD, false, 2, 10>(ctx, a, st);
    // d = 40 (config C3): ONE block of 16 warps per SM -- the 16 resident warps of an SM then work on 16 consecutive
    // groups (80 KB of contiguous rows); measured 7.55 ms against 7.86 ms with two blocks of 8 warps at n = 1e8
    if (!init && !opt && ssm->dim == 40) return l96_launch<40, false, 1, 10, 16>(ctx, a, st);
    L96_CASE(8) L96_CASE(16) L96_CASE(40)
    mb_set_error("pf_l96: unsupported dimension %d (compiled: 8, 16, 40; no CPU fallback)", ssm->dim);
    return MB_ERR_UNSUPPORTED;
}

extern "C" int mb_pf_l96_init(mb_ctx* ctx, const mb_ssm* ssm, float* x, int64_t n, int64_t n_total, const float* y0,
                              float* lw, uint64_t seed, int64_t gid0, double ess_threshold, mb_control* ctl,
                              mb_hist* hist, mb_comm* comm, mb_stream_t stream) {
    MB_REQUIRE(ctx && ssm && x && y0 && lw && ctl && n > 0, "mb_pf_l96_init: bad arguments");
    MB_REQUIRE(ssm->kind == MB_SSM_LORENZ96 && ssm->dim_obs == ssm->dim, "mb_pf_l96_init: Lorenz-96 with H = I only");
    MB_REQUIRE((gid0 & 1) == 0, "mb_pf_l96_init: gid0 must be even (Philox streams are keyed on particle pairs)");
    L96Args a{};
    a.x_in = x; a.x_out = x; a.n = n; a.y = y0; a.lw = lw; a.gid0 = gid0;
    a.tail.n_total = n_total; a.tail.t = 0; a.tail.ess_threshold = ess_threshold; a.tail.seed = seed;
    a.tail.ctl = ctl; a.tail.hist = hist;
    if (comm) { a.tail.comm = *mb_comm_dev(comm); a.tail.has_comm = a.tail.comm.world > 1; }
    return l96_dispatch(ctx, ssm, a, true, mb_s(stream));
}

extern "C" int mb_pf_l96_step(mb_ctx* ctx, const mb_ssm* ssm, const float* x_in, float* x_out, int64_t n,
                              int64_t n_total, const int32_t* anc, const float* y, float* lw, uint64_t seed, uint32_t t,
                              int64_t gid0, double ess_threshold, mb_control* ctl, mb_hist* hist, const mb_shard* sh,
                              mb_comm* comm, mb_stream_t stream) {
    MB_REQUIRE(ctx && ssm && x_in && x_out && anc && y && lw && ctl && n > 0 && x_in != x_out,
               "mb_pf_l96_step: bad arguments");
    MB_REQUIRE(ssm->kind == MB_SSM_LORENZ96 && ssm->dim_obs == ssm->dim && ssm->substeps >= 1,
               "mb_pf_l96_step: Lorenz-96 with H = I only");
    MB_REQUIRE((gid0 & 1) == 0, "mb_pf_l96_step: gid0 must be even (Philox streams are keyed on particle pairs)");
    L96Args a{};
    a.x_in = x_in; a.x_out = x_out; a.n = n; a.anc = anc; a.y = y; a.lw = lw; a.gid0 = gid0;
    a.tail.n_total = n_total; a.tail.t = t; a.tail.ess_threshold = ess_threshold; a.tail.seed = seed;
    a.tail.ctl = ctl; a.tail.hist = hist;
    if (sh && sh->world > 1) {
        a.sharded = 1; a.n_local = sh->n_local; a.world = sh->world; a.rank = sh->rank;
        for (int r = 0; r < sh->world; ++r) a.x_peers[r] = sh->x_peers[r];
    }
    if (comm) { a.tail.comm = *mb_comm_dev(comm); a.tail.has_comm = a.tail.comm.world > 1; }
    return l96_dispatch(ctx, ssm, a, false, mb_s(stream));
}

// ------------------------------------------------------------------------------------------------
// Weighted mean / variance of every coordinate of a ROW-MAJOR (n, d) population under weights exp(lw - ctl->wmax) / s1
// (diagnostics of the filter: the per-step moments returned instead of the reference's stacked (T, n, d) history,
// ssm/filtering.py:317-322).  One warp per 32 particles, one lane per coordinate (two when D > 32); particles whose
// weight underflows fp32 relative to the maximum (exp(lw - wmax) < 2^-50) are skipped together with their row,
// so a collapsed population costs 4 B per particle instead of 4 (D + 1).
#define TM_THREADS 256
#define TM_UNROLL 8
__global__ void __launch_bounds__(TM_THREADS)
rows_moments_kernel(const float* __restrict__ x, int64_t n, int d, const float* __restrict__ lw, const mb_control* ctl,
                    const float* __restrict__ shift, double* partials /*[gridDim.x][1 + 2 d]*/) {
    extern __shared__ double sm[];                   // [warps][1 + 2 d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float wmax = (float)ctl->wmax;
    const int c0 = lane, c1 = lane + 32;
    // shifts (numerical conditioning of the second moment): the caller's, or particle 0 of this population
    const float sh0 = (c0 < d) ? (shift ? shift[c0] : x[c0]) : 0.f, sh1 = (c1 < d) ? (shift ? shift[c1] : x[c1]) : 0.f;
    double s0 = 0.0, a0 = 0.0, b0 = 0.0, a1 = 0.0, b1 = 0.0;
    const int64_t ngroups = (n + 31) >> 5;
    // TM_UNROLL groups per trip: their weight loads are issued back to back (a collapsed population is nothing but this
    // 4-byte stream, and one 128-byte load in flight per warp left the kernel latency bound at 10 % of the HBM peak);
    // the groups are then consumed in the same order as before, so the sums are bit-identical to the one-group loop
    const int64_t gstride = (int64_t)gridDim.x * nw;
    for (int64_t grp0 = (int64_t)blockIdx.x * nw + warp; grp0 < ngroups; grp0 += gstride * TM_UNROLL) {
        float ev[TM_UNROLL];
#pragma unroll
        for (int u = 0; u < TM_UNROLL; ++u) {
            const int64_t grp = grp0 + (int64_t)u * gstride;
            const int64_t i = grp * 32 + lane;
            ev[u] = (grp < ngroups && i < n) ? __ldcs(lw + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < TM_UNROLL; ++u) {
            const int64_t grp = grp0 + (int64_t)u * gstride;
            const float dl = ev[u] - wmax;
            const bool in_range = grp < ngroups && grp * 32 + lane < n;
            const float e = (in_range && (dl > -34.6f || dl != dl)) ? __expf(dl) : 0.f;
            unsigned mask = __ballot_sync(MB_FULL, e != 0.f);
            while (mask) {                                             // (rows of four particles in flight at once: slower,
                const int j = __ffs(mask) - 1;                         //  0.64 vs 0.54 ms at n = 1e8 -- 80 registers)
                mask &= mask - 1;
                const double ej = (double)__shfl_sync(MB_FULL, e, j);
                const float* row = x + (grp * 32 + j) * d;
                s0 += ej;
                if (c0 < d) { const double v = (double)(row[c0] - sh0); a0 += ej * v; b0 += ej * v * v; }
                if (c1 < d) { const double v = (double)(row[c1] - sh1); a1 += ej * v; b1 += ej * v * v; }
            }
        }
    }
    double* mine = sm + (size_t)warp * (1 + 2 * d);
    if (lane == 0) mine[0] = s0;
    if (c0 < d) { mine[1 + c0] = a0; mine[1 + d + c0] = b0; }
    if (c1 < d) { mine[1 + c1] = a1; mine[1 + d + c1] = b1; }
    __syncthreads();
    for (int k = threadIdx.x; k < 1 + 2 * d; k += blockDim.x) {
        double acc = 0.0;
        for (int w = 0; w < nw; ++w) acc += sm[(size_t)w * (1 + 2 * d) + k];
        partials[(size_t)blockIdx.x * (1 + 2 * d) + k] = acc;
    }
}

__global__ void rows_moments_finish_kernel(const float* __restrict__ x, int d, const double* partials, int nblocks,
                                           double* mean, double* var) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= d) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int b = 0; b < nblocks; ++b) {
        const double* p = partials + (size_t)b * (1 + 2 * d);
        s0 += p[0]; s1 += p[1 + col]; s2 += p[1 + d + col];
    }
    const double m = s1 / s0;
    mean[col] = m + (double)x[col];
    if (var) var[col] = s2 / s0 - m * m;
}

// raw sums of one shard: sums[0] = sum e, sums[1 + k] = sum e (x_k - shift_k), sums[1 + d + k] = sum e (x_k - shift_k)^2
// with e = exp(lw - ctl->wmax) (ctl->wmax is the GLOBAL maximum of a sharded population), blocks merged in fixed order
__global__ void rows_moment_sums_finish_kernel(int d, const double* partials, int nblocks, double* sums) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 1 + 2 * d) return;
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += partials[(size_t)b * (1 + 2 * d) + k];
    sums[k] = acc;
}

static int rows_moments_launch(mb_ctx* ctx, const float* x, int64_t n, int d, const float* lw, const mb_control* ctl,
                               const float* shift, cudaStream_t st, int64_t* grid_out) {
    const int64_t ngroups = (n + 31) >> 5;
    int64_t grid = (ngroups + (TM_THREADS / 32) - 1) / (TM_THREADS / 32);
    if (grid > (int64_t)ctx->sms * 8) grid = (int64_t)ctx->sms * 8;
    const size_t bytes = (size_t)grid * (1 + 2 * d) * sizeof(double);
    if (mb_ensure_scratch(ctx, bytes) != MB_OK) return MB_ERR_CUDA;
    rows_moments_kernel<<<(unsigned)grid, TM_THREADS, (TM_THREADS / 32) * (1 + 2 * d) * sizeof(double), st>>>(
        x, n, d, lw, ctl, shift, (double*)ctx->scratch);
    MB_CHECK_LAUNCH();
    *grid_out = grid;
    return MB_OK;
}

extern "C" int mb_weighted_moment_sums_rows(mb_ctx* ctx, const float* x, int64_t n, int d, const float* lw,
                                            const mb_control* ctl, const float* shift, double* sums, mb_stream_t stream) {
    MB_REQUIRE(ctx && x && lw && ctl && shift && sums && n > 0 && d > 0 && d <= 64,
               "mb_weighted_moment_sums_rows: bad arguments (d <= 64, shift required)");
    int64_t grid;
    cudaStream_t st = mb_s(stream);
    int rc = rows_moments_launch(ctx, x, n, d, lw, ctl, shift, st, &grid);
    if (rc != MB_OK) return rc;
    rows_moment_sums_finish_kernel<<<(1 + 2 * d + 63) / 64, 64, 0, st>>>(d, (const double*)ctx->scratch, (int)grid, sums);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

extern "C" int mb_weighted_moments_rows(mb_ctx* ctx, const float* x, int64_t n, int d, const float* lw,
                                        const mb_control* ctl, double* mean, double* var, mb_stream_t stream) {
    MB_REQUIRE(ctx && x && lw && ctl && mean && n > 0 && d > 0 && d <= 64, "mb_weighted_moments_rows: bad arguments (d <= 64)");
    int64_t grid;
    cudaStream_t st = mb_s(stream);
    int rc = rows_moments_launch(ctx, x, n, d, lw, ctl, nullptr, st, &grid);
    if (rc != MB_OK) return rc;
    rows_moments_finish_kernel<<<(d + 63) / 64, 64, 0, st>>>(x, d, (const double*)ctx->scratch, (int)grid, mean, var);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// ------------------------------------------------------------------------------------------------
// Gather of a ROW-MAJOR population by ancestor (cdict.__getitem__, core.py:46-56, as used by resample_particles,
// ssm/filtering.py:202-217): x_out[i, :] = x_in[anc[i], :].  One warp per 32 outputs, the same two TMA modes as the
// step kernel: one bulk copy of the span of source rows when it is short (sorted ancestors), else one D*4-byte bulk
// copy per ancestor; every lane then bulk-stores its row from the window.  staged == 0: plain per-element loads and stores
// (comparison path of the tests and of scratch/c3_bench.py).
#define GR_WARPS 4
#define GR_ROWS 64

template <int D>
__global__ void __launch_bounds__(GR_WARPS * 32)
gather_rows_kernel(const int32_t* __restrict__ anc, int64_t n_out, const float* __restrict__ src, int64_t n_src,
                   float* __restrict__ dst, int staged) {
    extern __shared__ __align__(128) float stage[];                  // [GR_WARPS][GR_ROWS][D]
    __shared__ __align__(8) unsigned long long bar[GR_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* win = stage + (size_t)warp * GR_ROWS * D;
    const uint32_t bar_a = smem_u32(&bar[warp]), win_a = smem_u32(win);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t phase = 0;
    const int64_t ngroups = (n_out + 31) >> 5;
    for (int64_t grp = (int64_t)blockIdx.x * GR_WARPS + warp; grp < ngroups; grp += (int64_t)gridDim.x * GR_WARPS) {
        const int64_t i = grp * 32 + lane;
        const int64_t a = (i < n_out) ? (int64_t)anc[i] : -1;
        if (!staged) {
            if (a >= 0) for (int k = 0; k < D; ++k) dst[i * D + k] = __ldg(src + a * D + k);
            continue;
        }
        int64_t lo = (a < 0) ? INT64_MAX : a, hi = a;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, (int64_t)__shfl_xor_sync(MB_FULL, lo, o));
            hi = max(hi, (int64_t)__shfl_xor_sync(MB_FULL, hi, o));
        }
        if (hi < 0) continue;
        const bool window = hi - lo < GR_ROWS;
        const int nvalid = __popc(__ballot_sync(MB_FULL, a >= 0));
        if (window) {
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)((hi - lo + 1) * D * 4);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(win_a), "l"(src + lo * D), "r"(bytes), "r"(bar_a) : "memory");
            }
        } else {
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)(nvalid * D * 4)) : "memory");
            __syncwarp();
            if (a >= 0)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(win_a + (uint32_t)(lane * D * 4)), "l"(src + a * D), "r"((uint32_t)(D * 4)), "r"(bar_a) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
        phase ^= 1;
        // every lane stores its own ancestor's row straight from the window (no shuffle through registers)
        if (a >= 0) {
            const uint32_t row_a = win_a + (uint32_t)((window ? (int)(a - lo) : lane) * D * 4);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(dst + i * D), "r"(row_a), "r"((uint32_t)(D * 4)) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the window is refilled by the next group
        __syncwarp();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

extern "C" int mb_gather_rows(mb_ctx* ctx, const int32_t* anc, int64_t n_out, int d, const float* src_rows,
                              int64_t n_src, float* dst_rows, int staged, mb_stream_t stream) {
    MB_REQUIRE(ctx && anc && src_rows && dst_rows && n_out > 0 && n_src > 0 && src_rows != dst_rows,
               "mb_gather_rows: bad arguments");
    const int64_t ngroups = (n_out + 31) >> 5;
    int64_t grid = (ngroups + GR_WARPS - 1) / GR_WARPS;
    if (grid > (int64_t)ctx->sms * 4) grid = (int64_t)ctx->sms * 4;
    cudaStream_t st = mb_s(stream);
#define GR_CASE(DD)                                                                                            \
    if (d == DD) {                                                                                             \
        const size_t smem = (size_t)GR_WARPS * GR_ROWS * DD * sizeof(float);                                   \
        MB_CUDA(cudaFuncSetAttribute(gather_rows_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gather_rows_kernel<DD><<<(unsigned)grid, GR_WARPS * 32, smem, st>>>(anc, n_out, src_rows, n_src, dst_rows, staged); \
        MB_CHECK_LAUNCH();                                                                                     \
        return MB_OK;                                                                                          \
    }
    GR_CASE(8) GR_CASE(16) GR_CASE(40)
    mb_set_error("mb_gather_rows: unsupported dimension %d (compiled: 8, 16, 40)", d);
    return MB_ERR_UNSUPPORTED;
}
