// svgd.cu -- K8 (SVGD interaction), K9 (pairwise-distance bandwidth), K10 (adagrad).
//
//   phi_i = (1/n) sum_j [ -k(x_j, x_i) g_j + grad_x k(x_j, x_i) ]          transport/svgd.py:18-32
//         = [ -K G + (X o rowsum(K) - K X) / h^2 ]_i / n,  K_ij = exp(-|x_i - x_j|^2 / (2 h^2))  kernels.py:90-102
// i.e. an attention-shaped contraction with Q = K = X, V = [G, X, 1] and an un-normalised exp.
// The n x n matrix is never materialised: j-tiles stream through shared memory.
//
// variant 0 (this file): fp32 SIMT tiles -- the reference-precision implementation every other variant
// is validated against.  variant 1 (svgd_tc.cu): tcgen05 tensor-core kernel.
#include "select.cuh"

#define SV_THREADS 128
#define SV_TI 32          // rows i per block
#define SV_TJ 64          // columns j per tile

template <int DP> struct SvCfg {
    static constexpr int VP = ((2 * DP + 1 + 15) / 16) * 16;   // row of V = [G (DP) | X (DP) | 1 | pad]
    static constexpr int CW = VP / 4;                           // output columns per warp-quarter
};

// Stage one j-tile of V = [G | X | 1] (zero padded) and |x_j|^2 into shared memory.
template <int DP>
__device__ __forceinline__ void sv_load_tile(const float* __restrict__ X, const float* __restrict__ G, int n, int d,
                                             int j0, float* Vj, float* sqj, bool with_g) {
    constexpr int VP = SvCfg<DP>::VP;
    for (int idx = threadIdx.x; idx < SV_TJ * VP; idx += SV_THREADS) {
        const int j = idx / VP, c = idx - j * VP;
        const int gj = j0 + j;
        float v = 0.f;
        if (gj < n) {
            if (c < DP) { if (with_g && c < d) v = G[(int64_t)gj * d + c]; }
            else if (c < 2 * DP) { if (c - DP < d) v = X[(int64_t)gj * d + (c - DP)]; }
            else if (c == 2 * DP) v = 1.f;
        }
        Vj[idx] = v;
    }
    __syncthreads();
    if (threadIdx.x < SV_TJ) {
        const float* xr = Vj + threadIdx.x * VP + DP;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < DP; ++k) s = fmaf(xr[k], xr[k], s);
        sqj[threadIdx.x] = s;
    }
    __syncthreads();
}

// squared distances of this thread's row i against 16 columns of the staged tile (warp-uniform j ->
// shared-memory broadcast reads); D^2 = |x_i|^2 + |x_j|^2 - 2 x_i.x_j, clamped at 0.
template <int DP>
__device__ __forceinline__ void sv_dist16(const float (&xi)[DP], float sqi, const float* Vj, const float* sqj, int jb,
                                          float (&d2)[16]) {
    constexpr int VP = SvCfg<DP>::VP;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
        const float4* xr = reinterpret_cast<const float4*>(Vj + (jb + jj) * VP + DP);
        float s = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < DP / 4; ++k4) {
            const float4 v = xr[k4];
            s = fmaf(xi[4 * k4 + 0], v.x, s); s = fmaf(xi[4 * k4 + 1], v.y, s);
            s = fmaf(xi[4 * k4 + 2], v.z, s); s = fmaf(xi[4 * k4 + 3], v.w, s);
        }
        d2[jj] = fabsf(fmaxf(sqi + sqj[jb + jj] - 2.f * s, 0.f));   // fabsf: never -0.0 (keys are compared as bits)
    }
}

template <int DP>
__device__ __forceinline__ float sv_load_row(const float* __restrict__ X, int n, int d, int gi, float (&xi)[DP]) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < DP; ++k) {
        xi[k] = (gi < n && k < d) ? X[(int64_t)gi * d + k] : 0.f;
        s = fmaf(xi[k], xi[k], s);
    }
    return s;
}

// ------------------------------------------------------------------------------------------------ K8 (fp32)
template <int DP>
__global__ void __launch_bounds__(SV_THREADS)
svgd_phi_simt_kernel(const float* __restrict__ X, const float* __restrict__ G, int n, int d,
                     const float* __restrict__ bandwidth, float* __restrict__ phi) {
    constexpr int VP = SvCfg<DP>::VP, CW = SvCfg<DP>::CW;
    extern __shared__ __align__(16) float smem[];
    float* Vj = smem;                               // [SV_TJ][VP]
    float* Ks = Vj + SV_TJ * VP;                    // [SV_TI][SV_TJ + 1]
    float* sqj = Ks + SV_TI * (SV_TJ + 1);          // [SV_TJ]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * SV_TI;
    const int gi = i0 + lane;
    const float h = bandwidth[0];
    const float inv2h2 = 0.5f / (h * h);

    float xi[DP];
    const float sqi = sv_load_row<DP>(X, n, d, gi, xi);
    float acc[CW];
#pragma unroll
    for (int c = 0; c < CW; ++c) acc[c] = 0.f;

    for (int j0 = 0; j0 < n; j0 += SV_TJ) {
        sv_load_tile<DP>(X, G, n, d, j0, Vj, sqj, true);
        // phase 1: K_ij for row i = lane, columns [16 warp, 16 warp + 16)
        float d2[16];
        sv_dist16<DP>(xi, sqi, Vj, sqj, warp * 16, d2);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int j = warp * 16 + jj;
            Ks[lane * (SV_TJ + 1) + j] = (j0 + j < n) ? __expf(-d2[jj] * inv2h2) : 0.f;
        }
        __syncthreads();
        // phase 2: O[i][c] += sum_j K_ij V_j[c], row i = lane, columns [CW warp, CW warp + CW)
#pragma unroll 4
        for (int j = 0; j < SV_TJ; ++j) {
            const float kij = Ks[lane * (SV_TJ + 1) + j];
            const float4* vr = reinterpret_cast<const float4*>(Vj + j * VP + warp * CW);
#pragma unroll
            for (int c4 = 0; c4 < CW / 4; ++c4) {
                const float4 v = vr[c4];
                acc[4 * c4 + 0] = fmaf(kij, v.x, acc[4 * c4 + 0]); acc[4 * c4 + 1] = fmaf(kij, v.y, acc[4 * c4 + 1]);
                acc[4 * c4 + 2] = fmaf(kij, v.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(kij, v.w, acc[4 * c4 + 3]);
            }
        }
        __syncthreads();
    }
    // epilogue: gather the four column quarters of each row through shared memory
    float* Os = Vj;                                 // [SV_TI][VP]
#pragma unroll
    for (int c = 0; c < CW; ++c) Os[lane * VP + warp * CW + c] = acc[c];
    __syncthreads();
    const float invh2 = 1.f / (h * h), invn = 1.f / (float)n;
    for (int idx = threadIdx.x; idx < SV_TI * d; idx += SV_THREADS) {
        const int r = idx / d, k = idx - r * d;
        if (i0 + r < n) {
            const float* o = Os + r * VP;
            const float xik = X[(int64_t)(i0 + r) * d + k];
            phi[(int64_t)(i0 + r) * d + k] = (-o[k] + (xik * o[2 * DP] - o[DP + k]) * invh2) * invn;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K9
// mode 1 (mean): sum of all n^2 distances in fp64, deterministic two-level reduction.
// mode 0 (median): 3 radix-select passes over the n^2 squared distances (recomputed, never stored) + 1 pass
// for the next-larger value; median of the even count = mean of the two middle distances (np.median).
template <int DP, int MODE /*0 sum, 1 hist, 2 next*/, int SHIFT, int BITS, int HIGH_BITS>
__global__ void __launch_bounds__(SV_THREADS)
pairdist_kernel(const float* __restrict__ X, int n, int d, double* partials, SelectState* st, uint32_t* hist) {
    constexpr int VP = SvCfg<DP>::VP;
    extern __shared__ __align__(16) float smem[];
    float* Vj = smem;
    float* sqj = Vj + SV_TJ * VP;
    uint32_t* sh = reinterpret_cast<uint32_t*>(sqj + SV_TJ);       // [1 << BITS] (MODE 1)
    __shared__ double red[SV_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gi = blockIdx.x * SV_TI + lane;
    float xi[DP];
    const float sqi = sv_load_row<DP>(X, n, d, gi, xi);
    if (MODE == 1) {
        for (int i = threadIdx.x; i < (1 << BITS); i += SV_THREADS) sh[i] = 0;
    }
    const uint32_t prefix = (MODE == 1) ? st->prefix : 0u;
    const uint32_t sel = (MODE == 2) ? st->sel_key : 0u;
    double sum = 0.0;
    unsigned long long cnt = 0;
    uint32_t mg = 0xffffffffu;
    for (int j0 = 0; j0 < n; j0 += SV_TJ) {
        sv_load_tile<DP>(X, nullptr, n, d, j0, Vj, sqj, false);
        float d2[16];
        sv_dist16<DP>(xi, sqi, Vj, sqj, warp * 16, d2);
        if (gi < n) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                if (j0 + warp * 16 + jj < n) {
                    if (MODE == 0) sum += (double)sqrtf(d2[jj]);
                    else {
                        const uint32_t k = __float_as_uint(d2[jj]);     // non-negative floats order like their bits
                        if (MODE == 1) {
                            bool match = true;
                            if (HIGH_BITS > 0) match = (k >> (32 - HIGH_BITS)) == (prefix >> (32 - HIGH_BITS));
                            if (match) atomicAdd(&sh[(k >> SHIFT) & ((1u << BITS) - 1)], 1u);
                        } else {
                            if (k <= sel) ++cnt; else mg = min(mg, k);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (MODE == 0) {
        sum = block_sum_d(sum, red);
        if (threadIdx.x == 0) partials[blockIdx.x] = sum;
    } else if (MODE == 1) {
        __syncthreads();
        for (int i = threadIdx.x; i < (1 << BITS); i += SV_THREADS)
            if (sh[i]) atomicAdd(&hist[i], sh[i]);
    } else {
        for (int o = 16; o > 0; o >>= 1) {
            cnt += __shfl_down_sync(MB_FULL, cnt, o);
            mg = min(mg, __shfl_down_sync(MB_FULL, mg, o));
        }
        if (lane == 0) {
            atomicAdd((unsigned long long*)&st->count_le, cnt);
            atomicMin(&st->min_gt_key, mg);
        }
    }
}

__global__ void pairdist_mean_finish(const double* partials, int nblocks, int n, float* h) {
    if (threadIdx.x || blockIdx.x) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[b];
    h[0] = (float)(s / ((double)n * (double)n) / sqrt(2.0 * log((double)n)));       // kernels.py:227-229
}

__global__ void pairdist_median_rank(SelectState* st, int n, double* frac) {
    const double N2 = (double)n * (double)n;
    const double pos = 0.5 * (N2 - 1.0);
    const double lo = floor(pos);
    st->prefix = 0; st->rank = (int64_t)lo; st->count_le = 0; st->min_gt_key = 0xffffffffu; st->sel_key = 0;
    frac[0] = pos - lo; frac[1] = lo;
}

__global__ void pairdist_median_finish(const SelectState* st, const double* frac, int n, float* h) {
    const double vlo = sqrt((double)__uint_as_float(st->sel_key));
    double vhi = vlo;
    if (frac[0] > 0.0 && st->count_le == (int64_t)frac[1] + 1 && st->min_gt_key != 0xffffffffu)
        vhi = sqrt((double)__uint_as_float(st->min_gt_key));
    const double med = vlo * (1.0 - frac[0]) + vhi * frac[0];
    h[0] = (float)(med / sqrt(2.0 * log((double)n)));                                // kernels.py:220-224
}

// pick kernel for non-transformed keys (distances are >= 0 so the raw bits are already ordered)
template <int SHIFT, int BITS>
__global__ void pairdist_pick_kernel(SelectState* st, uint32_t* hist) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t r = st->rank;
        uint32_t b = 0;
        for (; b < (1u << BITS); ++b) {
            const int64_t c = hist[b];
            if (r < c) break;
            r -= c;
        }
        if (b == (1u << BITS)) b = (1u << BITS) - 1;
        st->rank = r;
        st->prefix |= (b << SHIFT);
        if (SHIFT == 0) { st->sel_key = st->prefix; st->count_le = 0; st->min_gt_key = 0xffffffffu; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (1 << BITS); i += blockDim.x) hist[i] = 0;
}

template <int DP>
static int svgd_launch_dp(mb_ctx* ctx, int what, const float* X, const float* G, int n, int d, const float* bw,
                          float* out, cudaStream_t st) {
    constexpr int VP = SvCfg<DP>::VP;
    const int grid = (n + SV_TI - 1) / SV_TI;
    if (what == 0) {                                 // phi
        const size_t sm = sizeof(float) * (SV_TJ * VP + SV_TI * (SV_TJ + 1) + SV_TJ);
        MB_CUDA(cudaFuncSetAttribute(svgd_phi_simt_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        svgd_phi_simt_kernel<DP><<<grid, SV_THREADS, sm, st>>>(X, G, n, d, bw, out);
        MB_CHECK_LAUNCH();
        return MB_OK;
    }
    const size_t sm_base = sizeof(float) * (SV_TJ * VP + SV_TJ);
    if (mb_ensure_scratch(ctx, (2u << 20) + sizeof(double) * grid) != MB_OK) return MB_ERR_CUDA;
    char* base = (char*)ctx->scratch + (1u << 20);
    SelectState* state = (SelectState*)base;
    double* frac = (double*)(base + 64);
    uint32_t* hist = (uint32_t*)(base + 256);
    double* partials = (double*)((char*)ctx->scratch + (2u << 20));
    if (what == 2) {                                 // mean
        pairdist_kernel<DP, 0, 0, 1, 0><<<grid, SV_THREADS, sm_base + 8, st>>>(X, n, d, partials, state, hist);
        MB_CHECK_LAUNCH();
        pairdist_mean_finish<<<1, 1, 0, st>>>(partials, grid, n, out);
        MB_CHECK_LAUNCH();
        return MB_OK;
    }
    // median
    MB_CUDA(cudaMemsetAsync(hist, 0, 2048 * sizeof(uint32_t), st));
    pairdist_median_rank<<<1, 1, 0, st>>>(state, n, frac);
    const size_t sm_h = sm_base + 2048 * sizeof(uint32_t);
    MB_CUDA(cudaFuncSetAttribute(pairdist_kernel<DP, 1, 21, 11, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_h));
    MB_CUDA(cudaFuncSetAttribute(pairdist_kernel<DP, 1, 10, 11, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_h));
    MB_CUDA(cudaFuncSetAttribute(pairdist_kernel<DP, 1, 0, 10, 22>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_h));
    pairdist_kernel<DP, 1, 21, 11, 0><<<grid, SV_THREADS, sm_h, st>>>(X, n, d, partials, state, hist);
    pairdist_pick_kernel<21, 11><<<1, 256, 0, st>>>(state, hist);
    pairdist_kernel<DP, 1, 10, 11, 11><<<grid, SV_THREADS, sm_h, st>>>(X, n, d, partials, state, hist);
    pairdist_pick_kernel<10, 11><<<1, 256, 0, st>>>(state, hist);
    pairdist_kernel<DP, 1, 0, 10, 22><<<grid, SV_THREADS, sm_h, st>>>(X, n, d, partials, state, hist);
    pairdist_pick_kernel<0, 10><<<1, 256, 0, st>>>(state, hist);
    pairdist_kernel<DP, 2, 0, 1, 0><<<grid, SV_THREADS, sm_base + 8, st>>>(X, n, d, partials, state, hist);
    pairdist_median_finish<<<1, 1, 0, st>>>(state, frac, n, out);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

static int svgd_dispatch(mb_ctx* ctx, int what, const float* X, const float* G, int n, int d, const float* bw,
                         float* out, cudaStream_t st) {
    if (d <= 4) return svgd_launch_dp<4>(ctx, what, X, G, n, d, bw, out, st);
    if (d <= 8) return svgd_launch_dp<8>(ctx, what, X, G, n, d, bw, out, st);
    if (d <= 16) return svgd_launch_dp<16>(ctx, what, X, G, n, d, bw, out, st);
    if (d <= 32) return svgd_launch_dp<32>(ctx, what, X, G, n, d, bw, out, st);
    if (d <= 52) return svgd_launch_dp<52>(ctx, what, X, G, n, d, bw, out, st);
    if (d <= 64) return svgd_launch_dp<64>(ctx, what, X, G, n, d, bw, out, st);
    mb_set_error("svgd: dim %d > 64 not built", d);
    return MB_ERR_UNSUPPORTED;
}

int mb_svgd_phi_tc(mb_ctx* ctx, const float* X, const float* G, int n, int d, const float* bandwidth, float* phi,
                   int row_begin, int row_count, cudaStream_t st);

extern "C" int mb_svgd_phi(mb_ctx* ctx, const float* X, const float* G, int n, int d, const float* bandwidth,
                           float* phi, int variant, mb_stream_t stream) {
    MB_REQUIRE(ctx && X && G && bandwidth && phi && n > 0 && d > 0, "mb_svgd_phi: bad arguments");
    if (variant == 0) return svgd_dispatch(ctx, 0, X, G, n, d, bandwidth, phi, mb_s(stream));
    if (variant == 1) return mb_svgd_phi_tc(ctx, X, G, n, d, bandwidth, phi, 0, n, mb_s(stream));
    mb_set_error("mb_svgd_phi: variant %d not built", variant);
    return MB_ERR_UNSUPPORTED;
}

// rows [row_begin, row_begin + row_count) of phi only: one rank's share of an ensemble sharded over GPUs (X and G are
// the whole, all-gathered ensemble; tensor-core variant, row_begin a multiple of 128)
extern "C" int mb_svgd_phi_rows(mb_ctx* ctx, const float* X, const float* G, int n, int d, const float* bandwidth,
                                float* phi, int row_begin, int row_count, mb_stream_t stream) {
    MB_REQUIRE(ctx && X && G && bandwidth && phi && n > 0 && d > 0, "mb_svgd_phi_rows: bad arguments");
    return mb_svgd_phi_tc(ctx, X, G, n, d, bandwidth, phi, row_begin, row_count, mb_s(stream));
}

int mb_pairdist_bandwidth_tc(mb_ctx* ctx, const float* X, int n, int d, int mode, float* h, cudaStream_t st);
int mb_pairdist_partial_tc(mb_ctx* ctx, const float* X, int n, int d, int mode, int itile_first, int itile_step, void* acc,
                           cudaStream_t st);
int mb_pairdist_finish_tc(mb_ctx* ctx, int mode, int n, const void* acc, float* h, cudaStream_t st);

extern "C" int mb_pairdist_partial(mb_ctx* ctx, const float* X, int n, int d, int mode, int share, int shares, void* acc,
                                   mb_stream_t stream) {
    MB_REQUIRE(ctx && X && acc && n > 1 && d > 0 && (mode == 0 || mode == 1) && shares >= 1 && share >= 0 && share < shares,
               "mb_pairdist_partial: bad arguments");
    return mb_pairdist_partial_tc(ctx, X, n, d, mode, share, shares, acc, mb_s(stream));
}

extern "C" int mb_pairdist_finish(mb_ctx* ctx, int mode, int n, const void* acc, float* h, mb_stream_t stream) {
    MB_REQUIRE(ctx && acc && h && n > 1 && (mode == 0 || mode == 1), "mb_pairdist_finish: bad arguments");
    return mb_pairdist_finish_tc(ctx, mode, n, acc, h, mb_s(stream));
}

extern "C" int mb_pairdist_bandwidth(mb_ctx* ctx, const float* X, int n, int d, int mode, float* h, int variant,
                                     mb_stream_t stream) {
    MB_REQUIRE(ctx && X && h && n > 1 && d > 0 && (mode == 0 || mode == 1), "mb_pairdist_bandwidth: bad arguments");
    if (variant == 1) return mb_pairdist_bandwidth_tc(ctx, X, n, d, mode, h, mb_s(stream));
    MB_REQUIRE(variant == 0, "mb_pairdist_bandwidth: variant not built");
    return svgd_dispatch(ctx, mode == 0 ? 1 : 2, X, nullptr, n, d, nullptr, h, mb_s(stream));
}

// Gaussian kernel value k(x, y) = exp(-|x - y|^2 / (2 h^2)) of ONE pair (kernels.py:90-95, the scalar Kernel.__call__ of
// the reference API); fp32 differences, fp64 accumulation, one warp
__global__ void gaussian_pair_kernel(const float* __restrict__ x, const float* __restrict__ y, int d, float h, float* out) {
    double s = 0.0;
    for (int k = threadIdx.x; k < d; k += 32) { const float df = (x[k] - y[k]) / h; s += (double)df * (double)df; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[0] = (float)exp(-0.5 * s);
}

extern "C" int mb_gaussian_kernel(mb_ctx* ctx, const float* x, const float* y, int d, float bandwidth, float* out,
                                  mb_stream_t stream) {
    MB_REQUIRE(ctx && x && y && out && d > 0 && bandwidth > 0.f, "mb_gaussian_kernel: bad arguments");
    gaussian_pair_kernel<<<1, 32, 0, mb_s(stream)>>>(x, y, d, bandwidth, out);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// ------------------------------------------------------------------------------------------------ K10
// jax.example_libraries.optimizers.adagrad, called with g = -phi (transport/svgd.py:138)
__global__ void adagrad_kernel(float* __restrict__ X, float* __restrict__ gsq, float* __restrict__ mom,
                               const float* __restrict__ phi, int64_t len, float step, float momentum) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = -phi[i];
        const float s = fmaf(g, g, gsq[i]);
        gsq[i] = s;
        const float inv = s > 0.f ? rsqrtf(s) : 0.f;
        const float m = fmaf(1.f - momentum, g * inv, momentum * mom[i]);
        mom[i] = m;
        X[i] = fmaf(-step, m, X[i]);
    }
}

extern "C" int mb_adagrad(mb_ctx* ctx, float* X, float* gsq, float* mom, const float* phi, int64_t len, float step,
                          float momentum, mb_stream_t stream) {
    MB_REQUIRE(ctx && X && gsq && mom && phi && len > 0, "mb_adagrad: bad arguments");
    int64_t grid = (len + 255) / 256;
    if (grid > (int64_t)ctx->sms * 16) grid = (int64_t)ctx->sms * 16;
    adagrad_kernel<<<(unsigned)grid, 256, 0, mb_s(stream)>>>(X, gsq, mom, phi, len, step, momentum);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// ------------------------------------------------------------------------------------------------
// Bayesian logistic regression (config C4; no such Scenario exists upstream): labels t in {0,1},
// features A (N x d), prior N(mean, 1/pscale^2 I):
//   U(w) = U_prior(w) + beta * sum_k [ softplus(a_k.w) - t_k a_k.w ],   grad = grad_prior + beta * A^T (sigma(A w) - t).
// variant 0: exact fp32, one thread per particle, data rows broadcast from shared memory.
#define LR_THREADS 128
#define LR_TILE 64
template <int DP>
__global__ void __launch_bounds__(LR_THREADS)
logistic_pg_simt_kernel(const float* __restrict__ A, const float* __restrict__ t, int N, int d, float prior_mean,
                        float prior_pscale, float beta, const float* __restrict__ W, int n, float* U, float* G) {
    __shared__ __align__(16) float sa[LR_TILE * DP];
    __shared__ float st[LR_TILE];
    const int i = blockIdx.x * LR_THREADS + threadIdx.x;
    float w[DP], g[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) { w[k] = (i < n && k < d) ? W[(int64_t)i * d + k] : 0.f; g[k] = 0.f; }
    float ul = 0.f;
    for (int j0 = 0; j0 < N; j0 += LR_TILE) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < LR_TILE * DP; idx += LR_THREADS) {
            const int j = idx / DP, k = idx - j * DP;
            sa[idx] = (j0 + j < N && k < d) ? A[(int64_t)(j0 + j) * d + k] : 0.f;
        }
        if (threadIdx.x < LR_TILE) st[threadIdx.x] = (j0 + threadIdx.x < N) ? t[j0 + threadIdx.x] : 0.f;
        __syncthreads();
        const int jn = min(LR_TILE, N - j0);
        for (int j = 0; j < jn; ++j) {
            const float4* ar = reinterpret_cast<const float4*>(sa + j * DP);
            float s = 0.f;
#pragma unroll
            for (int k4 = 0; k4 < DP / 4; ++k4) {
                const float4 v = ar[k4];
                s = fmaf(w[4 * k4], v.x, s); s = fmaf(w[4 * k4 + 1], v.y, s);
                s = fmaf(w[4 * k4 + 2], v.z, s); s = fmaf(w[4 * k4 + 3], v.w, s);
            }
            const float e = __expf(-fabsf(s));                        // stable softplus / sigmoid
            const float inv = 1.f / (1.f + e);
            const float sig = s >= 0.f ? inv : e * inv;
            ul += fmaxf(s, 0.f) + log1pf(e) - st[j] * s;
            const float r = sig - st[j];
#pragma unroll
            for (int k4 = 0; k4 < DP / 4; ++k4) {
                const float4 v = ar[k4];
                g[4 * k4] = fmaf(r, v.x, g[4 * k4]); g[4 * k4 + 1] = fmaf(r, v.y, g[4 * k4 + 1]);
                g[4 * k4 + 2] = fmaf(r, v.z, g[4 * k4 + 2]); g[4 * k4 + 3] = fmaf(r, v.w, g[4 * k4 + 3]);
            }
        }
    }
    if (i >= n) return;
    float up = 0.f;
    for (int k = 0; k < d; ++k) {
        const float rr = (w[k] - prior_mean) * prior_pscale;
        up = fmaf(0.5f * rr, rr, up);
        G[(int64_t)i * d + k] = fmaf(beta, g[k], rr * prior_pscale);
    }
    if (U) U[i] = fmaf(beta, ul, up);
}

int mb_logistic_potential_grad_tc(mb_ctx* ctx, const float* features, const float* labels, int N, int d, float prior_mean,
                                  float prior_pscale, float beta, const float* W, int n, float* U, float* G, cudaStream_t st);

extern "C" int mb_logistic_potential_grad(mb_ctx* ctx, const float* features, const float* labels, int N, int d,
                                          float prior_mean, float prior_pscale, double beta, const float* W, int n,
                                          float* U, float* G, int variant, mb_stream_t stream) {
    MB_REQUIRE(ctx && features && labels && W && G && N > 0 && d > 0 && n > 0, "mb_logistic_potential_grad: bad arguments");
    if (variant == 1)
        return mb_logistic_potential_grad_tc(ctx, features, labels, N, d, prior_mean, prior_pscale, (float)beta, W, n, U, G,
                                             mb_s(stream));
    MB_REQUIRE(variant == 0, "mb_logistic_potential_grad: variant not built");
    const int grid = (n + LR_THREADS - 1) / LR_THREADS;
    cudaStream_t st = mb_s(stream);
#define LR_CASE(DPV) logistic_pg_simt_kernel<DPV><<<grid, LR_THREADS, 0, st>>>(features, labels, N, d, prior_mean, prior_pscale, (float)beta, W, n, U, G)
    if (d <= 4) LR_CASE(4); else if (d <= 8) LR_CASE(8); else if (d <= 16) LR_CASE(16); else if (d <= 32) LR_CASE(32);
    else if (d <= 52) LR_CASE(52); else if (d <= 64) LR_CASE(64);
    else { mb_set_error("logistic regression: dim %d > 64 not built", d); return MB_ERR_UNSUPPORTED; }
    MB_CHECK_LAUNCH();
    return MB_OK;
}
