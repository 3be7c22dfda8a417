// ksd.cu -- kernelised Stein discrepancy of a (weighted) sample under the Gaussian kernel (metrics.ksd,
// metrics.py:88-130 with kernels.py:82-116): the n x n contraction
//   KSD^2 = sum_ij k0(x_i, x_j) w_i w_j / (sum w)^2,
//   k0(x, y) = sum_k d2k/dx_k dy_k + grad_x k . g_y + g_x . grad_y k + k g_x . g_y          (metrics.py:116-124)
//            = k [ (d h^2 - r^2) / h^4 + sgn (diff . g_x - diff . g_y) / h^2 + g_x . g_y ],   diff = x - y, r^2 = |diff|^2,
// k = exp(-r^2 / (2 h^2)).  sgn = +1 reproduces the reference literally (it contracts the kernel gradients with
// grad_POTENTIAL); sgn = -1 is the Stein kernel of the score -grad_potential, whose KSD vanishes for an exact sample.
// Exact fp32 SIMT: 64 x 64 pair tiles, operands staged in shared memory, 4 x 4 pairs per thread; k0 is symmetric, so
// only the tiles on or above the diagonal are evaluated (off-diagonal tiles count twice); per-block fp64 partials are
// merged in fixed order.
#include <math.h>
#include "common.cuh"

#define KS_TILE 64
#define KS_THREADS 256

struct KsdArgs {
    const float* X; const float* G; const float* lw; int n, d, ntile;
    float inv_h2, sgn; double* partials;
};

__global__ void __launch_bounds__(KS_THREADS) ksd_kernel(KsdArgs a) {
    extern __shared__ float sm[];                    // xi[d][64] | gi[d][64] | xj[d][64] | gj[d][64] | wi[64] | wj[64]
    const int d = a.d;
    float* xi = sm; float* gi = xi + d * KS_TILE; float* xj = gi + d * KS_TILE; float* gj = xj + d * KS_TILE;
    float* wi = gj + d * KS_TILE; float* wj = wi + KS_TILE;
    __shared__ double red[KS_THREADS / 32];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;          // 4 x 4 pairs: rows 4 ty .. +3, columns 4 tx .. +3
    const int64_t npairs = (int64_t)a.ntile * (a.ntile + 1) / 2;
    const float dh2 = (float)d * a.inv_h2;                           // d / h^2
    double acc = 0.0;
    for (int64_t p = blockIdx.x; p < npairs; p += gridDim.x) {
        // p -> (ti <= tj): row-major enumeration of the upper triangle
        int ti = (int)(((double)(2 * a.ntile + 1) - sqrt((double)(2 * a.ntile + 1) * (2 * a.ntile + 1) - 8.0 * (double)p)) * 0.5);
        while ((int64_t)ti * a.ntile - (int64_t)ti * (ti - 1) / 2 > p) --ti;
        while ((int64_t)(ti + 1) * a.ntile - (int64_t)(ti + 1) * ti / 2 <= p) ++ti;
        const int tj = ti + (int)(p - ((int64_t)ti * a.ntile - (int64_t)ti * (ti - 1) / 2));
        __syncthreads();
        for (int e = threadIdx.x; e < KS_TILE * d; e += KS_THREADS) {
            const int r = e / d, k = e - r * d;
            const int i = ti * KS_TILE + r, j = tj * KS_TILE + r;
            xi[k * KS_TILE + r] = i < a.n ? a.X[(int64_t)i * d + k] : 0.f;
            gi[k * KS_TILE + r] = i < a.n ? a.G[(int64_t)i * d + k] : 0.f;
            xj[k * KS_TILE + r] = j < a.n ? a.X[(int64_t)j * d + k] : 0.f;
            gj[k * KS_TILE + r] = j < a.n ? a.G[(int64_t)j * d + k] : 0.f;
        }
        if (threadIdx.x < KS_TILE) {
            const int i = ti * KS_TILE + threadIdx.x, j = tj * KS_TILE + threadIdx.x;
            wi[threadIdx.x] = i < a.n ? (a.lw ? __expf(a.lw[i]) : 1.f) : 0.f;      // weights = exp(log_weight), :108-111
            wj[threadIdx.x] = j < a.n ? (a.lw ? __expf(a.lw[j]) : 1.f) : 0.f;
        }
        __syncthreads();
        float r2[4][4], ci[4][4], cj[4][4], gg[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) { r2[u][v] = 0.f; ci[u][v] = 0.f; cj[u][v] = 0.f; gg[u][v] = 0.f; }
        for (int k = 0; k < d; ++k) {
            const float4 xa = *reinterpret_cast<const float4*>(xi + k * KS_TILE + 4 * ty);
            const float4 ga = *reinterpret_cast<const float4*>(gi + k * KS_TILE + 4 * ty);
            const float4 xb = *reinterpret_cast<const float4*>(xj + k * KS_TILE + 4 * tx);
            const float4 gb = *reinterpret_cast<const float4*>(gj + k * KS_TILE + 4 * tx);
            const float xav[4] = {xa.x, xa.y, xa.z, xa.w}, gav[4] = {ga.x, ga.y, ga.z, ga.w};
            const float xbv[4] = {xb.x, xb.y, xb.z, xb.w}, gbv[4] = {gb.x, gb.y, gb.z, gb.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float df = xav[u] - xbv[v];
                    r2[u][v] = fmaf(df, df, r2[u][v]);
                    ci[u][v] = fmaf(df, gav[u], ci[u][v]);
                    cj[u][v] = fmaf(df, gbv[v], cj[u][v]);
                    gg[u][v] = fmaf(gav[u], gbv[v], gg[u][v]);
                }
        }
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float kv = __expf(-0.5f * r2[u][v] * a.inv_h2);
                const float k0 = kv * (fmaf(-r2[u][v] * a.inv_h2, a.inv_h2, dh2) + a.sgn * (ci[u][v] - cj[u][v]) * a.inv_h2 + gg[u][v]);
                s = fmaf(k0, wi[4 * ty + u] * wj[4 * tx + v], s);     // padded rows / columns carry weight 0
            }
        acc += (double)s * (ti == tj ? 1.0 : 2.0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(MB_FULL, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < KS_THREADS / 32; ++w) t += red[w];
        a.partials[blockIdx.x] = t;
    }
}

// sum of weights (fixed order over 256 strided partials) and the final sqrt(sum k0 w w) / sum w
__global__ void __launch_bounds__(256) ksd_finish_kernel(const double* partials, int nblocks, const float* lw, int n, double* out) {
    __shared__ double sm[256];
    double s = 0.0, w = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) s += partials[b];
    for (int i = threadIdx.x; i < n; i += 256) w += lw ? (double)__expf(lw[i]) : 1.0;
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    const double tot = sm[0];
    __syncthreads();
    sm[threadIdx.x] = w;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) { out[0] = sqrt(fmax(tot, 0.0)) / sm[0]; out[1] = tot; out[2] = sm[0]; }
}

extern "C" int mb_ksd(mb_ctx* ctx, const float* X, const float* grad_potential, const float* log_weight, int n, int d,
                      float bandwidth, int reference_sign, double* out3, mb_stream_t stream) {
    MB_REQUIRE(ctx && X && grad_potential && out3 && n > 0 && d > 0 && d <= 128 && bandwidth > 0.f,
               "mb_ksd: bad arguments (d <= 128, bandwidth > 0)");
    cudaStream_t st = mb_s(stream);
    const int ntile = (n + KS_TILE - 1) / KS_TILE;
    const int64_t npairs = (int64_t)ntile * (ntile + 1) / 2;
    int grid = (int)(npairs < (int64_t)ctx->sms * 2 ? npairs : (int64_t)ctx->sms * 2);
    if (mb_ensure_scratch(ctx, (size_t)grid * sizeof(double)) != MB_OK) return MB_ERR_CUDA;
    KsdArgs a{X, grad_potential, log_weight, n, d, ntile, 1.f / (bandwidth * bandwidth), reference_sign ? 1.f : -1.f,
              (double*)ctx->scratch};
    const size_t smem = ((size_t)4 * d * KS_TILE + 2 * KS_TILE) * sizeof(float);
    MB_CUDA(cudaFuncSetAttribute(ksd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ksd_kernel<<<grid, KS_THREADS, smem, st>>>(a);
    MB_CHECK_LAUNCH();
    ksd_finish_kernel<<<1, 256, 0, st>>>((const double*)ctx->scratch, grid, log_weight, n, out3);
    MB_CHECK_LAUNCH();
    return MB_OK;
}
