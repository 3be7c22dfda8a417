// reduce.cu -- K2 (log-sum-exp / ESS), K3 (on-device adaptive tempering), column statistics,
// weighted moments and the radix-select quantile (K7).
//
// All kernels are HBM-bound streaming reductions: 128-bit loads, fp32 exponentials, fp64 accumulation,
// deterministic two-level reduction (per-block partials merged in fixed order by the last block / by
// every block after a grid sync) -- no floating-point atomics anywhere.
#include <cooperative_groups.h>
#include "common.cuh"
#include "comm.cuh"

const MbCommDev* mb_comm_dev(const mb_comm* c);

namespace cg = cooperative_groups;

#define RED_THREADS 256

// ------------------------------------------------------------------------------------------------
// per-thread online accumulator of (max, sum e^{w-max}, sum e^{2(w-max)})
struct LseAcc {
    float m;
    double s1, s2;
};

__device__ __forceinline__ void acc_rescale(LseAcc& a, float cm) {
    if (cm > a.m) {
        if (a.m != -INFINITY) {
            const double f = exp((double)a.m - (double)cm);
            a.s1 *= f;
            a.s2 *= f * f;
        }
        a.m = cm;
    }
}

__device__ __forceinline__ void acc_add4(LseAcc& a, float w0, float w1, float w2, float w3) {
    acc_rescale(a, fmaxf(fmaxf(w0, w1), fmaxf(w2, w3)));
    const bool any_nan = (w0 != w0) | (w1 != w1) | (w2 != w2) | (w3 != w3);
    if (a.m == -INFINITY && !any_nan) return;
    const float mm = (a.m == -INFINITY) ? 0.f : a.m;
    const float e0 = __expf(w0 - mm), e1 = __expf(w1 - mm), e2 = __expf(w2 - mm), e3 = __expf(w3 - mm);
    a.s1 += (double)((e0 + e1) + (e2 + e3));
    a.s2 += (double)((e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3));
}

__device__ __forceinline__ void acc_add1(LseAcc& a, float w) {
    acc_rescale(a, w);
    if (a.m == -INFINITY && w == w) return;
    const float mm = (a.m == -INFINITY) ? 0.f : a.m;
    const float e = __expf(w - mm);
    a.s1 += (double)e;
    a.s2 += (double)(e * e);
}

// One streaming pass over this block's share of the weights lw - dbeta*lik (lik may be NULL).
// Returns the block-level triple in thread 0.
__device__ __forceinline__ Lse3 lse_block_pass(const float* __restrict__ lw, const float* __restrict__ lik,
                                               float dbeta, int64_t n, Lse3* smem) {
    LseAcc a{-INFINITY, 0.0, 0.0};
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = (((uintptr_t)lw & 15) == 0) && (lik == nullptr || ((uintptr_t)lik & 15) == 0);
    int64_t done = 0;
    if (aligned) {
        const int64_t n4 = n >> 2;
        const float4* lw4 = reinterpret_cast<const float4*>(lw);
        const float4* lk4 = reinterpret_cast<const float4*>(lik);
        for (int64_t i = tid; i < n4; i += nthreads) {
            float4 w = __ldg(lw4 + i);
            if (lik) {
                const float4 l = __ldg(lk4 + i);
                w.x = fmaf(-dbeta, l.x, w.x); w.y = fmaf(-dbeta, l.y, w.y);
                w.z = fmaf(-dbeta, l.z, w.z); w.w = fmaf(-dbeta, l.w, w.w);
            }
            acc_add4(a, w.x, w.y, w.z, w.w);
        }
        done = n4 << 2;
    }
    for (int64_t i = done + tid; i < n; i += nthreads) {
        float w = lw[i];
        if (lik) w = fmaf(-dbeta, lik[i], w);
        acc_add1(a, w);
    }
    return lse3_block_reduce(Lse3{(double)a.m, a.s1, a.s2}, smem);
}

// merge `count` partial triples in a fixed order; result broadcast to all threads of the block
__device__ __forceinline__ Lse3 lse_merge_partials(const double* __restrict__ partials, int count, Lse3* smem) {
    Lse3 v = lse3_empty();
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const Lse3 p{partials[3 * i], partials[3 * i + 1], partials[3 * i + 2]};
        v = lse3_merge(v, p);
    }
    v = lse3_block_reduce(v, smem);
    __shared__ Lse3 bc;
    if (threadIdx.x == 0) bc = v;
    __syncthreads();
    v = bc;
    __syncthreads();
    return v;
}

__device__ __forceinline__ void write_out6(double* out6, const Lse3& r) {
    mb_control tmp;
    ctl_set_weights(&tmp, r);
    out6[0] = tmp.wmax; out6[1] = tmp.s1; out6[2] = tmp.s2;
    out6[3] = tmp.lse;  out6[4] = tmp.lse2; out6[5] = tmp.log_ess;
}

// ------------------------------------------------------------------------------------------------ K2
__global__ void __launch_bounds__(RED_THREADS)
lse_ess_kernel(const float* __restrict__ lw, const float* __restrict__ lik, float dbeta, int64_t n,
               double* partials, uint32_t* counter, double* out6) {
    __shared__ Lse3 smem[RED_THREADS / 32];
    __shared__ bool is_last;
    const Lse3 b = lse_block_pass(lw, lik, dbeta, n, smem);
    if (threadIdx.x == 0) {
        partials[3 * blockIdx.x] = b.m; partials[3 * blockIdx.x + 1] = b.s1; partials[3 * blockIdx.x + 2] = b.s2;
        __threadfence();
        const uint32_t t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const Lse3 r = lse_merge_partials(partials, gridDim.x, smem);
    if (threadIdx.x == 0) {
        write_out6(out6, r);
        *counter = 0;                                    // self-reset for the next launch
    }
}

static int reduce_grid(const mb_ctx* ctx, int64_t n) {
    int64_t blocks = (n + (int64_t)RED_THREADS * 16 - 1) / ((int64_t)RED_THREADS * 16);
    int64_t cap = (int64_t)ctx->sms * 8;
    if (cap > MB_MAX_PARTIAL_BLOCKS) cap = MB_MAX_PARTIAL_BLOCKS;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

extern "C" int mb_lse_ess(mb_ctx* ctx, const float* lw, const float* lik, double dbeta, int64_t n,
                          double* out6, mb_stream_t stream) {
    MB_REQUIRE(ctx && lw && out6 && n >= 0, "mb_lse_ess: bad arguments");
    const int grid = reduce_grid(ctx, n);
    lse_ess_kernel<<<grid, RED_THREADS, 0, mb_s(stream)>>>(lw, lik, (float)dbeta, n, ctx->partials,
                                                          ctx->counters + MB_CNT_REDUCE, out6);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// ------------------------------------------------------------------------------------------------ K3
// Persistent cooperative kernel: the whole regula-falsi search of utils.py:205-237 runs on the device,
// one pass + one grid sync per evaluation; every block merges the same partials in the same order, so
// all blocks take identical decisions without a second sync.  When the population fits (n <= grid x 256
// x 8) each thread keeps its 8 (lw, lik) pairs in REGISTERS for the whole search: the weights are read
// from memory once and written once however many evaluations the search needs.
struct TemperArgs {
    float* lw;
    const float* lik;
    int64_t n;
    mb_temper prm;
    int advance_iter;
    int64_t nan_denominator;
    int64_t n_total;
    mb_control* ctl;
    mb_hist* hist;
    double* partials;      // [2][MB_MAX_PARTIAL_BLOCKS][3]
    MbCommDev comm; int has_comm;
    double* gbuf;          // [2][8] global triple broadcast (sharded)
    unsigned long long* ll;   // resident variant: [2][MB_LL_BLOCKS][8] flag-in-data slots, then the sequence counter
};

#define TP_R 4             // float4 (lw, lik) pairs kept per thread in the resident variant

__device__ __forceinline__ double lse3_log_ess(const Lse3& r) {
    const double mm = (r.m == -INFINITY || r.m == INFINITY) ? 0.0 : r.m;
    return 2.0 * (log(r.s1) + mm) - (log(r.s2) + 2.0 * mm);
}

// two-stage merge of the per-block partials: global max first (cheap), then every partial is rescaled
// ONCE (one fp64 exp each, in parallel) and summed in a fixed order.  Result broadcast to the block.
__device__ __forceinline__ Lse3 lse_merge_partials_fast(const double* __restrict__ part, int count, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double m = -INFINITY;
    for (int i = threadIdx.x; i < count; i += blockDim.x) m = fmax(m, part[3 * i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, shfl_xor_d(m, o));
    if (lane == 0) sm[warp] = m;
    __syncthreads();
    double M = sm[0];
    for (int w = 1; w < nw; ++w) M = fmax(M, sm[w]);
    __syncthreads();
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const double mi = part[3 * i];
        if (mi != -INFINITY) {                          // NaN max propagates through exp()
            const double f = exp(mi - M);
            s1 += part[3 * i + 1] * f;
            s2 += part[3 * i + 2] * f * f;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += shfl_xor_d(s1, o); s2 += shfl_xor_d(s2, o); }
    if (lane == 0) { sm[warp] = s1; sm[nw + warp] = s2; }
    __syncthreads();
    double S1 = 0.0, S2 = 0.0;
    for (int w = 0; w < nw; ++w) { S1 += sm[w]; S2 += sm[nw + w]; }
    __syncthreads();
    return Lse3{M, S1, S2};
}

#define TP_THREADS_RES 512

__device__ __forceinline__ void st_gpu_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// merge of one triple per lane (empty lanes: lse3_empty()); the butterfly is symmetric, so every lane ends with the
// same bits.  NaN maxima propagate through exp().
__device__ __forceinline__ Lse3 lse3_warp_merge(Lse3 v) {
    double M = v.m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmax(M, shfl_xor_d(M, o));
    double s1 = 0.0, s2 = 0.0;
    if (v.m != -INFINITY) { const double f = exp(v.m - M); s1 = v.s1 * f; s2 = v.s2 * f * f; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += shfl_xor_d(s1, o); s2 += shfl_xor_d(s2, o); }
    return Lse3{M, s1, s2};
}

// merge of <= MB_LL_BLOCKS triples held in shared memory with ONE block barrier: every warp finds the global max
// redundantly (no barrier), thread t rescales triple t (one fp64 exp), warp butterflies, then every thread adds the
// (<= 10) warp partials in a fixed order.  `sm` needs 2 * 16 doubles.
__device__ __forceinline__ Lse3 lse_merge_smem_once(const double* sp, int count, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double M = -INFINITY;
    for (int i = lane; i < count; i += 32) M = fmax(M, sp[3 * i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmax(M, shfl_xor_d(M, o));
    const int nwu = (count + 31) >> 5;                  // warps that own triples
    if (warp < nwu) {
        double s1 = 0.0, s2 = 0.0;
        const int i = threadIdx.x;
        if (i < count) {
            const double mi = sp[3 * i];
            if (mi != -INFINITY) { const double f = exp(mi - M); s1 = sp[3 * i + 1] * f; s2 = sp[3 * i + 2] * f * f; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s1 += shfl_xor_d(s1, o); s2 += shfl_xor_d(s2, o); }
        if (lane == 0) { sm[warp] = s1; sm[16 + warp] = s2; }
    }
    __syncthreads();
    double S1 = 0.0, S2 = 0.0;
    for (int w = 0; w < nwu; ++w) { S1 += sm[w]; S2 += sm[16 + w]; }
    return Lse3{M, S1, S2};
}

// phase timing of the resident kernel (debug builds only: nvcc -DMB_TEMPER_TRACE): block 0 accumulates globaltimer
// deltas per phase and reports them through the history record (lse <- phase ns; see scratch/temper_trace.py)
#ifdef MB_TEMPER_TRACE
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TP_TRACE(k) do { const unsigned long long _t = gtimer(); tr_acc[k] += (double)(_t - tr_last); tr_last = _t; } while (0)
#else
#define TP_TRACE(k) do { } while (0)
#endif

template <bool RESIDENT>
__global__ void __launch_bounds__(RESIDENT ? TP_THREADS_RES : RED_THREADS, RESIDENT ? 1 : 2)
temper_adapt_kernel(TemperArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ Lse3 smem[16];
    __shared__ double smd[32];
    __shared__ float smf[16];
    __shared__ double xin[8], xout[5 * MB_MAX_WORLD];
    __shared__ double sparts[RESIDENT ? 3 * MB_LL_BLOCKS : 1];
    unsigned long long lseq = RESIDENT ? a.ll[2 * MB_LL_BLOCKS * 8] : 0ull;   // block 0 writes it back at the end
    unsigned long long xseq = a.has_comm ? *a.comm.seq : 0ull;   // read before the first grid sync; block 0 writes it back at the end
#ifdef MB_TEMPER_TRACE
    double tr_acc[6] = {0, 0, 0, 0, 0, 0};
    unsigned long long tr_last = gtimer();
#endif
    mb_control c0 = *a.ctl;                             // read before the first grid sync (see below)
    if (c0.done) return;                                // uniform over the grid: done only changes at the end
    if (a.advance_iter && c0.resampled) {               // the move kernel resampled: weights were reset to 0,
        const double nd = (double)a.n_total;            // ess to n (transport/smc.py:69-70)
        c0.wmax = 0.0; c0.s1 = nd; c0.s2 = nd;
        c0.lse = log(nd); c0.lse2 = log(nd); c0.log_ess = log(nd); c0.ess = nd;
    }
    const mb_temper& P = a.prm;
    const double beta = c0.beta;
    const int iter_new = c0.iter + (a.advance_iter ? 1 : 0);
    int parity = 0;
    double g_alpha_fx = (double)c0.alpha_fx, g_nan = (double)c0.nan_count;   // global sums when sharded
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = a.n >> 2;

    float4 rw[TP_R], rl[TP_R];
    if (RESIDENT) {                                     // host guarantees: 16 B aligned, n % 4 == 0, n4 <= TP_R * nthreads
#pragma unroll
        for (int r = 0; r < TP_R; ++r) {
            const int64_t i = tid + (int64_t)r * nthreads;
            if (i < n4) {
                rw[r] = reinterpret_cast<const float4*>(a.lw)[i];
                rl[r] = __ldg(reinterpret_cast<const float4*>(a.lik) + i);
            } else {
                rw[r] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                rl[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }

    auto evaluate = [&](double b) -> Lse3 {
        TP_TRACE(0);
        const float dbeta = (float)(b - beta);
        Lse3 blk;
        if (RESIDENT) {
            // registers hold the data: ONE exp per particle relative to the WARP max; the 16 warp triples meet in
            // shared memory (one block barrier) and warp 0 merges them with a butterfly -- only warp 0 needs the
            // block triple, it is the one that publishes it
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            float w[4 * TP_R];
#pragma unroll
            for (int r = 0; r < TP_R; ++r) {
                w[4 * r + 0] = fmaf(-dbeta, rl[r].x, rw[r].x); w[4 * r + 1] = fmaf(-dbeta, rl[r].y, rw[r].y);
                w[4 * r + 2] = fmaf(-dbeta, rl[r].z, rw[r].z); w[4 * r + 3] = fmaf(-dbeta, rl[r].w, rw[r].w);
            }
            float tm = -INFINITY;
#pragma unroll
            for (int k = 0; k < 4 * TP_R; ++k) tm = fmaxf(tm, w[k]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tm = fmaxf(tm, __shfl_xor_sync(MB_FULL, tm, o));
            const float mm = (tm == -INFINITY) ? 0.f : tm;
            float p1 = 0.f, p2 = 0.f;
#pragma unroll
            for (int k = 0; k < 4 * TP_R; ++k) { const float e = __expf(w[k] - mm); p1 += e; p2 = fmaf(e, e, p2); }
            double s1 = (double)p1, s2 = (double)p2;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s1 += shfl_xor_d(s1, o); s2 += shfl_xor_d(s2, o); }
            if (lane == 0) { sparts[3 * warp] = (double)tm; sparts[3 * warp + 1] = s1; sparts[3 * warp + 2] = s2; }
            __syncthreads();
            blk = lse3_empty();
            if (warp == 0) {
                const int nw = blockDim.x >> 5;
                blk = lse3_warp_merge(lane < nw ? Lse3{sparts[3 * lane], sparts[3 * lane + 1], sparts[3 * lane + 2]} : lse3_empty());
            }
            __syncthreads();                            // sparts is reused by the gather below
        } else {
            blk = lse_block_pass(a.lw, a.lik, dbeta, a.n, smem);
        }
        Lse3 r;
        if (RESIDENT) {
            // no grid-wide barrier: every block publishes its triple as six self-validating words {32 data bits,
            // 32-bit sequence flag} and every block gathers all of them (spinning on L2) -- barrier and data
            // exchange are the same memory operation.  Two parities: a slot is rewritten two evaluations later, which
            // its owner cannot reach before every block has gathered this one.  (A two-level gather through group
            // leaders was measured slower: the extra hop costs more than the polling traffic it saves.)
            TP_TRACE(1);
            ++lseq;
            const unsigned long long flag = (lseq & 0xffffffffull) << 32;
            unsigned long long* base = a.ll + (size_t)(lseq & 1ull) * MB_LL_BLOCKS * 8;
            if (threadIdx.x < 6) {
                const double v = threadIdx.x < 2 ? blk.m : (threadIdx.x < 4 ? blk.s1 : blk.s2);
                const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
                const unsigned long long half = (threadIdx.x & 1) ? (bits >> 32) : (bits & 0xffffffffull);
                st_gpu_u64(base + (size_t)blockIdx.x * 8 + threadIdx.x, flag | half);
            }
            for (int idx = threadIdx.x; idx < (int)gridDim.x * 6; idx += blockDim.x) {
                const int b = idx / 6, k = idx - b * 6;
                unsigned long long v;
                do { v = ld_gpu_u64(base + (size_t)b * 8 + k); } while ((v & 0xffffffff00000000ull) != flag);
                reinterpret_cast<unsigned int*>(sparts)[idx] = (unsigned int)(v & 0xffffffffull);
            }
            __syncthreads();
            TP_TRACE(2);
            r = lse_merge_smem_once(sparts, gridDim.x, smd);
            TP_TRACE(3);
        } else {
            double* part = a.partials + (size_t)parity * 3 * MB_MAX_PARTIAL_BLOCKS;
            if (threadIdx.x == 0) {
                part[3 * blockIdx.x] = blk.m; part[3 * blockIdx.x + 1] = blk.s1; part[3 * blockIdx.x + 2] = blk.s2;
            }
            grid.sync();
            r = lse_merge_partials_fast(part, gridDim.x, smd);
        }
        if (a.has_comm) {                                // sharded population: exchange the rank triples over NVLink
            // block 0 publishes, EVERY block receives for itself from the local mailbox: no second grid barrier
            ++xseq;
            if (threadIdx.x == 0) {
                xin[0] = r.m; xin[1] = r.s1; xin[2] = r.s2; xin[3] = (double)c0.alpha_fx; xin[4] = (double)c0.nan_count;
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                comm_exchange_warp(a.comm, xseq, blockIdx.x == 0, xin, 5, xout);
                if (threadIdx.x == 0) {
                    Lse3 g = lse3_empty();
                    double afx = 0.0, nanc = 0.0;
                    for (int q = 0; q < a.comm.world; ++q) {
                        g = lse3_merge(g, Lse3{xout[5 * q], xout[5 * q + 1], xout[5 * q + 2]});
                        afx += xout[5 * q + 3]; nanc += xout[5 * q + 4];
                    }
                    xin[0] = g.m; xin[1] = g.s1; xin[2] = g.s2; xin[3] = afx; xin[4] = nanc;
                }
            }
            __syncthreads();
            r = Lse3{xin[0], xin[1], xin[2]};
            g_alpha_fx = xin[3]; g_nan = xin[4];
            __syncthreads();
        }
        parity ^= 1;
        return r;
    };

    double b_new;
    Lse3 t_new;
    int it = 0;
    if (P.schedule != nullptr) {                         // transport/smc.py:123
        int idx = iter_new < P.schedule_len ? iter_new : P.schedule_len - 1;
        b_new = P.schedule[idx];
        t_new = evaluate(b_new);
    } else {                                             // transport/smc.py:311-326 + utils.py:205-237
        const double log_target = log(c0.ess * P.ess_retain);
        double b0 = beta, b1 = P.max_temperature;
        Lse3 t0{c0.wmax, c0.s1, c0.s2};
        double e0 = c0.log_ess - log_target;
        Lse3 t1 = evaluate(b1);
        double e1 = lse3_log_ess(t1) - log_target;
        const bool increasing = e1 > e0;
        while (!(fmin(fabs(e0), fabs(e1)) < P.tol || it >= P.max_search_iter || (e0 < 0 && e1 < 0) ||
                 (e0 > 0 && e1 > 0) || e0 != e0 || e1 != e1)) {
            const double x = b0 - e0 * (b1 - b0) / (e1 - e0);
            const Lse3 tx = evaluate(x);
            const double ex = lse3_log_ess(tx) - log_target;
            const bool upper = increasing ? (ex > 0) : (ex < 0);
            if (upper) { b1 = x; e1 = ex; t1 = tx; } else { b0 = x; e0 = ex; t0 = tx; }
            ++it;
        }
        // jnp.argmin(jnp.abs(evals)): first index on ties, NaN wins
        const bool pick0 = (e0 != e0) || (!(e1 != e1) && fabs(e0) <= fabs(e1));
        b_new = pick0 ? b0 : b1;
        t_new = pick0 ? t0 : t1;
    }

    // weight update  lw += -(beta' - beta) * lik   (transport/smc.py:201-203, 367-373)
    {
        const float dbeta = (float)(b_new - beta);
        if (dbeta != 0.f) {
            if (RESIDENT) {
#pragma unroll
                for (int r = 0; r < TP_R; ++r) {
                    const int64_t i = tid + (int64_t)r * nthreads;
                    if (i < n4)
                        reinterpret_cast<float4*>(a.lw)[i] =
                            make_float4(fmaf(-dbeta, rl[r].x, rw[r].x), fmaf(-dbeta, rl[r].y, rw[r].y),
                                        fmaf(-dbeta, rl[r].z, rw[r].z), fmaf(-dbeta, rl[r].w, rw[r].w));
                }
            } else {
                const bool aligned = (((uintptr_t)a.lw & 15) == 0) && (((uintptr_t)a.lik & 15) == 0);
                int64_t done = 0;
                if (aligned) {
                    float4* lw4 = reinterpret_cast<float4*>(a.lw);
                    const float4* lk4 = reinterpret_cast<const float4*>(a.lik);
                    for (int64_t i = tid; i < n4; i += nthreads) {
                        float4 w = lw4[i];
                        const float4 l = __ldg(lk4 + i);
                        w.x = fmaf(-dbeta, l.x, w.x); w.y = fmaf(-dbeta, l.y, w.y);
                        w.z = fmaf(-dbeta, l.z, w.z); w.w = fmaf(-dbeta, l.w, w.w);
                        lw4[i] = w;
                    }
                    done = n4 << 2;
                }
                for (int64_t i = done + tid; i < a.n; i += nthreads) a.lw[i] = fmaf(-dbeta, a.lik[i], a.lw[i]);
            }
        }
    }

    TP_TRACE(4);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        mb_control c = *a.ctl;                           // unchanged since the start of the kernel
        c.lse = c0.lse;
        const double lse_prev = c0.lse;
        ctl_set_weights(&c, t_new);
        c.log_z = c.log_z + (c.lse - lse_prev);          // transport/smc.py:212-215
        c.beta = b_new;
        c.iter = iter_new;
        c.search_iters = it;
        c.resample = (c.ess <= P.ess_resample * (double)a.n_total) ? 1 : 0;          // smc.py:298-301
        const double nan_frac = a.nan_denominator > 0 ? g_nan / (double)a.nan_denominator : 0.0;
        c.done = (b_new >= P.max_temperature || iter_new >= P.max_iter || nan_frac > 0.1) ? 1 : 0;  // :171-175
        c.nan_count = 0;
        if (a.advance_iter) c.alpha_mean = g_alpha_fx / 4294967296.0 / (double)a.n_total;
        c.alpha_fx = 0;
        *a.ctl = c;
        if (a.has_comm) *a.comm.seq = xseq;             // every block read it before the first exchange
        if (RESIDENT) a.ll[2 * MB_LL_BLOCKS * 8] = lseq;
        if (a.hist && iter_new < MB_HIST_MAX) {
            mb_hist h;
            h.beta = c.beta; h.ess = c.ess; h.log_z = c.log_z; h.alpha_mean = c.alpha_mean; h.lse = c.lse;
            h.resampled = c.resampled; h.search_iters = it;
#ifdef MB_TEMPER_TRACE
            h.beta = tr_acc[0]; h.ess = tr_acc[1]; h.log_z = tr_acc[2]; h.alpha_mean = tr_acc[3]; h.lse = tr_acc[4];
#endif
            a.hist[iter_new] = h;
        }
    }
}

extern "C" int mb_temper_adapt(mb_ctx* ctx, float* lw, const float* lik, int64_t n, const mb_temper* prm,
                               int advance_iter, int64_t nan_denominator, int64_t n_total, mb_control* ctl,
                               mb_hist* hist, mb_comm* comm, mb_stream_t stream) {
    MB_REQUIRE(ctx && lw && lik && prm && ctl && n > 0, "mb_temper_adapt: bad arguments");
    static int bps[2] = {0, 0};
    if (bps[0] == 0) {
        MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps[0], temper_adapt_kernel<false>, RED_THREADS, 0));
        MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps[1], temper_adapt_kernel<true>, TP_THREADS_RES, 0));
        if (bps[0] < 1 || bps[1] < 1) { mb_set_error("temper kernel cannot be resident"); return MB_ERR_CUDA; }
    }
    const int64_t need_res = (n + (int64_t)TP_THREADS_RES * 4 * TP_R - 1) / ((int64_t)TP_THREADS_RES * 4 * TP_R);
    const int64_t need = (n + (int64_t)RED_THREADS * 4 * TP_R - 1) / ((int64_t)RED_THREADS * 4 * TP_R);
    int64_t cap_res = (int64_t)bps[1] * ctx->sms;
    if (cap_res > MB_LL_BLOCKS) cap_res = MB_LL_BLOCKS;
    const bool resident = (n % 4 == 0) && (((uintptr_t)lw & 15) == 0) && (((uintptr_t)lik & 15) == 0) && need_res <= cap_res;
    int64_t grid = resident ? need_res : (int64_t)bps[0] * ctx->sms;
    if (!resident && grid > need) grid = need;
    if (grid > MB_MAX_PARTIAL_BLOCKS) grid = MB_MAX_PARTIAL_BLOCKS;
    if (grid < 1) grid = 1;
    TemperArgs args{lw, lik, n, *prm, advance_iter, nan_denominator, n_total > 0 ? n_total : n, ctl, hist, ctx->partials};
    args.has_comm = 0;
    args.ll = ctx->ll_slots;
    args.gbuf = ctx->partials + 2 * 3 * MB_MAX_PARTIAL_BLOCKS;          // 64 spare doubles after the partials
    if (comm) { args.comm = *mb_comm_dev(comm); args.has_comm = args.comm.world > 1; }
    void* kargs[] = {&args};
    void* fn = resident ? (void*)temper_adapt_kernel<true> : (void*)temper_adapt_kernel<false>;
    MB_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(resident ? TP_THREADS_RES : RED_THREADS), kargs, 0,
                                        mb_s(stream)));
    return MB_OK;
}

// ------------------------------------------------------------------------------------------------
// column statistics: per-dimension mean and ddof=1 variance over all n particles (abc/smc.py:97),
// or weighted by exp(lw - wmax) when lw != NULL.  One block column-slab per blockIdx.y = column.
__global__ void __launch_bounds__(RED_THREADS)
colstats_kernel(const float* __restrict__ x, int64_t ld, int64_t n, const float* __restrict__ lw,
                const mb_control* ctl, double* partials /*[d][gridDim.x][3]*/) {
    __shared__ double smem[RED_THREADS / 32];
    const int col = blockIdx.y;
    const float* xc = x + (int64_t)col * ld;
    const float shift = xc[0];
    const float wmax = lw ? (float)ctl->wmax : 0.f;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = xc[i] - shift;
        const float w = lw ? __expf(lw[i] - wmax) : 1.f;
        if (w > 0.f) { s0 += w; s1 += (double)w * v; s2 += (double)w * v * v; }
    }
    s0 = block_sum_d(s0, smem); s1 = block_sum_d(s1, smem); s2 = block_sum_d(s2, smem);
    if (threadIdx.x == 0) {
        double* p = partials + ((size_t)col * gridDim.x + blockIdx.x) * 3;
        p[0] = s0; p[1] = s1; p[2] = s2;
    }
}

__global__ void colstats_finish_kernel(const float* __restrict__ x, int64_t ld, const double* partials, int nblocks,
                                       int weighted, double* mean, double* var) {
    const int col = blockIdx.x;
    if (threadIdx.x != 0) return;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int b = 0; b < nblocks; ++b) {
        const double* p = partials + ((size_t)col * nblocks + b) * 3;
        s0 += p[0]; s1 += p[1]; s2 += p[2];
    }
    const double shift = (double)x[(int64_t)col * ld];
    const double m = s1 / s0;
    mean[col] = m + shift;
    if (var) var[col] = weighted ? (s2 / s0 - m * m) : (s2 - s0 * m * m) / (s0 - 1.0);
}

static int colstats_impl(mb_ctx* ctx, const float* x, int64_t ld, int64_t n, int d, const float* lw,
                         const mb_control* ctl, double* mean, double* var, cudaStream_t st) {
    int gx = reduce_grid(ctx, n);
    if (gx > 256) gx = 256;
    const size_t bytes = (size_t)d * gx * 3 * sizeof(double);
    if (mb_ensure_scratch(ctx, bytes) != MB_OK) return MB_ERR_CUDA;
    colstats_kernel<<<dim3(gx, d), RED_THREADS, 0, st>>>(x, ld, n, lw, ctl, (double*)ctx->scratch);
    MB_CHECK_LAUNCH();
    colstats_finish_kernel<<<d, 32, 0, st>>>(x, ld, (const double*)ctx->scratch, gx, lw != nullptr, mean, var);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

extern "C" int mb_colstats(mb_ctx* ctx, const float* x, int64_t ld, int64_t n, int d, double* mean, double* var,
                           mb_stream_t stream) {
    MB_REQUIRE(ctx && x && mean && n > 1 && d > 0, "mb_colstats: bad arguments");
    return colstats_impl(ctx, x, ld, n, d, nullptr, nullptr, mean, var, mb_s(stream));
}

// Robbins-Monro stepsize adaptation of RMMetropolisedSMCSampler.adapt (transport/smc.py:406-421) on the device:
//   alpha_mean = sum_i softmax(log_weight)_i alpha_i ;  log stepsize += rm_stepsize (alpha_mean - target)
// The stepsize lives in ctl->aux1 (the move kernel reads it when mb_move.stepsize <= 0), ctl->aux0 remembers the
// iteration already adapted so that replays after termination change nothing; stepsize_hist[iter] records the chain.
__global__ void rm_adapt_kernel(mb_control* ctl, const double* alpha_mean, double rm_stepsize, double target,
                                double init_stepsize, double* stepsize_hist) {
    if (init_stepsize > 0.0) {
        ctl->aux1 = init_stepsize; ctl->aux0 = (double)ctl->iter;
        if (stepsize_hist) stepsize_hist[ctl->iter] = init_stepsize;
        return;
    }
    const int it = ctl->iter;
    if ((int)ctl->aux0 == it) return;                                  // this iteration has been adapted (run is over)
    const double eps = exp(log(ctl->aux1) + rm_stepsize * (*alpha_mean - target));
    ctl->aux1 = eps; ctl->aux0 = (double)it;
    if (stepsize_hist && it >= 0 && it < MB_HIST_MAX) stepsize_hist[it] = eps;
}

extern "C" int mb_rm_adapt(mb_ctx* ctx, const float* alpha, const float* lw, int64_t n, double rm_stepsize, double target,
                           double init_stepsize, mb_control* ctl, double* stepsize_hist, mb_stream_t stream) {
    MB_REQUIRE(ctx && ctl && (init_stepsize > 0.0 || (alpha && lw && n > 0)), "mb_rm_adapt: bad arguments");
    cudaStream_t st = mb_s(stream);
    double* mean = reinterpret_cast<double*>(ctx->counters + MB_CNT_RM_MEAN);
    if (!(init_stepsize > 0.0)) {
        const int rc = colstats_impl(ctx, alpha, n, n, 1, lw, ctl, mean, nullptr, st);
        if (rc != MB_OK) return rc;
    }
    rm_adapt_kernel<<<1, 1, 0, st>>>(ctl, mean, rm_stepsize, target, init_stepsize, stepsize_hist);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

extern "C" int mb_weighted_moments(mb_ctx* ctx, const float* x, int64_t ld, int64_t n, int d, const float* lw,
                                   const mb_control* ctl, double* mean, double* var, mb_stream_t stream) {
    MB_REQUIRE(ctx && x && lw && ctl && mean && n > 0 && d > 0, "mb_weighted_moments: bad arguments");
    return colstats_impl(ctx, x, ld, n, d, lw, ctl, mean, var, mb_s(stream));
}

// ------------------------------------------------------------------------------------------------ K7
// Exact order statistics by 3-pass radix select on the order-preserving integer image of fp32
// (11 + 11 + 10 bits), then one pass for the next-larger value: linear-interpolated quantile exactly
// as jnp.quantile (abc/smc.py:166) without sorting.  Integer histograms -> deterministic.
#include "select.cuh"

int mb_quantile_impl(mb_ctx* ctx, const float* v, int64_t n, const double* q_dev, double q_host, double* out3,
                     cudaStream_t st) {
    // scratch: SelectState | frac[2] | hist[2048]
    const size_t bytes = 256 + 2048 * sizeof(uint32_t);
    if (mb_ensure_scratch(ctx, (1u << 20) + bytes) != MB_OK) return MB_ERR_CUDA;
    char* base = (char*)ctx->scratch + (1u << 20);       // [0, 1 MiB) is the colstats partial area
    SelectState* state = (SelectState*)base;
    double* frac = (double*)(base + 64);
    uint32_t* hist = (uint32_t*)(base + 256);
    MB_CUDA(cudaMemsetAsync(hist, 0, 2048 * sizeof(uint32_t), st));
    const int grid = reduce_grid(ctx, n);
    select_rank_kernel<<<1, 1, 0, st>>>(state, q_dev, q_host, n, frac);
    select_hist_kernel<21, 11, 0><<<grid, RED_THREADS, 0, st>>>(v, n, state, hist);
    select_pick_kernel<21, 11><<<1, 256, 0, st>>>(state, hist);
    select_hist_kernel<10, 11, 11><<<grid, RED_THREADS, 0, st>>>(v, n, state, hist);
    select_pick_kernel<10, 11><<<1, 256, 0, st>>>(state, hist);
    select_hist_kernel<0, 10, 22><<<grid, RED_THREADS, 0, st>>>(v, n, state, hist);
    select_pick_kernel<0, 10><<<1, 256, 0, st>>>(state, hist);
    select_next_kernel<<<grid, RED_THREADS, 0, st>>>(v, n, state);
    select_finish_dev_kernel<<<1, 1, 0, st>>>(state, frac, out3);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

extern "C" int mb_quantile(mb_ctx* ctx, const float* v, int64_t n, double q, double* out, mb_stream_t stream) {
    MB_REQUIRE(ctx && v && out && n > 0, "mb_quantile: bad arguments");
    return mb_quantile_impl(ctx, v, n, nullptr, q, out, mb_s(stream));
}
