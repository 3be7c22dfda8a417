// pf_common.cuh -- tail shared by the bootstrap-filter step kernels (propagate.cu, pf_l96.cu): grid-wide merge of the
// per-block (max, sum, sumsq) triples, optional cross-GPU exchange of the rank triples, and the control-block /
// history update of one filter step (ssm/filtering.py:287-311; log-evidence convention of transport/smc.py:160,212-215).
#pragma once
#include "common.cuh"
#include "comm.cuh"

struct PfTail {
    int64_t n_total;
    uint32_t t;
    double ess_threshold;
    uint64_t seed;
    mb_control* ctl; mb_hist* hist;
    double* partials; uint32_t* counter;
    MbCommDev comm; int has_comm;
};

// Called by EVERY thread of every block with the thread's accumulated triple.  smem: >= blockDim/32 Lse3.
template <bool INIT>
__device__ __forceinline__ void pf_finish(const PfTail& a, Lse3 mine, bool resample, Lse3* smem) {
    __shared__ bool is_last;
    const Lse3 b = lse3_block_reduce(mine, smem);
    if (threadIdx.x == 0) {
        a.partials[3 * blockIdx.x] = b.m; a.partials[3 * blockIdx.x + 1] = b.s1; a.partials[3 * blockIdx.x + 2] = b.s2;
        __threadfence();
        is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    Lse3 v = lse3_empty();
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)
        v = lse3_merge(v, Lse3{a.partials[3 * i], a.partials[3 * i + 1], a.partials[3 * i + 2]});
    v = lse3_block_reduce(v, smem);
    if (a.has_comm) {                                                  // global LSE/ESS: exchange the rank triples
        __shared__ double xin[3], xout[3 * MB_MAX_WORLD];
        if (threadIdx.x == 0) { xin[0] = v.m; xin[1] = v.s1; xin[2] = v.s2; }
        __syncthreads();
        if (threadIdx.x < 32) comm_allgather_warp(a.comm, xin, 3, xout);
        __syncthreads();
        if (threadIdx.x == 0) {
            v = lse3_empty();
            for (int r = 0; r < a.comm.world; ++r) v = lse3_merge(v, Lse3{xout[3 * r], xout[3 * r + 1], xout[3 * r + 2]});
        }
    }
    if (threadIdx.x == 0) {
        *a.counter = 0;
        mb_control c;
        if (INIT) { memset(&c, 0, sizeof(c)); c.seed = a.seed; } else c = *a.ctl;
        const double nd = (double)a.n_total;
        const double lse_prev = (INIT || resample) ? log(nd) : c.lse;          // log Z convention, SURVEY 8c
        ctl_set_weights(&c, v);
        c.log_z = (INIT ? 0.0 : c.log_z) + (c.lse - lse_prev);
        c.iter = (int32_t)a.t;
        c.resampled = resample ? 1 : 0;
        c.resample = (c.ess < a.ess_threshold * nd) ? 1 : 0;                   // filtering.py:287 (strict <)
        c.done = 0;
        *a.ctl = c;
        if (a.hist && a.t < MB_HIST_MAX) {
            mb_hist h;
            h.beta = 0.0; h.ess = c.ess; h.log_z = c.log_z; h.alpha_mean = 0.0; h.lse = c.lse;
            h.resampled = c.resampled; h.search_iters = 0;
            a.hist[a.t] = h;
        }
    }
}
