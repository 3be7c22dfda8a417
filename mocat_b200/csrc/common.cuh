// common.cuh -- context, error handling and warp/block reduction helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/mocat_b200.h"

#define MB_WARP 32
#define MB_FULL 0xffffffffu

// ---- host-side error plumbing -----------------------------------------------------------------
void mb_set_error(const char* fmt, ...);

#define MB_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            mb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return MB_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define MB_CHECK_LAUNCH() MB_CUDA(cudaGetLastError())

#define MB_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            mb_set_error(__VA_ARGS__);        \
            return MB_ERR_ARG;                \
        }                                     \
    } while (0)

// ---- context ----------------------------------------------------------------------------------
struct mb_ctx {
    int device;
    int sms;
    int max_smem_optin;
    // workspace (device)
    double*   partials;        // 3 * MB_MAX_PARTIAL_BLOCKS doubles, x2 parity buffers
    uint32_t* counters;        // small zero-initialised counter block (self-resetting kernels)
    // scan tile status
    int32_t*  scan_flag;
    double*   scan_agg;
    double*   scan_incl;
    int64_t   scan_tiles_cap;
    uint32_t  scan_epoch;
    // generic scratch
    void*     scratch;
    size_t    scratch_bytes;
    unsigned long long ws_generation;   // bumped whenever scratch / scan buffers move: captured graphs hold the old pointers
    // intra-GPU flag-in-data exchange of the resident tempering kernel: [2][MB_LL_BLOCKS][8] words + sequence counter
    unsigned long long* ll_slots;
    // CUDA-graph conditional capture (mb_cond_begin / mb_cond_end)
    cudaStream_t body_stream;
    int          cond_active;
};

#define MB_MAX_PARTIAL_BLOCKS 4096
#define MB_LL_BLOCKS 320          // >= 2 resident blocks x 148 SMs
// counters layout
#define MB_CNT_REDUCE   0    // last-block-done counter of the reduction kernels
#define MB_CNT_SCAN_TILE 1   // dynamic tile id of the scan
#define MB_CNT_SCAN_DONE 2
#define MB_CNT_MOVE     3
#define MB_CNT_MISC     4
#define MB_CNT_SCAN_EPOCH 5
#define MB_CNT_BW_FALLBACK 8 // (64-bit, slots 8-9) median-bracket misses of the tcgen05 bandwidth kernel
#define MB_CNT_RM_MEAN 16      // (fp64) weighted mean acceptance of the Robbins-Monro adaptation
#define MB_NUM_COUNTERS 64

int mb_ensure_scratch(mb_ctx* ctx, size_t bytes);
int mb_ensure_scan(mb_ctx* ctx, int64_t tiles);

static inline cudaStream_t mb_s(mb_stream_t s) { return (cudaStream_t)s; }

// ---- device helpers -----------------------------------------------------------------------------
#ifdef __CUDACC__

// (max, sum exp(w-max), sum exp(2(w-max))) triple and its associative merge.
struct Lse3 {
    double m, s1, s2;
};

__device__ __forceinline__ Lse3 lse3_empty() { return Lse3{-INFINITY, 0.0, 0.0}; }

__device__ __forceinline__ Lse3 lse3_merge(const Lse3& a, const Lse3& b) {
    // NaN in either max propagates (fmax would drop it)
    double m = (a.m != a.m || b.m != b.m) ? (a.m + b.m) : fmax(a.m, b.m);
    if (m == -INFINITY) return Lse3{m, 0.0, 0.0};
    double fa = (a.m == -INFINITY) ? 0.0 : exp(a.m - m);
    double fb = (b.m == -INFINITY) ? 0.0 : exp(b.m - m);
    return Lse3{m, a.s1 * fa + b.s1 * fb, a.s2 * fa * fa + b.s2 * fb * fb};
}

__device__ __forceinline__ double shfl_down_d(double v, int o) { return __shfl_down_sync(MB_FULL, v, o); }
__device__ __forceinline__ double shfl_xor_d(double v, int o) { return __shfl_xor_sync(MB_FULL, v, o); }

__device__ __forceinline__ Lse3 lse3_warp_reduce(Lse3 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Lse3 w{shfl_down_d(v.m, o), shfl_down_d(v.s1, o), shfl_down_d(v.s2, o)};
        v = lse3_merge(v, w);
    }
    return v;   // valid in lane 0
}

// block-wide merge; result valid in thread 0.  smem: at least (blockDim/32) Lse3.
__device__ __forceinline__ Lse3 lse3_block_reduce(Lse3 v, Lse3* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = lse3_warp_reduce(v);
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < nw) ? smem[lane] : lse3_empty();
        v = lse3_warp_reduce(v);
    }
    __syncthreads();
    return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_down_d(v, o);
    return v;
}

__device__ __forceinline__ double block_sum_d(double v, double* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum_d(v);
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < nw) ? smem[lane] : 0.0;
        v = warp_sum_d(v);
    }
    __syncthreads();
    return v;
}

// fill the derived fields of a control block from (wmax, s1, s2)
__device__ __forceinline__ void ctl_set_weights(mb_control* c, const Lse3& r) {
    c->wmax = r.m; c->s1 = r.s1; c->s2 = r.s2;
    double mm = (r.m == -INFINITY || r.m == INFINITY) ? 0.0 : r.m;     // jax logsumexp: non-finite max -> 0
    c->lse = log(r.s1) + mm;
    c->lse2 = log(r.s2) + 2.0 * mm;
    c->log_ess = 2.0 * c->lse - c->lse2;
    c->ess = exp(c->log_ess);
}

__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }

#endif  // __CUDACC__
