// select.cuh -- radix-select building blocks shared by the quantile (reduce.cu) and the
// pairwise-distance median (svgd.cu).  Integer histograms only -> deterministic.
#pragma once
#include "common.cuh"

#define SEL_THREADS 256

struct SelectState {
    uint32_t prefix;       // bits decided so far (high bits)
    uint32_t pad;
    int64_t rank;          // remaining rank inside the current bucket
    int64_t count_le;      // #values <= selected (pass 4)
    uint32_t min_gt_key;   // smallest key > selected
    uint32_t sel_key;
};

__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    const uint32_t b = k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu);
    return __uint_as_float(b);
}

template <int SHIFT, int BITS, int HIGH_BITS>
static __global__ void __launch_bounds__(SEL_THREADS)
select_hist_kernel(const float* __restrict__ v, int64_t n, const SelectState* st, uint32_t* hist) {
    __shared__ uint32_t sh[1 << BITS];
    for (int i = threadIdx.x; i < (1 << BITS); i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t prefix = st->prefix;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = f2key(v[i]);
        bool match = true;
        if (HIGH_BITS > 0) match = (k >> (32 - HIGH_BITS)) == (prefix >> (32 - HIGH_BITS));
        if (match) atomicAdd(&sh[(k >> SHIFT) & ((1u << BITS) - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (1 << BITS); i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

template <int SHIFT, int BITS>
static __global__ void __launch_bounds__(256) select_pick_kernel(SelectState* st, uint32_t* hist) {
    // 256 threads x (bins/256) consecutive bins: block-wide exclusive scan, the thread whose range holds the rank
    // publishes bucket and remaining rank; the histogram is cleared for the next pass
    constexpr int NB = 1 << BITS, PER = NB / 256;
    static_assert(NB % 256 == 0, "bins per thread");
    __shared__ long long wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint32_t c[PER];
    long long loc = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { c[k] = hist[tid * PER + k]; loc += c[k]; }
    long long inc = loc;
    for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(MB_FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[w] = inc;
    const long long r0 = st->rank;
    const uint32_t prefix0 = st->prefix;
    __syncthreads();
    long long cum = inc - loc, total = 0;
    for (int k = 0; k < 8; ++k) { if (k < w) cum += wsum[k]; total += wsum[k]; }
    uint32_t b = 0xffffffffu;
    long long r = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        if (b == 0xffffffffu && r0 >= cum && r0 < cum + (long long)c[k]) { b = (uint32_t)(tid * PER + k); r = r0 - cum; }
        cum += c[k];
    }
    if (r0 >= total && tid == 255) { b = NB - 1; r = r0 - (total - (long long)c[PER - 1]); }   // rank beyond the data: last bucket
    if (b != 0xffffffffu) {
        st->rank = r;
        const uint32_t pf = prefix0 | (b << SHIFT);
        st->prefix = pf;
        if (SHIFT == 0) { st->sel_key = pf; st->count_le = 0; st->min_gt_key = 0xffffffffu; }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) hist[tid * PER + k] = 0;
}

static __global__ void __launch_bounds__(SEL_THREADS)
select_next_kernel(const float* __restrict__ v, int64_t n, SelectState* st) {
    const uint32_t sel = st->sel_key;
    unsigned long long cnt = 0;
    uint32_t mg = 0xffffffffu;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = f2key(v[i]);
        if (k <= sel) ++cnt; else mg = min(mg, k);
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_down_sync(MB_FULL, cnt, o);
        mg = min(mg, __shfl_down_sync(MB_FULL, mg, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long*)&st->count_le, cnt);
        atomicMin(&st->min_gt_key, mg);
    }
}

static __global__ void select_init_kernel(SelectState* st, int64_t rank) {
    st->prefix = 0; st->rank = rank; st->count_le = 0; st->min_gt_key = 0xffffffffu; st->sel_key = 0;
}

// out = v[lo]*(1-frac) + v[hi]*frac with lo = floor(pos), hi = ceil(pos); lo_rank selected exactly.
static __global__ void select_finish_kernel(const SelectState* st, int64_t lo_rank, double frac, int need_hi, double* out) {
    const double vlo = (double)key2f(st->sel_key);
    double vhi = vlo;
    if (need_hi && st->count_le == lo_rank + 1 && st->min_gt_key != 0xffffffffu) vhi = (double)key2f(st->min_gt_key);
    out[0] = vlo * (1.0 - frac) + vhi * frac;
    out[1] = vlo;
    out[2] = vhi;
}

// q may come from the device (q_dev != NULL: q = q_dev[0]) -- used by mb_abc_adapt where q depends on ess.
static __global__ void select_rank_kernel(SelectState* st, const double* q_dev, double q_host, int64_t n, double* frac_out) {
    const double q = q_dev ? q_dev[0] : q_host;
    double pos = q * (double)(n - 1);
    if (!(pos >= 0.0)) pos = 0.0;
    if (pos > (double)(n - 1)) pos = (double)(n - 1);
    const double lo = floor(pos);
    st->prefix = 0; st->rank = (int64_t)lo; st->count_le = 0; st->min_gt_key = 0xffffffffu; st->sel_key = 0;
    frac_out[0] = pos - lo;      // fraction
    frac_out[1] = lo;            // lo rank as double
}

static __global__ void select_finish_dev_kernel(const SelectState* st, const double* frac, double* out) {
    const double vlo = (double)key2f(st->sel_key);
    double vhi = vlo;
    const int64_t lo_rank = (int64_t)frac[1];
    if (frac[0] > 0.0 && st->count_le == lo_rank + 1 && st->min_gt_key != 0xffffffffu)
        vhi = (double)key2f(st->min_gt_key);
    out[0] = vlo * (1.0 - frac[0]) + vhi * frac[0];
    out[1] = vlo;
    out[2] = vhi;
}

