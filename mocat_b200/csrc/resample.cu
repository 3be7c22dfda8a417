// resample.cu -- K4 (exact-fp64 CDF by single-pass decoupled look-back scan), K5 (ancestor search),
// K6 (SoA gather).  Replaces the O(n^2) Gumbel-max `random.categorical` + `cdict.__getitem__` of the
// reference (transport/smc.py:61-71, ssm/filtering.py:196-199, core.py:46-56).
//
// Exact-fp64 convention: weights are quantised to multiples of 2^-52,
//     q_i = rint(w_i * scale) * 2^-52,   scale = 2^52 (1 + 2^-24) / sum(w)   (or caller supplied),
// so every fp64 partial sum (< 2) is exactly representable: fp64 addition is associative on these
// values, the scan result is independent of tile order / look-back timing / GPU sharding and equals
// the sequential numpy cumsum bit for bit.  cdf = min(cumsum, 1), cdf[n-1] = 1.
#include "common.cuh"
#include "rng.cuh"
#include "comm.cuh"

const MbCommDev* mb_comm_dev(const mb_comm* c);

#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

#define ST_INVALID 0
#define ST_AGG 1
#define ST_INCL 2

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_d(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

struct ScanArgs {
    const float* in;          // lw (log mode) or w (linear mode)
    int64_t n;
    int log_mode;             // 1: w = exp(lw - ctl->wmax), scale from ctl->s1
    double scale;             // linear mode
    const mb_control* ctl;    // may be NULL in linear mode
    int predicated;           // 1: run only if ctl->resample && !ctl->done
    int raw;                  // 1: rank-relative CDF for the sharded path: no clamp at 1, last element not forced
    double* cdf;
    int32_t* flag; double* agg; double* incl;
    uint32_t* epoch_ptr;      // device-side launch epoch (incremented by the kernel itself: graph-replay safe)
    uint32_t* tile_counter; uint32_t* done_counter;
    int64_t num_tiles;
};

__global__ void __launch_bounds__(SCAN_THREADS)
scan_cdf_kernel(ScanArgs a) {
    if (a.predicated && (a.ctl->done || !a.ctl->resample)) return;
    __shared__ double warp_tot[SCAN_THREADS / 32];
    __shared__ double tile_prefix_s;
    __shared__ unsigned tile_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    double scale;
    float wmax = 0.f;
    if (a.log_mode) {
        wmax = (float)a.ctl->wmax;
        if (!(wmax > -INFINITY) || wmax == INFINITY) wmax = 0.f;
        scale = 4503599627370496.0 * (1.0 + 5.9604644775390625e-8) / a.ctl->s1;
    } else {
        scale = a.scale;
    }
    const uint32_t epoch = (*a.epoch_ptr) & 0x1fffffffu;     // every block reads it before any block can exit
    const int st_base = (int)(epoch << 2);

    while (true) {
        if (threadIdx.x == 0) tile_s = atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        const int64_t tile = tile_s;
        if (tile >= a.num_tiles) break;
        const int64_t base = tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;

        // ---- load 16 consecutive items per thread (4 x 128-bit), quantise
        double q[SCAN_ITEMS];
        const bool full = (base + SCAN_ITEMS <= a.n) && (((uintptr_t)a.in & 15) == 0);
        if (full) {
            const float4* p = reinterpret_cast<const float4*>(a.in + base);
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
                const float4 v = __ldcs(p + k);
                q[4 * k + 0] = v.x; q[4 * k + 1] = v.y; q[4 * k + 2] = v.z; q[4 * k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k)
                q[k] = (base + k < a.n) ? (double)a.in[base + k] : (a.log_mode ? -INFINITY : 0.0);
        }
        double run = 0.0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            float w = (float)q[k];
            if (a.log_mode) {
                w = __expf(w - wmax);
                if (w != w) w = 0.f;
            }
            const double qq = rint((double)w * scale) * 2.220446049250313e-16;
            run += qq;
            q[k] = run;                                   // thread-local inclusive
        }
        // ---- warp + block exclusive offsets (exact arithmetic: any association is the same)
        double incl_w = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(MB_FULL, incl_w, o);
            if (lane >= o) incl_w += t;
        }
        if (lane == 31) warp_tot[warp] = incl_w;
        __syncthreads();
        double warp_off = 0.0, tile_total = 0.0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w) {
            const double t = warp_tot[w];
            if (w < warp) warp_off += t;
            tile_total += t;
        }
        const double thread_off = warp_off + (incl_w - run);

        // ---- decoupled look-back (warp 0)
        if (warp == 0) {
            double prefix = 0.0;
            if (tile == 0) {
                if (lane == 0) {
                    a.incl[0] = tile_total;
                    __threadfence();
                    st_release(a.flag + 0, st_base | ST_INCL);
                }
            } else {
                if (lane == 0) {
                    a.agg[tile] = tile_total;
                    __threadfence();
                    st_release(a.flag + tile, st_base | ST_AGG);
                }
                int64_t look = tile - 1;
                while (true) {
                    const int64_t idx = look - lane;
                    int f = st_base | ST_INCL;               // lanes before tile 0 behave as "inclusive 0"
                    double val = 0.0;
                    if (idx >= 0) {
                        do { f = ld_acquire(a.flag + idx); } while ((f >> 2) != (int)epoch || (f & 3) == ST_INVALID);
                        val = ((f & 3) == ST_INCL) ? ld_relaxed_d(a.incl + idx) : ld_relaxed_d(a.agg + idx);
                    }
                    const unsigned incl_mask = __ballot_sync(MB_FULL, (f & 3) == ST_INCL);
                    const int first = incl_mask ? (__ffs(incl_mask) - 1) : 32;
                    double contrib = (lane <= first) ? val : 0.0;
                    contrib = warp_sum_d(contrib);
                    contrib = __shfl_sync(MB_FULL, contrib, 0);
                    prefix += contrib;
                    if (incl_mask) break;
                    look -= 32;
                }
                if (lane == 0) {
                    a.incl[tile] = prefix + tile_total;
                    __threadfence();
                    st_release(a.flag + tile, st_base | ST_INCL);
                }
            }
            if (lane == 0) tile_prefix_s = prefix;
        }
        __syncthreads();
        const double off = tile_prefix_s + thread_off;

        // ---- write cdf = min(off + local, 1); last element forced to 1
        if (base + SCAN_ITEMS <= a.n && (((uintptr_t)a.cdf & 15) == 0)) {
            double2* o2 = reinterpret_cast<double2*>(a.cdf + base);
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS / 2; ++k) {
                double c0 = off + q[2 * k], c1 = off + q[2 * k + 1];
                if (!a.raw) {
                    c0 = fmin(c0, 1.0); c1 = fmin(c1, 1.0);
                    if (base + 2 * k + 1 == a.n - 1) c1 = 1.0;
                }
                __stcs(o2 + k, make_double2(c0, c1));
            }
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k)
                if (base + k < a.n)
                    a.cdf[base + k] = a.raw ? (off + q[k]) : ((base + k == a.n - 1) ? 1.0 : fmin(off + q[k], 1.0));
        }
        __syncthreads();
    }
    // every block fetches exactly one terminating tile id; the last block to exit resets the counters
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(a.done_counter, 1u);
        if (t == gridDim.x - 1) {
            *a.done_counter = 0;
            *a.tile_counter = 0;
            *a.epoch_ptr = epoch + 1;
        }
    }
}

static int scan_launch(mb_ctx* ctx, ScanArgs& a, cudaStream_t st) {
    a.num_tiles = (a.n + SCAN_TILE - 1) / SCAN_TILE;
    MB_REQUIRE(a.num_tiles < 0xffff0000ll, "scan: too many tiles");
    if (mb_ensure_scan(ctx, a.num_tiles) != MB_OK) return MB_ERR_CUDA;
    a.flag = ctx->scan_flag; a.agg = ctx->scan_agg; a.incl = ctx->scan_incl;
    a.epoch_ptr = ctx->counters + MB_CNT_SCAN_EPOCH;
    a.tile_counter = ctx->counters + MB_CNT_SCAN_TILE;
    a.done_counter = ctx->counters + MB_CNT_SCAN_DONE;
    int64_t grid = (int64_t)ctx->sms * 6;                 // resident-sized grid, tiles fetched dynamically
    if (grid > a.num_tiles) grid = a.num_tiles;
    if (grid > 0xffff) grid = 0xffff;
    if (grid < 1) grid = 1;
    scan_cdf_kernel<<<(unsigned)grid, SCAN_THREADS, 0, st>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

extern "C" int mb_cumsum_lw(mb_ctx* ctx, const float* lw, int64_t n, const mb_control* ctl, int flags,
                            double* cdf, mb_stream_t stream) {
    const int force = flags & 1;
    MB_REQUIRE(ctx && lw && ctl && cdf && n > 0, "mb_cumsum_lw: bad arguments");
    ScanArgs a{};
    a.in = lw; a.n = n; a.log_mode = 1; a.scale = 0; a.ctl = ctl; a.predicated = force ? 0 : 1; a.cdf = cdf;
    a.raw = (flags & 2) ? 1 : 0;
    return scan_launch(ctx, a, mb_s(stream));
}

extern "C" int mb_cumsum_f32(mb_ctx* ctx, const float* w, int64_t n, double scale, double* cdf, mb_stream_t stream) {
    MB_REQUIRE(ctx && w && cdf && n > 0 && scale > 0, "mb_cumsum_f32: bad arguments");
    ScanArgs a{};
    a.in = w; a.n = n; a.log_mode = 0; a.scale = scale; a.ctl = nullptr; a.predicated = 0; a.cdf = cdf;
    return scan_launch(ctx, a, mb_s(stream));
}

// ------------------------------------------------------------------------------------------------ K5
// upper_bound: smallest j in [lo, hi) with cdf[j] > u (hi if none)
__device__ __forceinline__ int64_t upper_bound_g(const double* __restrict__ cdf, int64_t lo, int64_t hi, double u) {
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(cdf + mid) > u) hi = mid; else lo = mid + 1;
    }
    return lo;
}
__device__ __forceinline__ int upper_bound_s(const double* s, int lo, int hi, double u) {
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (s[mid] > u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

#define ANC_THREADS 256
#define ANC_ITEMS 8
#define ANC_BLOCK_OUT (ANC_THREADS * ANC_ITEMS)
#define ANC_SMEM_CAP 5120        // doubles (40 KiB)

struct AncArgs {
    const double* cdf; int64_t n;
    const double* u;            // NULL => Philox
    uint64_t seed; uint32_t step; int64_t gid0;
    int32_t* anc; int64_t n_out;
    const mb_control* ctl;
};

// Systematic: u_i = (i + u0)/n_out is monotone in i, so a block of outputs maps to one contiguous
// window of the CDF: two full binary searches per block find the window, it is staged in shared
// memory (coalesced), and the per-output searches run there.
__global__ void __launch_bounds__(ANC_THREADS)
ancestors_systematic_kernel(AncArgs a) {
    if (a.ctl && (a.ctl->done || !a.ctl->resample)) return;
    __shared__ double win[ANC_SMEM_CAP];
    __shared__ int64_t w_lo, w_hi;
    double u0;
    if (a.u) u0 = a.u[0];
    else { const Philox4 r = philox_raw(a.ctl ? a.ctl->seed : a.seed, 0ull, (a.ctl ? (uint32_t)(a.ctl->iter + 1) : a.step), MB_P_RESAMPLE, 0u); u0 = u53(r.x, r.y); }
    const double nd = (double)a.n_out;
    for (int64_t b0 = (int64_t)blockIdx.x * ANC_BLOCK_OUT; b0 < a.n_out; b0 += (int64_t)gridDim.x * ANC_BLOCK_OUT) {
        const int64_t b1 = min(b0 + (int64_t)ANC_BLOCK_OUT, a.n_out);
        if (threadIdx.x == 0) w_lo = upper_bound_g(a.cdf, 0, a.n, ((double)b0 + u0) / nd);
        if (threadIdx.x == 32) w_hi = upper_bound_g(a.cdf, 0, a.n, ((double)(b1 - 1) + u0) / nd);
        __syncthreads();
        const int64_t lo = w_lo;
        const int64_t hi = min(w_hi, a.n - 1);            // inclusive upper end of the window
        const int64_t len = hi - lo + 1;
        const bool staged = len <= ANC_SMEM_CAP;
        if (staged) {
            for (int64_t k = threadIdx.x; k < len; k += ANC_THREADS) win[k] = __ldg(a.cdf + lo + k);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ANC_ITEMS; ++k) {
            const int64_t i = b0 + (int64_t)k * ANC_THREADS + threadIdx.x;
            if (i < b1) {
                const double u = ((double)i + u0) / nd;
                int64_t j;
                if (staged) j = lo + upper_bound_s(win, 0, (int)len, u);
                else j = upper_bound_g(a.cdf, lo, hi + 1, u);
                if (j > a.n - 1) j = a.n - 1;
                a.anc[i] = (int32_t)j;
            }
        }
        __syncthreads();
    }
}

// Multinomial: n_out iid uniforms (unsorted) -> independent binary searches over the whole CDF.
__global__ void __launch_bounds__(ANC_THREADS)
ancestors_multinomial_kernel(AncArgs a) {
    if (a.ctl && (a.ctl->done || !a.ctl->resample)) return;
    const uint32_t step = a.ctl ? (uint32_t)(a.ctl->iter + 1) : a.step;   // device-side step: graph-capturable
    const uint64_t seed = a.ctl ? a.ctl->seed : a.seed;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_out; i += (int64_t)gridDim.x * blockDim.x) {
        double u;
        if (a.u) u = a.u[i];
        else { const Philox4 r = philox_raw(seed, (uint64_t)(a.gid0 + i), step, MB_P_RESAMPLE, 0u); u = u53(r.x, r.y); }
        int64_t j = upper_bound_g(a.cdf, 0, a.n, u);
        if (j > a.n - 1) j = a.n - 1;
        a.anc[i] = (int32_t)j;
    }
}

extern "C" int mb_ancestors(mb_ctx* ctx, const double* cdf, int64_t n, int mode, const double* u, uint64_t seed,
                            uint32_t step, int64_t gid0, int32_t* anc, int64_t n_out, const mb_control* ctl,
                            mb_stream_t stream) {
    MB_REQUIRE(ctx && cdf && anc && n > 0 && n_out > 0 && n <= 0x7fffffffll, "mb_ancestors: bad arguments");
    AncArgs a{cdf, n, u, seed, step, gid0, anc, n_out, ctl};
    if (mode == MB_RESAMPLE_SYSTEMATIC) {
        int64_t grid = (n_out + ANC_BLOCK_OUT - 1) / ANC_BLOCK_OUT;
        if (grid > (int64_t)ctx->sms * 16) grid = (int64_t)ctx->sms * 16;
        ancestors_systematic_kernel<<<(unsigned)grid, ANC_THREADS, 0, mb_s(stream)>>>(a);
    } else if (mode == MB_RESAMPLE_MULTINOMIAL) {
        int64_t grid = (n_out + ANC_THREADS - 1) / ANC_THREADS;
        if (grid > (int64_t)ctx->sms * 32) grid = (int64_t)ctx->sms * 32;
        ancestors_multinomial_kernel<<<(unsigned)grid, ANC_THREADS, 0, mb_s(stream)>>>(a);
    } else {
        mb_set_error("mb_ancestors: unknown mode %d", mode);
        return MB_ERR_ARG;
    }
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// ------------------------------------------------------------------------------------------------
// Stratified-exact multinomial resampling: n iid uniforms are equivalent in law to
//   (1) the counts N_b of uniforms falling into B equal strata of [0,1)  -- obtained as an integer histogram
//       of first-stage Philox uniforms (deterministic), and
//   (2) N_b fresh iid uniforms inside stratum b                          -- second-stage Philox uniforms.
// Output slot g (global) takes stratum s(g) = upper_bound(offsets, g) - 1 and u_g = (s + v_g)/B.  The u's come
// out sorted by stratum, so (like systematic resampling) a block of outputs maps to one contiguous CDF window
// that is staged in shared memory, ancestors are nearly sorted and the fused gather stays coalesced.  The law
// is exactly Cat(softmax(w))^n as in the reference (transport/smc.py:65-67).
struct StrataArgs {
    uint32_t* hist; uint32_t* offsets; int B;
    int64_t n_out; int64_t gid0;
    uint64_t seed; uint32_t step;
    const mb_control* ctl;
};

__global__ void __launch_bounds__(256) strata_hist_kernel(StrataArgs a) {
    if (a.ctl && (a.ctl->done || !a.ctl->resample)) return;
    const uint32_t step = a.ctl ? (uint32_t)(a.ctl->iter + 1) : a.step;
    const uint64_t seed = a.ctl ? a.ctl->seed : a.seed;
    const double Bd = (double)a.B;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_out; i += (int64_t)gridDim.x * blockDim.x) {
        const Philox4 r = philox_raw(seed, (uint64_t)(a.gid0 + i), step, MB_P_RESAMPLE, 0u);
        const int b = (int)(u53(r.x, r.y) * Bd);                      // exact: B is a power of two
        atomicAdd(a.hist + b, 1u);
    }
}

// exclusive scan of the B strata counts -> offsets[B+1]: chunk totals, single-block scan of the totals,
// per-chunk rescan (B <= 2^24 -> <= 4096 chunks of 4096 counts)
#define SS_CHUNK 4096
__global__ void __launch_bounds__(1024) strata_chunk_sum_kernel(StrataArgs a, uint32_t* chunk_tot) {
    if (a.ctl && (a.ctl->done || !a.ctl->resample)) return;
    __shared__ uint32_t wsum[32];
    const int base = blockIdx.x * SS_CHUNK;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int i = base + k * 1024 + threadIdx.x; if (i < a.B) s += a.hist[i]; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(MB_FULL, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = wsum[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(MB_FULL, w, o);
        if (threadIdx.x == 0) chunk_tot[blockIdx.x] = w;
    }
}

// DIRECT (few chunks, the n ~ 1e6 regime): every block sums the counts in front of its chunk itself, which saves the
// chunk-total kernel (one graph node per step); otherwise the prefix comes from the chunk totals.
template <bool DIRECT>
__global__ void __launch_bounds__(1024) strata_chunk_scan_kernel(StrataArgs a, const uint32_t* chunk_tot, int nchunks) {
    if (a.ctl && (a.ctl->done || !a.ctl->resample)) return;
    // every block scans the (<= 4096) chunk totals up to its own chunk in shared memory, then its chunk
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t pre = 0;
    if (DIRECT) {
        const uint4* h4 = reinterpret_cast<const uint4*>(a.hist);      // chunk starts are multiples of 4096 counts
        for (int i = threadIdx.x; i < (int)blockIdx.x * (SS_CHUNK / 4); i += 1024) {
            const uint4 v = h4[i];
            pre += v.x + v.y + v.z + v.w;
        }
    } else {
        for (int i = threadIdx.x; i < (int)blockIdx.x; i += 1024) pre += chunk_tot[i];
    }
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(MB_FULL, pre, o);
    if (lane == 0) wsum[warp] = pre;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = wsum[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(MB_FULL, w, o);
        if (threadIdx.x == 0) base_s = w;
    }
    __syncthreads();
    uint32_t carry = base_s;
    const int base = blockIdx.x * SS_CHUNK;
    for (int k = 0; k < 4; ++k) {
        const int i = base + k * 1024 + threadIdx.x;
        const uint32_t v = (i < a.B) ? a.hist[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(MB_FULL, x, o); if (lane >= o) x += t; }
        __syncthreads();
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(MB_FULL, w, o); if (lane >= o) w += t; }
            wsum[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = x + (warp ? wsum[warp - 1] : 0u) + carry;
        if (i < a.B) a.offsets[i] = incl - v;
        carry += wsum[31];
        if (i == a.B - 1) a.offsets[a.B] = incl;
    }
}

struct SortedAncArgs {
    int stratified;                      // 0: systematic, 1: stratified-exact multinomial
    const double* cdf; int64_t n;        // single-GPU: materialised clamped CDF
    mb_shard sh; int sharded;            // sharded: rank-relative CDFs of all ranks + totals
    const uint32_t* offsets; int B;      // strata offsets (stratified)
    uint32_t* hist_clear;                // stratified: the consumed counts are zeroed for the next resampling
    uint64_t seed; uint32_t step; int64_t gid0; int64_t n_out; int64_t n_total_out;
    int32_t* anc;
    const mb_control* ctl;
};

struct GlobalCdf {                       // unified view of the (possibly sharded) global CDF
    const double* cdf; int64_t n;
    bool sharded; int W; int64_t nl;
    const double* peers[MB_MAX_WORLD];
    double off[MB_MAX_WORLD + 1];

    __device__ __forceinline__ double at(int64_t j) const {           // min(C_j, 1), cdf[n-1] = 1
        if (!sharded) return __ldg(cdf + j);
        if (j >= n - 1) return 1.0;
        const int r = (int)(j / nl);
        return fmin(off[r] + __ldg(peers[r] + (j - (int64_t)r * nl)), 1.0);
    }
    __device__ __forceinline__ int64_t upper_bound(int64_t lo, int64_t hi, double u) const {
        if (sharded) {                                                // narrow to one rank first (<= 8 compares)
            for (int r = 0; r < W; ++r) {
                const int64_t r0 = (int64_t)r * nl, r1 = r0 + nl;
                if (r1 <= lo) continue;
                if (r0 >= hi) break;
                if (fmin(off[r + 1], 1.0) > u || r == W - 1) { lo = max(lo, r0); hi = min(hi, r1); break; }
                lo = r1;
            }
        }
        while (lo < hi) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (at(mid) > u) hi = mid; else lo = mid + 1;
        }
        return lo;
    }
};

__global__ void __launch_bounds__(ANC_THREADS)
ancestors_sorted_kernel(SortedAncArgs a) {
    if (a.ctl && (a.ctl->done || !a.ctl->resample)) return;
    __shared__ double win[ANC_SMEM_CAP];
    __shared__ uint32_t soff[1024];
    __shared__ int64_t w_lo, w_hi;
    __shared__ int s_lo_s, s_hi_s;
    const uint32_t step = a.ctl ? (uint32_t)(a.ctl->iter + 1) : a.step;
    const uint64_t seed = a.ctl ? a.ctl->seed : a.seed;
    if (a.hist_clear)
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.B; i += (int64_t)gridDim.x * blockDim.x)
            a.hist_clear[i] = 0u;
    GlobalCdf G;
    G.cdf = a.cdf; G.n = a.sharded ? a.sh.n_total : a.n; G.sharded = a.sharded != 0; G.W = a.sh.world; G.nl = a.sh.n_local;
    if (G.sharded) {
        G.off[0] = 0.0;
        for (int r = 0; r < G.W; ++r) { G.peers[r] = a.sh.cdf_peers[r]; G.off[r + 1] = G.off[r] + a.sh.totals[r]; }
    }
    double u0 = 0.0;
    if (!a.stratified) { const Philox4 r = philox_raw(seed, 0ull, step, MB_P_RESAMPLE, 0u); u0 = u53(r.x, r.y); }
    const double nd = (double)a.n_total_out, Bd = (double)a.B;

    for (int64_t b0 = (int64_t)blockIdx.x * ANC_BLOCK_OUT; b0 < a.n_out; b0 += (int64_t)gridDim.x * ANC_BLOCK_OUT) {
        const int64_t b1 = min(b0 + (int64_t)ANC_BLOCK_OUT, a.n_out);
        const int64_t g_first = a.gid0 + b0, g_last = a.gid0 + b1 - 1;
        // ---- u-range of the block -> CDF window
        if (threadIdx.x == 0) {
            double ulo;
            if (a.stratified) {
                int lo = 0, hi = a.B;                                  // stratum(g) = upper_bound(offsets, g) - 1
                while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int64_t)a.offsets[mid + 1] > g_first) hi = mid; else lo = mid + 1; }
                s_lo_s = lo;
                ulo = (double)lo / Bd;
            } else ulo = ((double)g_first + u0) / nd;
            w_lo = G.upper_bound(0, G.n, ulo);
        }
        if (threadIdx.x == 32) {
            double uhi;
            if (a.stratified) {
                int lo = 0, hi = a.B;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int64_t)a.offsets[mid + 1] > g_last) hi = mid; else lo = mid + 1; }
                s_hi_s = lo;
                uhi = (double)(lo + 1) / Bd;                           // every u of the block is < uhi
            } else uhi = ((double)g_last + u0) / nd;
            w_hi = G.upper_bound(0, G.n, uhi);
        }
        __syncthreads();
        const int64_t lo = w_lo, hi = min(w_hi, G.n - 1);
        const int64_t len = hi - lo + 1;
        const bool staged = len <= ANC_SMEM_CAP;
        const int s_lo = s_lo_s, s_hi = s_hi_s;
        const bool soff_staged = a.stratified && (s_hi - s_lo + 2) <= 1024;
        if (staged) for (int64_t k = threadIdx.x; k < len; k += ANC_THREADS) win[k] = G.at(lo + k);
        if (soff_staged) for (int k = threadIdx.x; k < s_hi - s_lo + 2; k += ANC_THREADS) soff[k] = a.offsets[s_lo + k];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ANC_ITEMS; ++k) {
            const int64_t i = b0 + (int64_t)k * ANC_THREADS + threadIdx.x;
            if (i < b1) {
                const int64_t g = a.gid0 + i;
                double u;
                if (a.stratified) {
                    int l = s_lo, h = s_hi;                           // stratum of g inside [s_lo, s_hi]
                    if (soff_staged) { while (l < h) { const int mid = (l + h) >> 1; if ((int64_t)soff[mid + 1 - s_lo] > g) h = mid; else l = mid + 1; } }
                    else { while (l < h) { const int mid = (l + h) >> 1; if ((int64_t)a.offsets[mid + 1] > g) h = mid; else l = mid + 1; } }
                    const Philox4 r = philox_raw(seed, (uint64_t)g, step, MB_P_RESAMPLE, 1u);
                    u = ((double)l + u53(r.x, r.y)) / Bd;
                } else u = ((double)g + u0) / nd;
                int64_t j;
                if (staged) j = lo + upper_bound_s(win, 0, (int)len, u);
                else j = G.upper_bound(lo, hi + 1, u);
                if (j > G.n - 1) j = G.n - 1;
                a.anc[i] = (int32_t)j;
            }
        }
        __syncthreads();
    }
}

static int strata_B(int64_t n_total_out) {
    int B = 1;
    while ((int64_t)B * 32 <= n_total_out && B < (1 << 24)) B <<= 1;   // largest power of two <= n/16
    return B;
}

extern "C" int mb_strata_count(int64_t n_total_out) { return strata_B(n_total_out); }

// first stage: histogram of this rank's outputs over the B strata (hist must hold B uint32; it is zeroed here)
extern "C" int mb_strata_hist(mb_ctx* ctx, int64_t n_out, int64_t gid0, int B, uint64_t seed, uint32_t step,
                              const mb_control* ctl, uint32_t* hist, int clear, mb_stream_t stream) {
    MB_REQUIRE(ctx && hist && n_out > 0 && B >= 1 && (B & (B - 1)) == 0, "mb_strata_hist: bad arguments");
    if (clear) MB_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * B, mb_s(stream)));
    StrataArgs a{hist, nullptr, B, n_out, gid0, seed, step, ctl};
    int64_t grid = (n_out + 255) / 256;
    if (grid > (int64_t)ctx->sms * 16) grid = (int64_t)ctx->sms * 16;
    strata_hist_kernel<<<(unsigned)grid, 256, 0, mb_s(stream)>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// second stage: exclusive scan -> offsets[B+1]; then ancestors for this rank's outputs.
// sh == NULL: single GPU (cdf = materialised clamped CDF of n particles); else the sharded global CDF.
extern "C" int mb_ancestors_sorted(mb_ctx* ctx, const double* cdf, int64_t n, const mb_shard* sh, int mode,
                                   uint32_t* hist, uint32_t* offsets, int B, uint64_t seed, uint32_t step,
                                   int64_t gid0, int64_t n_total_out, int32_t* anc, int64_t n_out,
                                   const mb_control* ctl, mb_stream_t stream) {
    MB_REQUIRE(ctx && anc && n_out > 0 && n_total_out >= n_out && (cdf || sh), "mb_ancestors_sorted: bad arguments");
    MB_REQUIRE(mode == MB_RESAMPLE_SYSTEMATIC || (hist && offsets && B >= 1), "mb_ancestors_sorted: strata buffers missing");
    SortedAncArgs a{};
    a.stratified = (mode == MB_RESAMPLE_MULTINOMIAL) ? 1 : 0;
    a.cdf = cdf; a.n = n;
    if (sh) { a.sh = *sh; a.sharded = 1; }
    a.offsets = offsets; a.B = B; a.seed = seed; a.step = step; a.gid0 = gid0; a.n_out = n_out;
    a.n_total_out = n_total_out; a.anc = anc; a.ctl = ctl;
    if (a.stratified) {
        StrataArgs sa{hist, offsets, B, n_out, gid0, seed, step, ctl};
        a.hist_clear = hist;
        const int nchunks = (B + SS_CHUNK - 1) / SS_CHUNK;
        uint32_t* chunk_tot = reinterpret_cast<uint32_t*>((char*)ctx->scratch + (3u << 20));    // <= 16 KiB of the scratch
        if (nchunks <= 32 && (((uintptr_t)hist & 15) == 0)) {
            strata_chunk_scan_kernel<true><<<nchunks, 1024, 0, mb_s(stream)>>>(sa, chunk_tot, nchunks);
        } else {
            strata_chunk_sum_kernel<<<nchunks, 1024, 0, mb_s(stream)>>>(sa, chunk_tot);
            strata_chunk_scan_kernel<false><<<nchunks, 1024, 0, mb_s(stream)>>>(sa, chunk_tot, nchunks);
        }
        MB_CHECK_LAUNCH();
    }
    int64_t grid = (n_out + ANC_BLOCK_OUT - 1) / ANC_BLOCK_OUT;
    if (grid > (int64_t)ctx->sms * 16) grid = (int64_t)ctx->sms * 16;
    ancestors_sorted_kernel<<<(unsigned)grid, ANC_THREADS, 0, mb_s(stream)>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// sharded: global strata counts = sum of the ranks' local histograms, read over NVLink after a mailbox barrier
struct StrataSumArgs {
    const uint32_t* peers[MB_MAX_WORLD]; int world; int B; uint32_t* out; const mb_control* ctl;
};
__global__ void __launch_bounds__(256) strata_sum_kernel(StrataSumArgs a) {
    if (a.ctl && (a.ctl->done || !a.ctl->resample)) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.B; i += gridDim.x * blockDim.x) {
        uint32_t s = 0;
        for (int r = 0; r < a.world; ++r) s += a.peers[r][i];
        a.out[i] = s;
    }
}
__global__ void strata_barrier_kernel(MbCommDev c, const mb_control* ctl) {
    if (ctl && (ctl->done || !ctl->resample)) return;               // identical decision on every rank
    __shared__ double in[1], out[MB_MAX_WORLD];
    if (threadIdx.x == 0) in[0] = 0.0;
    __syncwarp();
    comm_allgather_warp(c, in, 1, out);
}

extern "C" int mb_strata_reduce(mb_ctx* ctx, mb_comm* comm, const void* const* hist_peers, int world, int B,
                                uint32_t* hist_out, int barrier, const mb_control* ctl, mb_stream_t stream) {
    MB_REQUIRE(ctx && comm && hist_peers && hist_out && world >= 1 && world <= MB_MAX_WORLD && B >= 1,
               "mb_strata_reduce: bad arguments");
    if (barrier) {                                   // 0: the caller exchanged (mb_comm_allgather) after its histogram kernel
        strata_barrier_kernel<<<1, 32, 0, mb_s(stream)>>>(*mb_comm_dev(comm), ctl);
        MB_CHECK_LAUNCH();
    }
    StrataSumArgs a{};
    for (int r = 0; r < world; ++r) a.peers[r] = (const uint32_t*)hist_peers[r];
    a.world = world; a.B = B; a.out = hist_out; a.ctl = ctl;
    int grid = (B + 255) / 256;
    if (grid > ctx->sms * 8) grid = ctx->sms * 8;
    strata_sum_kernel<<<grid, 256, 0, mb_s(stream)>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// ---- sharded ancestors: global search over the ranks' relative CDFs (peer-mapped)
struct ShardAncArgs {
    mb_shard sh;
    int mode;
    uint64_t seed; uint32_t step;
    int32_t* anc; int64_t n_out;
    const mb_control* ctl;
};

__global__ void __launch_bounds__(ANC_THREADS)
ancestors_sharded_kernel(ShardAncArgs a) {
    if (a.ctl && (a.ctl->done || !a.ctl->resample)) return;
    const uint32_t step = a.ctl ? (uint32_t)(a.ctl->iter + 1) : a.step;
    const uint64_t seed = a.ctl ? a.ctl->seed : a.seed;
    const int W = a.sh.world;
    double off[MB_MAX_WORLD + 1];                       // exclusive prefix of the (exact) rank totals
    off[0] = 0.0;
    for (int r = 0; r < W; ++r) off[r + 1] = off[r] + a.sh.totals[r];
    double u0 = 0.0;
    if (a.mode == MB_RESAMPLE_SYSTEMATIC) { const Philox4 r = philox_raw(seed, 0ull, step, MB_P_RESAMPLE, 0u); u0 = u53(r.x, r.y); }
    const int64_t g0 = (int64_t)a.sh.rank * a.sh.n_local;
    const int64_t nl = a.sh.n_local, nt = a.sh.n_total;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_out; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = g0 + i;
        double u;
        if (a.mode == MB_RESAMPLE_SYSTEMATIC) u = ((double)g + u0) / (double)nt;
        else { const Philox4 r = philox_raw(seed, (uint64_t)g, step, MB_P_RESAMPLE, 0u); u = u53(r.x, r.y); }
        // owner: first rank whose last global CDF value min(off_r + T_r, 1) exceeds u
        int r = 0;
        while (r < W - 1 && !(fmin(off[r + 1], 1.0) > u)) ++r;
        const double* c = a.sh.cdf_peers[r];
        const double o = off[r];
        int64_t lo = 0, hi = nl;
        while (lo < hi) {                               // smallest j with min(o + c[j], 1) > u (exact fp64 add)
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (fmin(o + __ldg(c + mid), 1.0) > u) hi = mid; else lo = mid + 1;
        }
        int64_t j = (int64_t)r * nl + lo;
        if (j > nt - 1) j = nt - 1;                     // cdf[n-1] = 1 convention
        a.anc[i] = (int32_t)j;
    }
}

extern "C" int mb_ancestors_sharded(mb_ctx* ctx, const mb_shard* sh, int mode, uint64_t seed, uint32_t step,
                                    int32_t* anc, int64_t n_out, const mb_control* ctl, mb_stream_t stream) {
    MB_REQUIRE(ctx && sh && anc && n_out > 0 && sh->world >= 1 && sh->world <= MB_MAX_WORLD && sh->totals &&
                   sh->n_total <= 0x7fffffffll, "mb_ancestors_sharded: bad arguments");
    ShardAncArgs a{*sh, mode, seed, step, anc, n_out, ctl};
    int64_t grid = (n_out + ANC_THREADS - 1) / ANC_THREADS;
    if (grid > (int64_t)ctx->sms * 32) grid = (int64_t)ctx->sms * 32;
    ancestors_sharded_kernel<<<(unsigned)grid, ANC_THREADS, 0, mb_s(stream)>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// ------------------------------------------------------------------------------------------------ K6
__global__ void __launch_bounds__(256)
gather_state_kernel(const int32_t* __restrict__ anc, int64_t n_out, int ncols, const float* __restrict__ src,
                    int64_t ld_src, float* __restrict__ dst, int64_t ld_dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = anc[i];
        for (int c = 0; c < ncols; ++c) dst[(int64_t)c * ld_dst + i] = __ldg(src + (int64_t)c * ld_src + a);
    }
}

extern "C" int mb_gather_state(mb_ctx* ctx, const int32_t* anc, int64_t n_out, int ncols, const float* src,
                               int64_t ld_src, float* dst, int64_t ld_dst, mb_stream_t stream) {
    MB_REQUIRE(ctx && anc && src && dst && n_out > 0 && ncols > 0, "mb_gather_state: bad arguments");
    int64_t grid = (n_out + 255) / 256;
    if (grid > (int64_t)ctx->sms * 32) grid = (int64_t)ctx->sms * 32;
    gather_state_kernel<<<(unsigned)grid, 256, 0, mb_s(stream)>>>(anc, n_out, ncols, src, ld_src, dst, ld_dst);
    MB_CHECK_LAUNCH();
    return MB_OK;
}
