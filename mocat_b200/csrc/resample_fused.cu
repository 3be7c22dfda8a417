// resample_fused.cu -- systematic resampling without a materialised CDF (production path of the particle filter).
//
// Replaces `random.categorical` + gather index generation of the reference (transport/smc.py:61-71,
// ssm/filtering.py:196-199) for sorted-uniform (systematic) resampling.  Round 1 wrote an fp64 CDF (8 B/particle),
// read it back in a second kernel (8 B) and searched it per OUTPUT: 24 B/particle and two latency-bound kernels
// (25 % / 14 % of the HBM roofline).  Here the weights are read twice (4 + 4 B) and the ancestors written once (4 B):
//
//   pass A  rf_tile_sums_kernel   integer weight e_i of every particle, summed per 4096-particle tile; the last block
//                                 scans the tile sums (exclusive prefix) and publishes the shard total.
//   (sharded: the shard totals are exchanged with mb_comm_allgather -- one 8-byte word per rank)
//   pass B  rf_ancestors_kernel   every rank scans ITS OWN source tiles: re-derives e_i, block-scans the tile, turns every
//                                 particle's inclusive cumulative weight C_j into the NUMBER OF OUTPUTS BELOW IT c_j
//                                 (closed form, see rf_count), drops a head marker at output c_{j-1} for every particle
//                                 with offspring, max-scans the markers in shared memory and stores the ancestors with
//                                 coalesced 128 B writes -- into the ancestor array of the rank that OWNS the output
//                                 (peer stores over NVLink; the step kernel then fetches the ancestor's row, one
//                                 contiguous TMA copy, from whichever GPU owns it: the redistribution all-to-all of
//                                 the north star is fused into the gather).
//   pass C  rf_heavy_kernel       tiles with more than RF_INLINE outputs (collapsed weights: a handful of particles own
//                                 all the offspring) are queued by pass B as records {tile, C before, C after}.  After
//                                 a barrier every rank collects the records of ALL ranks and fills the part of each
//                                 heavy tile's output range that falls into its own slots ("pull": the work is
//                                 balanced however few particles own the offspring; their state is fetched by the step
//                                 kernel, once per run of equal ancestors).
//
// Exact arithmetic (DESIGN.md "resampling convention"): e_i = rint(w_i 2^K) as uint64 (w_i = exp(lw_i - max lw) <= 1
// in log mode, the caller's weight in linear mode; K = min(40, 63 - ceil(log2 n_total)) so that the total S < 2^63).
// Integer sums are associative: tile order, block scheduling and the sharding over GPUs cannot change a bit.
// Systematic resampling is then evaluated in EXACT rational arithmetic: with u0 = k0 / 2^32,
//     a_i = min{ j : (i + u0) / n < C_j / S }   <=>   c_j = #{ i : (i 2^32 + k0) S < C_j n 2^32 },  a_i = j for c_{j-1} <= i < c_j
// (128-bit integer comparison; an fp64 estimate decides except within 1e-5 of an integer).  oracle/core.py
// `ancestors_systematic_exact` restates this with Python integers; the two agree bit for bit.
#include <stdlib.h>
#include "common.cuh"
#include "rng.cuh"
#include "comm.cuh"

#define RF_THREADS 256
#define RF_ITEMS 16
#define RF_TILE (RF_THREADS * RF_ITEMS)          // 4096 particles
#define RF_CHUNK_ITEMS 24
#define RF_CHUNK (RF_THREADS * RF_CHUNK_ITEMS)   // 6144 outputs filled per shared-memory pass
#define RF_INLINE (3 * RF_CHUNK)                 // tiles with more outputs go to the heavy worklist
#define RF_HEAVY_CHUNK 8192                      // outputs per work item of the heavy pass

typedef unsigned long long u64;

struct RfHeader {                                // first 64 bytes of the caller's workspace
    u64 local_total;                             // sum of e_i over this shard (input of the totals exchange)
    unsigned heavy_count;                        // number of heavy records of this shard, reset by pass A
    unsigned done_counter;                       // last-block detection of pass A (self resetting)
    unsigned heavy_total;                        // records gathered from all ranks (pass C)
    unsigned pad0;
    u64 pad[5];
};

struct RfHeavy {                                 // a source tile with more than RF_INLINE outputs
    u64 Cex, Cin;                                // cumulative weight before / after the tile (global)
    unsigned T, pad;                             // global tile id: owner rank * ntiles + local tile
};

struct RfHeavyJob {                              // a heavy tile as seen by ONE rank: its share of the tile's outputs
    u64 Cex;                                     // cumulative weight before the tile
    u64 rot;                                     // heavy outputs of the jobs in front of this one (block partition)
    unsigned T, o_lo, o_hi, pad;                 // global tile id; output slots [o_lo, o_hi) of this rank
};

struct RfArgs {
    RfHeader* hdr; u64* prefix;                       // prefix[ntiles + 1]: exclusive tile prefix, last = shard total
    RfHeavy* heavy; RfHeavyJob* jobs;                 // own records [ntiles + 1]; this rank's jobs [world (ntiles + 1)]
    const float* in; int64_t n; int64_t ntiles;       // this shard's weights; ntiles tiles per shard
    int log_mode; float scale;                        // e = rint(w * scale), scale = 2^K
    const mb_control* ctl; int predicated;
    long long k0;                                     // >= 0: caller supplied u0 bits; < 0: Philox(ctl->seed, ctl->iter + 1)
    const u64* totals; int rank, world;               // shard totals (device, [world]) or NULL (single shard)
    int64_t n_local, n_total;
    int32_t* anc_peers[MB_MAX_WORLD];                 // ancestor array of every rank (peer mapped); [0] = anc for one shard
    const float* in_peers[MB_MAX_WORLD];              // weights of every rank (pass C reads the heavy tiles of any rank)
    const void* ws_peers[MB_MAX_WORLD];               // workspace of every rank (its heavy records)
};

__device__ __forceinline__ float rf_wmax(const RfArgs& a) {
    float wmax = 0.f;
    if (a.log_mode) {
        wmax = (float)a.ctl->wmax;
        if (!(wmax > -INFINITY) || wmax == INFINITY) wmax = 0.f;      // jax logsumexp convention: non-finite max -> 0
    }
    return wmax;
}

// Integer weight e = rint(x) as uint64.  Three bit-identical formulations, measured at n = 1e8 (pass A / pass B + C, ms;
// gpurun_out/call6, profiles/README.md):  0 = the 64-bit conversion instruction (XU pipe)        0.128 / 0.63 (collapsed), 0.68 (flat)
//                                          1 = mantissa shift (branchy, ALU)                      0.176 / 0.62, 0.72
//                                          2 = the split below (branch-free, FMA / ALU pipes)     0.172 / 0.62, 0.72
// The conversion wins pass A by 45 us -- the XU pipe is not what bounds it -- and is the default; the others stay
// selectable with -DMB_RF_RINT_MODE for the record.
// Split: x = hr 2^20 + lo with hr = rint(x 2^-20) and lo = x - hr 2^20 EXACT in fp32 (|lo| <= 2^19), both
// turned into integers by adding 1.5 * 2^23 (round-to-nearest-even lands in the mantissa); hr 2^20 is even, so the tie
// rule of rint(lo) is that of rint(x).  NaN, negative, zero -> 0 (as the saturating conversion gives); x >= 2^42 (never
// in log mode, where x <= 2^40) takes the conversion instruction.
__device__ __forceinline__ u64 rf_rint_u64(float x) {
    if (x >= 4398046511104.f) return __float2ull_rn(x);
    const float MAGIC = 12582912.f;                                   // 1.5 * 2^23
    const float th = fmaf(x, 9.5367431640625e-7f, MAGIC);             // x 2^-20 + magic (one rounding: RNE to an integer)
    const int hi = (int)(__float_as_uint(th) - 0x4B400000u);
    const float hr = th - MAGIC;
    const float lo = fmaf(-hr, 1048576.f, x);
    const int li = (int)(__float_as_uint(lo + MAGIC) - 0x4B400000u);
    const long long e = ((long long)hi << 20) + (long long)li;
    return (x > 0.f) ? (u64)e : 0ull;
}

// compile-time experiment switch -DMB_RF_RINT_MODE (0: the conversion instruction, 1: mantissa shift, 2: the split above);
// the three agree bit for bit
__device__ __forceinline__ u64 rf_rint_shift(float x) {
    if (!(x > 0.f)) return 0ull;
    if (x < 8388608.f) return (u64)(__float_as_uint(x + 8388608.f) - 0x4B000000u);
    const uint32_t bits = __float_as_uint(x);
    const int sh = (int)(bits >> 23) - 150;
    if (sh >= 40) return __float2ull_rn(x);
    return (u64)((bits & 0x7fffffu) | 0x800000u) << sh;
}
#ifndef MB_RF_RINT_MODE
#define MB_RF_RINT_MODE 0
#endif
__device__ __forceinline__ u64 rf_weight(float v, bool log_mode, float wmax, float scale) {
    const float w = log_mode ? __expf(v - wmax) : v;
    const float x = w * scale;
#if MB_RF_RINT_MODE == 0
    return __float2ull_rn(x);                                         // NaN, negative -> 0 (saturating conversion)
#elif MB_RF_RINT_MODE == 1
    return rf_rint_shift(x);
#else
    return rf_rint_u64(x);                                            // NaN, negative -> 0
#endif
}

// 16 consecutive weights of this thread -> integer weights e[] (zero beyond n)
__device__ __forceinline__ void rf_load(const RfArgs& a, const float* in, int64_t base, float wmax, u64 (&e)[RF_ITEMS]) {
    if (base + RF_ITEMS <= a.n && (((uintptr_t)in & 15) == 0)) {
        const float4* p = reinterpret_cast<const float4*>(in + base);
#pragma unroll
        for (int k = 0; k < RF_ITEMS / 4; ++k) {
            const float4 v = __ldg(p + k);
            e[4 * k + 0] = rf_weight(v.x, a.log_mode, wmax, a.scale); e[4 * k + 1] = rf_weight(v.y, a.log_mode, wmax, a.scale);
            e[4 * k + 2] = rf_weight(v.z, a.log_mode, wmax, a.scale); e[4 * k + 3] = rf_weight(v.w, a.log_mode, wmax, a.scale);
        }
    } else {
#pragma unroll
        for (int k = 0; k < RF_ITEMS; ++k) e[k] = (base + k < a.n) ? rf_weight(in[base + k], a.log_mode, wmax, a.scale) : 0ull;
    }
}

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MB_FULL, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------ pass A
__global__ void __launch_bounds__(RF_THREADS) rf_tile_sums_kernel(RfArgs a) {
    if (a.predicated && (a.ctl->done || !a.ctl->resample)) return;
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float wmax = rf_wmax(a);
    // one WARP per tile (32 float4 per lane, 8 in flight): no block barrier, no shared memory in the streaming part
    for (int64_t tile = (int64_t)blockIdx.x * (RF_THREADS / 32) + warp; tile < a.ntiles; tile += (int64_t)gridDim.x * (RF_THREADS / 32)) {
        u64 s = 0;
        const int64_t tb = tile * RF_TILE;
        if (tb + RF_TILE <= a.n && (((uintptr_t)a.in & 15) == 0)) {
            const float4* p = reinterpret_cast<const float4*>(a.in + tb) + lane;
#pragma unroll 8
            for (int q = 0; q < RF_TILE / 128; ++q) {
                const float4 v = __ldcs(p + q * 32);
                s += rf_weight(v.x, a.log_mode, wmax, a.scale) + rf_weight(v.y, a.log_mode, wmax, a.scale) +
                     rf_weight(v.z, a.log_mode, wmax, a.scale) + rf_weight(v.w, a.log_mode, wmax, a.scale);
            }
        } else {
            for (int q = 0; q < RF_TILE / 32; ++q) {
                const int64_t i = tb + q * 32 + lane;
                if (i < a.n) s += rf_weight(a.in[i], a.log_mode, wmax, a.scale);
            }
        }
        s = warp_sum_u64(s);
        if (lane == 0) a.prefix[tile] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(&a.hdr->done_counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // exclusive scan of the tile sums by the last block, in place: coalesced chunks of 256 x 8 sums (the loads of the
    // next chunk are in flight while this one is scanned); integer sums, so the association is free
    __shared__ u64 wtot[RF_THREADS / 32];
    constexpr int PER = 8;
    u64 carry = 0, v[PER], nx[PER];
    auto load = [&](int64_t base, u64 (&dst)[PER]) {
        const int64_t idx = base + (int64_t)threadIdx.x * PER;
#pragma unroll
        for (int k = 0; k < PER; ++k) dst[k] = (idx + k < a.ntiles) ? __ldcg(a.prefix + idx + k) : 0ull;
    };
    load(0, v);
    for (int64_t base = 0; base < a.ntiles; base += RF_THREADS * PER) {
        if (base + RF_THREADS * PER < a.ntiles) load(base + RF_THREADS * PER, nx);
        u64 tsum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) tsum += v[k];
        u64 incl = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(MB_FULL, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        u64 run = carry + incl - tsum, total = 0;
#pragma unroll
        for (int w = 0; w < RF_THREADS / 32; ++w) { const u64 t = wtot[w]; if (w < warp) run += t; total += t; }
        const int64_t idx = base + (int64_t)threadIdx.x * PER;
#pragma unroll
        for (int k = 0; k < PER; ++k) { if (idx + k < a.ntiles) a.prefix[idx + k] = run; run += v[k]; }
        carry += total;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; ++k) v[k] = nx[k];
    }
    if (threadIdx.x == 0) {
        a.prefix[a.ntiles] = carry;
        a.hdr->local_total = carry;
        a.hdr->heavy_count = 0;
        a.hdr->done_counter = 0;
    }
}

// ------------------------------------------------------------------------------------------------ counts
struct RfSys {                     // systematic grid seen from the cumulative weights
    u64 S; u64 n; unsigned k0; double rho, u0;
};

// is (i 2^32 + k0) S < C n 2^32 ?   (exact, 128-bit)
__device__ __forceinline__ bool rf_below(const RfSys& g, u64 i, u64 C) {
    const u64 A = (i << 32) | (u64)g.k0;
    const u64 lh = __umul64hi(A, g.S), ll = A * g.S;
    const u64 xh = __umul64hi(C, g.n), xl = C * g.n;
    const u64 rh = (xh << 32) | (xl >> 32), rl = xl << 32;
    return lh < rh || (lh == rh && ll < rl);
}

// c(C) = #{ i in [0, n) : (i + u0)/n < C/S }: fp64 estimate, exact fix-up when the estimate is within 1e-5 of an integer
__device__ __forceinline__ unsigned rf_count(const RfSys& g, u64 C) {
    if (C == 0) return 0u;
    if (C >= g.S) return (unsigned)g.n;
    const double t = (double)C * g.rho - g.u0;            // outputs i < t lie below C
    double ce = ceil(t);
    if (ce < 0.0) ce = 0.0;
    if (ce > (double)g.n) ce = (double)g.n;
    u64 c = (u64)ce;
    const double fr = t - floor(t);
    if (fr < 1e-5 || fr > 1.0 - 1e-5) {                   // rare: settle with integers
        while (c > 0 && !rf_below(g, c - 1, C)) --c;
        while (c < g.n && rf_below(g, c, C)) ++c;
    }
    return (unsigned)c;
}

// systematic grid + the cumulative weight in front of every rank (offs[r], r < world; written by thread 0)
__device__ __forceinline__ RfSys rf_grid_setup(const RfArgs& a, u64* offs) {
    RfSys g;
    u64 S = 0;
    if (a.totals) {
        for (int r = 0; r < a.world; ++r) { if (threadIdx.x == 0) offs[r] = S; S += a.totals[r]; }
    } else {
        if (threadIdx.x == 0) offs[0] = 0;
        S = a.prefix[a.ntiles];
    }
    g.S = S; g.n = (u64)a.n_total;
    if (a.k0 >= 0) g.k0 = (unsigned)a.k0;
    else g.k0 = philox_raw(a.ctl->seed, 0ull, (uint32_t)(a.ctl->iter + 1), MB_P_RESAMPLE, 0u).x;
    g.rho = (S > 0) ? (double)g.n / (double)S : 0.0;
    g.u0 = (double)g.k0 * 2.3283064365386963e-10;
    return g;
}

// block-wide scan of this thread's 16 integer weights -> exclusive offset of the thread inside the tile
__device__ __forceinline__ u64 rf_block_exclusive(u64 thread_total, u64* warp_tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 incl = thread_total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(MB_FULL, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    u64 off = 0;
#pragma unroll
    for (int w = 0; w < RF_THREADS / 32; ++w) if (w < warp) off += warp_tot[w];
    __syncthreads();
    return off + incl - thread_total;
}

// ------------------------------------------------------------------------------------------------ pass B
__global__ void __launch_bounds__(RF_THREADS) rf_ancestors_kernel(RfArgs a) {
    if (a.predicated && (a.ctl->done || !a.ctl->resample)) return;
    __shared__ u64 warp_tot[RF_THREADS / 32];
    __shared__ int buf[RF_CHUNK];
    __shared__ int wmaxs[RF_THREADS / 32];
    __shared__ u64 offs[MB_MAX_WORLD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float wmax = rf_wmax(a);
    const RfSys g = rf_grid_setup(a, offs);
    __syncthreads();
    if (g.S == 0) {                                       // all weights zero: legacy convention cdf[n-1] = 1
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x)
            a.anc_peers[a.world <= 1 ? 0 : a.rank][i] = (int32_t)(a.n_total - 1);
        return;
    }
    const int64_t gid0 = (int64_t)a.rank * a.n_local;     // global id of this shard's first particle
    const u64 offset = offs[a.rank];

    for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        // the tile's output range follows from the tile prefix alone (block-uniform): a tile without offspring -- nearly
        // every tile of a collapsed population -- is skipped WITHOUT reading its weights, a heavy one is only recorded
        const u64 Cex = offset + a.prefix[tile], Cin = offset + a.prefix[tile + 1];
        const unsigned o_lo = rf_count(g, Cex), o_hi = rf_count(g, Cin);
        if (o_hi == o_lo) continue;                       // no offspring in this tile
        if (o_hi - o_lo > RF_INLINE) {                    // collapsed weights: every rank fills its own share later (pass C)
            if (threadIdx.x == 0) {
                RfHeavy h;
                h.Cex = Cex; h.Cin = Cin;
                h.T = (unsigned)((int64_t)a.rank * a.ntiles + tile); h.pad = 0;
                a.heavy[atomicAdd(&a.hdr->heavy_count, 1u)] = h;
            }
            continue;
        }
        u64 e[RF_ITEMS];
        rf_load(a, a.in, tile * RF_TILE + (int64_t)threadIdx.x * RF_ITEMS, wmax, e);
        u64 tot = 0;
#pragma unroll
        for (int k = 0; k < RF_ITEMS; ++k) tot += e[k];
        u64 C = Cex + rf_block_exclusive(tot, warp_tot);
        // outputs below the cumulative weight: c_prev at the thread's exclusive prefix, then after every particle
        // (a thread whose 16 particles have no offspring between them -- the rule in the light tail of a collapsed
        // population -- settles with the two counts at the ends of its range)
        unsigned c[RF_ITEMS + 1];
        c[0] = rf_count(g, C);
        const bool barren = tot == 0 || rf_count(g, C + tot) == c[0];
#pragma unroll
        for (int k = 0; k < RF_ITEMS; ++k) {
            C += e[k];
            c[k + 1] = (barren || e[k] == 0) ? c[k] : rf_count(g, C);
        }
        const int32_t base = (int32_t)(gid0 + tile * RF_TILE) - 1;
        for (unsigned chunk_lo = o_lo; chunk_lo < o_hi; chunk_lo += RF_CHUNK) {
#pragma unroll
            for (int q2 = 0; q2 < RF_CHUNK_ITEMS; ++q2) buf[q2 * RF_THREADS + threadIdx.x] = 0;
            __syncthreads();
            // head marker (local particle index + 1) at the first output of every particle whose offspring meet the
            // chunk; a particle that started before the chunk marks the chunk's first slot (at most one does)
#pragma unroll
            for (int k = 0; k < RF_ITEMS; ++k) {
                const unsigned first = max(c[k], chunk_lo);
                if (c[k + 1] > first && first - chunk_lo < RF_CHUNK)
                    buf[first - chunk_lo] = threadIdx.x * RF_ITEMS + k + 1;
            }
            __syncthreads();
            // max-scan of the markers (local indices increase with the output slot): thread t owns 24 consecutive slots
            int v[RF_CHUNK_ITEMS];
            int run = 0;
#pragma unroll
            for (int q2 = 0; q2 < RF_CHUNK_ITEMS; ++q2) { run = max(run, buf[threadIdx.x * RF_CHUNK_ITEMS + q2]); v[q2] = run; }
            int incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(MB_FULL, incl, o); if (lane >= o) incl = max(incl, t); }
            if (lane == 31) wmaxs[warp] = incl;
            int excl = __shfl_up_sync(MB_FULL, incl, 1);
            if (lane == 0) excl = 0;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < RF_THREADS / 32; ++w) if (w < warp) excl = max(excl, wmaxs[w]);
#pragma unroll
            for (int q2 = 0; q2 < RF_CHUNK_ITEMS; ++q2) buf[threadIdx.x * RF_CHUNK_ITEMS + q2] = max(v[q2], excl);
            __syncthreads();
            const unsigned cnt = min((unsigned)RF_CHUNK, o_hi - chunk_lo);
            // destination of the run of outputs [chunk_lo, chunk_lo + cnt): at most one shard boundary inside (the host
            // requires n_local >= RF_CHUNK for sharded calls); one 64-bit division per chunk instead of one per output
            int r0 = 0;
            int64_t split = INT64_MAX;
            if (a.world > 1) { r0 = (int)((int64_t)chunk_lo / a.n_local); split = (int64_t)(r0 + 1) * a.n_local; }
#pragma unroll
            for (int q2 = 0; q2 < RF_CHUNK_ITEMS; ++q2) {
                const unsigned sl = q2 * RF_THREADS + threadIdx.x;
                if (sl >= cnt) continue;
                const int64_t o = (int64_t)chunk_lo + sl;
                const int r = (o < split) ? r0 : r0 + 1;
                a.anc_peers[r][o - (int64_t)r * a.n_local] = base + buf[sl];   // slot inside the owner's array
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------ pass C
// collect the heavy records of every rank (written by their pass B, complete after the caller's barrier) and turn them
// into this rank's JOBS: the part of each heavy tile's output range that falls into the rank's own slots, with the
// counts evaluated once here instead of by every block of pass C
__global__ void __launch_bounds__(RF_THREADS) rf_gather_heavy_kernel(RfArgs a) {
    if (a.predicated && (a.ctl->done || !a.ctl->resample)) return;
    __shared__ unsigned cnt[MB_MAX_WORLD + 1];
    __shared__ u64 offs[MB_MAX_WORLD];
    __shared__ unsigned njobs;
    const RfSys g = rf_grid_setup(a, offs);
    if (threadIdx.x < a.world)
        cnt[threadIdx.x + 1] = reinterpret_cast<const RfHeader*>(a.ws_peers[threadIdx.x])->heavy_count;   // peer reads in parallel
    if (threadIdx.x == 0) { cnt[0] = 0; njobs = 0; }
    __syncthreads();
    if (threadIdx.x == 0) for (int q = 0; q < a.world; ++q) cnt[q + 1] += cnt[q];
    __syncthreads();
    const int64_t o_begin = (int64_t)a.rank * a.n_local, o_end = o_begin + a.n;      // this rank's output slots
    const int64_t off_heavy = (int64_t)sizeof(RfHeader) + (int64_t)sizeof(u64) * (a.ntiles + 1);
    for (int q = 0; q < a.world; ++q) {
        const RfHeavy* src = reinterpret_cast<const RfHeavy*>(reinterpret_cast<const char*>(a.ws_peers[q]) + off_heavy);
        const unsigned hq = cnt[q + 1] - cnt[q];
        for (unsigned i = threadIdx.x; i < hq; i += blockDim.x) {
            const RfHeavy h = src[i];
            const unsigned o_lo = max(rf_count(g, h.Cex), (unsigned)o_begin), o_hi = min(rf_count(g, h.Cin), (unsigned)o_end);
            if (o_hi <= o_lo) continue;                   // none of this tile's offspring live here
            RfHeavyJob j;
            j.Cex = h.Cex; j.rot = 0; j.T = h.T; j.o_lo = o_lo; j.o_hi = o_hi; j.pad = 0;
            a.jobs[atomicAdd(&njobs, 1u)] = j;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                               // running count of heavy outputs (the order of the jobs is
        u64 rot = 0;                                      // irrelevant: every slot is written once, with an exact value)
        for (unsigned i = 0; i < njobs; ++i) {
            a.jobs[i].rot = rot;
            rot += (u64)(a.jobs[i].o_hi - a.jobs[i].o_lo);
        }
        a.hdr->heavy_total = njobs;
        a.hdr->pad[0] = rot;                              // heavy outputs of this rank in total
    }
}

__global__ void __launch_bounds__(RF_THREADS, 4) rf_heavy_kernel(RfArgs a) {
    if (a.predicated && (a.ctl->done || !a.ctl->resample)) return;
    const unsigned H = a.hdr->heavy_total;
    if (H == 0) return;
    __shared__ u64 warp_tot[RF_THREADS / 32];
    __shared__ unsigned cs[RF_TILE + 1];                  // cs[j] = outputs below particle j's EXCLUSIVE prefix; cs[4096] = o_hi
    __shared__ u64 offs[MB_MAX_WORLD];
    const float wmax = rf_wmax(a);
    const RfSys g = rf_grid_setup(a, offs);
    const int64_t o_begin = (int64_t)a.rank * a.n_local;
    int32_t* anc = a.anc_peers[a.world <= 1 ? 0 : a.rank];
    __syncthreads();
    // every block fills ONE contiguous share of this rank's heavy outputs (in the order of the jobs): it meets one or two
    // jobs and derives their tiles once -- the first version rotated 8192-output work items over the blocks, so that
    // every block re-derived (load, scan, counts) nearly every heavy tile for a single item of work
    const u64 Htot = a.hdr->pad[0];
    const u64 R_lo = Htot * blockIdx.x / gridDim.x, R_hi = Htot * (blockIdx.x + 1) / gridDim.x;
    if (R_lo >= R_hi) return;
    unsigned en = 0;                                      // last job that starts at or before R_lo (rot increases with the job)
    for (unsigned lo_j = 0, hi_j = H; lo_j < hi_j;) {
        const unsigned mid = (lo_j + hi_j) >> 1;
        if (a.jobs[mid].rot <= R_lo) { en = mid; lo_j = mid + 1; } else hi_j = mid;
    }
    for (; en < H; ++en) {
        const RfHeavyJob h = a.jobs[en];
        if (h.rot >= R_hi) break;                         // block-uniform
        const u64 s_lo = max(R_lo, h.rot), s_hi = min(R_hi, h.rot + (u64)(h.o_hi - h.o_lo));
        if (s_lo >= s_hi) continue;
        const unsigned w_lo = h.o_lo + (unsigned)(s_lo - h.rot), w_hi = h.o_lo + (unsigned)(s_hi - h.rot);
        const int q = (int)(h.T / (unsigned)a.ntiles);
        const int64_t tile = (int64_t)h.T - (int64_t)q * a.ntiles;
        u64 e[RF_ITEMS];
        rf_load(a, a.in_peers[q], tile * RF_TILE + (int64_t)threadIdx.x * RF_ITEMS, wmax, e);
        u64 tot = 0;
#pragma unroll
        for (int k = 0; k < RF_ITEMS; ++k) tot += e[k];
        u64 C = h.Cex + rf_block_exclusive(tot, warp_tot);
        unsigned cprev = rf_count(g, C);
#pragma unroll
        for (int k = 0; k < RF_ITEMS; ++k) {
            cs[threadIdx.x * RF_ITEMS + k] = cprev;
            C += e[k];
            if (e[k] != 0) cprev = rf_count(g, C);
        }
        if (threadIdx.x == RF_THREADS - 1) cs[RF_TILE] = cprev;
        __syncthreads();
        const int32_t base = (int32_t)((int64_t)q * a.n_local + tile * RF_TILE);
        // ancestor = last particle j with cs[j] <= o among those with offspring: upper_bound(cs, o) - 1.  A thread's
        // outputs increase, so does j: the previous answer is tried first (one shared-memory read) -- with collapsed
        // weights a handful of particles own nearly every output and the binary search runs a few times per share
        // `nb` = first output that no longer belongs to the current answer: the loop body is a compare and a store
        int lo = 0;
        unsigned nb = 0;                                  // forces the first search
        int32_t val = 0;
        int32_t* const out = anc - (int64_t)(unsigned)o_begin;
        unsigned o = w_lo + threadIdx.x;
        for (; o < w_hi; ) {
            if (o >= nb) {
                int hi = RF_TILE;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (cs[mid + 1] > o) hi = mid; else lo = mid + 1; }
                nb = cs[lo + 1];
                val = base + lo;
            }
            const unsigned o3 = o + 3 * RF_THREADS;
            if (o3 < nb && o3 < w_hi) {                   // four outputs of the same ancestor
                out[o] = val; out[o + RF_THREADS] = val; out[o + 2 * RF_THREADS] = val; out[o3] = val;
                o += 4 * RF_THREADS;
            } else {
                out[o] = val;
                o += RF_THREADS;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ host
static int rf_scale_bits(int64_t n_total) {
    int lg = 0;
    while (((int64_t)1 << lg) < n_total) ++lg;
    int K = 63 - lg;
    return K > 40 ? 40 : K;
}

extern "C" size_t mb_rs_workspace_bytes(int64_t n) {
    // header | own tile prefix | own heavy records | this rank's jobs on the heavy tiles of all ranks (MB_MAX_WORLD shards)
    const int64_t ntiles = (n + RF_TILE - 1) / RF_TILE;
    return sizeof(RfHeader) + sizeof(u64) * (size_t)(ntiles + 1) + sizeof(RfHeavy) * (size_t)(ntiles + 1) +
           sizeof(RfHeavyJob) * (size_t)(ntiles + 1) * MB_MAX_WORLD;
}

static void rf_fill(RfArgs& a, void* ws, const float* in, int64_t n, int64_t n_total, int log_mode, const mb_control* ctl,
                    int force) {
    a.ntiles = (n + RF_TILE - 1) / RF_TILE;
    a.hdr = (RfHeader*)ws;
    a.prefix = (u64*)((char*)ws + sizeof(RfHeader));
    a.heavy = (RfHeavy*)(a.prefix + a.ntiles + 1);
    a.jobs = (RfHeavyJob*)(a.heavy + (a.ntiles + 1));
    a.in = in; a.n = n; a.n_total = n_total; a.log_mode = log_mode;
    a.scale = ldexpf(1.f, rf_scale_bits(n_total));
    a.ctl = ctl; a.predicated = (ctl && !force) ? 1 : 0;
}

extern "C" int mb_rs_tile_sums(mb_ctx* ctx, void* ws, const float* in, int64_t n, int64_t n_total, int log_mode,
                               const mb_control* ctl, int force, mb_stream_t stream) {
    MB_REQUIRE(ctx && ws && in && n > 0 && n_total >= n && n_total < 0x7fffffffll, "mb_rs_tile_sums: bad arguments");
    MB_REQUIRE(!log_mode || ctl, "mb_rs_tile_sums: log mode needs the control block (max log-weight)");
    RfArgs a{};
    rf_fill(a, ws, in, n, n_total, log_mode, ctl, force);
    int64_t grid = (a.ntiles + RF_THREADS / 32 - 1) / (RF_THREADS / 32);   // one warp per tile
    if (grid > (int64_t)ctx->sms * 8) grid = (int64_t)ctx->sms * 8;
    rf_tile_sums_kernel<<<(unsigned)grid, RF_THREADS, 0, mb_s(stream)>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

static int rf_shard_args(RfArgs& a, void* ws, const float* in, int64_t n, int64_t n_total, int64_t k0,
                         const unsigned long long* totals, const mb_shard* sh, int32_t* anc, const char* who) {
    a.k0 = k0;
    a.rank = 0; a.world = 1; a.n_local = n; a.anc_peers[0] = anc; a.in_peers[0] = in; a.ws_peers[0] = ws;
    if (sh && sh->world > 1) {
        if (!(totals && sh->n_local == n && sh->n_total == n_total)) { mb_set_error("%s: sharded call needs the shard totals", who); return MB_ERR_ARG; }
        if (n < RF_HEAVY_CHUNK) { mb_set_error("%s: shards of a sharded population hold at least 8192 particles", who); return MB_ERR_ARG; }
        a.totals = totals; a.rank = sh->rank; a.world = sh->world; a.n_local = sh->n_local;
        for (int r = 0; r < sh->world; ++r) {
            if (!sh->anc_peers[r] || !sh->lw_peers[r] || !sh->ws_peers[r]) { mb_set_error("%s: anc_peers / lw_peers / ws_peers missing", who); return MB_ERR_ARG; }
            a.anc_peers[r] = sh->anc_peers[r]; a.in_peers[r] = sh->lw_peers[r]; a.ws_peers[r] = sh->ws_peers[r];
        }
    } else if (n_total != n) {
        mb_set_error("%s: n_total != n needs a shard description", who);
        return MB_ERR_ARG;
    }
    return MB_OK;
}

extern "C" int mb_rs_ancestors(mb_ctx* ctx, void* ws, const float* in, int64_t n, int64_t n_total, int log_mode,
                               const mb_control* ctl, int force, int64_t k0, const unsigned long long* totals,
                               const mb_shard* sh, int32_t* anc, mb_stream_t stream) {
    MB_REQUIRE(ctx && ws && in && anc && n > 0 && n_total >= n && n_total < 0x7fffffffll, "mb_rs_ancestors: bad arguments");
    MB_REQUIRE(!log_mode || ctl, "mb_rs_ancestors: log mode needs the control block");
    MB_REQUIRE(k0 >= 0 || ctl, "mb_rs_ancestors: k0 < 0 draws u0 from Philox(ctl->seed, ctl->iter + 1)");
    MB_REQUIRE(k0 <= 0xffffffffll, "mb_rs_ancestors: k0 is a 32-bit fraction");
    RfArgs a{};
    rf_fill(a, ws, in, n, n_total, log_mode, ctl, force);
    int rc = rf_shard_args(a, ws, in, n, n_total, k0, totals, sh, anc, "mb_rs_ancestors");
    if (rc != MB_OK) return rc;
    cudaStream_t st = mb_s(stream);
    int64_t grid = a.ntiles;
    if (grid > (int64_t)ctx->sms * 6) grid = (int64_t)ctx->sms * 6;
    rf_ancestors_kernel<<<(unsigned)grid, RF_THREADS, 0, st>>>(a);
    MB_CHECK_LAUNCH();
    if (a.world > 1) return MB_OK;                        // sharded: mb_rs_heavy after a barrier over the ranks
    rf_gather_heavy_kernel<<<1, RF_THREADS, 0, st>>>(a);
    MB_CHECK_LAUNCH();
    rf_heavy_kernel<<<(unsigned)(ctx->sms * 4), RF_THREADS, 0, st>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

extern "C" int mb_rs_heavy(mb_ctx* ctx, void* ws, const float* in, int64_t n, int64_t n_total, int log_mode,
                           const mb_control* ctl, int force, int64_t k0, const unsigned long long* totals,
                           const mb_shard* sh, int32_t* anc, mb_stream_t stream) {
    MB_REQUIRE(ctx && ws && in && anc && sh && sh->world > 1, "mb_rs_heavy: sharded populations only (a single shard is finished by mb_rs_ancestors)");
    RfArgs a{};
    rf_fill(a, ws, in, n, n_total, log_mode, ctl, force);
    int rc = rf_shard_args(a, ws, in, n, n_total, k0, totals, sh, anc, "mb_rs_heavy");
    if (rc != MB_OK) return rc;
    cudaStream_t st = mb_s(stream);
    rf_gather_heavy_kernel<<<1, RF_THREADS, 0, st>>>(a);
    MB_CHECK_LAUNCH();
    rf_heavy_kernel<<<(unsigned)(ctx->sms * 4), RF_THREADS, 0, st>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}
