// svgd_tc.cu -- K8 on the 5th-generation tensor cores: the SVGD interaction
//     phi = [ -K G + (X o rowsum(K) - K X)/h^2 ] / n,   K_ij = exp(-|x_i - x_j|^2 / (2 h^2))      (transport/svgd.py:18-32)
// as a FlashAttention-shaped tcgen05 pipeline (Q = K = X, V = [G, X, 1], un-normalised exp):
//
//   MMA1 (tcgen05.mma kind::f16, bf16 x bf16 -> fp32 in TMEM):  S = A_i B_j^T = exp2 argument -|x_i-x_j|^2 log2e/(2h^2)
//        coordinates centred and scaled by sc = sqrt(log2 e)/h; squared norms folded into the K dimension:
//        A_i = [x_i, 1, 1, s_i^hi, s_i^lo, 0..],  B_j = [x_j, s_j^hi, s_j^lo, 1, 1, ..],  s = -1/2 |x|^2 split in two bf16;
//   softmax warps (tcgen05.ld):  P = exp2(S) -> bf16 -> back into TMEM (tcgen05.st), never shared memory;
//   MMA2:  O += P W_j  (A operand = P from TMEM, B = W^T tile from smem; O stays in TMEM for the whole j loop);
//   finish kernel:  phi_i = ( -O_G + (x_i O_1 - O_X)/h^2 ) / n.
//
// ONE operand tile per j-tile serves both GEMMs.  W^T[c][j] (rows c, 128 j's, SWIZZLE_128B, two 64-wide panels) holds
//   rows 0..d-1: scaled x_j | d, d+1: s_j^hi, s_j^lo | d+2, d+3: 1 | d+4..2d+3: g_j | zero padding to NV rows.
// MMA2 reads it as the K-major B operand (K = j).  MMA1 reads rows 0..63 of the SAME bytes as an MN-major B operand
// (N = j contiguous, K = row): B_j above is exactly column j of those rows (the g rows meet zero columns of A).
// This removed the separate X_j tile: the kernel is bound by L2->SM bandwidth (every CTA streams all of W^T), and
// bytes per tile went 44 KiB -> 28 KiB.  Tiles are prepared once per iteration in the UMMA canonical layout, so the
// producer warp moves them with plain 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), 4 stages deep.
// X is centred (distances are translation invariant) before rounding to bf16 ("SVGD numerics").
//
// Warp roles (608 threads, 1 CTA/SM): warp 0 = TMA producer, warp 1 = MMA1 issuer + TMEM allocator, warp 2 = MMA2
// issuer, warps 3-18 = softmax / epilogue in two ping-pong groups: each thread owns one row
// (TMEM lane) and one 64-column half of every S tile, so every SM sub-partition has two warps to overlap the
// MUFU.EX2 issue interval of one with the FMUL / pack / store instructions of the other.
#include <cuda_bf16.h>
#include "common.cuh"

#define TC_BM 128
#define TC_BN 128
#define TC_K 64
#define TC_STAGES 6
#define TC_THREADS 608                               // 3 control warps + 16 softmax warps (4 per SM sub-partition)
#define TC_SM_THREADS 512
#define TC_SPLIT 4                                   // j range split: tiles*4 CTAs on 148 SMs -> < 2 % wave tail
#define TC_TILE_X_BYTES (TC_BM * TC_K * 2)          // 16384
#define TC_P_BYTES (TC_BM * TC_BN * 2)              // 32768

// ---- PTX helpers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(acc) : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (sm_100 descriptor version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);               // start address, LBO = 16 B (unused)
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);               // SBO = 1024 B, version 1, SWIZZLE_128B
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
// same bytes viewed MN-major (N contiguous): 64-element rows repeat every `lbo` bytes along N, 8-row K groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo) {
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N, int b_mn_major = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
#define TMEM_LD16(taddr, v)                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),     \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
                 : "r"(taddr))
#define TMEM_LD32(taddr, v)                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"    \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                            \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),     \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
                 : "r"(taddr))
#define TMEM_ST32(taddr, v)                                                                                        \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16," \
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"                                   \
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),       \
                   "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),     \
                   "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory")
#define TMEM_ST16(taddr, v)                                                                                        \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory")
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]^T : the P operand never touches shared memory
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc),
        "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// exp2 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax on [-0.5, 0.5], max rel. error 7.5e-5, far below
// the bf16 rounding of P): a fraction of the exponentials is taken off the MUFU pipe, which otherwise paces the
// whole pipeline (n^2 = 1.07e9 exps per iteration at 16 lanes/clk/SM).
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.f);
    const float xi = x + 12582912.f;                   // 1.5 * 2^23: integer part lands in the low mantissa bits
    const float f = x - (xi - 12582912.f);
    const float p = fmaf(fmaf(fmaf(0.05517167f, f, 0.24261113f), f, 0.69326097f), f, 0.99992806f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(xi) << 23));
}
#ifndef TC_POLY_NUM
#define TC_POLY_NUM 3                                  // of every 16 exponentials, this many use ex2_poly (measured: 0:0.417 2:.. 3:0.378 4:0.389 5:0.400 8:0.446 ms)
#endif
__device__ __forceinline__ constexpr bool tc_use_poly(int e) { return ((e * TC_POLY_NUM) & 15) < TC_POLY_NUM && TC_POLY_NUM > 0; }

// byte offset of element (row r, k) inside a K-major SWIZZLE_128B panel of 64 bf16 per row
__host__ __device__ __forceinline__ uint32_t sw128_off(int r, int k) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

// ---- operand preparation ------------------------------------------------------------------------
// column means, coalesced: 64 column lanes x 4 row groups per block, fp64 partials, fixed-order final sum
__global__ void __launch_bounds__(256) svgd_tc_colsum_kernel(const float* __restrict__ X, int n, int d, double* part) {
    __shared__ double sm[4][64];
    const int c = threadIdx.x & 63, rg = threadIdx.x >> 6;
    double s = 0.0;
    if (c < d) {                                                       // 8 independent loads in flight per thread
        int i = blockIdx.x * 4 + rg;
        const int step = gridDim.x * 4;
        for (; i + 7 * step < n; i += 8 * step) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = X[(int64_t)(i + u * step) * d + c];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (double)v[u];
        }
        for (; i < n; i += step) s += (double)X[(int64_t)i * d + c];
    }
    sm[rg][c] = s;
    __syncthreads();
    if (rg == 0) part[blockIdx.x * 64 + c] = sm[0][c] + sm[1][c] + sm[2][c] + sm[3][c];
}
__global__ void __launch_bounds__(1024) svgd_tc_colmean_kernel(const double* part, int nblocks, int n, int d, float* mean) {
    __shared__ double sm[16][64];
    const int c = threadIdx.x & 63, g = threadIdx.x >> 6;              // 16 block groups per column, fixed order
    double s = 0.0;
#pragma unroll 8
    for (int b = g; b < nblocks; b += 16) s += __ldcg(part + b * 64 + c);
    sm[g][c] = s;
    __syncthreads();
    if (g == 0 && c < d) {
        double t = 0.0;
        for (int k = 0; k < 16; ++k) t += sm[k][c];
        mean[c] = (float)(t / (double)n);
    }
}

struct TcPrepArgs {
    const float* X; const float* G; const float* mean; const float* bandwidth; int n, d, n_pad, NV;
    uint8_t* XA; uint8_t* WT; float* xs; float* saux;   // xs: bf16-rounded scaled centred x as fp32 (n_pad x d); saux: (hi, lo)
};

// pass 1, one thread per (row, 16-byte chunk of 8 k's): A tiles, rounded scaled coordinates, split -|x|^2/2
__global__ void __launch_bounds__(256) svgd_tc_rows_kernel(TcPrepArgs a) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)a.n_pad * 8) return;
    const int row = (int)(tid >> 3), ch = (int)(tid & 7);
    const int tile = row / TC_BM, r = row % TC_BM;
    const bool valid = row < a.n;
    const float sc = a.bandwidth ? sqrtf(1.4426950408889634f) / a.bandwidth[0] : 1.f;   // NULL: plain centred coordinates
    float sq = 0.f;                                    // from the ROUNDED coordinates, so that D_ii = 0 up to the hi/lo split
    if (valid)
        for (int k = 0; k < a.d; ++k) {
            const float v = __bfloat162float(__float2bfloat16_rn((a.X[(int64_t)row * a.d + k] - a.mean[k]) * sc));
            sq = fmaf(v, v, sq);
        }
    const float s = valid ? -0.5f * sq : -1.0e30f;     // padded rows: exp2(-huge) = 0
    const __nv_bfloat16 shi = __float2bfloat16_rn(s);
    const __nv_bfloat16 slo = __float2bfloat16_rn(s - __bfloat162float(shi));
    if (ch == 0) { a.saux[2 * row] = __bfloat162float(shi); a.saux[2 * row + 1] = __bfloat162float(slo); }
    __nv_bfloat16 va[8];
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f), one = __float2bfloat16_rn(1.f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = ch * 8 + e;
        __nv_bfloat16 xa = zero;
        if (k < a.d) {
            xa = __float2bfloat16_rn(valid ? (a.X[(int64_t)row * a.d + k] - a.mean[k]) * sc : 0.f);
            a.xs[(int64_t)row * a.d + k] = __bfloat162float(xa);
        } else if (k == a.d || k == a.d + 1) xa = one;
        else if (k == a.d + 2) xa = valid ? shi : zero;
        else if (k == a.d + 3) xa = valid ? slo : zero;
        va[e] = xa;
    }
    *reinterpret_cast<uint4*>(a.XA + (int64_t)tile * TC_TILE_X_BYTES + sw128_off(r, ch * 8)) = *reinterpret_cast<uint4*>(va);
}

// pass 2, one thread per (tile, row c, 16-byte chunk of 8 j's): the W^T tiles
__global__ void __launch_bounds__(256) svgd_tc_wt_kernel(TcPrepArgs a) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)(a.n_pad / TC_BN) * a.NV * 16) return;
    const int q = (int)(tid & 15);
    const int c = (int)((tid >> 4) % a.NV);
    const int tile = (int)((tid >> 4) / a.NV);
    __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int j = tile * TC_BN + q * 8 + e;          // j < n_pad always
        float f = 0.f;
        if (c < a.d) f = a.xs[(int64_t)j * a.d + c];
        else if (c == a.d) f = a.saux[2 * j];
        else if (c == a.d + 1) f = a.saux[2 * j + 1];
        else if (c == a.d + 2 || c == a.d + 3) f = 1.f;
        else if (c < 2 * a.d + 4 && a.G) f = (j < a.n) ? a.G[(int64_t)j * a.d + (c - a.d - 4)] : 0.f;
        v[e] = __float2bfloat16_rn(f);
    }
    const uint32_t off = (uint32_t)(q >> 3) * (uint32_t)(a.NV * 128) + sw128_off(c, (q & 7) * 8);
    *reinterpret_cast<uint4*>(a.WT + (int64_t)tile * (a.NV * 256) + off) = *reinterpret_cast<uint4*>(v);
}

// ---- main kernel --------------------------------------------------------------------------------
struct TcArgs {
    const uint8_t* XA; const uint8_t* WT; const float* xs;
    const float* bandwidth; float* phi;
    int n, d, n_pad, NV;
    float* opart;            // [TC_SPLIT][n_pad][NV] partial O
    float* upart;            // logistic mode: [TC_SPLIT][n_pad][4] partial row sums of softplus
    int ncols_pad;           // padded number of columns (j): = n_pad for the SVGD interaction, the data count for the logits
    int itile0;              // first 128-row tile of this launch (a rank of a sharded ensemble computes its own rows only)
    int row_begin, row_end;  // rows whose phi is written by the finish kernel
};

// MODE 0: P = exp2(S) (SVGD interaction).  MODE 1: S is the logit w_i . a_j of a logistic regression; P = sigmoid(S)
// and the row sums of softplus(S) are accumulated on the side (the potential), O = P A is the data part of the gradient.
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) svgd_phi_tc_kernel(TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t stage_bytes = (uint32_t)a.NV * 256u;               // one W^T tile
    uint8_t* sXA = smem;
    uint8_t* sStage = sXA + TC_TILE_X_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + TC_STAGES * stage_bytes);
    uint64_t* full = bars;                    // [TC_STAGES]
    uint64_t* empty = bars + TC_STAGES;       // [TC_STAGES]
    uint64_t* s_full = bars + 2 * TC_STAGES;  // [2]
    uint64_t* s_empty = s_full + 2;           // [2]
    uint64_t* p_full = s_empty + 2;           // [2]
    uint64_t* p_empty = p_full + 2;           // [2]
    uint64_t* o_full = p_empty + 2;
    uint64_t* xa_full = o_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(xa_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T_all = a.ncols_pad / TC_BN;
    const int itile = a.itile0 + blockIdx.x / TC_SPLIT, part = blockIdx.x % TC_SPLIT;
    const int t_begin = (int)(((int64_t)T_all * part) / TC_SPLIT), t_end = (int)(((int64_t)T_all * (part + 1)) / TC_SPLIT);
    const int T = t_end - t_begin;                                     // this CTA's j tiles: [t_begin, t_end)

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 2); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(s_full + b, 1); mbar_init(s_empty + b, TC_SM_THREADS / 2);
            mbar_init(p_full + b, TC_SM_THREADS / 2); mbar_init(p_empty + b, 1);
        }
        mbar_init(o_full, 1); mbar_init(xa_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                                   // TMEM: 512 columns (S x2, O)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t tmem_S = tmem, tmem_O = tmem + 256u, tmem_P = tmem + 384u;   // P[2]: 128 x 128 bf16 = 64 columns each

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(xa_full, TC_TILE_X_BYTES);
            bulk_g2s(sXA, a.XA + (int64_t)itile * TC_TILE_X_BYTES, TC_TILE_X_BYTES, xa_full);
            for (int t = 0; t < T; ++t) {
                const int st = t % TC_STAGES, k = t / TC_STAGES;
                mbar_wait(empty + st, (uint32_t)((k & 1) ^ 1));
                mbar_expect_tx(full + st, stage_bytes);
                uint8_t* dst = sStage + (size_t)st * stage_bytes;
                bulk_g2s(dst, a.WT + (int64_t)(t_begin + t) * stage_bytes, stage_bytes, full + st);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA1 issuer: S(t) = A_i B_t^T, runs ahead of the softmax groups =====================
        // (separate issuing threads for the two GEMMs: MMA1(t+2) must not wait behind the p_full(t) that gates MMA2(t),
        //  otherwise both softmax groups idle for the MMA latency and then contend for the MUFU pipe in phase)
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc_bf16(TC_BM, TC_BN, 1);
            const uint32_t lbo = (uint32_t)a.NV * 128u;                // panel (64 j's) stride of the W^T tile
            const uint32_t aXA = smem_u32(sXA);
            mbar_wait(xa_full, 0);
            for (int t = 0; t < T; ++t) {
                const int st = t % TC_STAGES;
                mbar_wait(full + st, (uint32_t)((t / TC_STAGES) & 1));
                if (t >= 2) mbar_wait(s_empty + (t & 1), (uint32_t)(((t >> 1) - 1) & 1));
                tc_fence_after();
                const uint32_t bW = smem_u32(sStage + (size_t)st * stage_bytes);
                const uint32_t dS = tmem_S + (uint32_t)(t & 1) * TC_BN;
#pragma unroll
                for (int ks = 0; ks < TC_K / 16; ++ks)                 // B = rows 16ks..16ks+15 of W^T, MN-major
                    tc_mma_bf16(dS, umma_desc(aXA + ks * 32), umma_desc_mn(bW + ks * 2048, lbo), idesc1, ks > 0 ? 1u : 0u);
                tc_commit(s_full + (t & 1));
                tc_commit(empty + st);                                 // first of the two releases of the stage
            }
        }
    } else if (warp == 2) {
        // ===================== MMA2 issuer: O += P(t) W_t =====================
        if (lane == 0) {
            const uint32_t idesc2 = umma_idesc_bf16(TC_BM, a.NV);
            for (int t = 0; t < T; ++t) {
                const int st = t % TC_STAGES;
                mbar_wait(full + st, (uint32_t)((t / TC_STAGES) & 1));
                mbar_wait(p_full + (t & 1), (uint32_t)((t >> 1) & 1));
                tc_fence_after();
                const uint32_t bVT = smem_u32(sStage + (size_t)st * stage_bytes);
#pragma unroll
                for (int ks = 0; ks < TC_BN / 16; ++ks) {                // A = P from TMEM (8 columns per K=16 step)
                    const uint32_t pb = bVT + (uint32_t)(ks >> 2) * (uint32_t)(a.NV * 128) + (uint32_t)(ks & 3) * 32;
                    tc_mma_bf16_ts(tmem_O, tmem_P + (uint32_t)(t & 1) * 64u + (uint32_t)ks * 8u, umma_desc(pb), idesc2,
                                   (t > 0 || ks > 0) ? 1u : 0u);
                }
                tc_commit(empty + st);                                 // second release: stage smem free
                tc_commit(p_empty + (t & 1));                          // P[t&1] free
            }
            if (T > 0) tc_commit(o_full);
        }
    } else {
        // ===================== softmax + epilogue (warps 2..9) =====================
        const int wq = warp & 3;                                       // TMEM lane quarter this warp may access
        // two ping-pong groups of 8 warps: group g exponentiates the tiles of parity g (S[g] -> P[g]), so the MUFU phase
        // of one group overlaps the barrier / TMEM load / TMEM store phases of the other and the XU pipe stays busy
        const int idx = warp - 3;
        const int grp = idx >> 3;                                      // tile parity handled by this warp
        const int ch = (idx >> 2) & 1;                                 // 64-column half of the S tile
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        float usum = 0.f;
        for (int t = grp; t < T; t += 2) {
            const int u = t >> 1;                                      // use index of S[grp] / P[grp]
            mbar_wait(s_full + grp, (uint32_t)(u & 1));
            tc_fence_after();
            uint32_t packed[32], va[64];
            const uint32_t sbase = tmem_S + lane_addr + (uint32_t)(grp * TC_BN + ch * 64);
            TMEM_LD32(sbase, va);
            TMEM_LD32(sbase + 32u, (va + 32));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_empty + grp);                                // S[grp] may be overwritten by MMA1(t+2)
            if (MODE == 0) {
#pragma unroll
                for (int e = 0; e < 64; e += 2) {
                    const float p0 = tc_use_poly(e) ? ex2_poly(__uint_as_float(va[e])) : ex2f(__uint_as_float(va[e]));
                    const float p1 = tc_use_poly(e + 1) ? ex2_poly(__uint_as_float(va[e + 1])) : ex2f(__uint_as_float(va[e + 1]));
                    const __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
                    packed[e >> 1] = *reinterpret_cast<const uint32_t*>(&pk);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 64; e += 2) {
                    float p[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {                      // stable sigmoid / softplus: t = e^{-|s|} in (0, 1]
                        const float sv = __uint_as_float(va[e + q]);
                        const float t = ex2f(-1.4426950408889634f * fabsf(sv));
                        const float inv = rcp_approx(1.f + t);
                        p[q] = sv >= 0.f ? inv : t * inv;
                        usum += fmaxf(sv, 0.f) + 0.6931471805599453f * lg2_approx(1.f + t);
                    }
                    const __nv_bfloat162 pk = __floats2bfloat162_rn(p[0], p[1]);
                    packed[e >> 1] = *reinterpret_cast<const uint32_t*>(&pk);
                }
            }
            if (u >= 1) mbar_wait(p_empty + grp, (uint32_t)((u - 1) & 1));   // MMA2(t-2) has consumed P[grp]
            tc_fence_after();
            TMEM_ST32(tmem_P + lane_addr + (uint32_t)(grp * 64 + ch * 32), packed);   // this row's 64 bf16 = 32 packed columns
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_full + grp);
        }
        if (MODE == 1) a.upart[(((int64_t)part * a.n_pad + (int64_t)itile * TC_BM + row) << 2) + grp * 2 + ch] = usum;
        const int cq = idx >> 2;                                       // epilogue: 32-column quarter of O per warp
        // ---- partial O of this j range -> workspace (summed in fixed order by svgd_tc_finish_kernel)
        if (T > 0) {
            mbar_wait(o_full, 0);
            tc_fence_after();
        }
        float* orow = a.opart + ((int64_t)part * a.n_pad + (int64_t)itile * TC_BM + row) * a.NV;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int col = (cq * 2 + c) * 16;
            if (col < a.NV) {
                uint32_t v[16];
                if (T > 0) { TMEM_LD16(tmem_O + lane_addr + (uint32_t)col, v); tmem_ld_wait(); }
#pragma unroll
                for (int e = 0; e < 16; e += 4)
                    *reinterpret_cast<float4*>(orow + col + e) =
                        T > 0 ? make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                                            __uint_as_float(v[e + 3]))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

// phi_i = ( -O_G + (x_i O_1 - O_X)/h^2 ) / n  with O = sum of the TC_SPLIT partials (fixed order: deterministic).
// O columns follow the W^T rows: [0,d) = sum_j P x_j (scaled by sc), d+2 = sum_j P, [d+4, 2d+4) = sum_j P g_j.
__global__ void __launch_bounds__(256) svgd_tc_finish_kernel(TcArgs a) {
    const int64_t idx = (int64_t)a.row_begin * a.d + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)a.row_end * a.d) return;
    const int i = (int)(idx / a.d), k = (int)(idx % a.d);
    float og = 0.f, ox = 0.f, o1 = 0.f;
#pragma unroll
    for (int p = 0; p < TC_SPLIT; ++p) {
        const float* o = a.opart + ((int64_t)p * a.n_pad + i) * a.NV;
        ox += o[k]; o1 += o[a.d + 2]; og += o[a.d + 4 + k];
    }
    const float h = a.bandwidth[0];
    const float sc = sqrtf(1.4426950408889634f) / h;
    const float invn = 1.f / (float)a.n;
    a.phi[idx] = (-og + (a.xs[idx] * o1 - ox) / (sc * h * h)) * invn;
}

// ---- host ---------------------------------------------------------------------------------------
int mb_svgd_phi_tc(mb_ctx* ctx, const float* X, const float* G, int n, int d, const float* bandwidth, float* phi,
                   int row_begin, int row_count, cudaStream_t st) {
    MB_REQUIRE(d >= 1 && d + 4 <= TC_K, "svgd tcgen05 variant needs d <= 60");
    MB_REQUIRE(row_begin >= 0 && row_count > 0 && row_begin + row_count <= n && row_begin % TC_BM == 0,
               "svgd tcgen05 variant: the row range must start on a multiple of 128");
    int NV = ((2 * d + 4 + 15) / 16) * 16;
    if (NV < TC_K) NV = TC_K;                          // MMA1 reads rows 0..63 of every W^T tile
    const int n_pad = ((n + TC_BM - 1) / TC_BM) * TC_BM;
    const int tiles = n_pad / TC_BM;
    const size_t bx = (size_t)tiles * TC_TILE_X_BYTES, bw = (size_t)tiles * NV * 256;
    const size_t bs = (size_t)n_pad * d * 4, ba = (size_t)n_pad * 2 * 4, bo = (size_t)TC_SPLIT * n_pad * NV * 4;
    const size_t need = 1024 + bx + bw + bs + ba + bo + 8192;
    if (mb_ensure_scratch(ctx, (4u << 20) + need) != MB_OK) return MB_ERR_CUDA;
    uint8_t* base = (uint8_t*)ctx->scratch + (4u << 20);              // [0, 4 MiB) is used by the other kernels
    auto align = [](uint8_t* p) { return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023); };
    float* mean = reinterpret_cast<float*>(align(base));
    uint8_t* XA = align(reinterpret_cast<uint8_t*>(mean) + 256);
    uint8_t* WT = align(XA + bx);
    float* xs = reinterpret_cast<float*>(align(WT + bw));
    float* saux = reinterpret_cast<float*>(align(reinterpret_cast<uint8_t*>(xs) + bs));
    float* opart = reinterpret_cast<float*>(align(reinterpret_cast<uint8_t*>(saux) + ba));
    double* cpart = reinterpret_cast<double*>((char*)ctx->scratch + (3u << 20) + (64u << 10));   // 4 x 148 x 64 doubles
    svgd_tc_colsum_kernel<<<ctx->sms * 2, 256, 0, st>>>(X, n, d, cpart);
    svgd_tc_colmean_kernel<<<1, 1024, 0, st>>>(cpart, ctx->sms * 2, n, d, mean);
    TcPrepArgs p{X, G, mean, bandwidth, n, d, n_pad, NV, XA, WT, xs, saux};
    svgd_tc_rows_kernel<<<(unsigned)(((int64_t)n_pad * 8 + 255) / 256), 256, 0, st>>>(p);
    svgd_tc_wt_kernel<<<(unsigned)(((int64_t)tiles * NV * 16 + 255) / 256), 256, 0, st>>>(p);
    MB_CHECK_LAUNCH();
    const size_t smem = 1024 + TC_TILE_X_BYTES + (size_t)TC_STAGES * NV * 256 + 256;
    MB_CUDA(cudaFuncSetAttribute(svgd_phi_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int row_end = row_begin + row_count;
    const int itile0 = row_begin / TC_BM, itiles = (row_end + TC_BM - 1) / TC_BM - itile0;
    TcArgs a{XA, WT, xs, bandwidth, phi, n, d, n_pad, NV, opart, nullptr, n_pad, itile0, row_begin, row_end};
    svgd_phi_tc_kernel<0><<<itiles * TC_SPLIT, TC_THREADS, smem, st>>>(a);
    MB_CHECK_LAUNCH();
    svgd_tc_finish_kernel<<<(unsigned)(((int64_t)row_count * d + 255) / 256), 256, 0, st>>>(a);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

// =================================================================================================
// K9 on the tensor cores: bandwidth heuristics over the full n x n distance matrix (kernels.py:220-229,
// utils.py:437-439).  S = -|x_i - x_j|^2 / 2 comes from the same MMA1 as above (centred, unscaled bf16 coordinates,
// fp32 accumulation); the 16 consumer warps reduce every S tile straight out of TMEM:
//   mean   : sum of sqrt(-2 S)                            (fp64 per-thread accumulators, fixed-order reduction)
//   median : a bracket [lo, hi] of the median of |x_i-x_j|^2 is taken from 2^18 hashed sample pairs (exact fp32,
//            +-6 sigma of the sample quantile); ONE pass over the matrix counts the entries below the bracket and
//            histograms the <~ 2 % inside it into 2048 bins; the finish kernel walks the histogram to the two middle
//            ranks.  Resolution (hi-lo)/2048, far below the bf16 rounding of the coordinates (~1e-3 per entry, which
//            moves the median only at second order).  If the bracket misses (probability ~1e-9) the sample median is
//            used and MB_CNT_BW_FALLBACK is incremented.
#define DT_THREADS 576                               // producer + MMA warps, 16 consumer warps
#define DT_STAGES 8
#define DT_BINS 2048
#define DT_SAMPLES (1 << 18)

struct DtBracket { float bin_off, inv_bw; double lo_w, bw_d2, fallback_d2; };
struct DtArgs {
    const uint8_t* XA; const uint8_t* WT; int n, n_pad;
    const DtBracket* br; uint32_t* hist; unsigned long long* below; double* partials;
    int itile_first, itile_step;     // this launch covers the 128-row tiles itile_first + k itile_step (a rank's share)
};

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MODE>   // 0 median (count + bracket histogram), 1 mean
__global__ void __launch_bounds__(DT_THREADS, 1) svgd_dist_tc_kernel(DtArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t stage_bytes = TC_K * 256u;                      // W^T tile with NV = 64 rows
    uint8_t* sXA = smem;
    uint8_t* sStage = sXA + TC_TILE_X_BYTES;
    uint32_t* shist = reinterpret_cast<uint32_t*>(sStage + DT_STAGES * stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(shist + DT_BINS);
    uint64_t* full = bars;
    uint64_t* empty = bars + DT_STAGES;
    uint64_t* s_full = bars + 2 * DT_STAGES;
    uint64_t* s_empty = s_full + 2;
    uint64_t* xa_full = s_empty + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(xa_full + 1);
    double* red = reinterpret_cast<double*>(tmem_ptr + 2);            // [16]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T_all = a.n_pad / TC_BN;
    // SYMMETRY: D_ij = D_ji, so only the j tiles at or right of the diagonal are evaluated and every off-diagonal tile
    // counts twice -- half the n^2 distance evaluations of the full matrix.  CTA (itile, part) takes the part-th quarter
    // of [itile, T_all); CTAs are launched in order of decreasing work (longest first), which keeps the tail short.
    const int itile = a.itile_first + (blockIdx.x / TC_SPLIT) * a.itile_step, part = blockIdx.x % TC_SPLIT;
    const int T_row = T_all - itile;
    const int t_begin = itile + (int)(((int64_t)T_row * part) / TC_SPLIT), t_end = itile + (int)(((int64_t)T_row * (part + 1)) / TC_SPLIT);
    const int T = t_end - t_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < DT_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(s_full + b, 1); mbar_init(s_empty + b, TC_SM_THREADS / 2); }
        mbar_init(xa_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (MODE == 0)
        for (int i = threadIdx.x; i < DT_BINS; i += DT_THREADS) shist[i] = 0;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_S = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(xa_full, TC_TILE_X_BYTES);
            bulk_g2s(sXA, a.XA + (int64_t)itile * TC_TILE_X_BYTES, TC_TILE_X_BYTES, xa_full);
            for (int t = 0; t < T; ++t) {
                const int st = t % DT_STAGES, k = t / DT_STAGES;
                mbar_wait(empty + st, (uint32_t)((k & 1) ^ 1));
                mbar_expect_tx(full + st, stage_bytes);
                bulk_g2s(sStage + (size_t)st * stage_bytes, a.WT + (int64_t)(t_begin + t) * stage_bytes, stage_bytes, full + st);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc_bf16(TC_BM, TC_BN, 1);
            const uint32_t aXA = smem_u32(sXA);
            mbar_wait(xa_full, 0);
            for (int t = 0; t < T; ++t) {
                const int st = t % DT_STAGES;
                mbar_wait(full + st, (uint32_t)((t / DT_STAGES) & 1));
                if (t >= 2) mbar_wait(s_empty + (t & 1), (uint32_t)(((t >> 1) - 1) & 1));
                tc_fence_after();
                const uint32_t bW = smem_u32(sStage + (size_t)st * stage_bytes);
                const uint32_t dS = tmem_S + (uint32_t)(t & 1) * TC_BN;
#pragma unroll
                for (int ks = 0; ks < TC_K / 16; ++ks)
                    tc_mma_bf16(dS, umma_desc(aXA + ks * 32), umma_desc_mn(bW + ks * 2048, TC_K * 128u), idesc1, ks > 0 ? 1u : 0u);
                tc_commit(s_full + (t & 1));
                tc_commit(empty + st);
            }
        }
    } else {
        const int wq = warp & 3;
        const int idx = warp - 2;
        const int grp = idx >> 3, ch = (idx >> 2) & 1;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        const bool row_valid = itile * TC_BM + wq * 32 + lane < a.n;     // padded rows of the last i tile count nothing
        float bin_scale = 0.f, bin_off = 0.f;
        if (MODE == 0) { bin_scale = -a.br->inv_bw; bin_off = a.br->bin_off; }
        unsigned long long cnt = 0;
        double dsum = 0.0;
        for (int t = grp; t < T; t += 2) {
            const int u = t >> 1;
            mbar_wait(s_full + grp, (uint32_t)(u & 1));
            tc_fence_after();
            uint32_t va[64];
            const uint32_t sbase = tmem_S + lane_addr + (uint32_t)(grp * TC_BN + ch * 64);
            TMEM_LD32(sbase, va);
            TMEM_LD32(sbase + 32u, (va + 32));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_empty + grp);
            if (!row_valid) continue;
            const uint32_t wgt = (t_begin + t == itile) ? 1u : 2u;     // the diagonal tile holds both (i, j) and (j, i)
            if (MODE == 0) {
                // bin = round(D^2/bw - k0) through the 1.5*2^23 trick (one FFMA, no F2I on the XU pipe):
                // bin < 0: below the bracket (counted), 0 <= bin < DT_BINS: histogrammed, otherwise above (padding: +huge)
                uint32_t tcnt = 0;
#pragma unroll
                for (int e = 0; e < 64; ++e) {
                    const int b = __float_as_int(fmaf(__uint_as_float(va[e]), bin_scale, bin_off)) - 0x4B400000;
                    tcnt += (uint32_t)b >> 31;
                    if ((uint32_t)b < (uint32_t)DT_BINS) atomicAdd(&shist[b], wgt);
                }
                cnt += tcnt * wgt;
            } else {
                float ts = 0.f;
#pragma unroll
                for (int e = 0; e < 64; ++e) {
                    const float v = __uint_as_float(va[e]);
                    const float dd = sqrt_approx(fmaxf(-2.f * v, 0.f));
                    ts += (v > -1.0e29f) ? dd : 0.f;                   // padded columns carry -1e30
                }
                dsum += (double)ts * (double)wgt;
            }
        }
        if (MODE == 0) {
            unsigned long long c = cnt;
            for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(MB_FULL, c, o);
            if (lane == 0) reinterpret_cast<unsigned long long*>(red)[idx] = c;
        } else {
            for (int o = 16; o > 0; o >>= 1) dsum += __shfl_down_sync(MB_FULL, dsum, o);
            if (lane == 0) red[idx] = dsum;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (MODE == 0) {
        for (int i = threadIdx.x; i < DT_BINS; i += DT_THREADS)
            if (shist[i]) atomicAdd(a.hist + i, shist[i]);
        if (threadIdx.x == 0) {
            unsigned long long c = 0;
            for (int w = 0; w < 16; ++w) c += reinterpret_cast<unsigned long long*>(red)[w];
            atomicAdd(a.below, c);
        }
    } else if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 16; ++w) s += red[w];
        a.partials[blockIdx.x] = s;
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_S), "r"(256u));
    }
}

__device__ __forceinline__ uint32_t dt_mix(uint32_t x) {              // murmur3 finaliser
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
// squared distances (exact fp32) of DT_SAMPLES hashed pairs
__global__ void __launch_bounds__(256) svgd_dist_sample_kernel(const float* __restrict__ X, int n, int d, float* out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= DT_SAMPLES) return;
    const uint32_t i = dt_mix(2u * p + 1u) % (uint32_t)n, j = dt_mix((2u * p + 2u) ^ 0x9e3779b9u) % (uint32_t)n;
    const float* xi = X + (int64_t)i * d;
    const float* xj = X + (int64_t)j * d;
    float s = 0.f;
    for (int k = 0; k < d; ++k) { const float t = xi[k] - xj[k]; s = fmaf(t, t, s); }
    out[p] = s;
}
// Bracket of the median from the sample distances by ONE block (replaces two radix selects, 16 launches): min / max,
// a DT_SBINS-bin histogram of the samples over [min, max] in shared memory, and the outer edges of the bins that hold
// the sample quantiles q_lo, q_hi (one bin of slack on either side: the bin of a sample is computed in fp32); the
// result goes to q[0] (lo) and q[3] (hi), the inputs of svgd_dist_bracket_kernel.
#define DT_SBINS 16384
__global__ void __launch_bounds__(1024) svgd_dist_sample_bracket_kernel(const float* __restrict__ smp, int ns, double q_lo,
                                                                        double q_hi, double* q /*[2][3]*/) {
    extern __shared__ uint32_t sh[];                                   // [DT_SBINS]
    __shared__ float wmn[32], wmx[32];
    __shared__ uint32_t wsum[32];
    __shared__ int bins[2];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    float mn = INFINITY, mx = -INFINITY;
#pragma unroll 16
    for (int i = tid; i < ns; i += 1024) { const float v = __ldcg(smp + i); mn = fminf(mn, v); mx = fmaxf(mx, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(MB_FULL, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(MB_FULL, mx, o)); }
    if (lane == 0) { wmn[w] = mn; wmx[w] = mx; }
    if (tid == 0) { bins[0] = 0; bins[1] = DT_SBINS - 1; }
    for (int b = tid; b < DT_SBINS; b += 1024) sh[b] = 0;
    __syncthreads();
    mn = wmn[0]; mx = wmx[0];
#pragma unroll
    for (int k = 1; k < 32; ++k) { mn = fminf(mn, wmn[k]); mx = fmaxf(mx, wmx[k]); }
    const float scale = (mx > mn) ? (float)DT_SBINS / (mx - mn) : 0.f;
#pragma unroll 16
    for (int i = tid; i < ns; i += 1024) {
        const int b = min(DT_SBINS - 1, (int)((__ldcg(smp + i) - mn) * scale));
        atomicAdd(&sh[b], 1u);
    }
    __syncthreads();
    // exclusive scan: 16 consecutive bins per thread
    uint32_t c[16], loc = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { c[k] = sh[tid * 16 + k]; loc += c[k]; }
    uint32_t inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(MB_FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    uint32_t cum = inc - loc;
    for (int k = 0; k < w; ++k) cum += wsum[k];
    const uint32_t ranks[2] = {(uint32_t)floor(q_lo * (double)(ns - 1)), (uint32_t)ceil(q_hi * (double)(ns - 1))};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        for (int r = 0; r < 2; ++r)
            if (ranks[r] >= cum && ranks[r] < cum + c[k]) bins[r] = tid * 16 + k;
        cum += c[k];
    }
    __syncthreads();
    if (tid == 0) {
        const double bw = (mx > mn) ? ((double)mx - (double)mn) / (double)DT_SBINS : 0.0;
        q[0] = fmax((double)mn + (double)(bins[0] - 1) * bw, 0.0);
        q[3] = (double)mn + (double)(bins[1] + 2) * bw;
    }
}
__global__ void svgd_dist_bracket_kernel(const double* q /*[2][3]*/, DtBracket* br, uint32_t* hist, unsigned long long* below) {
    const int i = threadIdx.x;
    for (int b = i; b < DT_BINS; b += blockDim.x) hist[b] = 0;
    if (i == 0) {
        *below = 0ull;
        const double lo = q[0], hi = q[3];
        const double lo_w = lo * (1.0 - 2e-3), hi_w = hi * (1.0 + 2e-3) + 1e-30;
        // centred bins of width bw: bin b = round(D^2/bw - k0) covers [(k0 + b - 1/2) bw, (k0 + b + 1/2) bw); k0 integer so
        // that the kernel's offset -k0 + 1.5*2^23 is exact in fp32 (bw is widened if the bracket is too narrow for that)
        double bw = (hi_w - lo_w) / (DT_BINS - 2);
        if (lo_w / bw > 2097152.0) bw = lo_w / 2097152.0;
        const double k0 = floor(lo_w / bw);
        br->lo_w = (k0 - 0.5) * bw;
        br->bw_d2 = bw;
        br->fallback_d2 = 0.5 * (lo + hi);
        br->inv_bw = (float)(2.0 / bw);                                // S = -D^2/2  ->  D^2/bw = -S * (2/bw)
        br->bin_off = (float)(12582912.0 - k0);
    }
}
// 256 threads x 8 bins: block-wide exclusive scan of the histogram, then the two middle ranks are located in parallel
__global__ void __launch_bounds__(256) svgd_dist_median_finish(const DtBracket* br, const uint32_t* hist,
                                                               const unsigned long long* below, int n, float* h,
                                                               unsigned long long* fallback_counter) {
    __shared__ long long wsum[8];
    __shared__ double dist[2];
    __shared__ int found[2];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint32_t c[8];
    long long loc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k] = hist[tid * 8 + k]; loc += c[k]; }
    long long inc = loc;
    for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(MB_FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[w] = inc;
    if (tid < 2) found[tid] = 0;
    __syncthreads();
    long long base = inc - loc;
    for (int k = 0; k < w; ++k) base += wsum[k];
    const long long N2 = (long long)n * n, c_lo = (long long)*below;
    const long long ranks[2] = {(N2 - 1) / 2 - c_lo, N2 / 2 - c_lo};    // np.median: mean of the two middle entries
    long long cum = base;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        for (int r = 0; r < 2; ++r)
            if (ranks[r] >= cum && ranks[r] < cum + (long long)c[k]) {
                const double frac = ((double)(ranks[r] - cum) + 0.5) / (double)c[k];
                dist[r] = sqrt(fmax(br->lo_w + ((double)(tid * 8 + k) + frac) * br->bw_d2, 0.0));
                found[r] = 1;
            }
        cum += c[k];
    }
    __syncthreads();
    if (tid == 0) {
        double med;
        if (found[0] && found[1]) med = 0.5 * (dist[0] + dist[1]);
        else { med = sqrt(fmax(br->fallback_d2, 0.0)); atomicAdd(fallback_counter, 1ull); }
        h[0] = (float)(med / sqrt(2.0 * log((double)n)));
    }
}
__global__ void __launch_bounds__(256) svgd_dist_partial_sum(const double* partials, int nblocks, double* out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) s += partials[b];
    sm[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int k = 0; k < 256; ++k) t += sm[k]; out[0] = t; }
}
__global__ void __launch_bounds__(256) svgd_dist_mean_finish(const double* partials, int nblocks, int n, float* h) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) s += partials[b];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) h[0] = (float)(sm[0] / ((double)n * (double)n) / sqrt(2.0 * log((double)n)));
}

int mb_quantile_impl(mb_ctx* ctx, const float* v, int64_t n, const double* q_dev, double q_host, double* out3, cudaStream_t st);

static size_t dt_scratch_need(int n, int d) {
    const int n_pad = ((n + TC_BM - 1) / TC_BM) * TC_BM, tiles = n_pad / TC_BM;
    return (4u << 20) + 1024 + (size_t)tiles * TC_TILE_X_BYTES + (size_t)tiles * TC_K * 256 + (size_t)n_pad * d * 4 +
           (size_t)n_pad * 8 + (size_t)DT_SAMPLES * 4 + 1024 + (size_t)(tiles * TC_SPLIT) * 8 + 8192;
}

// One rank's share of the statistics: the 128-row tiles itile_first, itile_first + itile_step, ... (interleaved, because
// the symmetric evaluation gives tile row i only T_all - i tiles of work).  acc (caller-owned device memory,
// MB_PAIRDIST_ACC_BYTES): [0] bracket (identical on every rank: same hashed sample pairs), [64] entries below the
// bracket (u64), [72] sum of distances (fp64, mean mode), [1024] 2048 histogram counters (u32).  The ranks' `below`,
// `sum` and histogram add up (any all-reduce); mb_pairdist_finish turns the totals into the bandwidth.
int mb_pairdist_partial_tc(mb_ctx* ctx, const float* X, int n, int d, int mode, int itile_first, int itile_step,
                           void* acc, cudaStream_t st) {
    MB_REQUIRE(d >= 1 && d + 4 <= TC_K, "pairwise-distance tcgen05 variant needs d <= 60");
    MB_REQUIRE((int64_t)n * n >= 4 * (int64_t)DT_SAMPLES, "pairwise-distance tcgen05 variant needs n >= 1024 (use variant 0)");
    const int NV = TC_K;
    const int n_pad = ((n + TC_BM - 1) / TC_BM) * TC_BM;
    const int tiles = n_pad / TC_BM;
    MB_REQUIRE(itile_step >= 1 && itile_first >= 0 && itile_first < itile_step, "mb_pairdist_partial: bad tile assignment");
    const int my_tiles = itile_first < tiles ? (tiles - itile_first + itile_step - 1) / itile_step : 0;
    const int grid = my_tiles * TC_SPLIT;
    const size_t bx = (size_t)tiles * TC_TILE_X_BYTES, bw = (size_t)tiles * NV * 256;
    const size_t bs = (size_t)n_pad * d * 4, ba = (size_t)n_pad * 2 * 4, bsmp = (size_t)DT_SAMPLES * 4;
    const size_t bmisc = 1024 + (size_t)(tiles * TC_SPLIT) * 8;
    const size_t need = 1024 + bx + bw + bs + ba + bsmp + bmisc + 8192;
    (void)need;
    if (mb_ensure_scratch(ctx, dt_scratch_need(n, d)) != MB_OK) return MB_ERR_CUDA;
    uint8_t* base = (uint8_t*)ctx->scratch + (4u << 20);
    auto align = [](uint8_t* p) { return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023); };
    float* mean = reinterpret_cast<float*>(align(base));
    uint8_t* XA = align(reinterpret_cast<uint8_t*>(mean) + 256);
    uint8_t* WT = align(XA + bx);
    float* xs = reinterpret_cast<float*>(align(WT + bw));
    float* saux = reinterpret_cast<float*>(align(reinterpret_cast<uint8_t*>(xs) + bs));
    float* smp = reinterpret_cast<float*>(align(reinterpret_cast<uint8_t*>(saux) + ba));
    uint8_t* misc = align(reinterpret_cast<uint8_t*>(smp) + bsmp);
    double* qout = reinterpret_cast<double*>(misc);                    // [2][3]
    double* partials = reinterpret_cast<double*>(misc + 1024);
    DtBracket* br = reinterpret_cast<DtBracket*>(acc);
    unsigned long long* below = reinterpret_cast<unsigned long long*>((char*)acc + 64);
    double* dsum = reinterpret_cast<double*>((char*)acc + 72);
    uint32_t* hist = reinterpret_cast<uint32_t*>((char*)acc + 1024);
    double* cpart = reinterpret_cast<double*>((char*)ctx->scratch + (3u << 20) + (64u << 10));
    svgd_tc_colsum_kernel<<<ctx->sms * 2, 256, 0, st>>>(X, n, d, cpart);
    svgd_tc_colmean_kernel<<<1, 1024, 0, st>>>(cpart, ctx->sms * 2, n, d, mean);
    TcPrepArgs p{X, nullptr, mean, nullptr, n, d, n_pad, NV, XA, WT, xs, saux};
    svgd_tc_rows_kernel<<<(unsigned)(((int64_t)n_pad * 8 + 255) / 256), 256, 0, st>>>(p);
    svgd_tc_wt_kernel<<<(unsigned)(((int64_t)tiles * NV * 16 + 255) / 256), 256, 0, st>>>(p);
    MB_CHECK_LAUNCH();
    DtArgs a{XA, WT, n, n_pad, br, hist, below, partials, itile_first, itile_step};
    const size_t smem = 1024 + TC_TILE_X_BYTES + (size_t)DT_STAGES * NV * 256 + DT_BINS * 4 + 512;
    if (mode == 0) {
        svgd_dist_sample_kernel<<<DT_SAMPLES / 256, 256, 0, st>>>(X, n, d, smp);
        const double delta = 3.0 / sqrt((double)DT_SAMPLES);           // 6 sigma of the sample median's quantile level
        MB_CUDA(cudaFuncSetAttribute(svgd_dist_sample_bracket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SBINS * 4));
        svgd_dist_sample_bracket_kernel<<<1, 1024, DT_SBINS * 4, st>>>(smp, DT_SAMPLES, 0.5 - delta, 0.5 + delta, qout);
        svgd_dist_bracket_kernel<<<1, 256, 0, st>>>(qout, br, hist, below);
        if (grid > 0) {
            MB_CUDA(cudaFuncSetAttribute(svgd_dist_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            svgd_dist_tc_kernel<0><<<grid, DT_THREADS, smem, st>>>(a);
        }
    } else {
        if (grid > 0) {
            MB_CUDA(cudaFuncSetAttribute(svgd_dist_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            svgd_dist_tc_kernel<1><<<grid, DT_THREADS, smem, st>>>(a);
        }
        svgd_dist_partial_sum<<<1, 256, 0, st>>>(partials, grid, dsum);
    }
    MB_CHECK_LAUNCH();
    return MB_OK;
}

int mb_pairdist_finish_tc(mb_ctx* ctx, int mode, int n, const void* acc, float* h, cudaStream_t st) {
    const DtBracket* br = reinterpret_cast<const DtBracket*>(acc);
    const unsigned long long* below = reinterpret_cast<const unsigned long long*>((const char*)acc + 64);
    const double* dsum = reinterpret_cast<const double*>((const char*)acc + 72);
    const uint32_t* hist = reinterpret_cast<const uint32_t*>((const char*)acc + 1024);
    if (mode == 0)
        svgd_dist_median_finish<<<1, 256, 0, st>>>(br, hist, below, n, h, (unsigned long long*)(ctx->counters + MB_CNT_BW_FALLBACK));
    else
        svgd_dist_mean_finish<<<1, 256, 0, st>>>(dsum, 1, n, h);
    MB_CHECK_LAUNCH();
    return MB_OK;
}

int mb_pairdist_bandwidth_tc(mb_ctx* ctx, const float* X, int n, int d, int mode, float* h, cudaStream_t st) {
    // single GPU: the whole matrix in one share; the accumulator lives in the context's scratch at [3.5 MiB, +10 KiB)
    // (sized first, so that the pointer survives the call below)
    if (mb_ensure_scratch(ctx, dt_scratch_need(n, d)) != MB_OK) return MB_ERR_CUDA;
    void* acc = (char*)ctx->scratch + (3u << 20) + (512u << 10);
    int rc = mb_pairdist_partial_tc(ctx, X, n, d, mode, 0, 1, acc, st);
    if (rc != MB_OK) return rc;
    return mb_pairdist_finish_tc(ctx, mode, n, acc, h, st);
}

// =================================================================================================
// Config C4's target on the tensor cores (variant 1 of mb_logistic_potential_grad; bf16 operands, fp32 accumulation):
// the same pipeline with MODE 1.  A tile rows = particles w_i (bf16), W^T tiles = features (rows c < d of the tile are
// a_j[c]); S = W A^T are the logits, P = sigmoid(S), O = P A; the potential comes from the softplus row sums and
// b = A^T t is exact fp32:   U_lik = sum_j softplus(s_ij) - w_i . b,   grad_lik = O_i - b.
__global__ void __launch_bounds__(256) logit_rows_kernel(const float* __restrict__ W, int n, int d, int n_pad, uint8_t* XA) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)n_pad * 8) return;
    const int row = (int)(tid >> 3), ch = (int)(tid & 7);
    __nv_bfloat16 va[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = ch * 8 + e;
        va[e] = __float2bfloat16_rn((row < n && k < d) ? W[(int64_t)row * d + k] : 0.f);
    }
    *reinterpret_cast<uint4*>(XA + (int64_t)(row / TC_BM) * TC_TILE_X_BYTES + sw128_off(row % TC_BM, ch * 8)) =
        *reinterpret_cast<uint4*>(va);
}
__global__ void __launch_bounds__(256) logit_wt_kernel(const float* __restrict__ A, int N, int d, int N_pad, uint8_t* WT) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)(N_pad / TC_BN) * TC_K * 16) return;
    const int q = (int)(tid & 15), c = (int)((tid >> 4) % TC_K), tile = (int)((tid >> 4) / TC_K);
    __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int j = tile * TC_BN + q * 8 + e;
        v[e] = __float2bfloat16_rn((c < d && j < N) ? A[(int64_t)j * d + c] : 0.f);
    }
    const uint32_t off = (uint32_t)(q >> 3) * (uint32_t)(TC_K * 128) + sw128_off(c, (q & 7) * 8);
    *reinterpret_cast<uint4*>(WT + (int64_t)tile * (TC_K * 256) + off) = *reinterpret_cast<uint4*>(v);
}
// b = A^T t (exact fp32, constant of the scenario): 16 row groups x 64 columns, fixed-order merge
__global__ void __launch_bounds__(1024) logit_b_kernel(const float* __restrict__ A, const float* __restrict__ t, int N, int d, float* b) {
    __shared__ float sm[16][64];
    const int k = threadIdx.x & 63, g = threadIdx.x >> 6;
    float s = 0.f;
    if (k < d)
        for (int j = g; j < N; j += 16) s = fmaf(t[j], A[(int64_t)j * d + k], s);
    sm[g][k] = s;
    __syncthreads();
    if (g == 0 && k < d) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) acc += sm[q][k];
        b[k] = acc;
    }
}
struct LogitFinishArgs {
    const float* W; const float* b; const float* opart; const float* upart; int n, d, n_pad, N, N_pad;
    float prior_mean, prior_pscale, beta; float* U; float* G;
};
// one thread per (particle, coordinate): 4 particles x 64 coordinates per block, every access coalesced; the potential's
// sums over the coordinates are reduced by shuffles (fixed order)
__global__ void __launch_bounds__(256) logit_finish_kernel(LogitFinishArgs a) {
    __shared__ float red[4][2][2];
    const int k = threadIdx.x & 63, r = threadIdx.x >> 6, half = (threadIdx.x >> 5) & 1;
    const int i = blockIdx.x * 4 + r;
    float up = 0.f, wb = 0.f;
    if (i < a.n && k < a.d) {
        const float w = a.W[(int64_t)i * a.d + k], bk = a.b[k];
        float o = 0.f;
#pragma unroll
        for (int p = 0; p < TC_SPLIT; ++p) o += a.opart[((int64_t)p * a.n_pad + i) * TC_K + k];
        const float rr = (w - a.prior_mean) * a.prior_pscale;
        up = 0.5f * rr * rr;
        wb = w * bk;
        a.G[(int64_t)i * a.d + k] = fmaf(a.beta, o - bk, rr * a.prior_pscale);
    }
    if (!a.U) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { up += __shfl_xor_sync(MB_FULL, up, o); wb += __shfl_xor_sync(MB_FULL, wb, o); }
    if ((threadIdx.x & 31) == 0) { red[r][half][0] = up; red[r][half][1] = wb; }
    __syncthreads();
    if (k == 0 && i < a.n) {
        up = red[r][0][0] + red[r][1][0];
        wb = red[r][0][1] + red[r][1][1];
        float sp = 0.f;
        for (int p = 0; p < TC_SPLIT; ++p)
            for (int q = 0; q < 4; ++q) sp += a.upart[(((int64_t)p * a.n_pad + i) << 2) + q];
        sp -= (float)(a.N_pad - a.N) * 0.6931471805599453f;           // padded data columns: softplus(0) = ln 2 each
        a.U[i] = fmaf(a.beta, sp - wb, up);
    }
}

int mb_logistic_potential_grad_tc(mb_ctx* ctx, const float* features, const float* labels, int N, int d, float prior_mean,
                                  float prior_pscale, float beta, const float* W, int n, float* U, float* G, cudaStream_t st) {
    MB_REQUIRE(d >= 1 && d <= TC_K, "logistic tcgen05 variant needs d <= 64");
    const int NV = TC_K;
    const int n_pad = ((n + TC_BM - 1) / TC_BM) * TC_BM, N_pad = ((N + TC_BN - 1) / TC_BN) * TC_BN;
    const int tiles = n_pad / TC_BM, jtiles = N_pad / TC_BN;
    const size_t bx = (size_t)tiles * TC_TILE_X_BYTES, bw = (size_t)jtiles * NV * 256;
    const size_t bo = (size_t)TC_SPLIT * n_pad * NV * 4, bu = (size_t)TC_SPLIT * n_pad * 4 * 4;
    const size_t need = 1024 + bx + bw + bo + bu + 8192;
    if (mb_ensure_scratch(ctx, (4u << 20) + need) != MB_OK) return MB_ERR_CUDA;
    uint8_t* base = (uint8_t*)ctx->scratch + (4u << 20);
    auto align = [](uint8_t* p) { return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023); };
    float* b = reinterpret_cast<float*>(align(base));
    uint8_t* XA = align(reinterpret_cast<uint8_t*>(b) + 256);
    uint8_t* WT = align(XA + bx);
    float* opart = reinterpret_cast<float*>(align(WT + bw));
    float* upart = reinterpret_cast<float*>(align(reinterpret_cast<uint8_t*>(opart) + bo));
    logit_b_kernel<<<1, 1024, 0, st>>>(features, labels, N, d, b);
    logit_rows_kernel<<<(unsigned)(((int64_t)n_pad * 8 + 255) / 256), 256, 0, st>>>(W, n, d, n_pad, XA);
    logit_wt_kernel<<<(unsigned)(((int64_t)jtiles * NV * 16 + 255) / 256), 256, 0, st>>>(features, N, d, N_pad, WT);
    MB_CHECK_LAUNCH();
    const size_t smem = 1024 + TC_TILE_X_BYTES + (size_t)TC_STAGES * NV * 256 + 256;
    MB_CUDA(cudaFuncSetAttribute(svgd_phi_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TcArgs a{XA, WT, nullptr, nullptr, nullptr, n, d, n_pad, NV, opart, upart, N_pad};
    svgd_phi_tc_kernel<1><<<tiles * TC_SPLIT, TC_THREADS, smem, st>>>(a);
    MB_CHECK_LAUNCH();
    LogitFinishArgs f{W, b, opart, upart, n, d, n_pad, N, N_pad, prior_mean, prior_pscale, beta, U, G};
    logit_finish_kernel<<<(n + 3) / 4, 256, 0, st>>>(f);
    MB_CHECK_LAUNCH();
    return MB_OK;
}
