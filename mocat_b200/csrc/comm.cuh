// comm.cuh -- device-side exchange of small fp64 records between the GPUs of one NVSwitch domain.
//
// One process per GPU; every rank owns a mailbox in its own HBM that all peers have mapped through CUDA
// IPC.  An "allgather" of a few doubles uses a low-latency flag-in-data protocol (the shape of NCCL's LL):
// every double travels as two 8-byte words {32 data bits, 32-bit sequence flag}; an 8-byte store is atomic, so
// no fence and no separate ready flag is needed -- the sender's warp stores straight into its slot of EVERY
// peer's mailbox, a receiving warp spins (volatile loads of its own HBM) until every word of every slot
// carries the expected flag.  ~1 NVLink store latency instead of a NCCL launch; usable from inside a running
// kernel (the tempering search exchanges its (max, sum, sumsq) triple once per regula-falsi evaluation, and
// every block of the grid receives for itself, so no second grid-wide barrier is needed).
// Two parities per slot: a slot is rewritten only two exchanges later, by which time every reader has
// provably consumed it (a rank cannot be two exchanges ahead of a peer it has to hear from each time).
#pragma once
#include "common.cuh"

#define MB_MAIL_DOUBLES 6
#define MB_MAIL_WORDS (2 * MB_MAIL_DOUBLES)

struct MbMail {                       // 128 B
    unsigned long long w[16];
};

struct MbCommDev {                    // passed by value to kernels
    int rank, world;
    MbMail* box[MB_MAX_WORLD];        // box[r] = rank r's mailbox [2][MB_MAX_WORLD] (peer mapped; box[rank] local)
    unsigned long long* seq;          // local exchange counter (identical on all ranks by construction)
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// exchange number s (>= 1).  Called by all 32 lanes of ONE warp per participating block; `in`/`out` are visible to
// the whole warp (shared or global memory).  send: this warp also publishes the rank's record (exactly one warp
// per rank and exchange must).  in[nd] -> out[world][nd] (rank order), nd <= MB_MAIL_DOUBLES.
__device__ __forceinline__ void comm_exchange_warp(const MbCommDev& c, unsigned long long s, bool send, const double* in,
                                                   int nd, double* out) {
    const int lane = threadIdx.x & 31;
    const int par = (int)(s & 1ull);
    const unsigned long long flag = (s & 0xffffffffull) << 32;
    const int nw = 2 * nd;
    if (send) {
        for (int idx = lane; idx < c.world * nw; idx += 32) {
            const int r = idx / nw, k = idx - r * nw;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(in[k >> 1]);
            const unsigned long long half = (k & 1) ? (bits >> 32) : (bits & 0xffffffffull);
            st_sys_u64(&(c.box[r] + par * MB_MAX_WORLD + c.rank)->w[k], flag | half);
        }
    }
    for (int idx = lane; idx < c.world * nw; idx += 32) {
        const int r = idx / nw, k = idx - r * nw;
        const unsigned long long* p = &(c.box[c.rank] + par * MB_MAX_WORLD + r)->w[k];
        unsigned long long v;
        do { v = ld_sys_u64(p); } while ((v & 0xffffffff00000000ull) != flag);
        reinterpret_cast<unsigned int*>(out)[r * nw + k] = (unsigned int)(v & 0xffffffffull);   // little endian halves
    }
    __syncwarp();
}

// one-warp convenience: next exchange of the communicator, sequence counter advanced
__device__ __forceinline__ void comm_allgather_warp(const MbCommDev& c, const double* in, int nd, double* out) {
    const unsigned long long s = *c.seq + 1ull;
    __syncwarp();
    comm_exchange_warp(c, s, true, in, nd, out);
    if ((threadIdx.x & 31) == 0) *c.seq = s;
    __syncwarp();
}
#endif
