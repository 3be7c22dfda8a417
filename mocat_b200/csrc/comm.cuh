// comm.cuh -- device-side exchange of small fp64 records between the GPUs of one NVSwitch domain.
//
// One process per GPU; every rank owns a mailbox in its own HBM that all peers have mapped through CUDA
// IPC.  An "allgather" of a few doubles is: store my record (+ a sequence number, release.sys) into my slot
// of EVERY peer's mailbox, then spin (acquire.sys) on my own mailbox until all `world` slots carry the
// expected sequence number.  ~1 NVLink store latency instead of a NCCL launch; usable from inside a running
// kernel (the tempering search exchanges its (max, sum, sumsq) triple once per regula-falsi evaluation).
// Two parities per slot: a slot is rewritten only two exchanges later, by which time every reader has
// provably consumed it (a rank cannot be two exchanges ahead of a peer it has to hear from each time).
#pragma once
#include "common.cuh"

#define MB_MAIL_DOUBLES 6

struct MbMail {                       // 64 B
    double v[MB_MAIL_DOUBLES];
    unsigned long long seq;
    unsigned long long pad;
};

struct MbCommDev {                    // passed by value to kernels
    int rank, world;
    MbMail* box[MB_MAX_WORLD];        // box[r] = rank r's mailbox [2][MB_MAX_WORLD] (peer mapped; box[rank] local)
    unsigned long long* seq;          // local exchange counter (identical on all ranks by construction)
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by ONE thread per rank.  in[nd] -> out[world][nd] (rank order).  nd <= MB_MAIL_DOUBLES.
__device__ __forceinline__ void comm_allgather(const MbCommDev& c, const double* in, int nd, double* out) {
    const unsigned long long s = *c.seq + 1ull;
    const int par = (int)(s & 1ull);
    for (int r = 0; r < c.world; ++r) {
        MbMail* m = c.box[r] + par * MB_MAX_WORLD + c.rank;
        for (int k = 0; k < nd; ++k) m->v[k] = in[k];
    }
    __threadfence_system();
    for (int r = 0; r < c.world; ++r) st_release_sys_u64(&(c.box[r] + par * MB_MAX_WORLD + c.rank)->seq, s);
    for (int r = 0; r < c.world; ++r) {
        const MbMail* m = c.box[c.rank] + par * MB_MAX_WORLD + r;
        while (ld_acquire_sys_u64(&m->seq) != s) { }
        for (int k = 0; k < nd; ++k) out[r * nd + k] = m->v[k];
    }
    *c.seq = s;
}
#endif
