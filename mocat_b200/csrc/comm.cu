// comm.cu -- multi-GPU plumbing of the C-ABI: IPC-shareable device allocations, the peer-mapped mailbox
// communicator (comm.cuh) and a stand-alone small allgather kernel.  One process per GPU; handles are
// exchanged by the host (torch.distributed all_gather_object in mocat_b200/parallel.py).
#include <string.h>
#include "comm.cuh"

struct mb_comm {
    mb_ctx* ctx;
    int rank, world;
    MbMail* local;                      // [2][MB_MAX_WORLD]
    unsigned long long* seq;
    void* opened[MB_MAX_WORLD];         // IPC-opened peer mailboxes (NULL for self)
    MbCommDev dev;
};

extern "C" int mb_alloc(mb_ctx* ctx, size_t bytes, void** out) {
    MB_REQUIRE(ctx && out && bytes > 0, "mb_alloc: bad arguments");
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaMalloc(out, bytes));
    MB_CUDA(cudaMemset(*out, 0, bytes));
    return MB_OK;
}

extern "C" int mb_free(mb_ctx* ctx, void* p) {
    MB_REQUIRE(ctx, "mb_free: bad arguments");
    if (p) MB_CUDA(cudaFree(p));
    return MB_OK;
}

extern "C" int mb_ipc_get_handle(mb_ctx* ctx, void* dev_ptr, void* handle64_host) {
    MB_REQUIRE(ctx && dev_ptr && handle64_host, "mb_ipc_get_handle: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    MB_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle64_host, &h, 64);
    return MB_OK;
}

extern "C" int mb_ipc_open(mb_ctx* ctx, const void* handle64_host, void** out) {
    MB_REQUIRE(ctx && handle64_host && out, "mb_ipc_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, 64);
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return MB_OK;
}

extern "C" int mb_ipc_close(mb_ctx* ctx, void* p) {
    MB_REQUIRE(ctx, "mb_ipc_close: bad arguments");
    if (p) MB_CUDA(cudaIpcCloseMemHandle(p));
    return MB_OK;
}

extern "C" int mb_comm_create(mb_ctx* ctx, int rank, int world, mb_comm** out, void* handle64_host) {
    MB_REQUIRE(ctx && out && handle64_host && world >= 1 && world <= MB_MAX_WORLD && rank >= 0 && rank < world,
               "mb_comm_create: bad arguments");
    mb_comm* c = new mb_comm();
    memset(c, 0, sizeof(*c));
    c->ctx = ctx; c->rank = rank; c->world = world;
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaMalloc(&c->local, sizeof(MbMail) * 2 * MB_MAX_WORLD));
    MB_CUDA(cudaMemset(c->local, 0, sizeof(MbMail) * 2 * MB_MAX_WORLD));
    MB_CUDA(cudaMalloc(&c->seq, sizeof(unsigned long long)));
    MB_CUDA(cudaMemset(c->seq, 0, sizeof(unsigned long long)));
    int rc = mb_ipc_get_handle(ctx, c->local, handle64_host);
    if (rc != MB_OK) return rc;
    *out = c;
    return MB_OK;
}

// handles: world x 64 bytes in rank order (the entry of this rank is ignored)
extern "C" int mb_comm_connect(mb_comm* c, const void* handles_host) {
    MB_REQUIRE(c && handles_host, "mb_comm_connect: bad arguments");
    c->dev.rank = c->rank; c->dev.world = c->world; c->dev.seq = c->seq;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) { c->dev.box[r] = c->local; continue; }
        void* p = nullptr;
        int rc = mb_ipc_open(c->ctx, (const char*)handles_host + 64 * r, &p);
        if (rc != MB_OK) return rc;
        c->opened[r] = p;
        c->dev.box[r] = (MbMail*)p;
    }
    return MB_OK;
}

extern "C" void mb_comm_destroy(mb_comm* c) {
    if (!c) return;
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->opened[r]);
    cudaFree(c->local);
    cudaFree(c->seq);
    delete c;
}

const MbCommDev* mb_comm_dev(const mb_comm* c) { return &c->dev; }

__global__ void comm_allgather_kernel(MbCommDev c, const double* in, int nd, double* out, const mb_control* ctl) {
    if (ctl && (ctl->done || !ctl->resample)) return;               // identical decision on every rank
    comm_allgather_warp(c, in, nd, out);
}

// out[world][nd] <- in[nd] of every rank (device pointers), nd <= 6.  Every rank must call it the same
// number of times in the same order.  ctl != NULL: skipped unless the (replicated) control block asks for a
// resampling step.  Kernels of this rank that precede the call in stream order are complete on every rank's
// side of the exchange, so it doubles as the barrier before peer reads of their output.
extern "C" int mb_comm_allgather(mb_comm* c, const double* in, int nd, double* out, const mb_control* ctl,
                                 mb_stream_t stream) {
    MB_REQUIRE(c && in && out && nd >= 1 && nd <= MB_MAIL_DOUBLES, "mb_comm_allgather: bad arguments");
    comm_allgather_kernel<<<1, 32, 0, mb_s(stream)>>>(c->dev, in, nd, out, ctl);
    MB_CHECK_LAUNCH();
    return MB_OK;
}
