// abi.cu -- context management and error reporting of the C-ABI (include/mocat_b200.h).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[1024] = "";

void mb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* mb_last_error(void) { return g_err; }
extern "C" int mb_abi_version(void) { return MB_ABI_VERSION; }

extern "C" mb_ctx* mb_create(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        mb_set_error("mb_create: no usable CUDA device %d (%s); this library has no CPU fallback", device,
                     e == cudaSuccess ? "index out of range" : cudaGetErrorString(e));
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { mb_set_error("mb_create: cudaSetDevice failed"); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { mb_set_error("mb_create: no device properties"); return nullptr; }
    if (prop.major != 10) {
        mb_set_error("mb_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                     prop.minor);
        return nullptr;
    }
    mb_ctx* ctx = new mb_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->sms = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    bool ok = true;
    ok = ok && cudaMalloc(&ctx->partials, sizeof(double) * (3 * MB_MAX_PARTIAL_BLOCKS * 2 + 64)) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->counters, sizeof(uint32_t) * MB_NUM_COUNTERS) == cudaSuccess;
    ok = ok && cudaMemset(ctx->counters, 0, sizeof(uint32_t) * MB_NUM_COUNTERS) == cudaSuccess;
    {
        const uint32_t one = 1;
        ok = ok && cudaMemcpy(ctx->counters + MB_CNT_SCAN_EPOCH, &one, sizeof(one), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    ok = ok && cudaStreamCreateWithFlags(&ctx->body_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->ll_slots, sizeof(unsigned long long) * (2 * MB_LL_BLOCKS * 8 + 8)) == cudaSuccess;
    ok = ok && cudaMemset(ctx->ll_slots, 0, sizeof(unsigned long long) * (2 * MB_LL_BLOCKS * 8 + 8)) == cudaSuccess;
    ctx->scratch_bytes = 8u << 20;
    ok = ok && cudaMalloc(&ctx->scratch, ctx->scratch_bytes) == cudaSuccess;
    if (!ok) {
        mb_set_error("mb_create: workspace allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        mb_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

extern "C" void mb_destroy(mb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->partials);
    cudaFree(ctx->counters);
    cudaFree(ctx->scan_flag);
    cudaFree(ctx->scan_agg);
    cudaFree(ctx->scan_incl);
    cudaFree(ctx->scratch);
    cudaFree(ctx->ll_slots);
    if (ctx->body_stream) cudaStreamDestroy(ctx->body_stream);
    delete ctx;
}

extern "C" int mb_sm_count(mb_ctx* ctx) { return ctx ? ctx->sms : 0; }

extern "C" unsigned long long mb_workspace_generation(mb_ctx* ctx) { return ctx ? ctx->ws_generation : 0ull; }

int mb_ensure_scratch(mb_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return MB_OK;
    MB_CUDA(cudaDeviceSynchronize());
    cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t nb = bytes + (bytes >> 1);
    MB_CUDA(cudaMalloc(&ctx->scratch, nb));
    ctx->scratch_bytes = nb;
    ctx->ws_generation++;
    return MB_OK;
}

int mb_ensure_scan(mb_ctx* ctx, int64_t tiles) {
    if (tiles <= ctx->scan_tiles_cap) return MB_OK;
    MB_CUDA(cudaDeviceSynchronize());
    cudaFree(ctx->scan_flag); cudaFree(ctx->scan_agg); cudaFree(ctx->scan_incl);
    ctx->scan_flag = nullptr; ctx->scan_agg = nullptr; ctx->scan_incl = nullptr; ctx->scan_tiles_cap = 0;
    int64_t cap = tiles + (tiles >> 1) + 64;
    MB_CUDA(cudaMalloc(&ctx->scan_flag, sizeof(int32_t) * cap));
    MB_CUDA(cudaMalloc(&ctx->scan_agg, sizeof(double) * cap));
    MB_CUDA(cudaMalloc(&ctx->scan_incl, sizeof(double) * cap));
    MB_CUDA(cudaMemset(ctx->scan_flag, 0, sizeof(int32_t) * cap));
    ctx->scan_tiles_cap = cap;
    ctx->ws_generation++;
    ctx->scan_epoch = 0;
    return MB_OK;
}

// ------------------------------------------------------------------------------------------------
// Conditional section of a captured population step.  All resampling kernels are predicated on the device-side
// flags of the control block, so launching them unconditionally is always correct; but an early-exit kernel still
// costs a graph node (~2-3 us each, 6-9 of them per step).  While `stream` is being captured, the launches issued
// between mb_cond_begin and mb_cond_end (on the stream returned in *body_stream) are recorded into the body of a
// CUDA-graph conditional IF node whose condition (ctl->resample && !ctl->done) is set on the device by a one-thread
// kernel.  Outside capture the pair is a no-op and *body_stream = stream.
__global__ void cond_set_kernel(cudaGraphConditionalHandle h, const mb_control* ctl) {
    cudaGraphSetConditional(h, (ctl->resample != 0 && ctl->done == 0) ? 1u : 0u);
}

extern "C" int mb_cond_begin(mb_ctx* ctx, const mb_control* ctl, mb_stream_t stream, mb_stream_t* body_stream) {
    MB_REQUIRE(ctx && ctl && body_stream, "mb_cond_begin: bad arguments");
    MB_REQUIRE(!ctx->cond_active, "mb_cond_begin: conditional sections do not nest");
    cudaStream_t st = mb_s(stream);
    *body_stream = stream;
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    cudaGraph_t graph = nullptr;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    MB_CUDA(cudaStreamGetCaptureInfo_v2(st, &status, &id, &graph, &deps, &ndeps));
    if (status != cudaStreamCaptureStatusActive) return MB_OK;
    cudaGraphConditionalHandle handle;
    MB_CUDA(cudaGraphConditionalHandleCreate(&handle, graph, 0, cudaGraphCondAssignDefault));
    cond_set_kernel<<<1, 1, 0, st>>>(handle, ctl);
    MB_CHECK_LAUNCH();
    MB_CUDA(cudaStreamGetCaptureInfo_v2(st, &status, &id, &graph, &deps, &ndeps));
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = handle;
    p.conditional.type = cudaGraphCondTypeIf;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    MB_CUDA(cudaGraphAddNode(&node, graph, deps, ndeps, &p));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    MB_CUDA(cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
    MB_CUDA(cudaStreamBeginCaptureToGraph(ctx->body_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    ctx->cond_active = 1;
    *body_stream = (mb_stream_t)ctx->body_stream;
    return MB_OK;
}

extern "C" int mb_cond_end(mb_ctx* ctx, mb_stream_t stream) {
    MB_REQUIRE(ctx, "mb_cond_end: bad arguments");
    (void)stream;
    if (!ctx->cond_active) return MB_OK;
    ctx->cond_active = 0;
    cudaGraph_t g = nullptr;
    MB_CUDA(cudaStreamEndCapture(ctx->body_stream, &g));
    return MB_OK;
}
