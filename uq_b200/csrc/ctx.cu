// Context, memory, arrays and the per-kernel timing facility of libuqb200.so.
#include "common.cuh"
#include <stdarg.h>
#include <new>

int uqb_fail(uqb_ctx* ctx, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    if (ctx) vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
    return 1;
}

extern "C" int uqb_version(void) { return UQB_VERSION; }

extern "C" int uqb_ctx_create(int device, void* stream, uqb_ctx** out) {
    if (!out) return 1;
    *out = nullptr;
    uqb_ctx* ctx = new (std::nothrow) uqb_ctx();
    if (!ctx) return 1;
    *out = ctx;       // returned even on failure so that the caller can read the error text
    ctx->device = device;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return uqb_fail(ctx, "no CUDA device available (%s); libuqb200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= count) return uqb_fail(ctx, "device %d out of range (%d devices)", device, count);
    UQB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    UQB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return uqb_fail(ctx, "device %d is sm_%d%d; libuqb200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
        ctx->own_stream = false;
    } else {
        UQB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    // keep freed blocks in the stream-ordered pool instead of returning them to the OS
    cudaMemPool_t pool;
    UQB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thresh = UINT64_MAX;
    UQB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    return 0;
}

extern "C" void uqb_ctx_destroy(uqb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& r : ctx->pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ctx->free_events) cudaEventDestroy(e);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->span_a) { cudaEventDestroy(ctx->span_a); cudaEventDestroy(ctx->span_b); }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* uqb_last_error(const uqb_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int uqb_ctx_sync(uqb_ctx* ctx) {
    UQB_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" uint64_t uqb_ctx_launch_count(const uqb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- timing ------------------------------------------------------------------------------------
static cudaEvent_t get_event(uqb_ctx* ctx) {
    if (!ctx->free_events.empty()) {
        cudaEvent_t e = ctx->free_events.back();
        ctx->free_events.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void uqb_timer_begin(uqb_ctx* ctx, const char* name, uint64_t bytes) {
    if (!ctx->timing) return;
    uqb_timer_rec r;
    r.name = name;
    r.bytes = bytes;
    r.a = get_event(ctx);
    r.b = get_event(ctx);
    cudaEventRecord(r.a, ctx->stream);
    ctx->pending.push_back(r);
}

void uqb_timer_end(uqb_ctx* ctx) {
    if (!ctx->timing) return;
    cudaEventRecord(ctx->pending.back().b, ctx->stream);
}

void uqb_timer_add_bytes(uqb_ctx* ctx, uint64_t bytes) {
    if (!ctx->timing || ctx->pending.empty()) return;
    ctx->pending.back().bytes += bytes;
}

static void drain_timers(uqb_ctx* ctx) {
    if (ctx->pending.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto& r : ctx->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto& t = ctx->totals[r.name];
            t.launches += 1;
            t.ms += ms;
            t.bytes += r.bytes;
        }
        ctx->free_events.push_back(r.a);
        ctx->free_events.push_back(r.b);
    }
    ctx->pending.clear();
}

extern "C" int uqb_ctx_timing(uqb_ctx* ctx, int enable) {
    drain_timers(ctx);
    ctx->timing = enable != 0;
    return 0;
}

extern "C" int uqb_ctx_timing_reset(uqb_ctx* ctx) {
    drain_timers(ctx);
    ctx->totals.clear();
    return 0;
}

extern "C" int uqb_ctx_timing_report(uqb_ctx* ctx, char* names, uint64_t* launches, double* ms, uint64_t* bytes, int cap, int* n) {
    drain_timers(ctx);
    int i = 0;
    for (auto& kv : ctx->totals) {
        if (i >= cap) break;
        snprintf(names + (size_t)i * UQB_TIMER_NAME, UQB_TIMER_NAME, "%s", kv.first.c_str());
        launches[i] = kv.second.launches;
        ms[i] = kv.second.ms;
        bytes[i] = kv.second.bytes;
        i++;
    }
    *n = i;
    return 0;
}

// device time between two points of the context's stream (includes host gaps in between)
extern "C" int uqb_ctx_span_begin(uqb_ctx* ctx) {
    if (!ctx->span_a) { UQB_CUDA(cudaEventCreate(&ctx->span_a)); UQB_CUDA(cudaEventCreate(&ctx->span_b)); }
    UQB_CUDA(cudaEventRecord(ctx->span_a, ctx->stream));
    return 0;
}
extern "C" int uqb_ctx_span_end(uqb_ctx* ctx, double* ms) {
    if (!ctx->span_a) return uqb_fail(ctx, "span_end without span_begin");
    UQB_CUDA(cudaEventRecord(ctx->span_b, ctx->stream));
    UQB_CUDA(cudaEventSynchronize(ctx->span_b));
    float f = 0.f;
    UQB_CUDA(cudaEventElapsedTime(&f, ctx->span_a, ctx->span_b));
    *ms = f;
    return 0;
}

// ---- memory ------------------------------------------------------------------------------------
int uqb_dalloc(uqb_ctx* ctx, void** p, size_t nbytes) {
    *p = nullptr;
    size_t want = nbytes ? nbytes : 16;
    want = (want + 255) & ~(size_t)255;
    cudaError_t e = cudaMallocAsync(p, want, ctx->stream);
    if (e != cudaSuccess) {
        *p = nullptr;
        return uqb_fail(ctx, "device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    ctx->bytes_in_use += want;
    return 0;
}

int uqb_dfree(uqb_ctx* ctx, void* p, size_t nbytes) {
    if (!p) return 0;
    size_t want = nbytes ? nbytes : 16;
    want = (want + 255) & ~(size_t)255;
    ctx->bytes_in_use -= want;
    UQB_CUDA(cudaFreeAsync(p, ctx->stream));
    return 0;
}

int uqb_new_array(uqb_ctx* ctx, uint64_t n, uint32_t width, uqb_array** out) {
    uqb_array* a = new (std::nothrow) uqb_array();
    if (!a) return uqb_fail(ctx, "out of host memory");
    a->n = n;
    a->width = width;
    a->owned = true;
    // +64 slack so that vectorised tails may over-read/over-write safely
    int r = uqb_dalloc(ctx, &a->d, a->nbytes() + 64);
    if (r) { delete a; return r; }
    *out = a;
    return 0;
}

int uqb_pinned(uqb_ctx* ctx, size_t nbytes, void** out) {
    if (ctx->pinned_bytes < nbytes) {
        if (ctx->pinned) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->pinned); ctx->pinned = nullptr; ctx->pinned_bytes = 0; }
        size_t want = nbytes < (1u << 20) ? (1u << 20) : nbytes;
        UQB_CUDA(cudaHostAlloc(&ctx->pinned, want, cudaHostAllocDefault));
        ctx->pinned_bytes = want;
    }
    *out = ctx->pinned;
    return 0;
}

int uqb_readback(uqb_ctx* ctx, void* host, const void* dev, size_t nbytes) {
    void* stage;
    UQB_TRY(uqb_pinned(ctx, nbytes, &stage));
    UQB_CUDA(cudaMemcpyAsync(stage, dev, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    UQB_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(host, stage, nbytes);
    return 0;
}

extern "C" int uqb_host_alloc(uqb_ctx* ctx, uint64_t nbytes, void** out) {
    UQB_CUDA(cudaHostAlloc(out, nbytes ? nbytes : 1, cudaHostAllocDefault));
    return 0;
}

extern "C" int uqb_host_free(uqb_ctx* ctx, void* p) {
    UQB_CUDA(cudaFreeHost(p));
    return 0;
}

extern "C" int uqb_mem_info(uqb_ctx* ctx, uint64_t* in_use, uint64_t* dev_free, uint64_t* dev_total) {
    size_t f = 0, t = 0;
    UQB_CUDA(cudaMemGetInfo(&f, &t));
    if (in_use) *in_use = ctx->bytes_in_use;
    if (dev_free) *dev_free = f;
    if (dev_total) *dev_total = t;
    return 0;
}

// ---- arrays ------------------------------------------------------------------------------------
extern "C" int uqb_array_info(const uqb_array* a, uint64_t* n, uint32_t* width) {
    if (!a) return 1;
    if (n) *n = a->n;
    if (width) *width = a->width;
    return 0;
}

extern "C" void* uqb_array_device_ptr(const uqb_array* a) { return a ? a->d : nullptr; }

extern "C" int uqb_array_upload(uqb_ctx* ctx, const void* host, uint64_t n, uint32_t width, uqb_array** out) {
    uqb_array* a;
    UQB_TRY(uqb_new_array(ctx, n, width, &a));
    if (a->nbytes()) {
        UQB_CUDA(cudaMemcpyAsync(a->d, host, a->nbytes(), cudaMemcpyHostToDevice, ctx->stream));
        UQB_CUDA(cudaStreamSynchronize(ctx->stream));   // the host buffer may be pageable
    }
    *out = a;
    return 0;
}

extern "C" int uqb_array_download(uqb_ctx* ctx, const uqb_array* a, void* host, uint64_t nbytes) {
    if (nbytes > a->nbytes()) return uqb_fail(ctx, "download of %llu bytes from an array of %llu", (unsigned long long)nbytes, (unsigned long long)a->nbytes());
    if (nbytes) UQB_CUDA(cudaMemcpyAsync(host, a->d, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    UQB_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int uqb_array_free(uqb_ctx* ctx, uqb_array* a) {
    if (!a) return 0;
    int r = 0;
    if (a->owned) r = uqb_dfree(ctx, a->d, a->nbytes() + 64);
    delete a;
    return r;
}
