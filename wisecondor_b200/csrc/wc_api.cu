// wisecondor_b200 - context, error text and workspace management behind the C ABI.
#include "wc_common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void wc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* wc_last_error(void) { return g_err; }
extern "C" const char* wc_version(void) { return "wisecondor_b200 0.2 (sm_100a)"; }
extern "C" int wc_abi_version(void) { return WC_ABI_VERSION; }

extern "C" wc_ctx* wc_create(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        wc_set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        wc_set_error("device %d out of range (have %d)", device, ndev);
        return nullptr;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        wc_set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        return nullptr;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        wc_set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return nullptr;
    }
    if (prop.major != 10) {
        wc_set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
        return nullptr;
    }
    wc_ctx* ctx = new wc_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < 2 * WC_NPHASE; ++i) {
        if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) {
            wc_set_error("cudaEventCreate failed");
            delete ctx;
            return nullptr;
        }
    }
    for (int i = 0; i < WC_NPHASE; ++i) ctx->phase_ms[i] = 0.0;
    for (int i = 0; i < WC_NCOUNTER; ++i) ctx->counter[i] = 0;
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        ctx->encode_tiled = fn;
    return ctx;
}

extern "C" void wc_destroy(wc_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < WC_NBUF; ++i)
        if (ctx->buf[i].p) cudaFree(ctx->buf[i].p);
    for (int i = 0; i < 2 * WC_NPHASE; ++i) cudaEventDestroy(ctx->ev[i]);
    if (ctx->search_plan && ctx->search_plan_free) ctx->search_plan_free(ctx->search_plan);
    if (ctx->d2h_stream) { cudaStreamDestroy(ctx->d2h_stream); cudaEventDestroy(ctx->d2h_ev); }
    delete ctx;
}

extern "C" int wc_sm_count(const wc_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" double wc_last_phase_ms(wc_ctx* ctx, int which) {
    if (!ctx || which < 0 || which >= WC_NPHASE) return -1.0;
    if (ctx->timed_mask & (1u << which)) {      // recorded by an asynchronous call: wait for the end event now
        float ms = 0.f;
        cudaSetDevice(ctx->device);
        if (cudaEventSynchronize(ctx->ev[2 * which + 1]) == cudaSuccess &&
            cudaEventElapsedTime(&ms, ctx->ev[2 * which], ctx->ev[2 * which + 1]) == cudaSuccess)
            ctx->phase_ms[which] = ms;
        ctx->timed_mask &= ~(1u << which);
    }
    return ctx->phase_ms[which];
}

extern "C" long long wc_last_counter(const wc_ctx* ctx, int which) {
    if (!ctx || which < 0) return -1;
    if (which >= 16 && which < 16 + 64) {
        // (bin, sample) pairs the z-score pass `which - 16` of the last wc_zscore_batch computed (waits for the stream)
        const int pass = which - 16;
        if (pass >= ctx->zs_repeats) return -1;
        if (pass == 0 || ctx->zs_npairs_d == nullptr) return ctx->zs_all_pairs;
        int n = 0;
        cudaSetDevice(ctx->device);
        if (cudaMemcpy(&n, ctx->zs_npairs_d + pass, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return n > ctx->zs_pair_limit ? ctx->zs_all_pairs : n;      // long lists run as a full pass
    }
    if (which >= WC_NCOUNTER) return -1;
    return ctx->counter[which];
}

int wc_reserve(wc_ctx* ctx, int slot, size_t bytes, void** out) {
    if (slot < 0 || slot >= WC_NBUF) {
        wc_set_error("workspace slot %d out of range", slot);
        return WC_ERR_INTERNAL;
    }
    wc_buf& b = ctx->buf[slot];
    if (bytes == 0) bytes = 16;
    if (b.bytes < bytes) {
        if (b.p) {
            cudaDeviceSynchronize();
            cudaFree(b.p);
            b.p = nullptr;
            b.bytes = 0;
        }
        size_t want = bytes + bytes / 8;   // a little slack so slightly larger follow-up calls do not reallocate
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&b.p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            wc_set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            b.p = nullptr;
            return WC_ERR_NOMEM;
        }
        b.bytes = want;
    }
    *out = b.p;
    return WC_OK;
}

// ---- raw device buffers for hosts that do not bring an allocator of their own (the command line without PyTorch) ----
extern "C" int wc_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

extern "C" void* wc_dev_alloc(wc_ctx* ctx, size_t bytes) {
    if (!ctx) { wc_set_error("wc_dev_alloc: no context"); return nullptr; }
    void* p = nullptr;
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        wc_set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}

extern "C" int wc_dev_free(wc_ctx* ctx, void* p) {
    WC_CHECK_ARG(ctx != nullptr);
    if (!p) return WC_OK;
    WC_CUDA(cudaSetDevice(ctx->device));
    WC_CUDA(cudaFree(p));
    return WC_OK;
}

extern "C" int wc_copy_h2d(wc_ctx* ctx, void* dst_d, const void* src_h, size_t bytes) {
    WC_CHECK_ARG(ctx != nullptr && (bytes == 0 || (dst_d != nullptr && src_h != nullptr)));
    WC_CUDA(cudaSetDevice(ctx->device));
    if (bytes) WC_CUDA(cudaMemcpy(dst_d, src_h, bytes, cudaMemcpyHostToDevice));
    return WC_OK;
}

extern "C" int wc_copy_d2h(wc_ctx* ctx, void* dst_h, const void* src_d, size_t bytes) {
    WC_CHECK_ARG(ctx != nullptr && (bytes == 0 || (dst_h != nullptr && src_d != nullptr)));
    WC_CUDA(cudaSetDevice(ctx->device));
    if (bytes) WC_CUDA(cudaMemcpy(dst_h, src_d, bytes, cudaMemcpyDeviceToHost));     // waits for prior work on the null stream
    return WC_OK;
}

extern "C" int wc_dev_sync(wc_ctx* ctx) {
    WC_CHECK_ARG(ctx != nullptr);
    WC_CUDA(cudaSetDevice(ctx->device));
    WC_CUDA(cudaDeviceSynchronize());
    return WC_OK;
}
