// Context management and error plumbing of the C ABI (include/marie_b200.h).
#include "common.cuh"
#include <stdarg.h>

int mb_set_err(mb_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

void* mb_scratch(mb_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return ctx->scratch;
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = mb_align_up(bytes + bytes / 4, (size_t)1 << 20);
    if (cudaMalloc(&ctx->scratch, want) != cudaSuccess) {
        cudaGetLastError();
        mb_set_err(ctx, MB_ERR_OOM, "scratch allocation of %zu bytes failed", want);
        return nullptr;
    }
    ctx->scratch_bytes = want;
    return ctx->scratch;
}

void mb_free_models(mb_ctx* ctx);   // craft.cu / trocr.cu

extern "C" const char* mb_version(void) { return "marie_b200 0.1 (sm_100a)"; }

extern "C" int mb_init(int device, mb_ctx** out) {
    if (!out) return MB_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return MB_ERR_NO_DEVICE;   // no CPU fallback by design
    }
    if (device < 0 || device >= count) return MB_ERR_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MB_ERR_CUDA;
    if (prop.major != 10) return MB_ERR_NO_DEVICE;   // kernels are sm_100a only
    mb_ctx* ctx = new mb_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    MbDeviceGuard guard(ctx);            // the caller's current device is left as it was
    if (cudaMalloc(&ctx->dev_diag, 64) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return MB_ERR_CUDA;
    }
    cudaMemset(ctx->dev_diag, 0, 64);
    *out = ctx;
    return MB_OK;
}

extern "C" void mb_free(mb_ctx* ctx) {
    if (!ctx) return;
    MbDeviceGuard guard(ctx);
    mb_free_models(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->dev_diag) cudaFree(ctx->dev_diag);
    delete ctx;
}

extern "C" int mb_set_dtype(mb_ctx* ctx, int dtype) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    if (dtype != MB_DTYPE_BF16 && dtype != MB_DTYPE_F16) return mb_set_err(ctx, MB_ERR_ARG, "mb_set_dtype: unknown dtype %d", dtype);
    if ((dtype == MB_DTYPE_F16) != (ctx->f16 != 0)) mb_free_models(ctx);   // packed weights are dtype-specific: unload
    ctx->f16 = dtype == MB_DTYPE_F16;
    return MB_OK;
}
extern "C" int mb_get_dtype(const mb_ctx* ctx) { return ctx && ctx->f16 ? MB_DTYPE_F16 : MB_DTYPE_BF16; }

// Live timing of the tensor-core GEMM launches (bench.py roofline).  enable != 0 starts (and resets) the collection;
// mb_profile_read drains it: out3 = {sum of launch durations [ms], sum of algorithmic FLOPs, launches}.
extern "C" int mb_profile_enable(mb_ctx* ctx, int enable) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    mb_profile_drain(ctx);
    ctx->profile = enable != 0;
    ctx->prof_ms_total = ctx->prof_flops_total = 0;
    ctx->prof_launches = 0;
    return MB_OK;
}
extern "C" int mb_profile_read(mb_ctx* ctx, double* out3) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx || !out3) return MB_ERR_ARG;
    mb_profile_drain(ctx);
    out3[0] = ctx->prof_ms_total; out3[1] = ctx->prof_flops_total; out3[2] = (double)ctx->prof_launches;
    return MB_OK;
}

extern "C" const char* mb_last_error(const mb_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" unsigned long long mb_launch_count(const mb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// Reads (and clears) the diagnostic word kernels write before trapping (debug aid for tests).
extern "C" unsigned int mb_debug_diag(mb_ctx* ctx) {
    MbDeviceGuard _mb_guard(ctx);
    unsigned int v = 0;
    if (ctx && ctx->dev_diag) cudaMemcpy(&v, ctx->dev_diag, 4, cudaMemcpyDeviceToHost);
    return v;
}
