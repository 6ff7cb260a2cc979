// Shared declarations for libmarie_b200.so (sm_100a only).
// Host-side context, error plumbing and small device helpers used by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/marie_b200.h"

typedef __nv_bfloat16 bf16;

#define MB_NUM_SMS_DEFAULT 148

// Every C-ABI entry point funnels errors through this: negative code + message kept in the ctx.
struct mb_ctx {
    int device = 0;
    int num_sms = MB_NUM_SMS_DEFAULT;
    char err[512] = {0};
    // scratch owned by the context (grown on demand, never shrunk)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // model state (opaque here; defined in craft.cu / trocr.cu)
    struct CraftModel* craft = nullptr;
    struct TrocrModel* trocr = nullptr;
    struct RefineModel* refine = nullptr;   // refine.cu
    // device-side diagnostic word written by kernels before __trap()
    unsigned int* dev_diag = nullptr;
    // 16-bit element type of activations / weights: 0 = bf16, 1 = fp16 (default; mb_set_dtype)
    int f16 = 1;
    // optional live timing of the tap-GEMM launches (bench.py roofline): CUDA event pairs on the launch stream
    int profile = 0;
    std::vector<cudaEvent_t> prof_events;   // start/stop pairs
    std::vector<double> prof_flops;
    double prof_ms_total = 0, prof_flops_total = 0;
    unsigned long long prof_launches = 0;
    // kernel launch counter (bench.py "gpu_launches")
    unsigned long long launches = 0;
};

int mb_set_err(mb_ctx* ctx, int code, const char* fmt, ...);

// Every C-ABI entry point that touches the device selects the context's device for its own duration and restores the
// caller's (a Marie executor thread may have another device current; cudaMalloc / launches follow the calling thread).
struct MbDeviceGuard {
    int prev = -1;
    explicit MbDeviceGuard(const mb_ctx* ctx) {
        if (!ctx) return;
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; return; }
        if (prev != ctx->device) cudaSetDevice(ctx->device); else prev = -1;
    }
    ~MbDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
void* mb_scratch(mb_ctx* ctx, size_t bytes);   // returns nullptr + sets error on failure

#define MB_CUDA(ctx, expr)                                                             \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess)                                                         \
            return mb_set_err((ctx), MB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, \
                              #expr, cudaGetErrorString(_e));                          \
    } while (0)

#define MB_LAUNCH_CHECK(ctx)                                                            \
    do {                                                                               \
        (ctx)->launches++;                                                             \
        cudaError_t _e = cudaGetLastError();                                           \
        if (_e != cudaSuccess)                                                         \
            return mb_set_err((ctx), MB_ERR_CUDA, "%s:%d launch -> %s", __FILE__,       \
                              __LINE__, cudaGetErrorString(_e));                       \
    } while (0)

#define MB_REQUIRE(ctx, cond, ...)                                   \
    do {                                                             \
        if (!(cond)) return mb_set_err((ctx), MB_ERR_ARG, __VA_ARGS__); \
    } while (0)

static inline int mb_cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t mb_align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------------
// tap-GEMM (gemm_tc.cu): D = epilogue( sum_taps A_tap * W^T ), bf16 in, fp32 accumulate in TMEM.
// One kernel serves 3x3 / dilated / 1x1 convolutions over NHWC activations (implicit GEMM through
// shifted TMA boxes, zero padding by TMA out-of-bounds fill) and plain row-major GEMMs
// (n = h = 1, w = M rows).
// ---------------------------------------------------------------------------------------------
enum { MB_ACT_NONE = 0, MB_ACT_RELU = 1, MB_ACT_GELU = 2 };
enum { MB_OUT_BF16 = 0, MB_OUT_F32 = 1, MB_OUT_F32_PLANAR = 2 };

struct TapGemm {
    // activations: up to two NHWC sources concatenated along channels (k index = tap*(c0+c1)+c)
    const bf16* a0 = nullptr; int c0 = 0; int a0_ld = 0;   // a0_ld: pixel pitch in elements
    const bf16* a1 = nullptr; int c1 = 0; int a1_ld = 0;
    int n = 1, h = 1, w = 1;      // output (and input) spatial extent; plain GEMM: w = M
    int taps = 1;                 // 1 or 9 (3x3, padding = dilation)
    int dil = 1;
    const bf16* wgt = nullptr;    // [n_rows_w, taps*(c0+c1)] row-major (K contiguous)
    int n_rows_w = 0;             // rows present in wgt (>= n_out, padding rows are zero)
    int n_out = 0;                // columns written
    int block_n = 0;              // 0 = pick automatically
    const float* bias = nullptr;  // [n_out] fp32 or null
    int act = MB_ACT_NONE;
    const bf16* residual = nullptr; int res_ld = 0;  // added after activation
    void* out = nullptr; long long out_ld = 0; int out_mode = MB_OUT_BF16;
    long long out_plane = 0;      // MB_OUT_F32_PLANAR: elements between channel planes
    // block-diagonal batches (per-head projections): batch b reads A columns [b*a_col_stride, +c0), weight rows
    // [b*w_row_stride, +n_out) and writes output / bias / residual columns [b*out_col_stride, +n_out)
    int batches = 1, a_col_stride = 0, w_row_stride = 0, out_col_stride = 0;
    // LayerNorm folded into the GEMM (encoder qkv / fc1): A is the RAW residual stream, wgt holds W * gamma, and the
    // epilogue computes rstd_r * (acc - mean_r * ln_c[n]) + bias[n] with ln_c[n] = sum_k wgt[n, k] and
    // bias[n] = b[n] + sum_k beta_k W[n, k].  ln_stats: per row (-mean, rstd) pairs (trocr.cu ln_stats_kernel).
    const float* ln_stats = nullptr;
    const float* ln_c = nullptr;
    // Row statistics of the OUTPUT, for the LayerNorm that follows a residual GEMM (encoder proj / fc2): the TMA epilogue
    // adds up x and x^2 of the values it writes (fp32, before the 16-bit rounding) and stores one (sum, sum of squares)
    // pair per row, 256-column N tile and epilogue-warp half: stat_out[row * 2 * n_tiles + 2 * tile + half] as float2.
    // Needs block_n = 256 (set it), a residual, no activation, the TMA epilogue; saves the separate pass over the stream.
    float* stat_out = nullptr;
};
int mb_tap_gemm(mb_ctx* ctx, const TapGemm& p, cudaStream_t stream);
void mb_profile_drain(mb_ctx* ctx);

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float bf16_bits_to_f32(unsigned short b) {
    return __uint_as_float(((unsigned int)b) << 16);
}
__device__ __forceinline__ unsigned int pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<unsigned int*>(&v);
}
// Runtime-selected 16-bit element type (f16 != 0: IEEE half, else bfloat16).  Buffers are typed `bf16*` in the
// sources purely as "16-bit element"; all arithmetic is fp32.
__device__ __forceinline__ float2 unpack2(unsigned int v, int f16) {
    if (f16) return __half22float2(*reinterpret_cast<__half2*>(&v));
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}
__device__ __forceinline__ unsigned int pack2(float lo, float hi, int f16) {
    if (f16) {
        // saturate to +-65504 instead of overflowing to inf: one F2FP.SATFINITE (four FMNMX + F2FP before)
        unsigned int d;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
        return d;
    }
    return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ float load16(const bf16* p, int f16) {
    const unsigned short b = *reinterpret_cast<const unsigned short*>(p);
    if (f16) return __half2float(__ushort_as_half(b));
    return bf16_bits_to_f32(b);
}
__device__ __forceinline__ void store16(bf16* p, float v, int f16) {
    if (f16) {
        v = fminf(fmaxf(v, -65504.f), 65504.f);
        *reinterpret_cast<__half*>(p) = __float2half_rn(v);
    } else {
        *p = __float2bfloat16_rn(v);
    }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif
