// K2-K4: the CRAFT network (VGG16-BN trunk + U-Net decoder + classification head), NHWC bf16.
// Reference: CRAFT.forward (marie/models/craft/craft.py:59-81), vgg16_bn.forward
// (marie/models/craft/basenet/vgg16_bn.py:23-73), double_conv (craft.py:14-28).
//
// Every 3x3 / 1x1 convolution with C_in >= 64 runs on the tcgen05 tap-GEMM (gemm_tc.cu); BatchNorm is folded into
// weights/bias by the packer (marie-icr_b200/weights.py).  The U-Net `torch.cat` is never materialised: the
// 1x1 convolutions read their two sources through two TMA descriptors.  The small kernels here are the
// memory-bound glue: conv1_1 (C_in = 3, direct), 2x2 / 3x3 max-pool and the bilinear x2 up-sampling.
//
// Skip taps (SURVEY.md hard part 9): relu2_2 / relu3_2 / relu4_3 are post-ReLU (in-place ReLU of the next slice),
// relu5_3 = BN(conv5_2) without ReLU, fc6/fc7 have no activation.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "blob.cuh"

struct CraftLayer {
    const bf16* w = nullptr;
    const float* b = nullptr;
    int rows = 0, k = 0;
};

struct CraftModel {
    WeightBlob blob;
    const float* c11_w = nullptr;   // [27][64] fp32
    const float* c11_b = nullptr;
    CraftLayer L[32];
    // activation arena
    void* arena = nullptr;
    size_t arena_bytes = 0;
};

namespace {

enum {
    L_C12 = 0, L_C21, L_C22, L_C31, L_C32, L_C33, L_C41, L_C42, L_C43, L_C51, L_C52, L_FC6, L_FC7,
    L_U1A, L_U1B, L_U2A, L_U2B, L_U3A, L_U3B, L_U4A, L_U4B, L_H1, L_H2, L_H3, L_H4, L_H5, L_COUNT
};
const char* kLayerNames[L_COUNT] = {
    "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3", "conv5_1",
    "conv5_2", "fc6", "fc7", "upconv1a", "upconv1b", "upconv2a", "upconv2b", "upconv3a", "upconv3b", "upconv4a",
    "upconv4b", "cls1", "cls2", "cls3", "cls4", "cls5"};

// conv1_1: [n,h,w,4] bf16 -> [n,h,w,64] bf16, 3x3 pad 1, folded BN, ReLU.  One thread per output pixel.
__global__ void __launch_bounds__(128) conv1_1_kernel(const bf16* __restrict__ in, const float* __restrict__ wgt,
                                                       const float* __restrict__ bias, bf16* __restrict__ out, int n,
                                                       int h, int w, int f16) {
    __shared__ __align__(16) float sw[27 * 64];
    __shared__ __align__(16) float sb[64];
    for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i] = wgt[i];
    if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int img = blockIdx.z;
    float v[27];                               // (lanes beyond the row compute on zeros and store nothing)
    const uint2* src = reinterpret_cast<const uint2*>(in) + (long long)img * h * w;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int yy = y + ky - 1, xx = x + kx - 1;
            float a = 0.f, b = 0.f, c = 0.f;
            if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
                const uint2 p = __ldg(src + (long long)yy * w + xx);
                const float2 ab = unpack2(p.x, f16);
                a = ab.x;
                b = ab.y;
                c = unpack2(p.y, f16).x;
            }
            v[(ky * 3 + kx) * 3 + 0] = a;
            v[(ky * 3 + kx) * 3 + 1] = b;
            v[(ky * 3 + kx) * 3 + 2] = c;
        }
    }
    // the 128 output bytes of a pixel are produced 16 at a time; they are staged per warp in shared memory (XOR-swizzled
    // 16-byte chunks) and written out as 512 contiguous bytes per warp instruction instead of 32 scattered 16-byte pieces
    __shared__ __align__(16) uint4 stage[4][32][8];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
    for (int c8 = 0; c8 < 8; ++c8) {
        // packed fp32 arithmetic (FFMA2): two output channels per instruction, same per-channel operation order
        const ulonglong2 b01 = *reinterpret_cast<const ulonglong2*>(sb + c8 * 8);
        const ulonglong2 b23 = *reinterpret_cast<const ulonglong2*>(sb + c8 * 8 + 4);
        f32x2 p0 = b01.x, p1 = b01.y, p2 = b23.x, p3 = b23.y;
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const ulonglong2 w01 = *reinterpret_cast<const ulonglong2*>(sw + k * 64 + c8 * 8);
            const ulonglong2 w23 = *reinterpret_cast<const ulonglong2*>(sw + k * 64 + c8 * 8 + 4);
            const f32x2 vv = pk2(v[k], v[k]);
            p0 = fma2(vv, w01.x, p0); p1 = fma2(vv, w01.y, p1);
            p2 = fma2(vv, w23.x, p2); p3 = fma2(vv, w23.y, p3);
        }
        float4 a0, a1;
        upk2(p0, a0.x, a0.y); upk2(p1, a0.z, a0.w); upk2(p2, a1.x, a1.y); upk2(p3, a1.z, a1.w);
        uint4 r;
        r.x = pack2(fmaxf(a0.x, 0.f), fmaxf(a0.y, 0.f), f16);
        r.y = pack2(fmaxf(a0.z, 0.f), fmaxf(a0.w, 0.f), f16);
        r.z = pack2(fmaxf(a1.x, 0.f), fmaxf(a1.y, 0.f), f16);
        r.w = pack2(fmaxf(a1.z, 0.f), fmaxf(a1.w, 0.f), f16);
        stage[wid][lane][c8 ^ (lane & 7)] = r;
    }
    __syncwarp();
    const int x_warp = blockIdx.x * blockDim.x + wid * 32;            // first pixel of this warp
    const int n_px = min(32, w - x_warp);                             // pixels of the warp inside the row
    uint4* o = reinterpret_cast<uint4*>(out + (((long long)img * h + y) * w + x_warp) * 64);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int idx = k * 32 + lane, px = idx >> 3, ch = idx & 7;
        if (px < n_px) o[idx] = stage[wid][px][ch ^ (px & 7)];
    }
}

// max of two packed 16-bit floats (bf16 and fp16 share sign-magnitude ordering): compare as sign-folded integers
__device__ __forceinline__ uint32_t f16x2_max(uint32_t a, uint32_t b) {
    auto key = [](uint32_t v) -> int { return (v & 0x8000u) ? -(int)(v & 0x7FFFu) : (int)(v & 0x7FFFu); };
    const uint32_t lo = key(a & 0xFFFFu) >= key(b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
    const uint32_t hi = key(a >> 16) >= key(b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
}
__device__ __forceinline__ uint4 u4max(uint4 a, uint4 b) {
    return make_uint4(f16x2_max(a.x, b.x), f16x2_max(a.y, b.y), f16x2_max(a.z, b.z), f16x2_max(a.w, b.w));
}

// 2x2 stride-2 max-pool, NHWC, 8 channels (16 B) per thread
__global__ void maxpool2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long total, int oh,
                                int ow, int c8) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int c = (int)(i % c8);
        long long r = i / c8;
        const int x = (int)(r % ow); r /= ow;
        const int y = (int)(r % oh);
        const long long img = r / oh;
        const long long iw = 2LL * ow;
        const uint4* p = in + ((img * 2 * oh + 2 * y) * iw + 2 * x) * c8 + c;
        out[i] = u4max(u4max(__ldg(p), __ldg(p + c8)), u4max(__ldg(p + iw * c8), __ldg(p + iw * c8 + c8)));
    }
}

// 3x3 stride-1 pad-1 max-pool (padding never wins: implicit -inf)
__global__ void maxpool3s1_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long total, int h,
                                  int w, int c8) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int c = (int)(i % c8);
        long long r = i / c8;
        const int x = (int)(r % w); r /= w;
        const int y = (int)(r % h);
        const long long img = r / h;
        uint4 m = __ldg(in + i);
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= h) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= w) continue;
                m = u4max(m, __ldg(in + ((img * h + yy) * w + xx) * c8 + c));
            }
        }
        out[i] = m;
    }
}

// bilinear x2, align_corners=False (F.interpolate, craft.py:67,71,75): fp32 arithmetic in PyTorch's order
__global__ void upsample2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long total, int ih,
                                 int iw, int c8, int f16) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int oh = 2 * ih, ow = 2 * iw;
    for (; i < total; i += stride) {
        const int c = (int)(i % c8);
        long long r = i / c8;
        const int x = (int)(r % ow); r /= ow;
        const int y = (int)(r % oh);
        const long long img = r / oh;
        float sy = 0.5f * (y + 0.5f) - 0.5f, sx = 0.5f * (x + 0.5f) - 0.5f;
        if (sy < 0.f) sy = 0.f;
        if (sx < 0.f) sx = 0.f;
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < ih - 1 ? 1 : 0), x1 = x0 + (x0 < iw - 1 ? 1 : 0);
        const float ly = sy - y0, lx = sx - x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const uint4* base = in + img * ih * iw * c8 + c;
        const uint4 p00 = __ldg(base + ((long long)y0 * iw + x0) * c8), p01 = __ldg(base + ((long long)y0 * iw + x1) * c8);
        const uint4 p10 = __ldg(base + ((long long)y1 * iw + x0) * c8), p11 = __ldg(base + ((long long)y1 * iw + x1) * c8);
        const uint32_t a[4] = {p00.x, p00.y, p00.z, p00.w}, b[4] = {p01.x, p01.y, p01.z, p01.w};
        const uint32_t cc[4] = {p10.x, p10.y, p10.z, p10.w}, d[4] = {p11.x, p11.y, p11.z, p11.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 fa = unpack2(a[k], f16), fb = unpack2(b[k], f16), fc = unpack2(cc[k], f16), fd = unpack2(d[k], f16);
            const float lo = hy * (hx * fa.x + lx * fb.x) + ly * (hx * fc.x + lx * fd.x);
            const float hi = hy * (hx * fa.y + lx * fb.y) + ly * (hx * fc.y + lx * fd.y);
            o[k] = pack2(lo, hi, f16);
        }
        out[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

int grid_for(mb_ctx* ctx, long long total, int threads) {
    long long g = (total + threads - 1) / threads;
    const long long cap = (long long)ctx->num_sms * 32;
    return (int)(g < cap ? g : cap);
}

struct Act { bf16* p; int c; };

int conv(mb_ctx* ctx, const CraftModel* m, int layer, Act a0, Act a1, int n, int h, int w, int taps, int dil, int act,
         void* out, int out_c, int out_mode, long long out_plane, cudaStream_t s) {
    const CraftLayer& L = m->L[layer];
    TapGemm g;
    g.a0 = a0.p; g.c0 = a0.c; g.a0_ld = a0.c;
    g.a1 = a1.p; g.c1 = a1.p ? a1.c : 0; g.a1_ld = a1.c;
    g.n = n; g.h = h; g.w = w; g.taps = taps; g.dil = dil;
    g.wgt = L.w; g.n_rows_w = L.rows; g.n_out = out_c;
    g.bias = L.b; g.act = act;
    g.out = out; g.out_ld = out_mode == MB_OUT_F32_PLANAR ? 1 : out_c; g.out_mode = out_mode; g.out_plane = out_plane;
    if (L.k != taps * (g.c0 + g.c1))
        return mb_set_err(ctx, MB_ERR_STATE, "craft: layer %s expects K=%d, got %d", kLayerNames[layer], L.k,
                          taps * (g.c0 + g.c1));
    return mb_tap_gemm(ctx, g, s);
}

}  // namespace

void mb_free_craft(mb_ctx* ctx) {
    if (!ctx->craft) return;
    ctx->craft->blob.release();
    if (ctx->craft->arena) cudaFree(ctx->craft->arena);
    delete ctx->craft;
    ctx->craft = nullptr;
}

extern "C" int mb_load_craft(mb_ctx* ctx, const void* blob_host, size_t nbytes) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    mb_free_craft(ctx);
    CraftModel* m = new CraftModel();
    ctx->craft = m;
    int rc = m->blob.load(ctx, blob_host, nbytes);
    if (rc) { mb_free_craft(ctx); return rc; }
    const BlobTensor* w = m->blob.get("conv1_1.w");
    const BlobTensor* b = m->blob.get("conv1_1.b");
    if (!w || !b || w->dtype != 0 || w->nbytes != 27 * 64 * 4) {
        mb_free_craft(ctx);
        return mb_set_err(ctx, MB_ERR_ARG, "craft blob: conv1_1 missing or malformed");
    }
    m->c11_w = (const float*)w->dev;
    m->c11_b = (const float*)b->dev;
    for (int i = 0; i < L_COUNT; ++i) {
        const BlobTensor* lw = m->blob.get(std::string(kLayerNames[i]) + ".w");
        const BlobTensor* lb = m->blob.get(std::string(kLayerNames[i]) + ".b");
        if (!lw || !lb || lw->dtype != (ctx->f16 ? 3 : 1) || lb->dtype != 0 || lw->ndim != 2) {
            mb_free_craft(ctx);
            return mb_set_err(ctx, MB_ERR_ARG, "craft blob: layer %s missing, malformed or not packed for the context dtype (%s)", kLayerNames[i],
                              ctx->f16 ? "fp16" : "bf16");
        }
        m->L[i].w = (const bf16*)lw->dev;
        m->L[i].b = (const float*)lb->dev;
        m->L[i].rows = (int)lw->dims[0];
        m->L[i].k = (int)lw->dims[1];
    }
    return 0;
}

// x: [n, h, w, 4] bf16 NHWC (h, w multiples of 32) -> text/link [n, h/2, w/2] fp32
int mb_craft_forward_impl(mb_ctx* ctx, const bf16* x, int n, int h, int w, float* text, float* link,
                          bf16* feature_out, cudaStream_t s) {
    CraftModel* m = ctx->craft;
    if (!m) return mb_set_err(ctx, MB_ERR_STATE, "craft: weights not loaded (mb_load_craft)");
    MB_REQUIRE(ctx, h % 32 == 0 && w % 32 == 0 && n > 0, "craft: input must be a multiple of 32 (got %dx%d)", h, w);
    const long long P1 = (long long)n * h * w, P2 = P1 / 4, P4 = P1 / 16, P8 = P1 / 64, P16 = P1 / 256;
    // arena layout (elements of bf16); big ping-pong buffers A/B sized for the largest tensor (64 ch @ full res)
    size_t off = 0;
    auto take = [&](long long elems) { size_t o = off; off += mb_align_up((size_t)elems * 2, 1024); return o; };
    const size_t oA = take(P1 * 64), oB = take(P1 * 64);
    const size_t o_c22 = take(P2 * 128), o_c32 = take(P4 * 256), o_c42 = take(P8 * 512), o_c52 = take(P16 * 512);
    const size_t o_fc7 = take(P16 * 1024);
    if (off > m->arena_bytes) {
        if (m->arena) cudaFree(m->arena);
        m->arena = nullptr; m->arena_bytes = 0;
        if (cudaMalloc(&m->arena, off) != cudaSuccess) {
            cudaGetLastError();
            return mb_set_err(ctx, MB_ERR_OOM, "craft: activation arena of %zu bytes failed", off);
        }
        m->arena_bytes = off;
    }
    unsigned char* base = (unsigned char*)m->arena;
    bf16* A = (bf16*)(base + oA);
    bf16* B = (bf16*)(base + oB);
    bf16* c22 = (bf16*)(base + o_c22);
    bf16* c32 = (bf16*)(base + o_c32);
    bf16* c42 = (bf16*)(base + o_c42);
    bf16* c52 = (bf16*)(base + o_c52);
    bf16* fc7 = (bf16*)(base + o_fc7);
    const Act none = {nullptr, 0};
    int rc;
#define CONV(layer, in, cin, H, W, taps, dil, act, out, cout)                                                   \
    if ((rc = conv(ctx, m, layer, Act{in, cin}, none, n, H, W, taps, dil, act, out, cout, MB_OUT_BF16, 0, s))) return rc;
#define POOL2(in, out, OH, OW, C)                                                                               \
    {                                                                                                           \
        const long long tot = (long long)n * (OH) * (OW) * ((C) / 8);                                           \
        maxpool2_kernel<<<grid_for(ctx, tot, 256), 256, 0, s>>>((const uint4*)(in), (uint4*)(out), tot, OH, OW, (C) / 8); \
        MB_LAUNCH_CHECK(ctx);                                                                                   \
    }
#define UP2(in, out, IH, IW, C)                                                                                 \
    {                                                                                                           \
        const long long tot = (long long)n * (IH) * 2 * (IW) * 2 * ((C) / 8);                                   \
        upsample2_kernel<<<grid_for(ctx, tot, 256), 256, 0, s>>>((const uint4*)(in), (uint4*)(out), tot, IH, IW, (C) / 8, ctx->f16); \
        MB_LAUNCH_CHECK(ctx);                                                                                   \
    }
    {
        dim3 grid(mb_cdiv(w, 128), h, n);
        conv1_1_kernel<<<grid, 128, 0, s>>>(x, m->c11_w, m->c11_b, A, n, h, w, ctx->f16);
        MB_LAUNCH_CHECK(ctx);
    }
    CONV(L_C12, A, 64, h, w, 9, 1, MB_ACT_RELU, B, 64);
    POOL2(B, A, h / 2, w / 2, 64);
    CONV(L_C21, A, 64, h / 2, w / 2, 9, 1, MB_ACT_RELU, B, 128);
    CONV(L_C22, B, 128, h / 2, w / 2, 9, 1, MB_ACT_RELU, c22, 128);
    POOL2(c22, A, h / 4, w / 4, 128);
    CONV(L_C31, A, 128, h / 4, w / 4, 9, 1, MB_ACT_RELU, B, 256);
    CONV(L_C32, B, 256, h / 4, w / 4, 9, 1, MB_ACT_RELU, c32, 256);
    CONV(L_C33, c32, 256, h / 4, w / 4, 9, 1, MB_ACT_RELU, A, 256);
    POOL2(A, B, h / 8, w / 8, 256);
    CONV(L_C41, B, 256, h / 8, w / 8, 9, 1, MB_ACT_RELU, A, 512);
    CONV(L_C42, A, 512, h / 8, w / 8, 9, 1, MB_ACT_RELU, c42, 512);
    CONV(L_C43, c42, 512, h / 8, w / 8, 9, 1, MB_ACT_RELU, A, 512);
    POOL2(A, B, h / 16, w / 16, 512);
    CONV(L_C51, B, 512, h / 16, w / 16, 9, 1, MB_ACT_RELU, A, 512);
    CONV(L_C52, A, 512, h / 16, w / 16, 9, 1, MB_ACT_NONE, c52, 512);
    {
        const long long tot = (long long)n * (h / 16) * (w / 16) * (512 / 8);
        maxpool3s1_kernel<<<grid_for(ctx, tot, 256), 256, 0, s>>>((const uint4*)c52, (uint4*)A, tot, h / 16, w / 16, 512 / 8);
        MB_LAUNCH_CHECK(ctx);
    }
    CONV(L_FC6, A, 512, h / 16, w / 16, 9, 6, MB_ACT_NONE, B, 1024);
    CONV(L_FC7, B, 1024, h / 16, w / 16, 1, 1, MB_ACT_NONE, fc7, 1024);
    // U-Net decoder: 1x1 over cat(y, skip) through two TMA sources, then 3x3
    if ((rc = conv(ctx, m, L_U1A, Act{fc7, 1024}, Act{c52, 512}, n, h / 16, w / 16, 1, 1, MB_ACT_RELU, A, 512, MB_OUT_BF16, 0, s))) return rc;
    CONV(L_U1B, A, 512, h / 16, w / 16, 9, 1, MB_ACT_RELU, B, 256);
    UP2(B, A, h / 16, w / 16, 256);
    if ((rc = conv(ctx, m, L_U2A, Act{A, 256}, Act{c42, 512}, n, h / 8, w / 8, 1, 1, MB_ACT_RELU, B, 256, MB_OUT_BF16, 0, s))) return rc;
    CONV(L_U2B, B, 256, h / 8, w / 8, 9, 1, MB_ACT_RELU, A, 128);
    UP2(A, B, h / 8, w / 8, 128);
    if ((rc = conv(ctx, m, L_U3A, Act{B, 128}, Act{c32, 256}, n, h / 4, w / 4, 1, 1, MB_ACT_RELU, A, 128, MB_OUT_BF16, 0, s))) return rc;
    CONV(L_U3B, A, 128, h / 4, w / 4, 9, 1, MB_ACT_RELU, B, 64);
    UP2(B, A, h / 4, w / 4, 64);
    if ((rc = conv(ctx, m, L_U4A, Act{A, 64}, Act{c22, 128}, n, h / 2, w / 2, 1, 1, MB_ACT_RELU, B, 64, MB_OUT_BF16, 0, s))) return rc;
    // upconv4b: 64 -> 32 real channels, stored padded to 64 (zero rows in the weight matrix) = `feature`
    bf16* feat = feature_out ? feature_out : A;
    CONV(L_U4B, B, 64, h / 2, w / 2, 9, 1, MB_ACT_RELU, feat, 64);
    CONV(L_H1, feat, 64, h / 2, w / 2, 9, 1, MB_ACT_RELU, B, 64);
    bf16* t2 = feature_out ? A : c22;   // c22 is dead after upconv4a
    CONV(L_H2, B, 64, h / 2, w / 2, 9, 1, MB_ACT_RELU, t2, 64);
    CONV(L_H3, t2, 64, h / 2, w / 2, 9, 1, MB_ACT_RELU, B, 64);
    CONV(L_H4, B, 64, h / 2, w / 2, 1, 1, MB_ACT_RELU, t2, 64);
    // cls5: 2 output channels written as fp32 planes; text/link must be adjacent planes of one buffer
    MB_REQUIRE(ctx, link == text + P2, "craft: link plane must follow the text plane");
    if ((rc = conv(ctx, m, L_H5, Act{t2, 64}, none, n, h / 2, w / 2, 1, 1, MB_ACT_NONE, text, 2, MB_OUT_F32_PLANAR, P2, s))) return rc;
#undef CONV
#undef POOL2
#undef UP2
    return 0;
}

extern "C" int mb_craft_forward(mb_ctx* ctx, const void* x_dev, int n, int h, int w, float* scores_dev,
                                void* feature_dev, void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    const long long P2 = (long long)n * (h / 2) * (w / 2);
    return mb_craft_forward_impl(ctx, (const bf16*)x_dev, n, h, w, scores_dev, scores_dev + P2, (bf16*)feature_dev,
                                 (cudaStream_t)stream);
}
