// NVCC_FLAGS: -fmad=false
// K1 (page preprocessing) and K9 (crop -> 384x384 network input) — the two exact image resamplers.
//
// K1  reference: resize_aspect_ratio + normalizeMeanVariance (marie/models/craft/imgproc.py:45-73,26-32) and the
//     CHW/H2D pre-amble of get_prediction (marie/boxes/craft_box_processor.py:96-106).
//     cv2.resize(INTER_LINEAR) on u8 is a two-pass fixed-point filter: 11-bit coefficients from float32
//     fractions, horizontal sums kept as int, vertical combine ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2)>>2.
//     Output: NHWC bf16 with C padded to 4 (x,y beyond the resized image are the zero canvas -> -1.0).
// K9  reference: MemoryDataset.__getitem__ (BGR->RGB, marie/models/icr/memory_dataset.py:43-53) +
//     preprocess_image (PIL bicubic 384x384, ToTensor, Normalize(0.5,0.5);
//     marie/document/trocr_ocr_processor.py:95-101,116-125).  Pillow's resampler: double-precision bicubic
//     coefficients (support scaled by the down-scale factor), 22-bit fixed point, horizontal pass to a u8
//     intermediate, then vertical pass.  Coefficients are evaluated here in fp64 with the same operation order
//     (this file is compiled with -fmad=false so nothing is contracted).
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ K1
struct LinCoef { int ofs; short a0, a1; };

__global__ void lin_coef_kernel(LinCoef* __restrict__ tab, int ssize, int dsize) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dsize) return;
    const double scale = (double)ssize / (double)dsize;
    float fx = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(fx);
    fx -= (float)s;
    if (s < 0) { fx = 0.f; s = 0; }
    if (s >= ssize - 1) { fx = 0.f; s = ssize - 1; }
    LinCoef c;
    c.ofs = s;
    c.a0 = (short)__float2int_rn((1.f - fx) * 2048.f);
    c.a1 = (short)__float2int_rn(fx * 2048.f);
    tab[d] = c;
}

__global__ void page_preprocess_kernel(const uint8_t* __restrict__ pages, long long page_stride, int sh, int sw,
                                       const LinCoef* __restrict__ xtab, const LinCoef* __restrict__ ytab, int th,
                                       int tw, int oh, int ow, bf16* __restrict__ out, int f16) {
    __shared__ float lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = ((float)i - 127.5f) / 127.5f;
    __syncthreads();
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    const int page = blockIdx.z;
    if (ox >= ow) return;
    float v0 = -1.0f, v1 = -1.0f, v2 = -1.0f;   // zero canvas after (0 - 127.5) / 127.5
    if (oy < th && ox < tw) {
        const LinCoef cx = xtab[ox], cy = ytab[oy];
        const int x0 = cx.ofs, x1 = min(cx.ofs + 1, sw - 1);
        const int y0 = cy.ofs, y1 = min(cy.ofs + 1, sh - 1);
        const uint8_t* p = pages + (long long)page * page_stride;
        const uint8_t* r0 = p + ((long long)y0 * sw) * 3;
        const uint8_t* r1 = p + ((long long)y1 * sw) * 3;
        int res[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int s0 = r0[x0 * 3 + c] * cx.a0 + r0[x1 * 3 + c] * cx.a1;
            const int s1 = r1[x0 * 3 + c] * cx.a0 + r1[x1 * 3 + c] * cx.a1;
            int v = (((cy.a0 * (s0 >> 4)) >> 16) + ((cy.a1 * (s1 >> 4)) >> 16) + 2) >> 2;
            res[c] = min(max(v, 0), 255);
        }
        v0 = lut[res[0]]; v1 = lut[res[1]]; v2 = lut[res[2]];
    }
    uint2 o;
    o.x = pack2(v0, v1, f16);
    o.y = pack2(v2, 0.f, f16);
    reinterpret_cast<uint2*>(out)[((long long)page * oh + oy) * ow + ox] = o;
}

// ------------------------------------------------------------------------------------------------ K9
constexpr int OUT = 384;
constexpr int PREC_BITS = 32 - 8 - 2;
constexpr int TILE_ROWS = 16;           // output rows per CTA (= one row of 16x16 patches)
constexpr int K9_THREADS = 384;
constexpr int K9_SMEM_BYTES = 200 * 1024;   // large-crop launch (1 CTA / SM)
constexpr int K9_SMEM_SMALL = 44 * 1024;    // common case: 5 CTAs / SM

struct CropDesc { const uint8_t* base; int pitch; int w; int h; };

// worst-case dynamic shared memory a crop needs in crop_resize_kernel (tables + the widest band of source rows)
__host__ __device__ inline size_t k9_smem_need(int w, int h);

__device__ __forceinline__ double bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for one output index (box = (0, in_size)).
__device__ __forceinline__ void pil_coef(int in_size, int xx, int ksize, int* xmin_out, int* n_out, int* kk) {
    const double scale = (double)((float)in_size - 0.0f) / OUT;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 2.0 * filterscale;
    const double center = 0.0 + (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) ww += bicubic((x + xmin - center + 0.5) * ss);
    for (int x = 0; x < xmax; ++x) {
        double w = bicubic((x + xmin - center + 0.5) * ss);
        if (ww != 0.0) w /= ww;
        kk[x] = (w < 0) ? (int)(-0.5 + w * (double)(1 << PREC_BITS)) : (int)(0.5 + w * (double)(1 << PREC_BITS));
    }
    for (int x = xmax; x < ksize; ++x) kk[x] = 0;
    *xmin_out = xmin;
    *n_out = xmax;
}

__device__ __forceinline__ int pil_ksize(int in_size) {
    const double scale = (double)((float)in_size) / OUT;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(2.0 * filterscale) * 2 + 1;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= PREC_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// grid = (OUT / TILE_ROWS, n_crops).  layout 0: [N,3,384,384] (RGB planes); layout 1: patch rows
// [N*576, 768] with k = c*256 + py*16 + px (the A operand of the ViT patch-embedding GEMM).
__global__ void __launch_bounds__(K9_THREADS)
crop_resize_kernel(const CropDesc* __restrict__ crops, bf16* __restrict__ out, int layout, int* __restrict__ err,
                   int f16, int smem_limit, int small_only, int tiles_per_cta) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ float lut[256];
    __shared__ int s_r0, s_r1;
    const CropDesc cd = crops[blockIdx.y];
    const int w = cd.w, h = cd.h;
    if (w <= 0 || h <= 0) return;
    // two launches share this kernel: the common one with a small shared-memory budget (5 CTAs/SM) takes the crops
    // that fit it, the second one (200 KB, 1 CTA/SM) takes only the large ones
    const bool fits_small = k9_smem_need(w, h) <= (size_t)K9_SMEM_SMALL;
    if (small_only != (int)fits_small) return;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = (((float)i / 255.0f) - 0.5f) / 0.5f;

    const int ks_v = pil_ksize(h), ks_h = pil_ksize(w);
    const size_t table_bytes = (size_t)(TILE_ROWS * 2 + TILE_ROWS * ks_v + OUT * 2 + OUT * ks_h) * 4;
    if (table_bytes + (size_t)(ks_v + 2) * OUT * 3 > (size_t)smem_limit) {   // uniform across the CTA
        if (threadIdx.x == 0) atomicExch(err, 1);
        return;
    }
    // smem carve: vbounds[TILE_ROWS*2] | vk[TILE_ROWS*ks_v] | hbounds[OUT*2] | hk[OUT*ks_h] | tmp rows (u8)
    int* vb = reinterpret_cast<int*>(smem);
    int* vk = vb + TILE_ROWS * 2;
    int* hb = vk + TILE_ROWS * ks_v;
    int* hk = hb + OUT * 2;
    uint8_t* tmp = reinterpret_cast<uint8_t*>(hk + OUT * ks_h);

    // horizontal coefficients once per CTA, shared by its `tiles_per_cta` row tiles (the fp64 coefficient evaluation
    // was a third of the kernel's instructions when every 16-row tile recomputed all 384 columns)
    for (int xx = threadIdx.x; xx < OUT; xx += blockDim.x)
        pil_coef(w, xx, ks_h, &hb[xx * 2], &hb[xx * 2 + 1], hk + xx * ks_h);
    for (int tile = blockIdx.x * tiles_per_cta; tile < (blockIdx.x + 1) * tiles_per_cta; ++tile) {
    __syncthreads();                  // previous tile's passes are done with vb / vk / tmp
    if (threadIdx.x < TILE_ROWS) {
        const int yy = tile * TILE_ROWS + threadIdx.x;
        pil_coef(h, yy, ks_v, &vb[threadIdx.x * 2], &vb[threadIdx.x * 2 + 1], vk + threadIdx.x * ks_v);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int r0 = vb[0], r1 = vb[0] + vb[1];
        for (int i = 1; i < TILE_ROWS; ++i) {
            r0 = min(r0, vb[i * 2]);
            r1 = max(r1, vb[i * 2] + vb[i * 2 + 1]);
        }
        s_r0 = r0; s_r1 = r1;
    }
    __syncthreads();
    const int r0 = s_r0, nrows = s_r1 - s_r0;
    if (table_bytes + (size_t)nrows * OUT * 3 > (size_t)smem_limit) {
        if (threadIdx.x == 0) atomicExch(err, 1);
        return;
    }
    // horizontal pass into tmp[nrows][OUT][3] (u8).  Pillow skips it when the width already matches; the
    // bicubic coefficients are then the identity, so running it is equivalent.
    for (int idx = threadIdx.x; idx < nrows * OUT; idx += blockDim.x) {
        const int r = idx / OUT, xx = idx - r * OUT;
        const uint8_t* src = cd.base + (long long)(r0 + r) * cd.pitch + hb[xx * 2] * 3;
        const int n = hb[xx * 2 + 1];
        const int* k = hk + xx * ks_h;
        int a0 = 1 << (PREC_BITS - 1), a1 = a0, a2 = a0;
        for (int x = 0; x < n; ++x) {
            const int kv = k[x];
            a0 += src[x * 3 + 0] * kv;
            a1 += src[x * 3 + 1] * kv;
            a2 += src[x * 3 + 2] * kv;
        }
        uint8_t* t = tmp + (size_t)idx * 3;
        t[0] = clip8(a0); t[1] = clip8(a1); t[2] = clip8(a2);
    }
    __syncthreads();
    // vertical pass + BGR->RGB + normalise + pack: 8 consecutive output pixels per thread (24 B of the u8 band per
    // tap, three 16-byte stores)
    for (int item = threadIdx.x; item < TILE_ROWS * (OUT / 8); item += blockDim.x) {
        const int ty = item / (OUT / 8), xx0 = (item - ty * (OUT / 8)) * 8;
        const int yy = tile * TILE_ROWS + ty;
        const int ymin = vb[ty * 2] - r0, n = vb[ty * 2 + 1];
        const int* k = vk + ty * ks_v;
        int acc[24];
#pragma unroll
        for (int j = 0; j < 24; ++j) acc[j] = 1 << (PREC_BITS - 1);
        for (int y = 0; y < n; ++y) {
            const uint2* t = reinterpret_cast<const uint2*>(tmp + ((size_t)(ymin + y) * OUT + xx0) * 3);
            const uint2 q0 = t[0], q1 = t[1], q2 = t[2];
            const uint32_t wds[6] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y};
            const int kv = k[y];
#pragma unroll
            for (int j = 0; j < 24; ++j) acc[j] += (int)((wds[j >> 2] >> ((j & 3) * 8)) & 0xFFu) * kv;
        }
        float pr[8], pg[8], pb[8];
#pragma unroll
        for (int px = 0; px < 8; ++px) {
            pb[px] = lut[clip8(acc[3 * px])];
            pg[px] = lut[clip8(acc[3 * px + 1])];
            pr[px] = lut[clip8(acc[3 * px + 2])];
        }
        const long long n_img = blockIdx.y;
        bf16* o;
        long long plane;
        if (layout == 0) {
            o = out + n_img * 3 * OUT * OUT + (long long)yy * OUT + xx0;
            plane = (long long)OUT * OUT;
        } else {
            const int patch = (yy >> 4) * 24 + (xx0 >> 4);
            o = out + (n_img * 576 + patch) * 768 + (yy & 15) * 16 + (xx0 & 15);
            plane = 256;
        }
        *reinterpret_cast<uint4*>(o) = make_uint4(pack2(pr[0], pr[1], f16), pack2(pr[2], pr[3], f16),
                                                  pack2(pr[4], pr[5], f16), pack2(pr[6], pr[7], f16));
        *reinterpret_cast<uint4*>(o + plane) = make_uint4(pack2(pg[0], pg[1], f16), pack2(pg[2], pg[3], f16),
                                                          pack2(pg[4], pg[5], f16), pack2(pg[6], pg[7], f16));
        *reinterpret_cast<uint4*>(o + 2 * plane) = make_uint4(pack2(pb[0], pb[1], f16), pack2(pb[2], pb[3], f16),
                                                              pack2(pb[4], pb[5], f16), pack2(pb[6], pb[7], f16));
    }
}   // tile loop
}

__host__ __device__ inline size_t k9_smem_need(int w, int h) {
    const double sv = (double)((float)h) / OUT, sh = (double)((float)w) / OUT;
    const double fv = sv < 1.0 ? 1.0 : sv, fh = sh < 1.0 ? 1.0 : sh;
    const int ks_v = (int)ceil(2.0 * fv) * 2 + 1, ks_h = (int)ceil(2.0 * fh) * 2 + 1;
    const size_t tables = (size_t)(TILE_ROWS * 2 + TILE_ROWS * ks_v + OUT * 2 + OUT * ks_h) * 4;
    const size_t band = (size_t)(TILE_ROWS * sv + 2.0 * (2.0 * fv) + 4.0);    // source rows under 16 output rows + support
    return tables + (band > (size_t)(ks_v + 2) ? band : (size_t)(ks_v + 2)) * OUT * 3 + 64;
}

// rect (x, y, w, h) on a page -> crop descriptor of page[y:y+h+1, x:x+w+1] (numpy-style clipping), the
// `crop_poly_low` snippet of marie/boxes/craft_box_processor.py:42-73,524.
__global__ void crop_desc_kernel(const uint8_t* __restrict__ pages, long long page_stride, int page_h, int page_w,
                                 const int* __restrict__ rects, const int* __restrict__ page_idx, int n,
                                 CropDesc* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = rects[i * 4], y = rects[i * 4 + 1], w = rects[i * 4 + 2], h = rects[i * 4 + 3];
    CropDesc d;
    const int x1 = min(x + w + 1, page_w), y1 = min(y + h + 1, page_h);
    d.w = max(x1 - x, 0);
    d.h = max(y1 - y, 0);
    d.pitch = page_w * 3;
    d.base = pages + (long long)page_idx[i] * page_stride + ((long long)y * page_w + x) * 3;
    out[i] = d;
}

__global__ void crop_desc_packed_kernel(const uint8_t* __restrict__ buf, const long long* __restrict__ offsets,
                                        const int* __restrict__ hw, int n, CropDesc* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    CropDesc d;
    d.h = hw[2 * i]; d.w = hw[2 * i + 1];
    d.pitch = d.w * 3;
    d.base = buf + offsets[i];
    out[i] = d;
}

int launch_crop_resize(mb_ctx* ctx, const CropDesc* descs, int n, bf16* out, int layout, int* err, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        MB_CUDA(ctx, cudaFuncSetAttribute(crop_resize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          K9_SMEM_BYTES));
        attr_set = true;
    }
    MB_CUDA(ctx, cudaMemsetAsync(err, 0, sizeof(int), stream));
    constexpr int TPC = 4;            // row tiles per CTA in the common launch
    crop_resize_kernel<<<dim3(OUT / TILE_ROWS / TPC, n), K9_THREADS, K9_SMEM_SMALL, stream>>>(descs, out, layout, err, ctx->f16,
                                                                                              K9_SMEM_SMALL, 1, TPC);
    MB_LAUNCH_CHECK(ctx);
    // large crops: one CTA per crop walks all 24 tiles (an empty launch then costs n CTAs, not 24 n)
    crop_resize_kernel<<<dim3(1, n), K9_THREADS, K9_SMEM_BYTES, stream>>>(descs, out, layout, err, ctx->f16, K9_SMEM_BYTES, 0,
                                                                          OUT / TILE_ROWS);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace

// Device-to-device form used by the page pipeline (internal linkage across .cu files).
int mb_crops_from_rects(mb_ctx* ctx, const uint8_t* pages, long long page_stride, int page_h, int page_w,
                        const int* rects, const int* page_idx, int n, bf16* out, int layout, void* desc_scratch,
                        int* err_flag, cudaStream_t stream) {
    if (n == 0) return 0;
    CropDesc* descs = (CropDesc*)desc_scratch;
    crop_desc_kernel<<<mb_cdiv(n, 128), 128, 0, stream>>>(pages, page_stride, page_h, page_w, rects, page_idx, n, descs);
    MB_LAUNCH_CHECK(ctx);
    return launch_crop_resize(ctx, descs, n, out, layout, err_flag, stream);
}

extern "C" int mb_page_preprocess(mb_ctx* ctx, const uint8_t* pages_dev, int n_pages, int page_h, int page_w,
                                  int target_h, int target_w, int out_h, int out_w, void* out_dev, void* stream_) {
    if (!ctx) return MB_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    MB_REQUIRE(ctx, n_pages > 0 && page_h > 0 && page_w > 0, "page_preprocess: empty input");
    MB_REQUIRE(ctx, target_h <= out_h && target_w <= out_w && target_h > 0 && target_w > 0,
               "page_preprocess: target larger than canvas");
    LinCoef* tabs = (LinCoef*)mb_scratch(ctx, sizeof(LinCoef) * (size_t)(target_h + target_w) + 256);
    if (!tabs) return MB_ERR_OOM;
    LinCoef* xtab = tabs;
    LinCoef* ytab = tabs + target_w;
    lin_coef_kernel<<<mb_cdiv(target_w, 256), 256, 0, stream>>>(xtab, page_w, target_w);
    MB_LAUNCH_CHECK(ctx);
    lin_coef_kernel<<<mb_cdiv(target_h, 256), 256, 0, stream>>>(ytab, page_h, target_h);
    MB_LAUNCH_CHECK(ctx);
    dim3 grid(mb_cdiv(out_w, 256), out_h, n_pages);
    page_preprocess_kernel<<<grid, 256, 0, stream>>>(pages_dev, (long long)page_h * page_w * 3, page_h, page_w, xtab,
                                                     ytab, target_h, target_w, out_h, out_w, (bf16*)out_dev, ctx->f16);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int mb_pack_crops(mb_ctx* ctx, const uint8_t* pages_dev, int page_h, int page_w,
                             const int32_t* rects_dev, const int32_t* page_idx_dev, int n_crops, void* out_dev,
                             int layout, void* stream_) {
    if (!ctx) return MB_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_crops == 0) return 0;
    MB_REQUIRE(ctx, n_crops > 0 && (layout == 0 || layout == 1), "pack_crops: bad arguments");
    unsigned char* s = (unsigned char*)mb_scratch(ctx, sizeof(CropDesc) * (size_t)n_crops + 512);
    if (!s) return MB_ERR_OOM;
    int* err = (int*)s;
    int rc = mb_crops_from_rects(ctx, pages_dev, (long long)page_h * page_w * 3, page_h, page_w, rects_dev,
                                 page_idx_dev, n_crops, (bf16*)out_dev, layout, s + 256, err, stream);
    if (rc) return rc;
    int host_err = 0;
    MB_CUDA(ctx, cudaMemcpyAsync(&host_err, err, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MB_CUDA(ctx, cudaStreamSynchronize(stream));
    if (host_err) return mb_set_err(ctx, MB_ERR_ARG, "pack_crops: a crop is too large for the resampler workspace");
    return 0;
}

extern "C" int mb_pack_fragments(mb_ctx* ctx, const uint8_t* buf_dev, const long long* offsets_dev,
                                 const int32_t* hw_dev, int n_crops, void* out_dev, int layout, void* stream_) {
    if (!ctx) return MB_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_crops == 0) return 0;
    MB_REQUIRE(ctx, n_crops > 0 && (layout == 0 || layout == 1), "pack_fragments: bad arguments");
    unsigned char* s = (unsigned char*)mb_scratch(ctx, sizeof(CropDesc) * (size_t)n_crops + 512);
    if (!s) return MB_ERR_OOM;
    int* err = (int*)s;
    CropDesc* descs = (CropDesc*)(s + 256);
    crop_desc_packed_kernel<<<mb_cdiv(n_crops, 128), 128, 0, stream>>>(buf_dev, offsets_dev, hw_dev, n_crops, descs);
    MB_LAUNCH_CHECK(ctx);
    int rc = launch_crop_resize(ctx, descs, n_crops, (bf16*)out_dev, layout, err, stream);
    if (rc) return rc;
    int host_err = 0;
    MB_CUDA(ctx, cudaMemcpyAsync(&host_err, err, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MB_CUDA(ctx, cudaStreamSynchronize(stream));
    if (host_err) return mb_set_err(ctx, MB_ERR_ARG, "pack_fragments: a fragment is too large for the resampler workspace");
    return 0;
}
