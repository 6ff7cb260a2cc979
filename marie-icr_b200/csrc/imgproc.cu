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
//     Kernels: crop_resize_up_kernel (crops of at most 384 rows: one CTA per crop, dp2a vertical pass) and the tiled
//     crop_resize_kernel for everything else; crop_classify_kernel builds their work lists.
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

// One CTA = K1_COLS output columns x K1_ROWS output rows.  The source rows under the tile are staged in shared memory
// with 16-byte loads (rows of a packed u8 page are only byte aligned, so each row is fetched as the aligned 16-byte
// chunks that cover it and indexed with its own sub-chunk offset); a thread then owns one output column and walks
// down the tile, keeping the horizontal sums of the current source-row pair in registers (a source row shared by two
// output rows is filtered once).
constexpr int K1_COLS = 256;
constexpr int K1_ROWS = 16;
constexpr int K1_MAX_SRC_ROWS = 40;      // source rows under 16 output rows (down-scale factor <= 2.4)
constexpr int K1_MAX_ROW_BYTES = 2048;   // staged bytes per source row (down-scale factor <= 2.6 at 256 columns)

__global__ void __launch_bounds__(K1_COLS)
page_preprocess_kernel(const uint8_t* __restrict__ pages, long long page_stride, long long total_bytes, int sh, int sw,
                       const LinCoef* __restrict__ xtab, const LinCoef* __restrict__ ytab, int th, int tw, int oh,
                       int ow, bf16* __restrict__ out, int f16, int K1_ROW_BYTES) {
    extern __shared__ __align__(16) unsigned char k1_smem[];
    __shared__ unsigned short lut[256];
    __shared__ int row_off[K1_MAX_SRC_ROWS];
    __shared__ LinCoef ytile[K1_ROWS];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const float v = ((float)i - 127.5f) / 127.5f;
        lut[i] = f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
    }
    const int ox0 = blockIdx.x * K1_COLS, oy0 = blockIdx.y * K1_ROWS;
    const int page = blockIdx.z;
    const int ox = ox0 + threadIdx.x;
    const int oy_end = min(oy0 + K1_ROWS, oh);
    const unsigned short neg1 = f16 ? (unsigned short)0xBC00 : (unsigned short)0xBF80;
    uint2* orow = reinterpret_cast<uint2*>(out) + ((long long)page * oh + oy0) * ow + ox;
    const bool tile_has_src = (oy0 < th) && (ox0 < tw);
    const int n_src_rows_out = tile_has_src ? min(oy_end, th) - oy0 : 0;   // output rows of the tile that have source
    int ys0 = 0, nsrc = 0, xs0 = 0;
    if (tile_has_src) {
        if (threadIdx.x < n_src_rows_out) ytile[threadIdx.x] = ytab[oy0 + threadIdx.x];
        const int oy_last = oy0 + n_src_rows_out - 1, ox_last = min(ox0 + K1_COLS, tw) - 1;
        ys0 = ytab[oy0].ofs;
        nsrc = min(ytab[oy_last].ofs + 1, sh - 1) - ys0 + 1;
        xs0 = xtab[ox0].ofs;
        const int xs1 = min(xtab[ox_last].ofs + 1, sw - 1);
        const int nbytes = (xs1 - xs0 + 1) * 3;
        const uint8_t* pbase = pages + (long long)page * page_stride;
        // nsrc / nbytes exceed the staging area only for down-scale factors the launcher rejects.  Warp w stages rows
        // w, w+8, ...; lanes take the row's aligned 16-byte chunks, so a thread has several independent loads in flight
        // (a row-by-row loop over the whole CTA serialised ~20 global-memory round trips per tile).
        const int cpr = (nbytes + 30) >> 4;                 // chunks per row, enough for any sub-chunk offset
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int r = warp; r < nsrc; r += K1_COLS / 32) {
            const uint8_t* g = pbase + ((long long)(ys0 + r) * sw + xs0) * 3;
            const int sub = (int)((uintptr_t)g & 15);
            if (lane == 0) row_off[r] = sub;
            unsigned char* dst = k1_smem + r * K1_ROW_BYTES;
#pragma unroll 3
            for (int c = lane; c < cpr; c += 32) {
                const uint8_t* src = g - sub + c * 16;
                uint4 v;
                if (src >= pages && (src + 16) <= pages + total_bytes) {
                    v = *reinterpret_cast<const uint4*>(src);
                } else {                                   // first / last chunk of the whole buffer
                    unsigned char b[16];
#pragma unroll 1
                    for (int k = 0; k < 16; ++k) b[k] = (src + k >= pages && src + k < pages + total_bytes) ? src[k] : 0;
                    v = *reinterpret_cast<uint4*>(b);
                }
                *reinterpret_cast<uint4*>(dst + c * 16) = v;
            }
        }
    }
    __syncthreads();
    if (ox >= ow) return;
    const uint2 pad = make_uint2((unsigned)neg1 | ((unsigned)neg1 << 16), (unsigned)neg1);   // (-1, -1, -1, 0)
    int k = 0;                                              // next output row of the tile
    if (tile_has_src && ox < tw) {
        const LinCoef cx = xtab[ox];
        const int bx0 = (cx.ofs - xs0) * 3, bx1 = (min(cx.ofs + 1, sw - 1) - xs0) * 3;
        const int a0 = cx.a0, a1 = cx.a1;
        // walk the staged source rows once: row r is filtered horizontally (one code path), then every output row whose
        // LOWER source row is r is emitted from (previous, current) — output rows are ordered by source row
        int prev0 = 0, prev1 = 0, prev2 = 0;
        for (int r = 0; r < nsrc; ++r) {
            const unsigned char* row = k1_smem + r * K1_ROW_BYTES + row_off[r];
            const int c0 = row[bx0] * a0 + row[bx1] * a1;
            const int c1 = row[bx0 + 1] * a0 + row[bx1 + 1] * a1;
            const int c2 = row[bx0 + 2] * a0 + row[bx1 + 2] * a1;
            while (k < n_src_rows_out) {                    // CTA-uniform
                const LinCoef cy = ytile[k];
                if (min(cy.ofs + 1, sh - 1) - ys0 != r) break;
                const bool same = cy.ofs - ys0 == r;        // clamped bottom edge: both source rows are r
                const int t0 = same ? c0 : prev0, t1 = same ? c1 : prev1, t2 = same ? c2 : prev2;
                // every term is non-negative: only the upper side of cv2's saturate_cast can trigger
                const int v0 = (((cy.a0 * (t0 >> 4)) >> 16) + ((cy.a1 * (c0 >> 4)) >> 16) + 2) >> 2;
                const int v1 = (((cy.a0 * (t1 >> 4)) >> 16) + ((cy.a1 * (c1 >> 4)) >> 16) + 2) >> 2;
                const int v2 = (((cy.a0 * (t2 >> 4)) >> 16) + ((cy.a1 * (c2 >> 4)) >> 16) + 2) >> 2;
                uint2 o;
                o.x = (unsigned)lut[min(v0, 255)] | ((unsigned)lut[min(v1, 255)] << 16);
                o.y = (unsigned)lut[min(v2, 255)];
                *orow = o;
                orow += ow;
                ++k;
            }
            prev0 = c0; prev1 = c1; prev2 = c2;
        }
    }
    for (; oy0 + k < oy_end; ++k) {                         // canvas rows / columns beyond the resized page
        *orow = pad;
        orow += ow;
    }
}

// Fallback for down-scale factors beyond the staging area of page_preprocess_kernel (very long or very wide pages):
// one thread per output pixel straight from global memory, same arithmetic.
__global__ void page_preprocess_generic_kernel(const uint8_t* __restrict__ pages, long long page_stride, int sh, int sw,
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

// pil_coef for in_size <= OUT (filterscale = 1, support = 2, at most 5 taps): every weight evaluated once and kept in
// registers; same operations in the same order as the generic routine.
__device__ __forceinline__ void pil_coef_up(int in_size, int xx, int* xmin_out, int* n_out, int kk[5]) {
    const double scale = (double)((float)in_size - 0.0f) / OUT;
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = (int)(center - 2.0 + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + 2.0 + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double wv[5];
    double ww = 0.0;
#pragma unroll
    for (int x = 0; x < 5; ++x) {
        wv[x] = bicubic((x + xmin - center + 0.5) * 1.0);
        if (x < xmax) ww += wv[x];
    }
#pragma unroll
    for (int x = 0; x < 5; ++x) {
        double w = wv[x];
        if (ww != 0.0) w /= ww;
        const int k = (w < 0) ? (int)(-0.5 + w * (double)(1 << PREC_BITS)) : (int)(0.5 + w * (double)(1 << PREC_BITS));
        kk[x] = x < xmax ? k : 0;
    }
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
crop_resize_kernel(const CropDesc* __restrict__ crops, const int* __restrict__ list, const int* __restrict__ count,
                   bf16* __restrict__ out, int layout, int* __restrict__ err, int f16, int smem_limit,
                   int tiles_per_cta) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ float lut[256];
    __shared__ int s_r0, s_r1;
    // two launches share this kernel, each walking its own work list (crop_classify_kernel): the one with a small
    // shared-memory budget (5 CTAs/SM) and the one for large crops (200 KB, 1 CTA/SM)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = (((float)i / 255.0f) - 0.5f) / 0.5f;
    const int n_list = *count;
    for (int li = blockIdx.y; li < n_list; li += gridDim.y) {
    const int crop_id = list[li];
    const CropDesc cd = crops[crop_id];
    const int w = cd.w, h = cd.h;
    __syncthreads();                  // previous crop is done with the tables

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
        const long long n_img = crop_id;
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
    }   // crop loop
}

__host__ __device__ inline size_t k9_smem_need(int w, int h) {
    const double sv = (double)((float)h) / OUT, sh = (double)((float)w) / OUT;
    const double fv = sv < 1.0 ? 1.0 : sv, fh = sh < 1.0 ? 1.0 : sh;
    const int ks_v = (int)ceil(2.0 * fv) * 2 + 1, ks_h = (int)ceil(2.0 * fh) * 2 + 1;
    const size_t tables = (size_t)(TILE_ROWS * 2 + TILE_ROWS * ks_v + OUT * 2 + OUT * ks_h) * 4;
    const size_t band = (size_t)(TILE_ROWS * sv + 2.0 * (2.0 * fv) + 4.0);    // source rows under 16 output rows + support
    return tables + (band > (size_t)(ks_v + 2) ? band : (size_t)(ks_v + 2)) * OUT * 3 + 64;
}

// ------------------------------------------------------------------------------------------------ K9, common case
// Word crops are small (h <= ~60 rows) and are scaled UP vertically, so Pillow's vertical filter has at most five
// taps.  One CTA owns one crop: coefficient tables once, the horizontal pass once over all source rows (the tiled
// kernel above recomputes both per 64-row band), then the vertical pass as packed integer dot products:
//   * the u8 intermediate is stored row-PAIR interleaved, tmp[pair][x*3+c][2], so a 32-bit word holds two vertically
//     adjacent samples of two neighbouring values;
//   * every 22-bit coefficient k is split exactly as k = kh * 2^11 + kl (kl = k & 2047, kh = k >> 11, both fit s16),
//     and dp2a.{lo,hi}.s32.u32 accumulates kh- and kl-sums for two taps at once: sum k*v = (sum kh*v << 11) + sum kl*v.
// Results are bit-identical to the tap-by-tap form.  Crops that do not fit (tall / very wide) take the tiled kernel.
constexpr int UP_THREADS = 384;
constexpr int UP_SMEM = 112 * 1024;      // two CTAs per SM
constexpr int UP_MIN_PAIRS = 12;         // smallest row-pair band worth running (wide crops have large tables)
constexpr int UP_VROW_INTS = 8;          // p0, pairs, kh[3], kl[3]
constexpr int UP_PAIR_BYTES = OUT * 3 * 2;

__host__ __device__ inline int k9_ksize(int in_size) {
    const double scale = (double)((float)in_size) / OUT;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(2.0 * filterscale) * 2 + 1;
}

__host__ __device__ inline size_t k9_up_tables(int w) {
    return (size_t)OUT * UP_VROW_INTS * 4 + (size_t)OUT * 2 * 4 + (size_t)OUT * k9_ksize(w) * 4 + 512 * 2;
}

__host__ __device__ inline bool k9_fast_ok(int w, int h) {
    if (h > OUT || h <= 0 || w <= 0) return false;
    return k9_up_tables(w) + (size_t)UP_MIN_PAIRS * UP_PAIR_BYTES <= (size_t)UP_SMEM;
}

__device__ __forceinline__ int dp2a_lo_su(int a, unsigned b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(int a, unsigned b, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// grid-stride over the list of eligible crops (list[0..*count))
__global__ void __launch_bounds__(UP_THREADS, 2)
crop_resize_up_kernel(const CropDesc* __restrict__ crops, const int* __restrict__ list, const int* __restrict__ count,
                      bf16* __restrict__ out, int layout, int f16) {
    extern __shared__ __align__(16) unsigned char smem[];
    int* vrow = reinterpret_cast<int*>(smem);                       // [OUT][8]
    int* hb = vrow + OUT * UP_VROW_INTS;                            // [OUT][2]
    unsigned short* lut = reinterpret_cast<unsigned short*>(hb + OUT * 2);   // [512]: index (v >> 22) + 128, clamped
    int* hk = reinterpret_cast<int*>(lut + 512);                    // [OUT][ks_h]
    const int n_list = *count;
    for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
    const int crop = list[li];
    const CropDesc cd = crops[crop];
    const int w = cd.w, h = cd.h;
    const int ks_h = k9_ksize(w);
    unsigned char* tmp = reinterpret_cast<unsigned char*>(hk + OUT * ks_h);   // [pairs of a band][OUT*3][2]
    __syncthreads();                                                // previous crop is done with the tables
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        int u = i - 128;
        u = u < 0 ? 0 : (u > 255 ? 255 : u);
        const float v = (((float)u / 255.0f) - 0.5f) / 0.5f;
        lut[i] = f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
    }
    for (int t = threadIdx.x; t < 2 * OUT; t += blockDim.x) {
        if (t < OUT) {
            if (w <= OUT) {
                int kk[5], xmin, n;
                pil_coef_up(w, t, &xmin, &n, kk);
                hb[t * 2] = xmin; hb[t * 2 + 1] = n;
#pragma unroll
                for (int x = 0; x < 5; ++x) hk[t * 5 + x] = kk[x];
            } else {
                pil_coef(w, t, ks_h, &hb[t * 2], &hb[t * 2 + 1], hk + t * ks_h);
            }
        } else {
            const int yy = t - OUT;
            int kk[5], ymin, n;
            pil_coef_up(h, yy, &ymin, &n, kk);
            // taps ymin .. ymin+n-1 (n <= 5) -> row pairs p0 .. p0+pairs-1, slot s = row - 2*p0
            const int odd = ymin & 1;
            int k6[6];
            k6[0] = odd ? 0 : kk[0];
#pragma unroll
            for (int x = 1; x < 5; ++x) k6[x] = odd ? kk[x - 1] : kk[x];
            k6[5] = odd ? kk[4] : 0;
            int* v = vrow + yy * UP_VROW_INTS;
            v[0] = ymin >> 1;
            v[1] = (odd + n + 1) >> 1;
            for (int q = 0; q < 3; ++q) {
                const int a = k6[2 * q], b = k6[2 * q + 1];
                v[2 + q] = (int)(((unsigned)(a >> 11) & 0xFFFFu) | ((unsigned)(b >> 11) << 16));
                v[5 + q] = (int)((unsigned)(a & 2047) | ((unsigned)(b & 2047) << 16));
            }
        }
    }
    __syncthreads();
    // Output rows are processed in bands whose source row pairs fit the staging area (one band for h <= ~78).
    const int npair = (h + 1) >> 1;
    const int max_pairs = (int)((UP_SMEM - k9_up_tables(w)) / UP_PAIR_BYTES);
    const int xg = threadIdx.x % (OUT / 8), slot = threadIdx.x / (OUT / 8);
    const int xx0 = xg * 8;
    constexpr int SLOTS = UP_THREADS / (OUT / 8);          // output rows in flight
    for (int yb0 = 0; yb0 < OUT;) {
        const int pf = vrow[yb0 * UP_VROW_INTS];           // first source pair of the band
        int yb1 = yb0 + SLOTS;
        while (yb1 < OUT) {                                 // uniform across the CTA: same table, same arithmetic
            const int* vl = vrow + (yb1 + SLOTS - 1) * UP_VROW_INTS;
            if (vl[0] + vl[1] - pf > max_pairs) break;
            yb1 += SLOTS;
        }
        const int* vlast = vrow + (yb1 - 1) * UP_VROW_INTS;
        const int pl = min(vlast[0] + vlast[1], npair);    // pairs past the crop carry zero coefficients only
    // horizontal pass, two source rows per item -> tmp[pair - pf][xx*3+c][row&1]
    for (int idx = threadIdx.x; idx < (pl - pf) * OUT; idx += blockDim.x) {
        const int pr = pf + idx / OUT, xx = idx % OUT;
        const int r0 = 2 * pr, r1 = (2 * pr + 1 < h) ? 2 * pr + 1 : r0;     // the odd tail row is never referenced
        const uint8_t* s0 = cd.base + (long long)r0 * cd.pitch + hb[xx * 2] * 3;
        const uint8_t* s1 = cd.base + (long long)r1 * cd.pitch + hb[xx * 2] * 3;
        const int n = hb[xx * 2 + 1];
        const int* k = hk + xx * ks_h;
        int a0 = 1 << (PREC_BITS - 1), a1 = a0, a2 = a0, b0 = a0, b1 = a0, b2 = a0;
        if (ks_h == 5) {
            // up-scale: five coefficient slots (zero past n), taps clamped to the last tap so every load stays in the row
#pragma unroll
            for (int x = 0; x < 5; ++x) {
                const int kv = k[x];
                const int o = (x < n ? x : n - 1) * 3;
                a0 += s0[o + 0] * kv; a1 += s0[o + 1] * kv; a2 += s0[o + 2] * kv;
                b0 += s1[o + 0] * kv; b1 += s1[o + 1] * kv; b2 += s1[o + 2] * kv;
            }
        } else {
            for (int x = 0; x < n; ++x) {
                const int kv = k[x];
                a0 += s0[x * 3 + 0] * kv; a1 += s0[x * 3 + 1] * kv; a2 += s0[x * 3 + 2] * kv;
                b0 += s1[x * 3 + 0] * kv; b1 += s1[x * 3 + 1] * kv; b2 += s1[x * 3 + 2] * kv;
            }
        }
        unsigned short* t = reinterpret_cast<unsigned short*>(tmp + (size_t)(pr - pf) * UP_PAIR_BYTES + xx * 6);
        t[0] = (unsigned short)(clip8(a0) | (clip8(b0) << 8));
        t[1] = (unsigned short)(clip8(a1) | (clip8(b1) << 8));
        t[2] = (unsigned short)(clip8(a2) | (clip8(b2) << 8));
    }
    __syncthreads();
    // vertical pass: thread = 8 consecutive output pixels (24 values = 48 bytes of a pair row), SLOTS rows in flight
    for (int yy = yb0 + slot; yy < yb1; yy += SLOTS) {
        const int4 va = *reinterpret_cast<const int4*>(vrow + yy * UP_VROW_INTS);
        const int4 vb4 = *reinterpret_cast<const int4*>(vrow + yy * UP_VROW_INTS + 4);
        const int p0 = va.x - pf, pairs = va.y;
        const int kh[3] = {va.z, va.w, vb4.x};
        const int kl[3] = {vb4.y, vb4.z, vb4.w};
        int ah[24], al[24];
        const unsigned char* tb = tmp + (size_t)p0 * UP_PAIR_BYTES + xg * 48;
        {   // first pair: always present
            const uint4* t4 = reinterpret_cast<const uint4*>(tb);
            const uint4 w0 = t4[0], w1 = t4[1], w2 = t4[2];
            const unsigned wd[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                ah[2 * j] = dp2a_lo_su(kh[0], wd[j], 0);
                ah[2 * j + 1] = dp2a_hi_su(kh[0], wd[j], 0);
                al[2 * j] = dp2a_lo_su(kl[0], wd[j], 1 << (PREC_BITS - 1));
                al[2 * j + 1] = dp2a_hi_su(kl[0], wd[j], 1 << (PREC_BITS - 1));
            }
        }
#pragma unroll
        for (int q = 1; q < 3; ++q) {
            if (q < pairs) {
                const uint4* t4 = reinterpret_cast<const uint4*>(tb + (size_t)q * UP_PAIR_BYTES);
                const uint4 w0 = t4[0], w1 = t4[1], w2 = t4[2];
                const unsigned wd[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    ah[2 * j] = dp2a_lo_su(kh[q], wd[j], ah[2 * j]);
                    ah[2 * j + 1] = dp2a_hi_su(kh[q], wd[j], ah[2 * j + 1]);
                    al[2 * j] = dp2a_lo_su(kl[q], wd[j], al[2 * j]);
                    al[2 * j + 1] = dp2a_hi_su(kl[q], wd[j], al[2 * j + 1]);
                }
            }
        }
        unsigned short hv[24];
#pragma unroll
        for (int j = 0; j < 24; ++j) {
            const int tot = ah[j] * 2048 + al[j];
            hv[j] = lut[(tot >> PREC_BITS) + 128];     // |tot >> 22| < 128 beyond [0, 255]: the table clamps
        }
        bf16* o;
        long long plane;
        if (layout == 0) {
            o = out + (long long)crop * 3 * OUT * OUT + (long long)yy * OUT + xx0;
            plane = (long long)OUT * OUT;
        } else {
            const int patch = (yy >> 4) * 24 + (xx0 >> 4);
            o = out + ((long long)crop * 576 + patch) * 768 + (yy & 15) * 16 + (xx0 & 15);
            plane = 256;
        }
        // source order is B, G, R (value j = px*3 + c): plane 0 = R (c = 2), plane 1 = G, plane 2 = B
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int sc = 2 - c;
            uint4 q;
            q.x = (unsigned)hv[0 * 3 + sc] | ((unsigned)hv[1 * 3 + sc] << 16);
            q.y = (unsigned)hv[2 * 3 + sc] | ((unsigned)hv[3 * 3 + sc] << 16);
            q.z = (unsigned)hv[4 * 3 + sc] | ((unsigned)hv[5 * 3 + sc] << 16);
            q.w = (unsigned)hv[6 * 3 + sc] | ((unsigned)hv[7 * 3 + sc] << 16);
            *reinterpret_cast<uint4*>(o + c * plane) = q;
        }
    }
        yb0 = yb1;
        if (yb0 < OUT) __syncthreads();                     // the next band overwrites tmp
    }   // band loop
    }   // crop loop
}

// ------------------------------------------------------------------------------------------------ K9, common case (v2)
// The same filter, re-organised around the instruction count of the vertical pass (ncu, r01: 17.8 thread instructions per
// output value at 56 % issue utilisation; two CTAs / SM, three CTA-wide phases per crop):
//   * an up-scaled axis has at most FOUR taps (exhaustive over in_size = 1..384, tests/test_device_algorithms_cpu.py), so
//     the u8 intermediate is stored as ROW QUADS, tmp[q][x*3+c] = rows 4q..4q+3 of one value in one 32-bit word, and a tap
//     window at any row offset is one PRMT of two quad words;
//   * every coefficient k (|k| < 2^23) is split exactly as k = kh*2^16 + kl (kl = k & 0xFFFF unsigned, kh = k >> 16 a signed
//     byte); a value is two dp2a over the 16-bit halves (taps 0,1 / 2,3) chained through the accumulator, one dp4a over
//     the four high bytes and one shift-add: sum k*v = (sum kh*v << 16) + sum kl*v;
//   * consecutive output rows with the same first tap (a "run", 384 / h rows) share the window: a warp third owns a run,
//     builds the windows of its 4 px x 3 channels once and walks the run with one coefficient fetch per row;
//   * the horizontal pass keeps the thread's four coefficients in registers (thread = output column) and packs four rows
//     with two saturating cvt.pack;
//   * 72 KB and <= 56 registers: three CTAs per SM, so one crop's coefficient / horizontal phases run under another's
//     vertical pass.  Tall crops are processed in bands of quads (one quad of overlap).
// Bit-identical to the tap-by-tap form (and to Pillow): tests/test_imgproc_gpu.py.
constexpr int UP2_THREADS = OUT;           // thread t = output column t (horizontal pass) = output row t (tables)
constexpr int UP2_SMEM = 75 * 1024;        // three CTAs per SM
constexpr int UP2_QWORDS = OUT * 3;        // words per row quad
constexpr int UP2_FIXED = OUT * 16 + (OUT + 8) * 4 + 64;   // vtab | run_y0 | warp counts

__host__ __device__ inline size_t k9_up2_tables(int w) {
    return (size_t)UP2_FIXED + (w > OUT ? (size_t)OUT * 8 + (size_t)OUT * k9_ksize(w) * 4 : 0);
}
__host__ __device__ inline bool k9_up2_ok(int w, int h) {
    if (h > OUT || h <= 0 || w < 4) return false;
    return k9_up2_tables(w) + (size_t)6 * UP2_QWORDS * 4 <= (size_t)UP2_SMEM;      // at least five quads + the spare
}
__device__ __forceinline__ int dp2a_lo_uu(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_uu(unsigned a, unsigned b, int c) {
    int d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_su(unsigned a, unsigned b, int c) {
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// bytes (lo .. hi) = sat_u8(a0), sat_u8(a1), sat_u8(a2), sat_u8(a3)
__device__ __forceinline__ unsigned pack_sat4(int a0, int a1, int a2, int a3) {
    unsigned hi, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(a3), "r"(a2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a1), "r"(a0), "r"(hi));
    return d;
}

template <bool HWORD, int LAYOUT>
__global__ void __launch_bounds__(UP2_THREADS, 3)
crop_resize_up2_kernel(const CropDesc* __restrict__ crops, const int* __restrict__ list, const int* __restrict__ count,
                       bf16* __restrict__ out, int f16) {
    constexpr bool hword = HWORD;
    constexpr int layout = LAYOUT;
    extern __shared__ __align__(16) unsigned char smem[];
    int4* vtab = reinterpret_cast<int4*>(smem);                                  // [OUT] (kl taps 0,1 | kl taps 2,3 | kh bytes | ymin)
    int* run_y0 = reinterpret_cast<int*>(vtab + OUT);                            // [n_runs + 1] first output row of a run
    int* wcnt = run_y0 + OUT + 8;                                                // [16]
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int n_list = *count;
    for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
        const int crop = list[li];
        const CropDesc cd = crops[crop];
        const int w = cd.w, h = cd.h;
        const bool wide = w > OUT;
        const int ks_h = k9_ksize(w);
        int* hb = wcnt + 16;                                                     // wide crops only: [OUT][2], [OUT][ks_h]
        int* hk = hb + OUT * 2;
        unsigned* tmpw = reinterpret_cast<unsigned*>(smem + k9_up2_tables(w));
        const int QB = (int)((UP2_SMEM - k9_up2_tables(w)) / (UP2_QWORDS * 4)) - 1;   // quads of a band (+ one spare behind)
        __syncthreads();                                    // previous crop is done with the tables
        // ---- tables: vertical digits for output row t, horizontal coefficients for output column t
        int hx = 0, hk0 = 0, hk1 = 0, hk2 = 0, hk3 = 0;
        {
            int kk[5], ymin, n;
            pil_coef_up(h, t, &ymin, &n, kk);
            const unsigned l01 = (unsigned)(kk[0] & 0xFFFF) | ((unsigned)(kk[1] & 0xFFFF) << 16);
            const unsigned l23 = (unsigned)(kk[2] & 0xFFFF) | ((unsigned)(kk[3] & 0xFFFF) << 16);
            unsigned k2 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) k2 |= (unsigned)((kk[i] >> 16) & 255) << (8 * i);
            vtab[t] = make_int4((int)l01, (int)l23, (int)k2, ymin);
            if (!wide) {
                int xmin;
                pil_coef_up(w, t, &xmin, &n, kk);
                // the four-column window [hx, hx+4) stays inside the row (w >= 4): taps shifted, empty slots zero
                hx = min(xmin, w - 4);
                const int sh = xmin - hx;
                hk0 = sh == 0 ? kk[0] : 0;
                hk1 = sh == 0 ? kk[1] : (sh == 1 ? kk[0] : 0);
                hk2 = sh == 0 ? kk[2] : (sh == 1 ? kk[1] : (sh == 2 ? kk[0] : 0));
                hk3 = sh == 0 ? kk[3] : (sh == 1 ? kk[2] : (sh == 2 ? kk[1] : kk[0]));
            } else {
                pil_coef(w, t, ks_h, &hb[t * 2], &hb[t * 2 + 1], hk + t * ks_h);
            }
        }
        __syncthreads();
        // ---- runs of output rows with the same first tap (ymin is non-decreasing in y)
        {
            const bool first = t == 0 || vtab[t].w != vtab[t - 1].w;
            const unsigned bal = __ballot_sync(0xffffffffu, first);
            if (lane == 0) wcnt[warp] = __popc(bal);
            __syncthreads();
            int base = 0, total = 0;
#pragma unroll
            for (int i = 0; i < UP2_THREADS / 32; ++i) {
                const int c = wcnt[i];
                base += i < warp ? c : 0;
                total += c;
            }
            if (first) run_y0[base + __popc(bal & ((1u << lane) - 1u))] = t;
            if (t == 0) { run_y0[total] = OUT; wcnt[12] = total; }
        }
        __syncthreads();
        const int n_runs = wcnt[12];
        const int t3 = warp % 3, rsel = warp / 3;
        const int px4 = t3 * 128 + lane * 4;                 // this lane's four pixels in the vertical pass
        const uint8_t* hsrc = cd.base + (long long)hx * 3;
        for (int rb0 = 0; rb0 < n_runs;) {
            // band: quads [bq0, bq0 + QB) are computed; a run needs quads q0 and q0 + 1 of its window
            const int bq0 = vtab[run_y0[rb0]].w >> 2;
            int lo = rb0, hi = n_runs;                      // first run whose q0 > bq0 + QB - 2 (binary search, CTA-uniform)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((vtab[run_y0[mid]].w >> 2) <= bq0 + QB - 2) lo = mid + 1; else hi = mid;
            }
            const int rb1 = lo;
            const int nq = min(bq0 + QB, (h + 3) >> 2) - bq0;        // quads with source rows in this band
            // ---- horizontal pass: thread = output column, item = row quad
            if (!wide && hword) {
                // the window's twelve bytes (4 px x 3 channels from hsrc + row * pitch, any alignment) as aligned 32-bit
                // loads + funnel shifts instead of twelve LDG.U8: ncu (r02) had the byte loads at 2.25 sectors per request,
                // 38 % of the wavefronts of an L1-data-pipe-bound kernel.  Per channel the four taps are gathered into one
                // word (two PRMT) and weighted with the same exact 16 + 8 bit digit split as the vertical pass.
                const unsigned hl01 = (unsigned)(hk0 & 0xFFFF) | ((unsigned)(hk1 & 0xFFFF) << 16);
                const unsigned hl23 = (unsigned)(hk2 & 0xFFFF) | ((unsigned)(hk3 & 0xFFFF) << 16);
                const unsigned hkh = (unsigned)((hk0 >> 16) & 255) | ((unsigned)((hk1 >> 16) & 255) << 8) |
                                     ((unsigned)((hk2 >> 16) & 255) << 16) | ((unsigned)((hk3 >> 16) & 255) << 24);
                for (int q = 0; q < nq; ++q) {
                    int a[3][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int row = min((bq0 + q) * 4 + r, h - 1);
                        const uintptr_t ap = reinterpret_cast<uintptr_t>(hsrc + (long long)row * cd.pitch);
                        const unsigned* wp = reinterpret_cast<const unsigned*>(ap & ~(uintptr_t)3);
                        const unsigned sh = ((unsigned)ap & 3u) * 8u;
                        const unsigned w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
                        const unsigned w3 = sh ? __ldg(wp + 3) : 0u;     // aligned window: the twelve bytes end inside w2
                        const unsigned r0 = __funnelshift_r(w0, w1, sh), r1 = __funnelshift_r(w1, w2, sh), r2 = __funnelshift_r(w2, w3, sh);
                        // channel c: bytes c, 3 + c, 6 + c, 9 + c of (r0 | r1 | r2)
                        const unsigned v[3] = {__byte_perm(__byte_perm(r0, r1, 0x0630), r2, 0x5210),
                                               __byte_perm(__byte_perm(r0, r1, 0x0741), r2, 0x6210),
                                               __byte_perm(__byte_perm(r0, r1, 0x0052), r2, 0x7410)};
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            int dl = dp2a_lo_uu(hl01, v[c], 1 << (PREC_BITS - 1));
                            dl = dp2a_hi_uu(hl23, v[c], dl);
                            const int dh = dp4a_su(hkh, v[c], 0);
                            a[c][r] = ((dh << 16) + dl) >> PREC_BITS;
                        }
                    }
                    unsigned* d = tmpw + q * UP2_QWORDS + t * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) d[c] = pack_sat4(a[c][0], a[c][1], a[c][2], a[c][3]);
                }
            } else if (!wide) {
                for (int q = 0; q < nq; ++q) {
                    int a[3][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int row = min((bq0 + q) * 4 + r, h - 1);
                        const uint8_t* sp = hsrc + (long long)row * cd.pitch;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            a[c][r] = ((1 << (PREC_BITS - 1)) + sp[c] * hk0 + sp[3 + c] * hk1 + sp[6 + c] * hk2 + sp[9 + c] * hk3) >> PREC_BITS;
                    }
                    unsigned* d = tmpw + q * UP2_QWORDS + t * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) d[c] = pack_sat4(a[c][0], a[c][1], a[c][2], a[c][3]);
                }
            } else {
                const int xmin = hb[t * 2], n = hb[t * 2 + 1];
                const int* k = hk + t * ks_h;
                for (int q = 0; q < nq; ++q) {
                    int a[3][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int row = min((bq0 + q) * 4 + r, h - 1);
                        const uint8_t* sp = cd.base + (long long)row * cd.pitch + xmin * 3;
                        int a0 = 1 << (PREC_BITS - 1), a1 = a0, a2 = a0;
                        for (int x = 0; x < n; ++x) {
                            const int kv = k[x];
                            a0 += sp[x * 3 + 0] * kv; a1 += sp[x * 3 + 1] * kv; a2 += sp[x * 3 + 2] * kv;
                        }
                        a[0][r] = a0 >> PREC_BITS; a[1][r] = a1 >> PREC_BITS; a[2][r] = a2 >> PREC_BITS;
                    }
                    unsigned* d = tmpw + q * UP2_QWORDS + t * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) d[c] = pack_sat4(a[c][0], a[c][1], a[c][2], a[c][3]);
                }
            }
            __syncthreads();
            // ---- vertical pass: warp third = 128 px of a run
            for (int r = rb0 + rsel; r < rb1; r += 4) {
                const int y0 = run_y0[r], y1 = run_y0[r + 1];
                const int ymin = vtab[y0].w;
                const int off = ymin & 3;
                const uint4* tq = reinterpret_cast<const uint4*>(tmpw + ((ymin >> 2) - bq0) * UP2_QWORDS + px4 * 3);
                unsigned win[12];
                {
                    const uint4 l0 = tq[0], l1 = tq[1], l2 = tq[2];
                    const uint4 h0 = tq[UP2_QWORDS / 4], h1 = tq[UP2_QWORDS / 4 + 1], h2 = tq[UP2_QWORDS / 4 + 2];   // spare quad at the end
                    const unsigned sel = 0x3210u + 0x1111u * (unsigned)off;
                    const unsigned lw[12] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w, l2.x, l2.y, l2.z, l2.w};
                    const unsigned hw[12] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, h2.x, h2.y, h2.z, h2.w};
#pragma unroll
                    for (int j = 0; j < 12; ++j) win[j] = __byte_perm(lw[j], hw[j], sel);
                }
                for (int yy = y0; yy < y1; ++yy) {
                    const int4 kv = vtab[yy];
                    bf16* o;
                    long long plane;
                    if (layout == 0) {
                        o = out + (long long)crop * 3 * OUT * OUT + (long long)yy * OUT + px4;
                        plane = (long long)OUT * OUT;
                    } else {
                        const int patch = (yy >> 4) * 24 + (px4 >> 4);
                        o = out + ((long long)crop * 576 + patch) * 768 + (yy & 15) * 16 + (px4 & 15);
                        plane = 256;
                    }
                    // source order is B, G, R (value j = px*3 + c): plane 0 = R (c = 2), plane 1 = G, plane 2 = B
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float fv[4];
#pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            const unsigned wv = win[px * 3 + c];
                            int dl = dp2a_lo_uu((unsigned)kv.x, wv, 1 << (PREC_BITS - 1));
                            dl = dp2a_hi_uu((unsigned)kv.y, wv, dl);
                            const int dh = dp4a_su((unsigned)kv.z, wv, 0);
                            const int tot = (dh << 16) + dl;
                            // clip8, then ToTensor + Normalize(0.5, 0.5) as ONE fma: RN16(fma(u, fl32(2/255), -1)) equals
                            // RN16(((u / 255) - 0.5) / 0.5) for all 256 values, fp16 and bf16 (tests/test_device_algorithms_cpu.py);
                            // the table lookup this replaces was a third of the kernel's L1 wavefronts (ncu, r02)
                            fv[px] = fmaf((float)__vimin_s32_relu(tot >> PREC_BITS, 255), 2.0f / 255.0f, -1.0f);
                        }
                        *reinterpret_cast<uint2*>(o + (2 - c) * plane) = make_uint2(pack2(fv[0], fv[1], f16), pack2(fv[2], fv[3], f16));
                    }
                }
            }
            rb0 = rb1;
            if (rb0 < n_runs) __syncthreads();              // the next band overwrites tmp
        }
    }
}

// crops -> three work lists: [0] fast (crop_resize_up_kernel), [1] tiled / small smem, [2] tiled / large smem
__global__ void crop_classify_kernel(const CropDesc* __restrict__ crops, int n, int* __restrict__ counts,
                                     int* __restrict__ lists, int v1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int w = crops[i].w, h = crops[i].h;
    if (w <= 0 || h <= 0) return;
    const bool fast = v1 ? k9_fast_ok(w, h) : k9_up2_ok(w, h);
    const int cls = fast ? 0 : (k9_smem_need(w, h) <= (size_t)K9_SMEM_SMALL ? 1 : 2);
    lists[(long long)cls * n + atomicAdd(counts + cls, 1)] = i;
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

// scratch layout behind the descriptors: counts[4] | lists[3][n]
static size_t k9_list_bytes(int n) { return 64 + sizeof(int) * 3 * (size_t)n; }

int launch_crop_resize(mb_ctx* ctx, const CropDesc* descs, int n, bf16* out, int layout, int* err, int* lists_scratch,
                       cudaStream_t stream) {
    static bool attr_set = false;
    static int k9_v1 = 0, k9_hword = 1;
    if (!attr_set) {
        MB_CUDA(ctx, cudaFuncSetAttribute(crop_resize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          K9_SMEM_BYTES));
        MB_CUDA(ctx, cudaFuncSetAttribute(crop_resize_up_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UP_SMEM));
        MB_CUDA(ctx, cudaFuncSetAttribute(crop_resize_up2_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, UP2_SMEM));
        MB_CUDA(ctx, cudaFuncSetAttribute(crop_resize_up2_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UP2_SMEM));
        MB_CUDA(ctx, cudaFuncSetAttribute(crop_resize_up2_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, UP2_SMEM));
        MB_CUDA(ctx, cudaFuncSetAttribute(crop_resize_up2_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UP2_SMEM));
        const char* e = getenv("MB_K9_V1");
        k9_v1 = (e && e[0] == '1') ? 1 : 0;
        const char* hw = getenv("MB_K9_HBYTE");               // 1: the horizontal pass with byte loads (r02 form)
        k9_hword = (hw && hw[0] == '1') ? 0 : 1;
        attr_set = true;
    }
    int* counts = lists_scratch;
    int* lists = lists_scratch + 16;
    MB_CUDA(ctx, cudaMemsetAsync(err, 0, sizeof(int), stream));
    MB_CUDA(ctx, cudaMemsetAsync(counts, 0, 4 * sizeof(int), stream));
    crop_classify_kernel<<<mb_cdiv(n, 256), 256, 0, stream>>>(descs, n, counts, lists, k9_v1);
    MB_LAUNCH_CHECK(ctx);
    const int sms = ctx->num_sms;
    // common case: one CTA per crop (grid-stride), three resident per SM
    if (k9_v1)
        crop_resize_up_kernel<<<n < 2 * sms * 8 ? n : 2 * sms * 8, UP_THREADS, UP_SMEM, stream>>>(descs, lists, counts, out,
                                                                                                  layout, ctx->f16);
    else {
        const int g2 = n < 3 * sms * 4 ? n : 3 * sms * 4;
        if (k9_hword && layout == 1) crop_resize_up2_kernel<true, 1><<<g2, UP2_THREADS, UP2_SMEM, stream>>>(descs, lists, counts, out, ctx->f16);
        else if (k9_hword) crop_resize_up2_kernel<true, 0><<<g2, UP2_THREADS, UP2_SMEM, stream>>>(descs, lists, counts, out, ctx->f16);
        else if (layout == 1) crop_resize_up2_kernel<false, 1><<<g2, UP2_THREADS, UP2_SMEM, stream>>>(descs, lists, counts, out, ctx->f16);
        else crop_resize_up2_kernel<false, 0><<<g2, UP2_THREADS, UP2_SMEM, stream>>>(descs, lists, counts, out, ctx->f16);
    }
    MB_LAUNCH_CHECK(ctx);
    constexpr int TPC = 4;            // row tiles per CTA in the small-smem tiled launch
    const int gy = n < sms * 4 ? n : sms * 4;
    crop_resize_kernel<<<dim3(OUT / TILE_ROWS / TPC, gy), K9_THREADS, K9_SMEM_SMALL, stream>>>(
        descs, lists + n, counts + 1, out, layout, err, ctx->f16, K9_SMEM_SMALL, TPC);
    MB_LAUNCH_CHECK(ctx);
    // large crops: one CTA per crop walks all 24 tiles
    crop_resize_kernel<<<dim3(1, n < sms ? n : sms), K9_THREADS, K9_SMEM_BYTES, stream>>>(
        descs, lists + 2 * (size_t)n, counts + 2, out, layout, err, ctx->f16, K9_SMEM_BYTES, OUT / TILE_ROWS);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace

// Device-to-device form used by the page pipeline (internal linkage across .cu files).
int mb_crops_from_rects(mb_ctx* ctx, const uint8_t* pages, long long page_stride, int page_h, int page_w,
                        const int* rects, const int* page_idx, int n, bf16* out, int layout, void* desc_scratch,
                        int* err_flag, cudaStream_t stream) {
    if (n == 0) return 0;
    // desc_scratch: descriptors [n] | work lists (k9_list_bytes)
    CropDesc* descs = (CropDesc*)desc_scratch;
    int* lists = (int*)((unsigned char*)desc_scratch + mb_align_up(sizeof(CropDesc) * (size_t)n, 256));
    crop_desc_kernel<<<mb_cdiv(n, 128), 128, 0, stream>>>(pages, page_stride, page_h, page_w, rects, page_idx, n, descs);
    MB_LAUNCH_CHECK(ctx);
    return launch_crop_resize(ctx, descs, n, out, layout, err_flag, lists, stream);
}

extern "C" int mb_page_preprocess(mb_ctx* ctx, const uint8_t* pages_dev, int n_pages, int page_h, int page_w,
                                  int target_h, int target_w, int out_h, int out_w, void* out_dev, void* stream_) {
    MbDeviceGuard _mb_guard(ctx);
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
    // staging area limits (see page_preprocess_kernel): source rows / bytes under one 256 x 16 output tile; beyond them
    // (down-scale factors above ~2.3) the per-pixel kernel takes over
    if (!((double)page_h / target_h * K1_ROWS + 3 <= K1_MAX_SRC_ROWS &&
          ((double)page_w / target_w * K1_COLS + 3) * 3 + 32 <= K1_MAX_ROW_BYTES)) {
        dim3 ggrid(mb_cdiv(out_w, 256), out_h, n_pages);
        page_preprocess_generic_kernel<<<ggrid, 256, 0, stream>>>(pages_dev, (long long)page_h * page_w * 3, page_h, page_w, xtab,
                                                                  ytab, target_h, target_w, out_h, out_w, (bf16*)out_dev,
                                                                  ctx->f16);
        MB_LAUNCH_CHECK(ctx);
        return 0;
    }
    static bool k1_attr = false;
    if (!k1_attr) {
        MB_CUDA(ctx, cudaFuncSetAttribute(page_preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          K1_MAX_SRC_ROWS * K1_MAX_ROW_BYTES));
        k1_attr = true;
    }
    const int src_rows = (int)((double)page_h / target_h * K1_ROWS) + 3;
    dim3 grid(mb_cdiv(out_w, K1_COLS), mb_cdiv(out_h, K1_ROWS), n_pages);
    const long long pstride = (long long)page_h * page_w * 3;
    const int row_bytes = ((int)(((double)page_w / target_w * K1_COLS + 3) * 3) + 32 + 15) & ~15;
    page_preprocess_kernel<<<grid, K1_COLS, (size_t)src_rows * row_bytes, stream>>>(
        pages_dev, pstride, pstride * n_pages, page_h, page_w, xtab, ytab, target_h, target_w, out_h, out_w,
        (bf16*)out_dev, ctx->f16, row_bytes);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int mb_pack_crops(mb_ctx* ctx, const uint8_t* pages_dev, int page_h, int page_w,
                             const int32_t* rects_dev, const int32_t* page_idx_dev, int n_crops, void* out_dev,
                             int layout, void* stream_) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_crops == 0) return 0;
    MB_REQUIRE(ctx, n_crops > 0 && (layout == 0 || layout == 1), "pack_crops: bad arguments");
    unsigned char* s = (unsigned char*)mb_scratch(ctx, mb_align_up(sizeof(CropDesc) * (size_t)n_crops, 256) + 512 +
                                                           k9_list_bytes(n_crops));
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
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_crops == 0) return 0;
    MB_REQUIRE(ctx, n_crops > 0 && (layout == 0 || layout == 1), "pack_fragments: bad arguments");
    unsigned char* s = (unsigned char*)mb_scratch(ctx, mb_align_up(sizeof(CropDesc) * (size_t)n_crops, 256) + 512 +
                                                           k9_list_bytes(n_crops));
    if (!s) return MB_ERR_OOM;
    int* err = (int*)s;
    CropDesc* descs = (CropDesc*)(s + 256);
    int* lists = (int*)(s + 256 + mb_align_up(sizeof(CropDesc) * (size_t)n_crops, 256));
    crop_desc_packed_kernel<<<mb_cdiv(n_crops, 128), 128, 0, stream>>>(buf_dev, offsets_dev, hw_dev, n_crops, descs);
    MB_LAUNCH_CHECK(ctx);
    int rc = launch_crop_resize(ctx, descs, n_crops, (bf16*)out_dev, layout, err, lists, stream);
    if (rc) return rc;
    int host_err = 0;
    MB_CUDA(ctx, cudaMemcpyAsync(&host_err, err, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MB_CUDA(ctx, cudaStreamSynchronize(stream));
    if (host_err) return mb_set_err(ctx, MB_ERR_ARG, "pack_fragments: a fragment is too large for the resampler workspace");
    return 0;
}
