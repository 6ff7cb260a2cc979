// The CRAFT link refiner and its line branch.
//
//   RefineNet.forward (marie/models/craft/refinenet.py:15-66): cat(y, upconv4) [34 ch] -> three 3x3 conv+BN+ReLU (64 ch)
//     -> four ASPP branches (3x3 dilation 6 / 12 / 18 / 24 -> 128, 1x1 -> 128, 1x1 -> 1) -> sum of the four.
//   Line branch of get_prediction (marie/boxes/craft_box_processor.py:150-217): refined link > threshold,
//     MORPH_CLOSE 3x3, 4-connected components with stats -> [x, y, w, h] per label -> line_merge (host, lines.cu).
//
// Every convolution runs on the tap-GEMM (gemm_tc.cu; dilation = shifted TMA boxes with zero fill).  The 34-channel
// input is the 64-channel padded `feature` tensor of mb_craft_forward with the two score maps dropped into its unused
// channels 32 / 33 (the packer permutes the first layer's input channels accordingly); the four final 1x1 -> 1
// convolutions and their sum are ONE 1x1 convolution over the four branches' hidden states laid side by side (512 ch).
// The morphology works on bit planes (32 pixels per word); labelling reuses the run-based CCL of ccl.cu.
#include "common.cuh"
#include "blob.cuh"

int mb_ccl_run_planes(mb_ctx* ctx, int n_img, int h, int w, int* parent, int* rowcount, int* rowbase, unsigned* fg,
                      int* wmax, int* labels, int* n_labels, int* stats, int max_labels, int* overflow,
                      cudaStream_t stream);

struct RefineLayer {
    const bf16* w = nullptr;
    const float* b = nullptr;
    int rows = 0, k = 0;
};

struct RefineModel {
    WeightBlob blob;
    RefineLayer c[3];        // last_conv
    RefineLayer a[4][2];     // aspp k: 3x3 dilated, 1x1
    RefineLayer fin;         // 4 x (1x1 -> 1) side by side: [16, 512]
    void* arena = nullptr;
    size_t arena_bytes = 0;
};

namespace {

// feature[..., 32] = text, feature[..., 33] = link (16-bit), one thread per pixel
__global__ void refine_pack_kernel(bf16* __restrict__ feat, const float* __restrict__ text, const float* __restrict__ link,
                                   long long px, int f16) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < px; i += stride)
        *reinterpret_cast<unsigned int*>(feat + i * 64 + 32) = pack2(text[i], link[i], f16);
}

// link > threshold -> bit plane; one warp per image row
__global__ void __launch_bounds__(256)
line_threshold_kernel(const float* __restrict__ link, unsigned* __restrict__ bits, int rows, int w, int wd, float thr) {
    const int lane = threadIdx.x & 31;
    const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int rowi = warp0; rowi < rows; rowi += nwarps) {
        const float* lrow = link + (long long)rowi * w;
        for (int wx = 0; wx < wd; ++wx) {
            const int x = wx * 32 + lane;
            const float v = x < w ? __ldcs(lrow + x) : -INFINITY;
            const unsigned m = __ballot_sync(0xffffffffu, v > thr);
            if (lane == 0) bits[(long long)rowi * wd + wx] = m;
        }
    }
}

// 3x3 dilation (ERODE = false) or erosion (ERODE = true) of a bit plane with OpenCV's morphology border: pixels outside
// the image never win (background for the dilation, foreground for the erosion).  One thread per word.
template <bool ERODE>
__global__ void __launch_bounds__(256)
morph3x3_kernel(const unsigned* __restrict__ in, unsigned* __restrict__ out, int n_words, int h, int w, int wd) {
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    const unsigned tail_mask = (w & 31) ? (0xffffffffu >> (32 - (w & 31))) : 0xffffffffu;   // valid bits of a row's last word
    for (; wi < n_words; wi += stride) {
        const int rowi = wi / wd;
        const int wx = wi - rowi * wd;
        const int y = rowi % h;
        unsigned acc = ERODE ? 0xffffffffu : 0u;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= h) continue;                       // outside rows never win
            const unsigned* r = in + (long long)(rowi + dy) * wd;
            unsigned m = r[wx];
            unsigned left = wx > 0 ? r[wx - 1] : (ERODE ? 0xffffffffu : 0u);
            unsigned right = wx + 1 < wd ? r[wx + 1] : (ERODE ? 0xffffffffu : 0u);
            if (ERODE) {                                            // columns beyond w count as foreground
                if (wx == wd - 1) m |= ~tail_mask;
                if (wx + 1 == wd - 1) right |= ~tail_mask;
            }
            const unsigned l1 = (m << 1) | (left >> 31);            // pixel x-1 at bit x
            const unsigned r1 = (m >> 1) | (right << 31);           // pixel x+1 at bit x
            if (ERODE) acc &= m & l1 & r1;
            else acc |= m | l1 | r1;
        }
        if (wx == wd - 1) acc &= tail_mask;
        out[wi] = acc;
    }
}

// raw statistics [area, minx, miny, maxx, maxy, ...] -> cv2 layout [left, top, width, height, area]
__global__ void line_stats_kernel(const int* __restrict__ raw, const int* __restrict__ n_labels, int* __restrict__ cv,
                                  int n_img, int max_labels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_img * max_labels) return;
    const int img = i / max_labels, k = i - img * max_labels;
    int* o = cv + (long long)i * 5;
    if (k == 0 || k >= n_labels[img]) { o[0] = o[1] = o[2] = o[3] = o[4] = 0; return; }
    const int* s = raw + (long long)i * 8;
    o[0] = s[1]; o[1] = s[2]; o[2] = s[3] - s[1] + 1; o[3] = s[4] - s[2] + 1; o[4] = s[0];
}

int grid_for(mb_ctx* ctx, long long items, int threads, int per_sm = 8) {
    const long long want = (items + threads - 1) / threads;
    const long long cap = (long long)ctx->num_sms * per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

int rconv(mb_ctx* ctx, const RefineLayer& L, const bf16* in, int cin, int n, int h, int w, int taps, int dil, int act,
          void* out, int out_c, long long out_ld, int out_mode, long long out_plane, cudaStream_t s) {
    TapGemm g;
    g.a0 = in; g.c0 = cin; g.a0_ld = cin;
    g.n = n; g.h = h; g.w = w; g.taps = taps; g.dil = dil;
    g.wgt = L.w; g.n_rows_w = L.rows; g.n_out = out_c;
    g.bias = L.b; g.act = act;
    g.out = out; g.out_ld = out_ld; g.out_mode = out_mode; g.out_plane = out_plane;
    if (L.k != taps * cin) return mb_set_err(ctx, MB_ERR_STATE, "refine: layer expects K=%d, got %d", L.k, taps * cin);
    return mb_tap_gemm(ctx, g, s);
}

}  // namespace

void mb_free_refine(mb_ctx* ctx) {
    if (!ctx->refine) return;
    ctx->refine->blob.release();
    if (ctx->refine->arena) cudaFree(ctx->refine->arena);
    delete ctx->refine;
    ctx->refine = nullptr;
}

extern "C" int mb_load_refine(mb_ctx* ctx, const void* blob_host, size_t nbytes) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    mb_free_refine(ctx);
    RefineModel* m = new RefineModel();
    ctx->refine = m;
    int rc = m->blob.load(ctx, blob_host, nbytes);
    if (rc) { mb_free_refine(ctx); return rc; }
    bool ok = true;
    std::string missing;
    auto L = [&](const std::string& name, RefineLayer& l) {
        const BlobTensor* w = m->blob.get(name + ".w");
        const BlobTensor* b = m->blob.get(name + ".b");
        if (!w || !b || w->dtype != (ctx->f16 ? 3 : 1) || b->dtype != 0 || w->ndim != 2) { ok = false; missing = name; return; }
        l.w = (const bf16*)w->dev; l.b = (const float*)b->dev; l.rows = (int)w->dims[0]; l.k = (int)w->dims[1];
    };
    for (int i = 0; i < 3; ++i) L("ref.c" + std::to_string(i + 1), m->c[i]);
    for (int k = 0; k < 4; ++k) {
        L("ref.a" + std::to_string(k + 1) + "a", m->a[k][0]);
        L("ref.a" + std::to_string(k + 1) + "b", m->a[k][1]);
    }
    L("ref.final", m->fin);
    if (!ok) {
        mb_free_refine(ctx);
        return mb_set_err(ctx, MB_ERR_ARG, "refine blob: layer %s missing, malformed or not packed for the context dtype (%s)",
                          missing.c_str(), ctx->f16 ? "fp16" : "bf16");
    }
    return 0;
}

// feature_dev: [n, h, w, 64] 16-bit NHWC as written by mb_craft_forward (channels 32..63 zero); channels 32 / 33 are
// OVERWRITTEN with the score maps.  scores_dev: [2][n][h][w] fp32 (text, link).  link_out_dev: [n][h][w] fp32.
extern "C" int mb_refine_forward(mb_ctx* ctx, void* feature_dev, const float* scores_dev, int n, int h, int w,
                                 float* link_out_dev, void* stream_) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    RefineModel* m = ctx->refine;
    if (!m) return mb_set_err(ctx, MB_ERR_STATE, "refine: weights not loaded (mb_load_refine)");
    MB_REQUIRE(ctx, n > 0 && h > 0 && w > 0 && feature_dev && scores_dev && link_out_dev, "refine_forward: bad arguments");
    cudaStream_t s = (cudaStream_t)stream_;
    const long long P = (long long)n * h * w;
    size_t off = 0;
    auto take = [&](long long elems) { size_t o = off; off += mb_align_up((size_t)elems * 2, 1024); return o; };
    const size_t oA = take(P * 64), oB = take(P * 64), oT = take(P * 128), oH = take(P * 512);
    if (off > m->arena_bytes) {
        if (m->arena) cudaFree(m->arena);
        m->arena = nullptr; m->arena_bytes = 0;
        if (cudaMalloc(&m->arena, off) != cudaSuccess) {
            cudaGetLastError();
            return mb_set_err(ctx, MB_ERR_OOM, "refine: activation arena of %zu bytes failed", off);
        }
        m->arena_bytes = off;
    }
    unsigned char* base = (unsigned char*)m->arena;
    bf16* A = (bf16*)(base + oA);
    bf16* B = (bf16*)(base + oB);
    bf16* T = (bf16*)(base + oT);
    bf16* H = (bf16*)(base + oH);
    bf16* X = (bf16*)feature_dev;
    refine_pack_kernel<<<grid_for(ctx, P, 256), 256, 0, s>>>(X, scores_dev, scores_dev + P, P, ctx->f16);
    MB_LAUNCH_CHECK(ctx);
    int rc;
    if ((rc = rconv(ctx, m->c[0], X, 64, n, h, w, 9, 1, MB_ACT_RELU, A, 64, 64, MB_OUT_BF16, 0, s))) return rc;
    if ((rc = rconv(ctx, m->c[1], A, 64, n, h, w, 9, 1, MB_ACT_RELU, B, 64, 64, MB_OUT_BF16, 0, s))) return rc;
    if ((rc = rconv(ctx, m->c[2], B, 64, n, h, w, 9, 1, MB_ACT_RELU, A, 64, 64, MB_OUT_BF16, 0, s))) return rc;
    const int dil[4] = {6, 12, 18, 24};
    for (int k = 0; k < 4; ++k) {
        if ((rc = rconv(ctx, m->a[k][0], A, 64, n, h, w, 9, dil[k], MB_ACT_RELU, T, 128, 128, MB_OUT_BF16, 0, s))) return rc;
        // hidden state of branch k -> columns [128 k, 128 k + 128) of the 512-wide tensor
        if ((rc = rconv(ctx, m->a[k][1], T, 128, n, h, w, 1, 1, MB_ACT_RELU, H + k * 128, 128, 512, MB_OUT_BF16, 0, s))) return rc;
    }
    return rconv(ctx, m->fin, H, 512, n, h, w, 1, 1, MB_ACT_NONE, link_out_dev, 1, 1, MB_OUT_F32_PLANAR, P, s);
}

// Components of the refiner's line map: link > threshold, 3x3 closing, 4-connected labelling.
// labels_dev [n,h,w] i32 (may be null), n_labels_dev [n] (incl. background), stats_dev [n, max_labels, 5] in cv2's
// layout (row 0 zero).  Label order = raster order of each component's first pixel, like cv2.
extern "C" int mb_line_components(mb_ctx* ctx, const float* link_dev, int n, int h, int w, float link_threshold,
                                  int32_t* labels_dev, int32_t* n_labels_dev, int32_t* stats_dev, int max_labels,
                                  void* stream_) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream_;
    MB_REQUIRE(ctx, n > 0 && h > 0 && w > 0 && link_dev && n_labels_dev && stats_dev && max_labels > 1,
               "line_components: bad arguments");
    const int wd = (w + 31) / 32;
    const long long px = (long long)n * h * w, nw = (long long)n * h * wd, rows = (long long)n * h;
    MB_REQUIRE(ctx, nw < 0x7fffffffLL && (long long)h * w < 0x7fffffffLL, "line_components: batch too large");
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += mb_align_up(bytes, 256); return o; };
    const size_t o_parent = take(px * 4), o_labels = take(labels_dev ? 0 : px * 4), o_rowcount = take(rows * 4),
                 o_rowbase = take(rows * 4), o_raw = take((size_t)n * max_labels * 32), o_b0 = take(nw * 4),
                 o_b1 = take(nw * 4), o_wmax = take(nw * 4), o_ovf = take(256);
    unsigned char* sc = (unsigned char*)mb_scratch(ctx, off);
    if (!sc) return MB_ERR_OOM;
    unsigned* b0 = (unsigned*)(sc + o_b0);
    unsigned* b1 = (unsigned*)(sc + o_b1);
    int* ovf = (int*)(sc + o_ovf);
    int* labels = labels_dev ? labels_dev : (int*)(sc + o_labels);
    line_threshold_kernel<<<grid_for(ctx, rows * 32, 256), 256, 0, s>>>(link_dev, b0, (int)rows, w, wd, link_threshold);
    MB_LAUNCH_CHECK(ctx);
    morph3x3_kernel<false><<<grid_for(ctx, nw, 256), 256, 0, s>>>(b0, b1, (int)nw, h, w, wd);
    MB_LAUNCH_CHECK(ctx);
    morph3x3_kernel<true><<<grid_for(ctx, nw, 256), 256, 0, s>>>(b1, b0, (int)nw, h, w, wd);
    MB_LAUNCH_CHECK(ctx);
    int rc = mb_ccl_run_planes(ctx, n, h, w, (int*)(sc + o_parent), (int*)(sc + o_rowcount), (int*)(sc + o_rowbase), b0,
                               (int*)(sc + o_wmax), labels, n_labels_dev, (int*)(sc + o_raw), max_labels, ovf, s);
    if (rc) return rc;
    line_stats_kernel<<<mb_cdiv(n * max_labels, 256), 256, 0, s>>>((int*)(sc + o_raw), n_labels_dev, stats_dev, n, max_labels);
    MB_LAUNCH_CHECK(ctx);
    int host_ovf = 0;
    MB_CUDA(ctx, cudaMemcpyAsync(&host_ovf, ovf, sizeof(int), cudaMemcpyDeviceToHost, s));
    MB_CUDA(ctx, cudaStreamSynchronize(s));
    if (host_ovf) return mb_set_err(ctx, MB_ERR_STATE, "line_components: more components than max_labels=%d", max_labels);
    return 0;
}
