// NVCC_FLAGS: -fmad=false
// K6/K7: per-component box extraction, batched over pages.
//   phase 1  per-row extremes of {label == k and text > low_text} inside the component's bbox
//   phase 2  analytic (1+niter)^2 dilation, convex hull, rotating calipers, boxPoints, diamond fix, roll, coordinate
//            adjustment and the +4 px crop rect  (csrc/boxgeom.cuh)
// Phase 2 is a sequential algorithm (OpenCV's operation order is part of the contract), so the parallelism is across
// components: one THREAD per component with a compact workspace in local memory (box_extract_small_kernel, word-sized
// components: <= 64 rows, <= 40 hull vertices per chain); taller or vertex-rich components are queued for the
// warp-per-component kernel with the full-size workspace (box_extract_kernel).
// Reference: marie/models/craft/craft_utils.py:47-98,268-274; marie/boxes/craft_box_processor.py:499-521.
// The reference builds four full-image boolean masks per label (O(N*H*W)); here each component touches only its
// bounding box (O(sum of bbox areas)).
#include "common.cuh"
#include "boxgeom.cuh"

int mb_ccl_run(mb_ctx* ctx, const float* text, const float* link, int n_img, int h, int w, float low_text,
               float link_thr, int* parent, int* rowcount, int* rowbase, unsigned* fg, unsigned* tx, int* wmax,
               int* labels, int* n_labels, int* stats, int max_labels, int* overflow, cudaStream_t stream);

namespace {

constexpr int MAX_ROWS = 4096;   // heat-map height limit (4096 rows = 8192-pixel-high page)

struct BoxPlan {   // one per surviving component, in label order
    int label, x, y, w, h, area, niter;
    int sx, ex, sy, ey;
};

__device__ __forceinline__ float ordered_to_float(int i) {
    return __int_as_float(i ^ ((i >> 31) & 0x7fffffff));
}

// One block per image: filter labels (area >= 10, max text >= text_threshold), ordered compaction, ROI planning,
// and the cv2-layout stats [left, top, width, height, area].
__global__ void box_plan_kernel(const int* __restrict__ raw_stats, const int* __restrict__ n_labels,
                                int* __restrict__ cv_stats, BoxPlan* __restrict__ plans, int* __restrict__ mapper,
                                int* __restrict__ n_boxes, int* __restrict__ overflow, int max_labels,
                                int max_boxes, int img_h, int img_w, float text_threshold) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int img = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int nl = n_labels[img];
    if (nl > max_labels) {
        if (threadIdx.x == 0) atomicExch(overflow, 1);
        nl = max_labels;
    }
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nl; base += blockDim.x) {
        const int k = base + threadIdx.x;
        int valid = 0;
        int x = 0, y = 0, w = 0, h = 0, area = 0;
        if (k < nl && k >= 1) {
            const int* s = raw_stats + ((long long)img * max_labels + k) * 8;
            area = s[0];
            x = s[1]; y = s[2]; w = s[3] - s[1] + 1; h = s[4] - s[2] + 1;
            int* o = cv_stats + ((long long)img * max_labels + k) * 5;
            o[0] = x; o[1] = y; o[2] = w; o[3] = h; o[4] = area;
            valid = (area >= 10) && !(ordered_to_float(s[5]) < text_threshold);
        } else if (k == 0 && k < nl) {
            int* o = cv_stats + (long long)img * max_labels * 5;
            o[0] = o[1] = o[2] = o[3] = o[4] = 0;   // background row is not reproduced
        }
        int inc = valid;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int t = (lane < (int)(blockDim.x >> 5)) ? warp_tot[lane] : 0;
            int ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += u;
            }
            warp_tot[lane] = ti - t;
        }
        __syncthreads();
        const int slot = carry + warp_tot[wid] + inc - valid;
        if (valid) {
            if (slot < max_boxes) {
                BoxPlan p;
                p.label = k; p.x = x; p.y = y; p.w = w; p.h = h; p.area = area;
                // niter = int(sqrt(size * min(w, h) / (w * h)) * 2)   (int32 products, float64 division / sqrt)
                const int mn = w < h ? w : h;
                const int num = (int)((unsigned)area * (unsigned)mn);
                const int den = (int)((unsigned)w * (unsigned)h);
                p.niter = (int)(sqrt((double)num / (double)den) * 2);
                int sx = x - p.niter, ex = x + w + p.niter + 1, sy = y - p.niter, ey = y + h + p.niter + 1;
                if (sx < 0) sx = 0;
                if (sy < 0) sy = 0;
                if (ex >= img_w) ex = img_w;
                if (ey >= img_h) ey = img_h;
                p.sx = sx; p.ex = ex; p.sy = sy; p.ey = ey;
                plans[(long long)img * max_boxes + slot] = p;
                mapper[(long long)img * max_boxes + slot] = k;
            } else {
                atomicExch(overflow, 1);
            }
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = slot + valid;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_boxes[img] = carry < max_boxes ? carry : max_boxes;
}

struct BoxSmem {
    MbHullWork hull;
    short rowmin[MAX_ROWS];
    short rowmax[MAX_ROWS];
};

constexpr int SMALL_ROWS = 64;
constexpr int BOX_LANE_STRIDE = 4;
typedef MbHullWorkT<40, MbPtS, short> SmallHull;

__device__ __forceinline__ void box_write(const float* box, long long o, int img, const double* __restrict__ ratios,
                                          const int* __restrict__ page_hw, float* __restrict__ det,
                                          float* __restrict__ adj, int* __restrict__ rects) {
    float a[8];
    int rect[4];
    const double rw = ratios ? ratios[2 * img] : 1.0, rh = ratios ? ratios[2 * img + 1] : 1.0;
    const int ph = page_hw ? page_hw[2 * img] : 0x7fffffff, pw = page_hw ? page_hw[2 * img + 1] : 0x7fffffff;
    mb_adjust_and_rect(box, rw, rh, pw, ph, a, rect);
    for (int i = 0; i < 8; ++i) { det[o * 8 + i] = box[i]; adj[o * 8 + i] = a[i]; }
    for (int i = 0; i < 4; ++i) rects[o * 4 + i] = rect[i];
}

// One thread per (image, slot).  A run of consecutive `text > low_text` bits lies inside one foreground run, hence
// inside one component: one label fetch per bit group decides the whole group.
__global__ void __launch_bounds__(32)
box_extract_small_kernel(const int* __restrict__ labels, const unsigned* __restrict__ tx, int wd,
                         const BoxPlan* __restrict__ plans, const int* __restrict__ n_boxes, float* __restrict__ det,
                         float* __restrict__ adj, int* __restrict__ rects, int* __restrict__ big_list,
                         int* __restrict__ big_count, int n_img, int max_boxes, int img_h, int img_w,
                         const double* __restrict__ ratios, const int* __restrict__ page_hw) {
    // every BOX_LANE_STRIDE-th lane owns a box: 8 boxes per warp diverge far less than 32 and there are four times
    // as many warps to hide the dependent-load latency (ncu: 3 % of the warp slots active with 32 boxes per warp)
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid % BOX_LANE_STRIDE) return;
    const int id = tid / BOX_LANE_STRIDE;
    if (id >= n_img * max_boxes) return;
    const int img = id / max_boxes;
    const int b = id - img * max_boxes;
    if (b >= n_boxes[img]) return;
    const BoxPlan p = plans[id];
    if (p.h > SMALL_ROWS) {
        big_list[atomicAdd(big_count, 1)] = id;
        return;
    }
    const int* lab = labels + (long long)img * img_h * img_w;
    const unsigned* txi = tx + (long long)img * img_h * wd;
    short rowmin[SMALL_ROWS], rowmax[SMALL_ROWS];
    const int w0 = p.x >> 5, w1 = (p.x + p.w - 1) >> 5;
    for (int r = 0; r < p.h; ++r) {
        const long long row = (long long)(p.y + r) * img_w;
        const unsigned* trow = txi + (long long)(p.y + r) * wd;
        int mn = 0x7fff, mx = -1;
        for (int wx = w0; wx <= w1; ++wx) {
            unsigned t = trow[wx];
            const int lo = p.x - wx * 32, hi = p.x + p.w - 1 - wx * 32;
            if (lo > 0) t &= 0xffffffffu << lo;
            if (hi < 31) t &= 0xffffffffu >> (31 - hi);
            while (t) {
                const int bit = __ffs(t) - 1;
                const unsigned inv = ~(t >> bit);
                const int len = inv ? (__ffs(inv) - 1) : 32;
                const int x0 = wx * 32 + bit;
                if (lab[row + x0] == p.label) {
                    mn = x0 < mn ? x0 : mn;
                    mx = x0 + len - 1 > mx ? x0 + len - 1 : mx;
                }
                t = (bit + len >= 32) ? 0u : (t & (0xffffffffu << (bit + len)));
            }
        }
        rowmin[r] = (short)mn; rowmax[r] = (short)mx;
    }
    SmallHull hw;
    float box[8];
    const int rc = mb_component_box(&hw, rowmin, rowmax, p.y, p.h, p.sx, p.ex, p.sy, p.ey, p.niter, box);
    if (rc < 0) {                    // more hull vertices than the compact workspace holds
        big_list[atomicAdd(big_count, 1)] = id;
        return;
    }
    if (rc == 0)
        for (int i = 0; i < 8; ++i) box[i] = 0.f;       // cannot happen for components that passed the filters
    box_write(box, id, img, ratios, page_hw, det, adj, rects);
}

// One WARP per box, two warps per CTA, 32 CTAs per SM (all 64 warp slots): the per-box geometry (hull + rotating
// calipers) is a long single-thread dependency chain (~40 us), so throughput comes from how many boxes are in flight.
// The hull workspace lives in a per-warp slice of global scratch (the few dozen vertices actually touched stay in L1)
// instead of 64 KB of shared memory, which had limited the first version to 3 CTAs/SM (ncu: 14 % warps active).
// Phase 1 reads the `text > low_text` bit plane written by the labelling pass (32 pixels per word; label words are only
// fetched where a bit is set) instead of the fp32 text map.
constexpr int BOX_WARPS = 2;

__global__ void __launch_bounds__(32 * BOX_WARPS)
box_extract_kernel(const int* __restrict__ labels, const unsigned* __restrict__ tx, int wd,
                   const BoxPlan* __restrict__ plans, const int* __restrict__ big_list,
                   const int* __restrict__ big_count, float* __restrict__ det, float* __restrict__ adj,
                   int* __restrict__ rects, int* __restrict__ overflow, int max_boxes, int img_h, int img_w,
                   const double* __restrict__ ratios, const int* __restrict__ page_hw,
                   BoxSmem* __restrict__ workspace) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * BOX_WARPS + (threadIdx.x >> 5), ngw = gridDim.x * BOX_WARPS;
    BoxSmem* sm = workspace + gw;
    const int n_big = *big_count;
    // persistent warps stride over the queued (image, slot) pairs
    for (int qi = gw; qi < n_big; qi += ngw) {
    const int id = big_list[qi];
    const int img = id / max_boxes;
    __syncwarp();                    // the previous box's single-thread phase is done with the row buffers
    const BoxPlan p = plans[id];
    const int* lab = labels + (long long)img * img_h * img_w;
    const unsigned* txi = tx + (long long)img * img_h * wd;
    const int w0 = p.x >> 5, w1 = (p.x + p.w - 1) >> 5;
    for (int r = 0; r < p.h; ++r) {
        const long long row = (long long)(p.y + r) * img_w;
        const unsigned* trow = txi + (long long)(p.y + r) * wd;
        int mn = 0x7fff, mx = -1;
        for (int wx = w0; wx <= w1; ++wx) {
            const unsigned t = trow[wx];
            if (!t) continue;
            const int x = wx * 32 + lane;
            if (((t >> lane) & 1u) && x >= p.x && x < p.x + p.w && lab[row + x] == p.label) {
                mn = x < mn ? x : mn;
                mx = x > mx ? x : mx;
            }
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) { sm->rowmin[r] = (short)mn; sm->rowmax[r] = (short)mx; }
    }
    __syncwarp();
    if (lane == 0) {
        float box[8];
        const int rc = mb_component_box(&sm->hull, sm->rowmin, sm->rowmax, p.y, p.h, p.sx, p.ex, p.sy, p.ey, p.niter,
                                        box);
        if (rc <= 0) {
            if (rc < 0) atomicExch(overflow, 2);
            for (int i = 0; i < 8; ++i) box[i] = 0.f;   // cannot happen for components that passed the filters
        }
        box_write(box, id, img, ratios, page_hw, det, adj, rects);
    }
    }
}

}  // namespace

// Full score-map post-processing (getDetBoxes + adjustResultCoordinates + rect conversion), device resident.
extern "C" int mb_craft_post(mb_ctx* ctx, const float* text_dev, const float* link_dev, int n_img, int h, int w,
                             float text_threshold, float link_threshold, float low_text,
                             const double* ratios_dev, const int32_t* page_hw_dev, int32_t* labels_dev,
                             int32_t* n_labels_dev, int32_t* stats_dev, int max_labels, float* det_dev,
                             float* adj_dev, int32_t* rects_dev, int32_t* mapper_dev, int32_t* n_boxes_dev,
                             int max_boxes, void* stream_) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    MB_REQUIRE(ctx, n_img > 0 && h > 0 && w > 0, "craft_post: empty input");
    MB_REQUIRE(ctx, h <= MAX_ROWS && w <= 32767, "craft_post: heat map larger than %d rows / 32767 cols", MAX_ROWS);
    MB_REQUIRE(ctx, (long long)h * w < 0x7fffffffLL, "craft_post: heat map too large");
    MB_REQUIRE(ctx, max_labels > 1 && max_boxes > 0, "craft_post: bad capacities");
    const size_t px = (size_t)n_img * h * w;
    // scratch: parent[px] | rowcount[n*h] | rowbase[n*h] | raw_stats[n*max_labels*8] | plans | overflow
    size_t off = 0;
    const size_t o_parent = off; off += mb_align_up(px * 4, 256);
    const size_t o_rowcount = off; off += mb_align_up((size_t)n_img * h * 4, 256);
    const size_t o_rowbase = off; off += mb_align_up((size_t)n_img * h * 4, 256);
    const size_t o_raw = off; off += mb_align_up((size_t)n_img * max_labels * 8 * 4, 256);
    const size_t o_plans = off; off += mb_align_up((size_t)n_img * max_boxes * sizeof(BoxPlan), 256);
    const size_t o_ovf = off; off += 256;
    const long long slots = (long long)n_img * max_boxes;
    MB_REQUIRE(ctx, slots < 0x7fffffffLL, "craft_post: n_img * max_boxes too large");
    const long long want = (slots + BOX_WARPS - 1) / BOX_WARPS;
    const int box_grid = (int)(want < (long long)ctx->num_sms * 4 ? want : (long long)ctx->num_sms * 4);
    const size_t o_boxws = off; off += mb_align_up((size_t)box_grid * BOX_WARPS * sizeof(BoxSmem), 256);
    const size_t o_big = off; off += mb_align_up((size_t)slots * 4 + 256, 256);   // big_count | pad | big_list
    const int wd = (w + 31) / 32;
    const size_t o_fg = off; off += mb_align_up((size_t)n_img * h * wd * 4, 256);
    const size_t o_tx = off; off += mb_align_up((size_t)n_img * h * wd * 4, 256);
    const size_t o_wmax = off; off += mb_align_up((size_t)n_img * h * wd * 4, 256);
    unsigned char* s = (unsigned char*)mb_scratch(ctx, off);
    if (!s) return MB_ERR_OOM;
    int* parent = (int*)(s + o_parent);
    int* rowcount = (int*)(s + o_rowcount);
    int* rowbase = (int*)(s + o_rowbase);
    int* raw = (int*)(s + o_raw);
    BoxPlan* plans = (BoxPlan*)(s + o_plans);
    int* ovf = (int*)(s + o_ovf);

    unsigned* fg = (unsigned*)(s + o_fg);
    unsigned* tx = (unsigned*)(s + o_tx);
    int rc = mb_ccl_run(ctx, text_dev, link_dev, n_img, h, w, low_text, link_threshold, parent, rowcount, rowbase, fg, tx,
                        (int*)(s + o_wmax), labels_dev, n_labels_dev, raw, max_labels, ovf, stream);
    if (rc) return rc;
    box_plan_kernel<<<n_img, 1024, 0, stream>>>(raw, n_labels_dev, stats_dev, plans, mapper_dev, n_boxes_dev, ovf,
                                                max_labels, max_boxes, h, w, text_threshold);
    MB_LAUNCH_CHECK(ctx);
    int* big_count = (int*)(s + o_big);
    int* big_list = big_count + 64;
    MB_CUDA(ctx, cudaMemsetAsync(big_count, 0, sizeof(int), stream));
    box_extract_small_kernel<<<(int)((slots * BOX_LANE_STRIDE + 31) / 32), 32, 0, stream>>>(labels_dev, tx, wd, plans, n_boxes_dev, det_dev,
                                                                          adj_dev, rects_dev, big_list, big_count, n_img,
                                                                          max_boxes, h, w, ratios_dev, page_hw_dev);
    MB_LAUNCH_CHECK(ctx);
    box_extract_kernel<<<box_grid, 32 * BOX_WARPS, 0, stream>>>(labels_dev, tx, wd, plans, big_list, big_count, det_dev,
                                                                adj_dev, rects_dev, ovf, max_boxes, h, w, ratios_dev,
                                                                page_hw_dev, (BoxSmem*)(s + o_boxws));
    MB_LAUNCH_CHECK(ctx);
    int host_ovf = 0;
    MB_CUDA(ctx, cudaMemcpyAsync(&host_ovf, ovf, sizeof(int), cudaMemcpyDeviceToHost, stream));
    MB_CUDA(ctx, cudaStreamSynchronize(stream));
    if (host_ovf == 1)
        return mb_set_err(ctx, MB_ERR_STATE, "craft_post: more components than max_labels=%d / max_boxes=%d",
                          max_labels, max_boxes);
    if (host_ovf == 2) return mb_set_err(ctx, MB_ERR_STATE, "craft_post: convex-hull workspace overflow");
    return 0;
}
