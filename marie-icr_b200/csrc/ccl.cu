// K5: score-map binarisation + 4-connected component labelling + per-label statistics, batched over pages.
//
// Reference: marie/models/craft/craft_utils.py:32-38
//   text_score = text > low_text; link_score = link > link_threshold (strict, float32)
//   comb = clip(text_score + link_score, 0, 1); cv2.connectedComponentsWithStats(comb, connectivity=4)
// OpenCV numbers components in raster order of their first pixel.  Here: union-find with min-index roots
// (atomicMin label equivalence), then rank of each root among the roots of its image in raster order
// (row counts -> exclusive scan -> in-row prefix), which reproduces cv2's numbering exactly.
//
// All work is integer/byte traffic bound by HBM; kernels are grid-stride over n_img*H*W pixels.
#include "common.cuh"

namespace {

__device__ __forceinline__ int uf_find(const int* L, int a) {
    while (true) {
        int p = __ldcg(L + a);
        if (p == a) return a;
        a = p;
    }
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) {
            int old = atomicMin(L + b, a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(L + a, b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// parent[i] = i (pixel index inside its image) for foreground, -1 for background
__global__ void ccl_init_kernel(const float* __restrict__ text, const float* __restrict__ link,
                                int* __restrict__ parent, long long total, int hw, float low_text,
                                float link_thr) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const bool fg = (text[i] > low_text) || (link[i] > link_thr);
        parent[i] = fg ? (int)(i % hw) : -1;
    }
}

__global__ void ccl_merge_kernel(int* __restrict__ parent, long long total, int hw, int w) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        if (parent[i] < 0) continue;
        const int li = (int)(i % hw);
        int* L = parent + (i - li);
        const int x = li % w;
        const bool left = (x > 0) && (L[li - 1] >= 0);
        const bool up = (li >= w) && (L[li - w] >= 0);
        if (left) uf_union(L, li, li - 1);
        // the up-link is redundant when left and up-left are both foreground and already chained through the row above
        if (up && !(left && L[li - w - 1] >= 0)) uf_union(L, li, li - w);
    }
}

// parent[i] <- root; counts roots per image row
__global__ void ccl_flatten_kernel(int* __restrict__ parent, int* __restrict__ rowcount, long long total, int hw,
                                   int w, int h) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        if (parent[i] < 0) continue;
        const int li = (int)(i % hw);
        const long long img = i / hw;
        int* L = parent + (i - li);
        const int r = uf_find(L, li);
        if (r != li) L[li] = r;
        else atomicAdd(rowcount + img * h + li / w, 1);
    }
}

// one block per image: exclusive scan of the per-row root counts; n_labels = roots + 1 (background)
__global__ void ccl_rowscan_kernel(const int* __restrict__ rowcount, int* __restrict__ rowbase,
                                   int* __restrict__ n_labels, int h) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int img = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < h; base += blockDim.x) {
        const int y = base + threadIdx.x;
        const int v = (y < h) ? rowcount[(long long)img * h + y] : 0;
        int inc = v;
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
            warp_tot[lane] = ti - t;   // exclusive warp offsets
            if (lane == 31) warp_tot[31] = ti - t;
        }
        __syncthreads();
        const int excl = carry + warp_tot[wid] + inc - v;
        if (y < h) rowbase[(long long)img * h + y] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_labels[img] = carry + 1;
}

// one warp per image row: roots get their final id, encoded in place as -(id) - 2
__global__ void ccl_assign_kernel(int* __restrict__ parent, const int* __restrict__ rowbase, int n_img, int h,
                                  int w) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const long long rows = (long long)n_img * h;
    const long long rstride = (long long)gridDim.x * warps_per_block;
    for (; row < rows; row += rstride) {
        const int y = (int)(row % h);
        int* L = parent + (row / h) * (long long)h * w;
        int running = rowbase[row];
        for (int x0 = 0; x0 < w; x0 += 32) {
            const int x = x0 + lane;
            const int li = y * w + x;
            const bool is_root = (x < w) && (L[li] == li);
            const unsigned m = __ballot_sync(0xffffffffu, is_root);
            if (is_root) L[li] = -(running + __popc(m & ((1u << lane) - 1)) + 1) - 2;
            running += __popc(m);
        }
    }
}

__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}

// labels[i] = final id; per-label stats via warp-aggregated atomics.
// stats layout per label: [area, minx, miny, maxx, maxy, max_text(ordered int), 0, 0]
__global__ void ccl_finalize_kernel(const int* __restrict__ parent, const float* __restrict__ text,
                                    int* __restrict__ labels, int* __restrict__ stats, int* __restrict__ overflow,
                                    long long total, int hw, int w, int max_labels) {
    long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const long long total_up = (total + 31) / 32 * 32;   // keep warps converged for the match/reduce intrinsics
    for (long long i = i0; i < total_up; i += stride) {
        int id = 0;
        int li = 0;
        long long img = 0;
        if (i < total) {
            li = (int)(i % hw);
            img = i / hw;
            const int p = parent[i];
            if (p < -1) id = -p - 2;
            else if (p >= 0) id = -parent[(i - li) + p] - 2;
            labels[i] = id;
        }
        const unsigned fgmask = __ballot_sync(0xffffffffu, id > 0);
        if (id > 0) {
            // lanes of a warp may straddle two images only when hw % 32 != 0; fold the image into the key
            const long long key = img * (long long)max_labels + id;
            const unsigned m = __match_any_sync(fgmask, key);
            const int x = li % w, y = li / w;
            const int mnx = __reduce_min_sync(m, x), mxx = __reduce_max_sync(m, x);
            const int mny = __reduce_min_sync(m, y), mxy = __reduce_max_sync(m, y);
            const int mt = __reduce_max_sync(m, float_to_ordered(text[i]));
            if (lane == __ffs(m) - 1) {
                if (id < max_labels) {
                    int* s = stats + (img * max_labels + id) * 8;
                    atomicAdd(s + 0, __popc(m));
                    atomicMin(s + 1, mnx);
                    atomicMin(s + 2, mny);
                    atomicMax(s + 3, mxx);
                    atomicMax(s + 4, mxy);
                    atomicMax(s + 5, mt);
                } else {
                    atomicExch(overflow, 1);
                }
            }
        }
    }
}

__global__ void ccl_stats_init_kernel(int* __restrict__ stats, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int f = (int)(i & 7);
        int v = 0;
        if (f == 1 || f == 2) v = 0x7fffffff;
        else if (f == 3 || f == 4) v = -1;
        else if (f == 5) v = (int)0x80000000;
        stats[i] = v;
    }
}

}  // namespace

// Internal entry: labels + raw stats.  parent/rowcount/rowbase are caller-provided scratch.
int mb_ccl_run(mb_ctx* ctx, const float* text, const float* link, int n_img, int h, int w, float low_text,
               float link_thr, int* parent, int* rowcount, int* rowbase, int* labels, int* n_labels, int* stats,
               int max_labels, int* overflow, cudaStream_t stream) {
    const long long total = (long long)n_img * h * w;
    const int hw = h * w;
    const int threads = 256;
    const int grid = (int)((total + threads - 1) / threads < (long long)ctx->num_sms * 16
                               ? (total + threads - 1) / threads
                               : (long long)ctx->num_sms * 16);
    MB_CUDA(ctx, cudaMemsetAsync(rowcount, 0, sizeof(int) * (size_t)n_img * h, stream));
    MB_CUDA(ctx, cudaMemsetAsync(overflow, 0, sizeof(int), stream));
    const long long nstats = (long long)n_img * max_labels * 8;
    ccl_stats_init_kernel<<<(int)((nstats + 255) / 256 < 4096 ? (nstats + 255) / 256 : 4096), 256, 0, stream>>>(stats, nstats);
    MB_LAUNCH_CHECK(ctx);
    ccl_init_kernel<<<grid, threads, 0, stream>>>(text, link, parent, total, hw, low_text, link_thr);
    MB_LAUNCH_CHECK(ctx);
    ccl_merge_kernel<<<grid, threads, 0, stream>>>(parent, total, hw, w);
    MB_LAUNCH_CHECK(ctx);
    ccl_flatten_kernel<<<grid, threads, 0, stream>>>(parent, rowcount, total, hw, w, h);
    MB_LAUNCH_CHECK(ctx);
    ccl_rowscan_kernel<<<n_img, 1024, 0, stream>>>(rowcount, rowbase, n_labels, h);
    MB_LAUNCH_CHECK(ctx);
    {
        const long long rows = (long long)n_img * h;
        const int wpb = 8;
        const int g = (int)((rows + wpb - 1) / wpb < (long long)ctx->num_sms * 8 ? (rows + wpb - 1) / wpb
                                                                               : (long long)ctx->num_sms * 8);
        ccl_assign_kernel<<<g, wpb * 32, 0, stream>>>(parent, rowbase, n_img, h, w);
        MB_LAUNCH_CHECK(ctx);
    }
    ccl_finalize_kernel<<<grid, threads, 0, stream>>>(parent, text, labels, stats, overflow, total, hw, w,
                                                      max_labels);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}
