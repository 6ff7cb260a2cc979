// K5: score-map binarisation + 4-connected component labelling + per-label statistics, batched over pages.
//
// Reference: marie/models/craft/craft_utils.py:32-38
//   text_score = text > low_text; link_score = link > link_threshold (strict, float32)
//   comb = clip(text_score + link_score, 0, 1); cv2.connectedComponentsWithStats(comb, connectivity=4)
// OpenCV numbers components in raster order of their first pixel.
//
// The score maps are read ONCE (8 B/px) and reduced to two bit planes (1 bit/px each): fg = text|link and
// tx = text > low_text.  Everything between that pass and the final label write works on the bit planes, 32 pixels per
// word, with the maximal horizontal RUN as the union-find node (node id = pixel index of the run's first pixel; the
// parent array is only ever touched at run starts):
//   mask      text, link -> fg, tx bit words; parent[run start] = itself                     (streams 8 B/px)
//   merge     runs that overlap a run of the row above are united (atomicMin: the root is the smallest node id,
//             i.e. the component's first pixel in raster order)
//   flatten   parent[run start] <- root; roots counted per image row
//   rowscan   exclusive scan of the per-row root counts (one block per image)
//   assign    roots get their final id = raster rank + 1 (cv2's numbering), encoded in place as -(id) - 2
//   finalize  labels i32 written for every pixel (streams 4 B/px, one thread per word), statistics per run piece (atomics);
//             words with a single run take the max text score from the per-word maximum of the mask pass
// Page text is sparse, so the middle passes move a few bytes per 32 pixels.
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

__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}

// first pixel (x) of the run of `row` (bit words) that contains bit `b` of word `wx`
__device__ __forceinline__ int run_start(const unsigned* __restrict__ row, int wx, int b, unsigned m) {
    const unsigned inv = ~m & ((b == 0) ? 0u : (0xffffffffu >> (32 - b)));   // clear bits below b
    if (inv) return wx * 32 + (32 - __clz(inv));
    for (int j = wx - 1; j >= 0; --j) {
        const unsigned iv = ~row[j];
        if (iv) return j * 32 + (32 - __clz(iv));      // one past the highest clear bit (32 -> next word's bit 0)
    }
    return 0;
}

// one warp per image row (grid-stride over n_img * h rows), 32 pixels = one bit word per step; words of a row:
// wd = ceil(w / 32).  No 64-bit division anywhere on the per-word path.
__global__ void __launch_bounds__(256)
ccl_mask_kernel(const float* __restrict__ text, const float* __restrict__ link, unsigned* __restrict__ fg,
                unsigned* __restrict__ tx, int* __restrict__ wmax, int* __restrict__ parent, int rows, int h, int w,
                int wd, float low_text, float link_thr) {
    const int lane = threadIdx.x & 31;
    const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    constexpr int U = 4;                                  // words in flight per warp (loads issued before any ballot)
    for (int rowi = warp0; rowi < rows; rowi += nwarps) {
        const int y = rowi % h;
        const float* trow = text + (long long)rowi * w;
        const float* lrow = link + (long long)rowi * w;
        int* prow = parent + (long long)rowi * w;
        unsigned* frow = fg + (long long)rowi * wd;
        unsigned* xrow = tx + (long long)rowi * wd;
        int* mxrow = wmax + (long long)rowi * wd;
        unsigned prev = 0;                                // bit 31 of the previous word
        for (int wx0 = 0; wx0 < wd; wx0 += U) {
            float t[U], l[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int x = (wx0 + u) * 32 + lane;
                t[u] = l[u] = -INFINITY;
                if (x < w) { t[u] = __ldcs(trow + x); l[u] = __ldcs(lrow + x); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int wx = wx0 + u;
                if (wx >= wd) break;
                const bool tb = t[u] > low_text;
                const unsigned mt = __ballot_sync(0xffffffffu, tb);
                const bool fgb = tb || (l[u] > link_thr);
                const unsigned m = __ballot_sync(0xffffffffu, fgb);
                // max text score over the word's foreground pixels (ordered-int form): the statistics pass uses it for
                // words that hold a single run, i.e. a single component
                const int wm = m ? __reduce_max_sync(0xffffffffu, fgb ? float_to_ordered(t[u]) : (int)0x80000000) : (int)0x80000000;
                if (lane == 0) { frow[wx] = m; xrow[wx] = mt; mxrow[wx] = wm; }
                const unsigned starts = m & ~((m << 1) | prev);
                if ((starts >> lane) & 1u) prow[wx * 32 + lane] = y * w + wx * 32 + lane;
                prev = m >> 31;
            }
        }
    }
}

// one thread per word: unite every run segment of this word with the run of the row above it touches
__global__ void __launch_bounds__(256)
ccl_merge_kernel(const unsigned* __restrict__ fg, int* __restrict__ parent, int n_words, int h, int w, int wd) {
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    for (; wi < n_words; wi += stride) {
        const unsigned cur = fg[wi];
        if (!cur || wi < wd) continue;
        const unsigned up = fg[wi - wd];                 // for y == 0 this is another image's last row: rejected below
        unsigned both = cur & up;
        if (!both) continue;
        const int rowi = wi / wd;
        const int y = rowi % h;
        if (y == 0) continue;
        const int wx = wi - rowi * wd;
        const unsigned* crow = fg + (long long)rowi * wd;
        const unsigned* urow = crow - wd;
        int* L = parent + (rowi - y) * (long long)w;     // this image's parent plane
        // a segment that starts at bit 0 and continues the same pair of runs from the previous word is redundant
        if ((both & 1u) && wx > 0 && (crow[wx - 1] & urow[wx - 1] & 0x80000000u)) {
            const unsigned rest = ~both;                  // drop the leading group
            both = rest ? (both & ~((rest & (0u - rest)) - 1u)) : 0u;
        }
        while (both) {
            const int b = __ffs(both) - 1;
            const int cs = run_start(crow, wx, b, cur);
            const int us = run_start(urow, wx, b, up);
            uf_union(L, y * w + cs, (y - 1) * w + us);
            // clear this group of consecutive bits
            const unsigned shifted = both >> b;
            const unsigned inv = ~shifted;
            const int len = inv ? (__ffs(inv) - 1) : 32;
            both = (b + len >= 32) ? 0u : (both & (0xffffffffu << (b + len)));
        }
    }
}

// one thread per word: parent[run start] <- root; counts roots per image row
__global__ void __launch_bounds__(256)
ccl_flatten_kernel(const unsigned* __restrict__ fg, int* __restrict__ parent, int* __restrict__ rowcount,
                   int n_words, int h, int w, int wd) {
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    for (; wi < n_words; wi += stride) {
        const unsigned m = fg[wi];
        if (!m) continue;
        const int rowi = wi / wd;
        const int wx = wi - rowi * wd;
        const unsigned prev = wx > 0 ? (fg[wi - 1] >> 31) : 0u;
        unsigned starts = m & ~((m << 1) | prev);
        if (!starts) continue;
        const int y = rowi % h;
        int* L = parent + (rowi - y) * (long long)w;
        int roots = 0;
        while (starts) {
            const int b = __ffs(starts) - 1;
            starts &= starts - 1;
            const int li = y * w + wx * 32 + b;
            const int r = uf_find(L, li);
            if (r != li) L[li] = r;
            else ++roots;
        }
        if (roots) atomicAdd(rowcount + rowi, roots);
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

// one warp per image row that holds a root: roots get their final id, encoded in place as -(id) - 2
__global__ void ccl_assign_kernel(const unsigned* __restrict__ fg, int* __restrict__ parent,
                                  const int* __restrict__ rowcount, const int* __restrict__ rowbase, long long rows,
                                  int h, int w, int wd) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const long long rstride = (long long)gridDim.x * warps_per_block;
    for (; row < rows; row += rstride) {
        if (rowcount[row] == 0) continue;
        const int y = (int)(row % h);
        int* L = parent + (row - y) * (long long)w;
        const unsigned* mrow = fg + row * wd;
        int running = rowbase[row];
        for (int w0 = 0; w0 < wd; w0 += 32) {
            const int wx = w0 + lane;
            unsigned rootbits = 0;
            if (wx < wd) {
                const unsigned m = mrow[wx];
                const unsigned prev = wx > 0 ? (mrow[wx - 1] >> 31) : 0u;
                unsigned starts = m & ~((m << 1) | prev);
                while (starts) {
                    const int b = __ffs(starts) - 1;
                    starts &= starts - 1;
                    const int li = y * w + wx * 32 + b;
                    if (L[li] == li) rootbits |= 1u << b;
                }
            }
            const int cnt = __popc(rootbits);
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            int id = running + inc - cnt;                 // ids before this lane's first root
            while (rootbits) {
                const int b = __ffs(rootbits) - 1;
                rootbits &= rootbits - 1;
                ++id;
                L[y * w + wx * 32 + b] = -id - 2;
            }
            running += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
}

// One THREAD per 32-pixel word: the word's labels (eight 16-byte stores) and the statistics of every run piece in it.
// stats layout per label: [area, minx, miny, maxx, maxy, max_text(ordered int), 0, 0]
// (The first version used a warp per word — ballots, shuffles and a match per word: ncu 62.5 M warp instructions for
// 635 K words; empty words are now a few instructions of one thread, foreground words a 32-step bit scan.)
__global__ void __launch_bounds__(256)
ccl_finalize_kernel(const unsigned* __restrict__ fg, const int* __restrict__ wmax, const int* __restrict__ parent,
                    const float* __restrict__ text, int* __restrict__ labels, int* __restrict__ stats,
                    int* __restrict__ overflow, int n_words, int h, int w, int wd, int max_labels) {
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    const bool vec_ok = (w & 3) == 0;                             // rows (and words) are 16-byte aligned
    for (; wi < n_words; wi += stride) {
        const unsigned m = fg[wi];
        const int rowi = wi / wd;
        const int wx = wi - rowi * wd;
        const int x0 = wx * 32;
        const int valid = min(32, w - x0);
        int* lrow = labels + (long long)rowi * w + x0;
        if (m == 0) {
            if (vec_ok && valid == 32) {
#pragma unroll
                for (int q = 0; q < 8; ++q) __stcs(reinterpret_cast<int4*>(lrow) + q, make_int4(0, 0, 0, 0));
            } else {
                for (int b = 0; b < valid; ++b) lrow[b] = 0;
            }
            continue;
        }
        const int y = rowi % h;
        const int img = rowi / h;
        const int* L = parent + (long long)(rowi - y) * w;
        const unsigned gstarts = m & ~(m << 1);
        const bool single = (gstarts & (gstarts - 1)) == 0;
        int cur = 0, gstart = 0, gmax = (int)0x80000000;
        int out4[4];
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const bool on = (m >> b) & 1u;
            if (on && !((b > 0) && ((m >> (b - 1)) & 1u))) {          // a run piece starts here: resolve its label
                const int s0 = (b == 0) ? run_start(fg + (long long)rowi * wd, wx, 0, m) : x0 + b;
                const int p = L[y * w + s0];
                cur = (p < -1) ? (-p - 2) : (-L[p] - 2);
                gstart = b;
                gmax = (int)0x80000000;
            }
            if (on && !single && text != nullptr) gmax = max(gmax, float_to_ordered(text[(long long)rowi * w + x0 + b]));
            out4[b & 3] = on ? cur : 0;
            if ((b & 3) == 3) {
                if (vec_ok && valid == 32) {
                    __stcs(reinterpret_cast<int4*>(lrow) + (b >> 2), make_int4(out4[0], out4[1], out4[2], out4[3]));
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (b - 3 + k < valid) lrow[b - 3 + k] = out4[k];
                }
            }
            if (on && (b == 31 || !((m >> (b + 1)) & 1u))) {          // the piece ends here: its statistics
                if (cur < max_labels) {
                    int* s = stats + ((long long)img * max_labels + cur) * 8;
                    atomicAdd(s + 0, b - gstart + 1);
                    atomicMin(s + 1, x0 + gstart);
                    atomicMin(s + 2, y);
                    atomicMax(s + 3, x0 + b);
                    atomicMax(s + 4, y);
                    atomicMax(s + 5, single ? wmax[wi] : gmax);
                } else {
                    atomicExch(overflow, 1);
                }
            }
        }
    }
}

// parent[run start] = itself for a foreground bit plane that already exists (line components: refine.cu); also fills
// the per-word maximum with "no text statistics"
__global__ void __launch_bounds__(256)
ccl_init_runs_kernel(const unsigned* __restrict__ fg, int* __restrict__ wmax, int* __restrict__ parent, int n_words, int h,
                     int w, int wd) {
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    for (; wi < n_words; wi += stride) {
        wmax[wi] = (int)0x80000000;
        const unsigned m = fg[wi];
        if (!m) continue;
        const int rowi = wi / wd;
        const int wx = wi - rowi * wd;
        const unsigned prev = wx > 0 ? (fg[wi - 1] >> 31) : 0u;
        unsigned starts = m & ~((m << 1) | prev);
        const int y = rowi % h;
        while (starts) {
            const int b = __ffs(starts) - 1;
            starts &= starts - 1;
            parent[(long long)rowi * w + wx * 32 + b] = y * w + wx * 32 + b;
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

// Internal entry: labels + raw stats + the two bit planes.  parent / rowcount / rowbase / fg / tx are caller-provided
// scratch (fg, tx, wmax: n_img * h * ceil(w / 32) words each; tx is consumed by the box extraction).
int mb_ccl_run(mb_ctx* ctx, const float* text, const float* link, int n_img, int h, int w, float low_text,
               float link_thr, int* parent, int* rowcount, int* rowbase, unsigned* fg, unsigned* tx, int* wmax,
               int* labels, int* n_labels, int* stats, int max_labels, int* overflow, cudaStream_t stream) {
    const int wd = (w + 31) / 32;
    const long long n_words_ll = (long long)n_img * h * wd;
    const long long rows_ll = (long long)n_img * h;
    MB_REQUIRE(ctx, n_words_ll < 0x7fffffffLL, "ccl: batch too large (n_img * h * ceil(w / 32) must fit 31 bits)");
    const int n_words = (int)n_words_ll, rows = (int)rows_ll;
    const int threads = 256;
    auto grid_for = [&](long long items, int per_sm) {
        const long long want = (items + threads - 1) / threads;
        const long long cap = (long long)ctx->num_sms * per_sm;
        return (int)(want < cap ? want : cap);
    };
    MB_CUDA(ctx, cudaMemsetAsync(rowcount, 0, sizeof(int) * (size_t)rows, stream));
    MB_CUDA(ctx, cudaMemsetAsync(overflow, 0, sizeof(int), stream));
    const long long nstats = (long long)n_img * max_labels * 8;
    ccl_stats_init_kernel<<<grid_for(nstats, 8), threads, 0, stream>>>(stats, nstats);
    MB_LAUNCH_CHECK(ctx);
    if (text != nullptr && link != nullptr) {
        ccl_mask_kernel<<<grid_for((long long)rows * 32, 8), threads, 0, stream>>>(text, link, fg, tx, wmax, parent, rows, h,
                                                                                   w, wd, low_text, link_thr);
    } else {
        // the foreground plane is given (mb_ccl_run_planes): only the run nodes have to be created
        ccl_init_runs_kernel<<<grid_for(n_words, 8), threads, 0, stream>>>(fg, wmax, parent, n_words, h, w, wd);
    }
    MB_LAUNCH_CHECK(ctx);
    ccl_merge_kernel<<<grid_for(n_words, 8), threads, 0, stream>>>(fg, parent, n_words, h, w, wd);
    MB_LAUNCH_CHECK(ctx);
    ccl_flatten_kernel<<<grid_for(n_words, 8), threads, 0, stream>>>(fg, parent, rowcount, n_words, h, w, wd);
    MB_LAUNCH_CHECK(ctx);
    ccl_rowscan_kernel<<<n_img, 1024, 0, stream>>>(rowcount, rowbase, n_labels, h);
    MB_LAUNCH_CHECK(ctx);
    ccl_assign_kernel<<<grid_for((long long)rows * 32, 8), threads, 0, stream>>>(fg, parent, rowcount, rowbase, rows, h, w,
                                                                                 wd);
    MB_LAUNCH_CHECK(ctx);
    ccl_finalize_kernel<<<grid_for(n_words, 8), threads, 0, stream>>>(fg, wmax, parent, text, labels, stats, overflow,
                                                                      n_words, h, w, wd, max_labels);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

// Labelling of a foreground bit plane that the caller has already built (fg: n_img * h * ceil(w / 32) words).
int mb_ccl_run_planes(mb_ctx* ctx, int n_img, int h, int w, int* parent, int* rowcount, int* rowbase, unsigned* fg,
                      int* wmax, int* labels, int* n_labels, int* stats, int max_labels, int* overflow,
                      cudaStream_t stream) {
    return mb_ccl_run(ctx, nullptr, nullptr, n_img, h, w, 0.f, 0.f, parent, rowcount, rowbase, fg, nullptr, wmax, labels,
                      n_labels, stats, max_labels, overflow, stream);
}
