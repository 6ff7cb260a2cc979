// Sequential geometry of CRAFT box extraction, written once for host and device (MB_HD):
//   streaming convex-hull reduction over per-row extremes -> OpenCV-order convex hull (Sklansky) ->
//   rotating calipers (float32, OpenCV operation order) -> RotatedRect::points -> diamond fix -> roll.
// Reference semantics: marie/models/craft/craft_utils.py:74-93 (cv2.dilate + cv2.minAreaRect + cv2.boxPoints +
// diamond fix + roll) and :268-274 (adjustResultCoordinates); marie/boxes/craft_box_processor.py:499-521
// (int32 truncation, boundingRect, +4 px expansion).  OpenCV's algorithms (imgproc/convhull.cpp,
// rotcalipers.cpp) are restated from their published form; oracle/craft_post.py holds the same restatement in
// Python, pinned against cv2 4.13 and the reference.
//
// Floating point: every float32 expression below must round after each operation exactly like the x86-64 scalar
// code in OpenCV — this translation unit is compiled with -fmad=false (device) / -ffp-contract=off (host).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MB_HD __host__ __device__ __forceinline__
#else
#define MB_HD static inline
#endif

#define MB_HULL_CAP 512    // per chain; convex lattice polygons in a 4096^2 grid have < 1000 vertices

struct MbPt { int x, y; };
struct MbPtS { short x, y; };   // compact form (heat maps are at most 32767 wide / high)

// Workspace of one component.  CAP bounds each chain; PT / IX are the point and index storage types (the
// one-thread-per-box kernel keeps a small workspace in local memory: MbHullWorkT<40, MbPtS, short>).
template <int CAP, class PT, class IX>
struct MbHullWorkT {
    typedef PT pt_t;
    typedef IX ix_t;
    static const int cap = CAP;
    PT left[CAP];     // left chain (row minima), top to bottom
    PT right[CAP];    // right chain (row maxima)
    int nleft, nright;
    // merged + sorted vertex set and OpenCV hull scratch
    PT v[2 * CAP];
    int nv;
    IX order[2 * CAP];     // hull output: indices into v
    IX stack[2 * CAP + 4];
    IX stack2[2 * CAP + 4];
    float vx[2 * CAP], vy[2 * CAP], inv[2 * CAP];
    float hx[2 * CAP], hy[2 * CAP];
    int overflow;
};
typedef MbHullWorkT<MB_HULL_CAP, MbPt, int> MbHullWork;

template <class PT>
MB_HD MbPt mb_pt(PT p) { MbPt q; q.x = p.x; q.y = p.y; return q; }

MB_HD long long mb_cross(MbPt o, MbPt a, MbPt b) {
    return (long long)(a.x - o.x) * (b.y - o.y) - (long long)(a.y - o.y) * (b.x - o.x);
}

template <class W>
MB_HD void mb_hull_begin(W* w) { w->nleft = w->nright = 0; w->overflow = 0; }

// Feed rows in increasing y; (lo, hi) are the extreme x of the (dilated) set on that row.
template <class W>
MB_HD void mb_hull_push_row(W* w, int y, int lo, int hi) {
    MbPt pl = {lo, y}, pr = {hi, y};
    typename W::pt_t sl, sr;
    sl.x = lo; sl.y = y; sr.x = hi; sr.y = y;
    while (w->nleft >= 2 && mb_cross(mb_pt(w->left[w->nleft - 2]), mb_pt(w->left[w->nleft - 1]), pl) >= 0) w->nleft--;
    if (w->nleft < W::cap) w->left[w->nleft++] = sl; else w->overflow = 1;
    while (w->nright >= 2 && mb_cross(mb_pt(w->right[w->nright - 2]), mb_pt(w->right[w->nright - 1]), pr) <= 0) w->nright--;
    if (w->nright < W::cap) w->right[w->nright++] = sr; else w->overflow = 1;
}

MB_HD int mb_sign_ll(long long v) { return (v > 0) - (v < 0); }
MB_HD int mb_sign_i(int v) { return (v > 0) - (v < 0); }

// OpenCV Sklansky_ over points sorted by (x, y).  Returns the stack size; stack holds indices into arr.
template <class PT, class IX>
MB_HD int mb_sklansky(const PT* arr, int start, int end, IX* stack, int nsign, int sign2) {
    int incr = end > start ? 1 : -1;
    int pprev = start, pcur = pprev + incr, pnext = pcur + incr;
    int stacksize = 3;
    if (start == end || (arr[start].x == arr[end].x && arr[start].y == arr[end].y)) {
        stack[0] = start;
        return 1;
    }
    stack[0] = pprev; stack[1] = pcur; stack[2] = pnext;
    end += incr;
    while (pnext != end) {
        int cury = arr[pcur].y, nexty = arr[pnext].y;
        int by = nexty - cury;
        if (mb_sign_i(by) != nsign) {
            int ax = arr[pcur].x - arr[pprev].x;
            int bx = arr[pnext].x - arr[pcur].x;
            int ay = cury - arr[pprev].y;
            long long convexity = (long long)ay * bx - (long long)ax * by;
            if (mb_sign_ll(convexity) == sign2 && (ax != 0 || ay != 0)) {
                pprev = pcur; pcur = pnext; pnext += incr;
                stack[stacksize] = pnext; stacksize++;
            } else if (pprev == start) {
                pcur = pnext; stack[1] = pcur; pnext += incr; stack[2] = pnext;
            } else {
                stack[stacksize - 2] = pnext;
                pcur = pprev;
                pprev = stack[stacksize - 4];
                stacksize--;
            }
        } else {
            pnext += incr;
            stack[stacksize - 1] = pnext;
        }
    }
    return --stacksize;
}

template <class PT>
MB_HD long long mb_key(PT p) { return ((long long)p.y << 20) + p.x; }   // raster rank of the pixel

// Builds the OpenCV convexHull(points, clockwise=false) vertex order from the two chains.
// Result: w->order[0..n) indexes w->v; returns n.
template <class W>
MB_HD int mb_hull_finish(W* w) {
    typedef typename W::pt_t PT;
    typedef typename W::ix_t IX;
    // merge chains, dropping duplicates (first/last rows may contribute the same point twice)
    int nv = 0;
    for (int i = 0; i < w->nleft; ++i) w->v[nv++] = w->left[i];
    for (int i = 0; i < w->nright; ++i) {
        PT p = w->right[i];
        bool dup = false;
        for (int j = 0; j < w->nleft; ++j)
            if (w->left[j].x == p.x && w->left[j].y == p.y) { dup = true; break; }
        if (!dup) w->v[nv++] = p;
    }
    // insertion sort by (x, y)
    for (int i = 1; i < nv; ++i) {
        PT p = w->v[i];
        int j = i - 1;
        while (j >= 0 && (w->v[j].x > p.x || (w->v[j].x == p.x && w->v[j].y > p.y))) { w->v[j + 1] = w->v[j]; --j; }
        w->v[j + 1] = p;
    }
    w->nv = nv;
    const PT* arr = w->v;
    int total = nv;
    if (total == 0) return 0;
    int miny_ind = 0, maxy_ind = 0;
    for (int i = 1; i < total; ++i) {
        int y = arr[i].y;
        if (arr[miny_ind].y > y) miny_ind = i;
        if (arr[maxy_ind].y < y) maxy_ind = i;
    }
    IX* hull = w->order;
    int nout = 0;
    if (arr[0].x == arr[total - 1].x && arr[0].y == arr[total - 1].y) {
        hull[nout++] = 0;
        return nout;
    }
    IX* stack = w->stack;
    IX* tl_stack = stack;
    int tl_count = mb_sklansky(arr, 0, maxy_ind, tl_stack, -1, 1);
    IX* tr_stack = stack + tl_count;
    int tr_count = mb_sklansky(arr, total - 1, maxy_ind, tr_stack, -1, -1);
    {   // clockwise == false: swap
        IX* t = tl_stack; tl_stack = tr_stack; tr_stack = t;
        int c = tl_count; tl_count = tr_count; tr_count = c;
    }
    for (int i = 0; i < tl_count - 1; ++i) hull[nout++] = tl_stack[i];
    for (int i = tr_count - 1; i > 0; --i) hull[nout++] = tr_stack[i];
    int stop_idx = tr_count > 2 ? tr_stack[1] : tl_count > 2 ? tl_stack[tl_count - 2] : -1;

    IX* bl_stack = w->stack2;
    int bl_count = mb_sklansky(arr, 0, miny_ind, bl_stack, 1, -1);
    IX* br_stack = w->stack2 + bl_count;
    int br_count = mb_sklansky(arr, total - 1, miny_ind, br_stack, 1, 1);
    if (stop_idx >= 0) {
        int check_idx = bl_count > 2 ? bl_stack[1] : bl_count + br_count > 2 ? br_stack[2 - bl_count] : -1;
        if (check_idx == stop_idx ||
            (check_idx >= 0 && arr[check_idx].x == arr[stop_idx].x && arr[check_idx].y == arr[stop_idx].y)) {
            bl_count = bl_count < 2 ? bl_count : 2;
            br_count = br_count < 2 ? br_count : 2;
        }
    }
    for (int i = 0; i < bl_count - 1; ++i) hull[nout++] = bl_stack[i];
    for (int i = br_count - 1; i > 0; --i) hull[nout++] = br_stack[i];

    // cyclic shift so that original indices (raster ranks) ascend / descend where possible
    if (nout >= 3) {
        int min_idx = 0, max_idx = 0, lt = 0;
        int i;
        for (i = 1; i < nout; ++i) {
            long long idx = mb_key(arr[hull[i]]);
            lt += mb_key(arr[hull[i - 1]]) < idx;
            if (lt > 1 && lt <= i - 2) break;
            if (idx < mb_key(arr[hull[min_idx]])) min_idx = i;
            if (idx > mb_key(arr[hull[max_idx]])) max_idx = i;
        }
        int mmdist = max_idx - min_idx; if (mmdist < 0) mmdist = -mmdist;
        if ((mmdist == 1 || mmdist == nout - 1) && (lt <= 1 || lt >= nout - 2)) {
            int ascending = (max_idx + 1) % nout == min_idx;
            int i0 = ascending ? min_idx : max_idx, j = i0;
            if (i0 > 0) {
                IX* tmp = w->stack;
                for (i = 0; i < nout; ++i) {
                    int curr = hull[j];
                    tmp[i] = curr;
                    int next_j = j + 1 < nout ? j + 1 : 0;
                    long long ck = mb_key(arr[curr]), nk = mb_key(arr[hull[next_j]]);
                    if (i < nout - 1 && (ascending != (ck < nk))) break;
                    j = next_j;
                }
                if (i == nout)
                    for (i = 0; i < nout; ++i) hull[i] = tmp[i];
            }
        }
    }
    return nout;
}

// OpenCV rotatingCalipers(CALIPERS_MINAREARECT). px/py: hull points (float). out[6].
template <class W>
MB_HD void mb_rotating_calipers(W* w, int n, float* out) {
    const float* px = w->hx; const float* py = w->hy;
    float* vx = w->vx; float* vy = w->vy; float* inv = w->inv;
    float minarea = 3.402823466e+38f;
    int left = 0, bottom = 0, right = 0, top = 0;
    float pt0x = px[0], pt0y = py[0];
    float left_x = pt0x, right_x = pt0x, top_y = pt0y, bottom_y = pt0y;
    for (int i = 0; i < n; ++i) {
        if (pt0x < left_x) { left_x = pt0x; left = i; }
        if (pt0x > right_x) { right_x = pt0x; right = i; }
        if (pt0y > top_y) { top_y = pt0y; top = i; }
        if (pt0y < bottom_y) { bottom_y = pt0y; bottom = i; }
        int ni = (i + 1 < n) ? i + 1 : 0;
        float ptx = px[ni], pty = py[ni];
        double dx = (double)ptx - (double)pt0x;
        double dy = (double)pty - (double)pt0y;
        vx[i] = (float)dx;
        vy[i] = (float)dy;
        inv[i] = (float)(1. / sqrt(dx * dx + dy * dy));
        pt0x = ptx; pt0y = pty;
    }
    float orientation = 0.f;
    {
        double ax = vx[n - 1], ay = vy[n - 1];
        for (int i = 0; i < n; ++i) {
            double bx = vx[i], by = vy[i];
            double convexity = ax * by - ay * bx;
            if (convexity != 0) { orientation = (convexity > 0) ? 1.f : -1.f; break; }
            ax = bx; ay = by;
        }
    }
    float base_a = orientation, base_b = 0.f;
    int seq[4] = {bottom, right, top, left};
    int b_left = 0, b_bottom = 0;
    float b_a = 0.f, b_w = 0.f, b_b = 0.f, b_h = 0.f;
    for (int k = 0; k < n; ++k) {
        float dp0 = +base_a * vx[seq[0]] + base_b * vy[seq[0]];
        float dp1 = -base_b * vx[seq[1]] + base_a * vy[seq[1]];
        float dp2 = -base_a * vx[seq[2]] - base_b * vy[seq[2]];
        float dp3 = +base_b * vx[seq[3]] - base_a * vy[seq[3]];
        float maxcos = dp0 * inv[seq[0]];
        int main_element = 0;
        float c1 = dp1 * inv[seq[1]];
        if (c1 > maxcos) { main_element = 1; maxcos = c1; }
        float c2 = dp2 * inv[seq[2]];
        if (c2 > maxcos) { main_element = 2; maxcos = c2; }
        float c3 = dp3 * inv[seq[3]];
        if (c3 > maxcos) { main_element = 3; maxcos = c3; }
        int pindex = seq[main_element];
        float lead_x = vx[pindex] * inv[pindex];
        float lead_y = vy[pindex] * inv[pindex];
        switch (main_element) {
            case 0: base_a = lead_x; base_b = lead_y; break;
            case 1: base_a = lead_y; base_b = -lead_x; break;
            case 2: base_a = -lead_x; base_b = -lead_y; break;
            default: base_a = -lead_y; base_b = lead_x; break;
        }
        seq[main_element] += 1;
        if (seq[main_element] == n) seq[main_element] = 0;
        float dx = px[seq[1]] - px[seq[3]];
        float dy = py[seq[1]] - py[seq[3]];
        float width = dx * base_a + dy * base_b;
        dx = px[seq[2]] - px[seq[0]];
        dy = py[seq[2]] - py[seq[0]];
        float height = -dx * base_b + dy * base_a;
        float area = width * height;
        if (area <= minarea) {
            minarea = area;
            b_left = seq[3]; b_a = base_a; b_w = width; b_b = base_b; b_h = height; b_bottom = seq[0];
        }
    }
    float A1 = b_a, B1 = b_b, A2 = -b_b, B2 = b_a;
    float C1 = A1 * px[b_left] + py[b_left] * B1;
    float C2 = A2 * px[b_bottom] + py[b_bottom] * B2;
    float idet = 1.f / (A1 * B2 - A2 * B1);
    out[0] = (C1 * B2 - C2 * B1) * idet;
    out[1] = (A1 * C2 - A2 * C1) * idet;
    out[2] = A1 * b_w; out[3] = B1 * b_w;
    out[4] = A2 * b_h; out[5] = B2 * b_h;
}

// cv2.minAreaRect (4.13: angle in [-90,0)) + cv2.boxPoints on the finished hull; box[8] = 4 x (x,y).
template <class W>
MB_HD void mb_min_area_box(W* w, int nh, float* box) {
    for (int i = 0; i < nh; ++i) { w->hx[i] = (float)w->v[w->order[i]].x; w->hy[i] = (float)w->v[w->order[i]].y; }
    float cx, cy, bw, bh;
    double deg;
    if (nh > 2) {
        float o[6];
        mb_rotating_calipers(w, nh, o);
        cx = o[0] + (o[2] + o[4]) * 0.5f;
        cy = o[1] + (o[3] + o[5]) * 0.5f;
        bw = (float)sqrt((double)o[2] * o[2] + (double)o[3] * o[3]);
        bh = (float)sqrt((double)o[4] * o[4] + (double)o[5] * o[5]);
        deg = atan2((double)o[3], (double)o[2]) * 180 / 3.141592653589793238462643383279502884;
    } else if (nh == 2) {
        cx = (w->hx[0] + w->hx[1]) * 0.5f;
        cy = (w->hy[0] + w->hy[1]) * 0.5f;
        double dx = (double)(w->hx[1] - w->hx[0]), dy = (double)(w->hy[1] - w->hy[0]);
        bw = (float)sqrt(dx * dx + dy * dy);
        bh = 0.f;
        deg = atan2(dy, dx) * 180 / 3.141592653589793238462643383279502884;
    } else {
        cx = w->hx[0]; cy = w->hy[0]; bw = bh = 0.f; deg = 0.0;
    }
    while (deg >= 0) { deg -= 90; float t = bw; bw = bh; bh = t; }
    while (deg < -90) { deg += 90; float t = bw; bw = bh; bh = t; }
    float angle = (float)deg;
    double rad = (double)angle * 3.141592653589793238462643383279502884 / 180.;
    float b = (float)cos(rad) * 0.5f;
    float a = (float)sin(rad) * 0.5f;
    box[0] = cx - a * bh - b * bw;
    box[1] = cy + b * bh - a * bw;
    box[2] = cx + a * bh - b * bw;
    box[3] = cy - b * bh - a * bw;
    box[4] = 2 * cx - box[0];
    box[5] = 2 * cy - box[1];
    box[6] = 2 * cx - box[2];
    box[7] = 2 * cy - box[3];
}

// Diamond fix + roll (craft_utils.py:81-93).  l,r,t,b: extremes of the dilated point set.
MB_HD void mb_diamond_roll(float* box, int l, int r, int t, int b) {
    // np.linalg.norm on float32 vectors: sqrt(sum of squares) in float32
    float d0x = box[0] - box[2], d0y = box[1] - box[3];
    float d1x = box[2] - box[4], d1y = box[3] - box[5];
    float wn = sqrtf(d0x * d0x + d0y * d0y);
    float hn = sqrtf(d1x * d1x + d1y * d1y);
    // max(w,h) / (min(w,h) + 1e-5): np.float32 + python float -> float32 (NumPy 2 weak scalars)
    float mx = wn > hn ? wn : hn, mn = wn > hn ? hn : wn;
    float ratio = mx / (mn + 1e-5f);
    float dev = 1.0f - ratio;
    if (dev < 0) dev = -dev;
    if (dev <= 0.1f) {
        box[0] = (float)l; box[1] = (float)t;
        box[2] = (float)r; box[3] = (float)t;
        box[4] = (float)r; box[5] = (float)b;
        box[6] = (float)l; box[7] = (float)b;
    }
    // startidx = argmin(x+y) (first minimum), roll so it comes first
    int s = 0;
    float best = box[0] + box[1];
    for (int i = 1; i < 4; ++i) {
        float v = box[2 * i] + box[2 * i + 1];
        if (v < best) { best = v; s = i; }
    }
    float tmp[8];
    for (int i = 0; i < 4; ++i) { tmp[2 * i] = box[2 * ((i + s) & 3)]; tmp[2 * i + 1] = box[2 * ((i + s) & 3) + 1]; }
    for (int i = 0; i < 8; ++i) box[i] = tmp[i];
}

// adjustResultCoordinates (f32 *= f64 -> f32) then the rect of craft_box_processor.py:500-521.
MB_HD void mb_adjust_and_rect(const float* box, double sx, double sy, int img_w, int img_h, float* adj, int* rect) {
    int minx = 0, miny = 0, maxx = 0, maxy = 0;
    for (int i = 0; i < 4; ++i) {
        float ax = (float)((double)box[2 * i] * sx);
        float ay = (float)((double)box[2 * i + 1] * sy);
        adj[2 * i] = ax; adj[2 * i + 1] = ay;
        int ix = (int)ax, iy = (int)ay;   // astype(np.int32): truncation toward zero
        if (i == 0) { minx = maxx = ix; miny = maxy = iy; }
        else {
            minx = ix < minx ? ix : minx; maxx = ix > maxx ? ix : maxx;
            miny = iy < miny ? iy : miny; maxy = iy > maxy ? iy : maxy;
        }
    }
    int bw = maxx - minx + 1, bh = maxy - miny + 1;   // cv2.boundingRect
    rect[0] = minx - 2 > 0 ? minx - 2 : 0;
    rect[1] = miny - 2 > 0 ? miny - 2 : 0;
    rect[2] = bw + 4 < img_w ? bw + 4 : img_w;
    rect[3] = bh + 4 < img_h ? bh + 4 : img_h;
}

// One component: rowmin/rowmax hold the undilated segmap extremes for bbox rows [y0, y0+h) (empty row: min > max).
// ROI [sx,ex) x [sy,ey), dilation (1+niter)^2 with OpenCV's anchor.  Writes det box (heat-map coords).
template <class W>
MB_HD int mb_component_box(W* w, const short* rowmin, const short* rowmax, int y0, int h, int sx, int ex,
                           int sy, int ey, int niter, float* box) {
    const int anchor = (1 + niter) / 2;
    const int back = niter - anchor;
    mb_hull_begin(w);
    int l = 0x7fffffff, r = -1, t = 0x7fffffff, b = -1;
    for (int y = sy; y < ey; ++y) {
        int lo = 0x7fffffff, hi = -1;
        int r0 = y - anchor, r1 = y + back;
        if (r0 < y0) r0 = y0;
        if (r1 > y0 + h - 1) r1 = y0 + h - 1;
        for (int rr = r0; rr <= r1; ++rr) {
            int a = rowmin[rr - y0], c = rowmax[rr - y0];
            if (a <= c) { lo = a < lo ? a : lo; hi = c > hi ? c : hi; }
        }
        if (hi < 0) continue;
        lo = lo - back > sx ? lo - back : sx;
        hi = hi + anchor < ex - 1 ? hi + anchor : ex - 1;
        mb_hull_push_row(w, y, lo, hi);
        l = lo < l ? lo : l; r = hi > r ? hi : r;
        t = y < t ? y : t; b = y > b ? y : b;
    }
    if (w->nleft == 0) return 0;
    int nh = mb_hull_finish(w);
    mb_min_area_box(w, nh, box);
    mb_diamond_roll(box, l, r, t, b);
    return w->overflow ? -1 : 1;
}
