// K8: line grouping on the host (n <= a few thousand boxes per page; O(n^2) integer work with float64 IoU).
// Reference: find_line_number / line_merge / __line_merge (marie/boxes/line_processor.py:15-171) over
// find_overlap_vertical (marie/utils/overlap.py:42-103).  Visit order and comparison strictness follow the reference.
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace {
struct Box { long long x, y, w, h; };

// indices (ascending) of boxes overlapping `b` vertically, with their clamped vertical IoU
void vertical_overlaps(const Box& b, const std::vector<Box>& data, std::vector<int>& idx, std::vector<double>& iou) {
    idx.clear(); iou.clear();
    if (b.h <= 0) return;
    for (size_t i = 0; i < data.size(); ++i) {
        const Box& o = data[i];
        if (o.h <= 0) continue;
        if (o.x == b.x && o.y == b.y && o.w == b.w && o.h == b.h) continue;     // never match exact duplicates
        if (b.y < o.y + o.h && o.y < b.y + b.h) {
            const long long inter = std::min(b.y + b.h, o.y + o.h) - std::max(b.y, o.y);
            double v = (double)inter / (double)(b.h + o.h - inter);
            v = v > 1.0 ? 1.0 : (v < 0.0 ? 0.0 : v);
            idx.push_back((int)i);
            iou.push_back(v);
        }
    }
}

// argsort by y.  numpy's default argsort (what the reference calls) is not stable and its tie order depends on the
// CPU's SIMD dispatch (x86-simd-sort on AVX-512 / AVX2 hosts), so for boxes sharing the same y the reference itself
// is platform-dependent; ties are resolved here by original order (stable), which is numpy's answer whenever its sort
// happens to be stable and is exact for tie-free inputs.
void np_argsort(const std::vector<long long>& v, std::vector<int>& order) {
    order.resize(v.size());
    for (size_t i = 0; i < v.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return v[a] < v[b]; });
}

std::vector<Box> merge_pass(const std::vector<Box>& in, double min_iou) {
    std::vector<long long> ys(in.size());
    for (size_t i = 0; i < in.size(); ++i) ys[i] = in[i].y;
    std::vector<int> order;
    np_argsort(ys, order);
    std::vector<Box> boxes(in.size());
    for (size_t i = 0; i < in.size(); ++i) boxes[i] = in[order[i]];
    const int n = (int)boxes.size();
    std::vector<char> visited(n, 0);
    std::vector<Box> out;
    std::vector<int> idx, back_idx;
    std::vector<double> iou, back_iou;
    for (int i = 0; i < n; ++i) {
        if (visited[i]) continue;
        visited[i] = 1;
        vertical_overlaps(boxes[i], boxes, idx, iou);
        long long x0 = boxes[i].x, y0 = boxes[i].y, x1 = boxes[i].x + boxes[i].w, hmax = boxes[i].h;
        for (size_t k = 0; k < idx.size(); ++k) {
            const int j = idx[k];
            if (visited[j] || iou[k] < min_iou) continue;
            vertical_overlaps(boxes[j], boxes, back_idx, back_iou);
            if (back_idx.size() == idx.size()) {           // the candidate sees as many overlaps as the source
                visited[j] = 1;
                x0 = std::min(x0, boxes[j].x); y0 = std::min(y0, boxes[j].y);
                x1 = std::max(x1, boxes[j].x + boxes[j].w); hmax = std::max(hmax, boxes[j].h);
            }
        }
        out.push_back({x0, y0, x1 - x0, hmax});            // line height = tallest member, not the union
    }
    return out;
}
}  // namespace

// boxes_host [n,4] i32 (x,y,w,h) -> lines_out_host [<= n, 4] i32 sorted by y; returns count in *n_lines.
extern "C" int mb_line_merge(mb_ctx* ctx, const int32_t* boxes_host, int n, int32_t* lines_out_host, int* n_lines) {
    if (!n_lines || (n > 0 && (!boxes_host || !lines_out_host))) return mb_set_err(ctx, MB_ERR_ARG, "line_merge: null argument");
    *n_lines = 0;
    if (n <= 0) return 0;
    std::vector<Box> cur(n);
    for (int i = 0; i < n; ++i) cur[i] = {boxes_host[4 * i], boxes_host[4 * i + 1], boxes_host[4 * i + 2], boxes_host[4 * i + 3]};
    const double thr[7] = {0.8, 0.7, 0.6, 0.5, 0.4, 0.37, 0.35};
    int unchanged = 0;
    for (int p = 0; p < 7; ++p) {
        const size_t before = cur.size();
        cur = merge_pass(cur, thr[p]);
        if (cur.size() == before && ++unchanged > 2) break;
    }
    std::vector<char> drop(cur.size(), 0);
    for (size_t i = 0; i < cur.size(); ++i)
        for (size_t j = 0; j < cur.size(); ++j) {
            if (i == j) continue;
            const Box &a = cur[i], &b = cur[j];
            if (b.x > a.x && b.x + b.w < a.x + a.w && b.y > a.y && b.y + b.h < a.y + a.h) drop[j] = 1;   // strictly inside
        }
    std::vector<Box> kept;
    for (size_t i = 0; i < cur.size(); ++i) if (!drop[i]) kept.push_back(cur[i]);
    std::vector<long long> ys(kept.size());
    for (size_t i = 0; i < kept.size(); ++i) ys[i] = kept[i].y;
    std::vector<int> order;
    np_argsort(ys, order);
    for (size_t i = 0; i < kept.size(); ++i) {
        const Box& b = kept[order[i]];
        lines_out_host[4 * i] = (int32_t)b.x; lines_out_host[4 * i + 1] = (int32_t)b.y;
        lines_out_host[4 * i + 2] = (int32_t)b.w; lines_out_host[4 * i + 3] = (int32_t)b.h;
    }
    *n_lines = (int)kept.size();
    return 0;
}

// line id (1-based) per box, -1 when there are no lines (find_line_number, line_processor.py:15-45)
extern "C" int mb_find_line_numbers(mb_ctx* ctx, const int32_t* lines_host, int n_lines, const int32_t* boxes_host,
                                    int n_boxes, int32_t* ids_out_host) {
    if (n_boxes > 0 && (!boxes_host || !ids_out_host)) return mb_set_err(ctx, MB_ERR_ARG, "find_line_numbers: null argument");
    std::vector<Box> lines(n_lines > 0 ? n_lines : 0);
    for (int i = 0; i < n_lines; ++i) lines[i] = {lines_host[4 * i], lines_host[4 * i + 1], lines_host[4 * i + 2], lines_host[4 * i + 3]};
    std::vector<int> idx;
    std::vector<double> iou;
    for (int b = 0; b < n_boxes; ++b) {
        const Box box = {boxes_host[4 * b], boxes_host[4 * b + 1], boxes_host[4 * b + 2], boxes_host[4 * b + 3]};
        vertical_overlaps(box, lines, idx, iou);
        int line = -1;
        if (idx.size() == 1) line = idx[0] + 1;
        else if (idx.size() > 1) {
            double best = 0;
            for (size_t k = 0; k < idx.size(); ++k) if (iou[k] > best) { best = iou[k]; line = idx[k] + 1; }
        }
        if (line == -1) {
            long long best = 100000;
            const long long cy = box.y + box.h / 2;       // python floor division; heights are non-negative here
            for (int i = 0; i < n_lines; ++i) {
                long long d = cy - (lines[i].y + lines[i].h);
                d = d < 0 ? -d : d;
                if (d < best) { best = d; line = i + 1; }
            }
        }
        ids_out_host[b] = line;
    }
    return 0;
}
