"""CPU restatement of CRAFT post-processing (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows the reference line by line, but without its full-image masks per label:
  getDetBoxes_core            marie/models/craft/craft_utils.py:25-98
  getDetBoxes                 marie/models/craft/craft_utils.py:257-265
  adjustResultCoordinates     marie/models/craft/craft_utils.py:268-274
  box -> rect -> crop loop    marie/boxes/craft_box_processor.py:499-537, crop_poly_low :42-73

Two flavours are provided:
  * det_boxes_cv(...)      — uses the same cv2 calls as the reference (threshold, connectedComponentsWithStats,
                             dilate, minAreaRect, boxPoints) restricted to each component's ROI.  Pinned against the
                             reference's own function in tests/test_oracle_vs_reference.py and the committed goldens.
  * det_boxes_restated(...) — no cv2 geometry: union-find labelling in raster order, analytic rect dilation,
                             Sklansky hull + rotating calipers in float32 written out from OpenCV's published
                             algorithm (imgproc/convhull.cpp, rotcalipers.cpp; OpenCV 4.13 behaviour probed here).
                             This is the algorithm the CUDA kernels implement (csrc/ccl.cu, csrc/boxes.cu).
"""
import math

import cv2
import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------------------------- cv2 flavour
def label_maps(textmap, linkmap, link_threshold, low_text):
    """craft_utils.py:32-38 — strict '>' thresholds, OR, 4-connected labelling with stats."""
    _, text_score = cv2.threshold(textmap, low_text, 1, 0)
    _, link_score = cv2.threshold(linkmap, link_threshold, 1, 0)
    comb = np.clip(text_score + link_score, 0, 1)
    n_labels, labels, stats, _ = cv2.connectedComponentsWithStats(comb.astype(np.uint8), connectivity=4)
    return n_labels, labels, stats, text_score, link_score


def det_boxes_cv(textmap, linkmap, text_threshold, link_threshold, low_text):
    """Returns (det [N,4,2] f32, labels i32 [H,W], mapper [N]) exactly like getDetBoxes_core (:25-98)."""
    img_h, img_w = textmap.shape
    n_labels, labels, stats, text_score, link_score = label_maps(textmap, linkmap, link_threshold, low_text)
    det, mapper = [], []
    for k in range(1, n_labels):
        size = stats[k, cv2.CC_STAT_AREA]
        if size < 10:                                                   # :50
            continue
        x, y = int(stats[k, cv2.CC_STAT_LEFT]), int(stats[k, cv2.CC_STAT_TOP])
        w, h = int(stats[k, cv2.CC_STAT_WIDTH]), int(stats[k, cv2.CC_STAT_HEIGHT])
        comp = labels[y:y + h, x:x + w] == k
        if np.max(textmap[y:y + h, x:x + w][comp]) < text_threshold:    # :54 (float32 compare)
            continue
        niter = int(math.sqrt(int(size) * min(w, h) / (w * h)) * 2)    # :63
        sx, ex, sy, ey = x - niter, x + w + niter + 1, y - niter, y + h + niter + 1
        sx, sy = max(sx, 0), max(sy, 0)
        ex, ey = min(ex, img_w), min(ey, img_h)                         # :66-73
        roi_lab = labels[sy:ey, sx:ex]
        seg = np.where(roi_lab == k, 255, 0).astype(np.uint8)           # :57-58
        rm = np.logical_and(link_score[sy:ey, sx:ex] == 1, text_score[sy:ey, sx:ex] == 0)
        seg[rm] = 0                                                     # :60
        kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (1 + niter, 1 + niter))
        seg = cv2.dilate(seg, kernel)                                   # :74-75
        ys, xs = np.where(seg != 0)
        pts = np.stack([xs + sx, ys + sy], axis=1).astype(np.int64)     # :78 (x, y) in raster order
        rect = cv2.minAreaRect(pts)
        box = cv2.boxPoints(rect)
        bw, bh = np.linalg.norm(box[0] - box[1]), np.linalg.norm(box[1] - box[2])
        ratio = max(bw, bh) / (min(bw, bh) + 1e-5)
        if abs(1 - ratio) <= 0.1:                                       # :83-88 diamond fix
            l, r = pts[:, 0].min(), pts[:, 0].max()
            t, b = pts[:, 1].min(), pts[:, 1].max()
            box = np.array([[l, t], [r, t], [r, b], [l, b]], dtype=np.float32)
        start = box.sum(axis=1).argmin()                                # :91-92
        box = np.roll(box, 4 - start, 0)
        det.append(np.array(box))
        mapper.append(k)
    return det, labels, mapper


def adjust_result_coordinates(polys, ratio_w, ratio_h, ratio_net=2):
    """craft_utils.py:268-274 — in-place f32 *= (f64, f64)."""
    if len(polys) > 0:
        polys = np.array(polys)
        for k in range(len(polys)):
            if polys[k] is not None:
                polys[k] *= (ratio_w * ratio_net, ratio_h * ratio_net)
    return polys


def boxes_to_rects(bboxes, max_h, max_w):
    """craft_box_processor.py:499-521 — int32 truncation, boundingRect, (-2,-2,+4,+4) expansion with clamps."""
    rects = []
    for region in bboxes:
        region = np.array(region).astype(np.int32).reshape(-1, 2)
        x0, y0 = region[:, 0].min(), region[:, 1].min()
        bw, bh = region[:, 0].max() - x0 + 1, region[:, 1].max() - y0 + 1     # cv2.boundingRect of points
        rects.append([max(0, int(x0) - 2), max(0, int(y0) - 2), min(max_w, int(bw) + 4), min(max_h, int(bh) + 4)])
    return rects


def crop_rect(image, rect):
    """crop_poly_low (:42-73) on the axis-aligned expanded rect == image[y:y+h+1, x:x+w+1] (numpy clips)."""
    x, y, w, h = rect
    return image[y:y + h + 1, x:x + w + 1].copy()


# --------------------------------------------------------------------------------------------- restated flavour
def label_restated(fg):
    """4-connected labelling; ids in raster order of each component's first pixel (== cv2's numbering).
    Union-find with min-index roots, then rank of roots.  Returns (n_labels, labels i32, stats[n,5])."""
    h, w = fg.shape
    parent = np.arange(h * w, dtype=np.int64)
    fgf = fg.reshape(-1)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for i in np.flatnonzero(fgf):
        yy, xx = divmod(int(i), w)
        if xx > 0 and fgf[i - 1]:
            a, b = find(i), find(i - 1)
            if a != b:
                parent[max(a, b)] = min(a, b)
        if yy > 0 and fgf[i - w]:
            a, b = find(i), find(i - w)
            if a != b:
                parent[max(a, b)] = min(a, b)
    labels = np.zeros(h * w, np.int32)
    roots = {}
    for i in np.flatnonzero(fgf):
        r = find(int(i))
        if r not in roots:
            roots[r] = len(roots) + 1      # roots are met in raster order because root = min index
        labels[i] = roots[r]
    n = len(roots) + 1
    labels = labels.reshape(h, w)
    stats = np.zeros((n, 5), np.int32)
    for k in range(1, n):
        ys, xs = np.where(labels == k)
        stats[k] = [xs.min(), ys.min(), xs.max() - xs.min() + 1, ys.max() - ys.min() + 1, len(xs)]
    return n, labels, stats


def _sign(v):
    return (v > 0) - (v < 0)


def _sklansky(arr, start, end, nsign, sign2):
    incr = 1 if end > start else -1
    pprev, pcur, pnext = start, start + incr, start + 2 * incr
    if start == end or (arr[start][0] == arr[end][0] and arr[start][1] == arr[end][1]):
        return [start]
    stack = [pprev, pcur, pnext] + [0] * (len(arr) + 2)
    size = 3
    end += incr
    while pnext != end:
        cury, nexty = arr[pcur][1], arr[pnext][1]
        by = nexty - cury
        if _sign(by) != nsign:
            ax = arr[pcur][0] - arr[pprev][0]
            bx = arr[pnext][0] - arr[pcur][0]
            ay = cury - arr[pprev][1]
            conv = ay * bx - ax * by
            if _sign(conv) == sign2 and (ax != 0 or ay != 0):
                pprev, pcur = pcur, pnext
                pnext += incr
                stack[size] = pnext
                size += 1
            elif pprev == start:
                pcur = pnext
                stack[1] = pcur
                pnext += incr
                stack[2] = pnext
            else:
                stack[size - 2] = pnext
                pcur = pprev
                pprev = stack[size - 4]
                size -= 1
        else:
            pnext += incr
            stack[size - 1] = pnext
    return stack[:size - 1]


def convex_hull_restated(points, keys):
    """OpenCV convexHull(points, clockwise=False) order.  `keys[i]` is the original (raster) rank of point i;
    only its ordering matters (final cyclic-shift rule).  Returns indices into `points`."""
    total = len(points)
    order = sorted(range(total), key=lambda i: (points[i][0], points[i][1]))
    arr = [points[i] for i in order]
    miny = maxy = 0
    for i in range(1, total):
        if arr[miny][1] > arr[i][1]:
            miny = i
        if arr[maxy][1] < arr[i][1]:
            maxy = i
    if arr[0][0] == arr[-1][0] and arr[0][1] == arr[-1][1]:
        return [order[0]]
    tl = _sklansky(arr, 0, maxy, -1, 1)
    tr = _sklansky(arr, total - 1, maxy, -1, -1)
    tl, tr = tr, tl                                   # clockwise == False
    hull = [order[tl[i]] for i in range(len(tl) - 1)]
    hull += [order[tr[i]] for i in range(len(tr) - 1, 0, -1)]
    stop_idx = tr[1] if len(tr) > 2 else (tl[len(tl) - 2] if len(tl) > 2 else -1)
    bl = _sklansky(arr, 0, miny, 1, -1)
    br = _sklansky(arr, total - 1, miny, 1, 1)
    if stop_idx >= 0:
        check_idx = bl[1] if len(bl) > 2 else (br[2 - len(bl)] if len(bl) + len(br) > 2 else -1)
        if check_idx == stop_idx or (check_idx >= 0 and arr[check_idx][0] == arr[stop_idx][0]
                                     and arr[check_idx][1] == arr[stop_idx][1]):
            bl, br = bl[:min(len(bl), 2)], br[:min(len(br), 2)]
    hull += [order[bl[i]] for i in range(len(bl) - 1)]
    hull += [order[br[i]] for i in range(len(br) - 1, 0, -1)]
    nout = len(hull)
    if nout >= 3:
        hk = [keys[i] for i in hull]
        min_idx = max_idx = lt = 0
        i = 1
        while i < nout:
            lt += hk[i - 1] < hk[i]
            if 1 < lt <= i - 2:
                break
            if hk[i] < hk[min_idx]:
                min_idx = i
            if hk[i] > hk[max_idx]:
                max_idx = i
            i += 1
        mmdist = abs(max_idx - min_idx)
        if (mmdist == 1 or mmdist == nout - 1) and (lt <= 1 or lt >= nout - 2):
            ascending = (max_idx + 1) % nout == min_idx
            i0 = min_idx if ascending else max_idx
            if i0 > 0:
                j, st, ok = i0, [], True
                for i in range(nout):
                    st.append(hull[j])
                    nj = j + 1 if j + 1 < nout else 0
                    if i < nout - 1 and (ascending != (hk[j] < hk[nj])):
                        ok = False
                        break
                    j = nj
                if ok:
                    hull = st
    return hull


def rotating_calipers_restated(pts):
    """OpenCV rotatingCalipers(CALIPERS_MINAREARECT) in float32; pts = hull as (f32 x, f32 y)."""
    n = len(pts)
    minarea = f32(np.finfo(np.float32).max)
    vect, inv = [None] * n, [None] * n
    left = bottom = right = top = 0
    pt0 = pts[0]
    left_x = right_x = pt0[0]
    top_y = bottom_y = pt0[1]
    for i in range(n):
        if pt0[0] < left_x:
            left_x, left = pt0[0], i
        if pt0[0] > right_x:
            right_x, right = pt0[0], i
        if pt0[1] > top_y:
            top_y, top = pt0[1], i
        if pt0[1] < bottom_y:
            bottom_y, bottom = pt0[1], i
        pt = pts[i + 1 if i + 1 < n else 0]
        dx, dy = float(pt[0]) - float(pt0[0]), float(pt[1]) - float(pt0[1])
        vect[i] = (f32(dx), f32(dy))
        inv[i] = f32(1.0 / math.sqrt(dx * dx + dy * dy))
        pt0 = pt
    orientation = f32(0)
    ax, ay = float(vect[n - 1][0]), float(vect[n - 1][1])
    for i in range(n):
        bx, by = float(vect[i][0]), float(vect[i][1])
        conv = ax * by - ay * bx
        if conv != 0:
            orientation = f32(1.0) if conv > 0 else f32(-1.0)
            break
        ax, ay = bx, by
    base_a, base_b = orientation, f32(0)
    seq = [bottom, right, top, left]
    buf = None
    for _ in range(n):
        dp = [+base_a * vect[seq[0]][0] + base_b * vect[seq[0]][1],
              -base_b * vect[seq[1]][0] + base_a * vect[seq[1]][1],
              -base_a * vect[seq[2]][0] - base_b * vect[seq[2]][1],
              +base_b * vect[seq[3]][0] - base_a * vect[seq[3]][1]]
        maxcos = dp[0] * inv[seq[0]]
        main = 0
        for i in range(1, 4):
            c = dp[i] * inv[seq[i]]
            if c > maxcos:
                main, maxcos = i, c
        pidx = seq[main]
        lead_x, lead_y = vect[pidx][0] * inv[pidx], vect[pidx][1] * inv[pidx]
        if main == 0:
            base_a, base_b = lead_x, lead_y
        elif main == 1:
            base_a, base_b = lead_y, -lead_x
        elif main == 2:
            base_a, base_b = -lead_x, -lead_y
        else:
            base_a, base_b = -lead_y, lead_x
        seq[main] += 1
        if seq[main] == n:
            seq[main] = 0
        dx, dy = pts[seq[1]][0] - pts[seq[3]][0], pts[seq[1]][1] - pts[seq[3]][1]
        width = dx * base_a + dy * base_b
        dx, dy = pts[seq[2]][0] - pts[seq[0]][0], pts[seq[2]][1] - pts[seq[0]][1]
        height = -dx * base_b + dy * base_a
        area = width * height
        if area <= minarea:
            minarea = area
            buf = (seq[3], base_a, width, base_b, height, seq[0])
    a1, b1, a2, b2 = buf[1], buf[3], -buf[3], buf[1]
    c1 = a1 * pts[buf[0]][0] + pts[buf[0]][1] * b1
    c2 = a2 * pts[buf[5]][0] + pts[buf[5]][1] * b2
    idet = f32(1.0) / (a1 * b2 - a2 * b1)
    px, py = (c1 * b2 - c2 * b1) * idet, (a1 * c2 - a2 * c1) * idet
    return [px, py, a1 * buf[2], b1 * buf[2], a2 * buf[4], b2 * buf[4]]


def min_area_rect_restated(points, keys):
    """cv2.minAreaRect (OpenCV 4.13 behaviour: angle normalised into [-90, 0)) -> ((cx,cy),(w,h),angle) f32."""
    hull = convex_hull_restated(points, keys)
    hp = [(f32(points[i][0]), f32(points[i][1])) for i in hull]
    n = len(hp)
    if n > 2:
        o = rotating_calipers_restated(hp)
        cx = o[0] + (o[2] + o[4]) * f32(0.5)
        cy = o[1] + (o[3] + o[5]) * f32(0.5)
        w = f32(math.sqrt(float(o[2]) * float(o[2]) + float(o[3]) * float(o[3])))
        h = f32(math.sqrt(float(o[4]) * float(o[4]) + float(o[5]) * float(o[5])))
        d = math.atan2(float(o[3]), float(o[2])) * 180 / math.pi
    elif n == 2:
        cx, cy = (hp[0][0] + hp[1][0]) * f32(0.5), (hp[0][1] + hp[1][1]) * f32(0.5)
        dx, dy = float(hp[1][0] - hp[0][0]), float(hp[1][1] - hp[0][1])
        w, h = f32(math.sqrt(dx * dx + dy * dy)), f32(0)
        d = math.atan2(dy, dx) * 180 / math.pi
    else:
        cx, cy = hp[0]
        w = h = f32(0)
        d = 0.0
    while d >= 0:
        d -= 90
        w, h = h, w
    while d < -90:
        d += 90
        w, h = h, w
    return (cx, cy), (w, h), f32(d)


def box_points_restated(center, size, angle):
    """cv::RotatedRect::points (double cos/sin cast to float, then float arithmetic)."""
    ang = float(angle) * math.pi / 180.0
    b = f32(math.cos(ang)) * f32(0.5)
    a = f32(math.sin(ang)) * f32(0.5)
    cx, cy = center
    w, h = size
    p0 = (cx - a * h - b * w, cy + b * h - a * w)
    p1 = (cx + a * h - b * w, cy - b * h - a * w)
    p2 = (f32(2) * cx - p0[0], f32(2) * cy - p0[1])
    p3 = (f32(2) * cx - p1[0], f32(2) * cy - p1[1])
    return np.array([p0, p1, p2, p3], dtype=np.float32)


def hull_reduce_rows(cand):
    """Strictly convex vertices of the candidate set, processed in raster (y, then x) order: a left chain over the
    row minima and a right chain over the row maxima (Andrew's monotone chain along y).  Collinear points are
    dropped, matching what OpenCV's Sklansky scan keeps."""
    rows = {}
    for (x, y) in cand:
        lo, hi = rows.get(y, (x, x))
        rows[y] = (min(lo, x), max(hi, x))
    ys = sorted(rows)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    left, right = [], []
    for y in ys:
        pl, pr = (rows[y][0], y), (rows[y][1], y)
        while len(left) >= 2 and cross(left[-2], left[-1], pl) >= 0:
            left.pop()
        left.append(pl)
        while len(right) >= 2 and cross(right[-2], right[-1], pr) <= 0:
            right.pop()
        right.append(pr)
    out = list(left)
    for p in right:
        if p not in out:
            out.append(p)
    # the two chains meet on the first and last rows: drop points made redundant there (collinear on the flat edge)
    return out


def component_box_restated(rows, sx, ex, sy, ey, niter):
    """rows: dict y -> (xmin, xmax) of the undilated segmap (component ∧ text>low).  Applies the (1+niter)^2
    rect dilation analytically inside the ROI [sx,ex)x[sy,ey) and returns the 4x2 f32 box after the diamond fix
    and the roll (craft_utils.py:74-93).  Only per-row extremes of the dilated set are needed: every other
    pixel is interior to its row segment and cannot be a hull vertex."""
    anchor = (1 + niter) // 2
    back = niter - anchor
    cand = []
    for y in range(sy, ey):
        lo, hi = None, None
        for r in range(y - anchor, y + back + 1):          # source rows r light dst rows [r-back, r+anchor]
            if r in rows:
                a, b = rows[r]
                lo = a if lo is None else min(lo, a)
                hi = b if hi is None else max(hi, b)
        if lo is None:
            continue
        lo, hi = max(lo - back, sx), min(hi + anchor, ex - 1)
        cand.append((lo, y))
        if hi != lo:
            cand.append((hi, y))
    full_cand = cand
    cand = hull_reduce_rows(cand)      # streaming y-monotone reduction (what the CUDA kernel keeps in smem)
    keys = [p[1] * (1 << 20) + p[0] for p in cand]
    (c, s, ang) = min_area_rect_restated(cand, keys)
    box = box_points_restated(c, s, ang)
    bw = np.linalg.norm(box[0] - box[1])
    bh = np.linalg.norm(box[1] - box[2])
    ratio = max(bw, bh) / (min(bw, bh) + 1e-5)
    if abs(1 - ratio) <= 0.1:
        xs = [p[0] for p in full_cand]
        ys = [p[1] for p in full_cand]
        l, r, t, b = min(xs), max(xs), min(ys), max(ys)
        box = np.array([[l, t], [r, t], [r, b], [l, b]], dtype=np.float32)
    start = box.sum(axis=1).argmin()
    return np.roll(box, 4 - start, 0)


def iter_components(textmap, linkmap, text_threshold, link_threshold, low_text):
    """Yields (label, rows{y:(xmin,xmax)}, (x,y,w,h,size), (sx,ex,sy,ey), niter) for every component that passes
    the reference's filters (craft_utils.py:47-73), plus the labels image first."""
    img_h, img_w = textmap.shape
    text_fg = textmap > f32(low_text)
    link_fg = linkmap > f32(link_threshold)
    n_labels, labels, stats = label_restated(np.logical_or(text_fg, link_fg))
    yield labels, stats
    for k in range(1, n_labels):
        x, y, w, h, size = (int(v) for v in stats[k])
        if size < 10:
            continue
        comp = labels[y:y + h, x:x + w] == k
        if np.max(textmap[y:y + h, x:x + w][comp]) < f32(text_threshold):
            continue
        niter = int(math.sqrt(size * min(w, h) / (w * h)) * 2)
        roi = (max(x - niter, 0), min(x + w + niter + 1, img_w), max(y - niter, 0), min(y + h + niter + 1, img_h))
        seg = np.logical_and(comp, text_fg[y:y + h, x:x + w])
        rows = {}
        for r in range(h):
            xs = np.flatnonzero(seg[r])
            if len(xs):
                rows[y + r] = (x + int(xs[0]), x + int(xs[-1]))
        yield k, rows, (x, y, w, h, size), roi, niter


def det_boxes_restated(textmap, linkmap, text_threshold, link_threshold, low_text):
    """getDetBoxes_core without any cv2 geometry — the algorithm of the CUDA kernels."""
    it = iter_components(textmap, linkmap, text_threshold, link_threshold, low_text)
    labels, _ = next(it)
    det, mapper = [], []
    for k, rows, _, (sx, ex, sy, ey), niter in it:
        det.append(component_box_restated(rows, sx, ex, sy, ey, niter))
        mapper.append(k)
    return det, labels, mapper


def close3x3_restated(mask):
    """cv2.morphologyEx(MORPH_CLOSE, 3x3 rect, iterations=1) on a binary image: dilation then erosion; OpenCV's default
    morphology border never wins (outside pixels count as background for the dilation, as foreground for the erosion)."""
    m = np.asarray(mask, bool)
    h, w = m.shape
    p = np.zeros((h + 2, w + 2), bool)
    p[1:-1, 1:-1] = m
    d = np.zeros((h, w), bool)
    for dy in range(3):
        for dx in range(3):
            d |= p[dy:dy + h, dx:dx + w]
    p = np.ones((h + 2, w + 2), bool)
    p[1:-1, 1:-1] = d
    e = np.ones((h, w), bool)
    for dy in range(3):
        for dx in range(3):
            e &= p[dy:dy + h, dx:dx + w]
    return e


def line_components_cv(linkmap, link_threshold):
    """The line branch of get_prediction up to the component list (marie/boxes/craft_box_processor.py:161-205):
    link > threshold (the text map is thresholded there too, but `text_score_comb` is overwritten by link_score * 255),
    MORPH_CLOSE 3x3, 4-connected components -> (labels i32, boxes [[x, y, w, h], ...] in label order)."""
    import cv2
    _, link_score = cv2.threshold(linkmap, link_threshold, 1, 0)
    comb = link_score * 255
    kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3))
    line_img = cv2.morphologyEx(comb, cv2.MORPH_CLOSE, kernel, iterations=1)
    n, labels, stats, _ = cv2.connectedComponentsWithStats(line_img.astype(np.uint8), connectivity=4)
    boxes = [[int(stats[k, cv2.CC_STAT_LEFT]), int(stats[k, cv2.CC_STAT_TOP]), int(stats[k, cv2.CC_STAT_WIDTH]),
              int(stats[k, cv2.CC_STAT_HEIGHT])] for k in range(1, n)]
    return labels, boxes


def line_boxes(linkmap, link_threshold, ratio_w, ratio_h, ratio_net=2):
    """Line boxes of the refiner branch in page coordinates (craft_box_processor.py:161-217): components -> line_merge
    (marie/boxes/line_processor.py:48-171, restated in oracle/lines.py) -> int() scaling."""
    from oracle import lines as _lines
    _, boxes = line_components_cv(linkmap, link_threshold)
    merged = _lines.line_merge(boxes) if boxes else []
    return [[int(b[0] * ratio_w * ratio_net), int(b[1] * ratio_h * ratio_net), int(b[2] * ratio_w * ratio_net),
             int(b[3] * ratio_h * ratio_net)] for b in merged]
